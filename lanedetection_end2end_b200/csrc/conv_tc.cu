// tcgen05 (5th-gen tensor core) implicit-GEMM kernel for ERFNet's factorised 3-tap convolutions
// (non_bottleneck_1d: conv3x1 / conv1x3 with dilation, BP/Networks/ERFNet.py:29-37,44-53) and
// their input gradients, C in {64, 128}, NHWC fp32 activations, TF32 multiply / fp32 accumulate
// (the arithmetic cuDNN uses for the reference's fp32 convs on Ampere+).
//
//   out[n,y,x,co] = epi( sum_{t<3} sum_{ci} in[n, (y,x) + o_t along the conv axis, ci] * Wp[co][t*C + ci] )
//   with o_t in {-d, 0, +d} (forward) or {+d, 0, -d} (input gradient).
//
// GEMM view per CTA tile: M = 128 pixels = TA positions along the conv axis x TB across it
// (TA x TB = 16x8 or 8x16), N = 64 output channels, K = 3*C.
//  * B (all 3 taps of this CTA's 64 output channels, 48/96 KB) is TMA-loaded ONCE per CTA and
//    stays in shared memory (persistent CTA, one per SM); tiles stream through it.
//  * A: ONE TMA "slab" per 32-channel chunk: [(TA+2d) x TB pixels] x 32 ch, 128B-swizzled, with the
//    cross axis fastest in shared memory.  The three taps are three VIEWS of the same slab, shifted by
//    d*TB rows (TB is a multiple of 8, so every view starts on a 1024-byte boundary and keeps the 8-row
//    swizzle phase): the input crosses the L2->SM fabric (TA+2d)/TA times instead of 3 times -- the
//    first version of this kernel (conv_tc_v1.cu, one box per tap) was measured L2->SM bound
//    (216 MB / 35 us = 6.2 TB/s, profiles/r01).  Image borders (the conv zero padding, any dilation)
//    are TMA out-of-bounds fill.  Multi-stage mbarrier ring.
//  * MMA: tcgen05.mma.cta_group::1.kind::tf32, M=128 N=64 K=8, issued by one elected thread;
//    accumulators in TMEM, double-buffered (2 x 64 columns) so the epilogue of tile i overlaps
//    the MMAs of tile i+1.
//  * Epilogue: 4 warps, two phases per 32-channel half of the tile.  (1) tcgen05.ld 32x32b (thread = pixel
//    row) + bias + ReLU, staged into a padded shared-memory half-tile (18 KB).  (2) the half-tile is walked
//    row-major so that 8 consecutive threads cover one pixel's 32 channels (one 128-byte line): ReLU-backward
//    mask, residual-gradient add, column sums / sums of squares (bias gradient, BatchNorm statistics) and the
//    output stores are fully coalesced 128-bit accesses.  The mask / residual operands of a tile are
//    prefetched into registers BEFORE waiting for its accumulator, so their latency hides behind the MMAs.  (Writing
//    straight from the pixel-row registers made every store instruction touch 32 different 512-byte-strided
//    rows: ~8K L1 wavefronts per tile, 5x the MMA time -- measured in profiles/r01.)
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue.
#include "conv_tc_common.cuh"

// variant 1 (conv_tc_v1.cu)
int lf_conv1d_tc_supported_v1(int N, int H, int W, int C);
int lf_conv1d_tc_v1(const LfConvTcArgs* args, lf_stream_t stream_);

namespace lf {

static int g_tc_variant = 2;  // 2 = halo slab (this file), 1 = one box per tap (conv_tc_v1.cu)
static int g_tc_debug = 0;    // timing experiments only: bit0 = skip the epilogue body, bit1 = skip the TMA loads


constexpr int TC_B_ATOM_BYTES = TC_BN * 128;  // 8 KB

// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 <<17, M>>4 <<24
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);

template <int C, bool AHEAD = false>
struct TcCfg {
    static constexpr int KCHUNKS = C / TC_KCH;  // 32-channel chunks (= slabs per tile)
    static constexpr int B_BYTES = 3 * KCHUNKS * TC_B_ATOM_BYTES;
    // Epilogue warp groups of 4 warps (one per TMEM lane quadrant).  At C=64 a tile has half the MMA work of a
    // C=128 tile but the same 128x64 output, so one group cannot drain the accumulator as fast as the tensor
    // pipe fills it (measured: epilogue-bound at 1.5-2.4x the MMA floor); two groups each take one 32-channel
    // half with their own staging tile.  At C=128 the shared memory goes to slab stages instead.
    // The residual-add launches (AHEAD) of C=128 also run two groups: they are epilogue-bound (two extra operand
    // tensors per tile) and the second group's 18 KB staging tile costs one slab stage.
    static constexpr int EPI_GROUPS = (C == 64 || AHEAD) ? 2 : 1;
    static constexpr int HALVES = 2 / EPI_GROUPS;  // 32-channel halves each group walks per tile
    static constexpr int THREADS = 64 + 128 * EPI_GROUPS;
    static constexpr int STG_BYTES = EPI_GROUPS * TC_STG_BYTES;
};

// AHEAD (C=64 only): the residual-gradient operands (add_src, add_mask) of tile i+1 are fetched while tile i is
// processed.  A separate instantiation because the extra 64 live registers cost the plain / mask-only launches
// ~10 us each through spills (measured), while the add launches gain ~25 us.
template <int C, bool AHEAD>
__global__ void __launch_bounds__(TcCfg<C, AHEAD>::THREADS, 1)
conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    pdl_trigger();
    using Cfg = TcCfg<C, AHEAD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    uint8_t* sA = smem + Cfg::B_BYTES;
    float* stg = reinterpret_cast<float*>(sA + (size_t)a.stages * a.stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + Cfg::STG_BYTES);
    uint64_t* full = bars;                       // [TC_MAX_STAGES]
    uint64_t* empty = bars + TC_MAX_STAGES;      // [TC_MAX_STAGES]
    uint64_t* bfull = bars + 2 * TC_MAX_STAGES;  // [1]
    uint64_t* tfull = bfull + 1;                 // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_half = blockIdx.x % a.n_halves;
    const int cta_m = blockIdx.x / a.n_halves;
    const int m_stride = gridDim.x / a.n_halves;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(bfull, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4 * Cfg::EPI_GROUPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * TC_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0) {
        // ================= TMA producer (whole warp converged, one elected lane issues) =================
        const bool leader = elect_one();
        if (leader) {
            mbar_arrive_expect_tx(bfull, Cfg::B_BYTES);
            for (int kb = 0; kb < 3 * Cfg::KCHUNKS; ++kb)
                tma_load_2d(&tmB, bfull, sB + kb * TC_B_ATOM_BYTES, kb * TC_KCH, n_half * TC_BN);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = cta_m; mt < a.total_m_tiles; mt += m_stride) {
            const int ta = mt % a.tiles_a;
            const int tb = (mt / a.tiles_a) % a.tiles_b;
            const int n = mt / (a.tiles_a * a.tiles_b);
            const int a0 = ta * a.TA - a.dil, b0 = tb * a.TB;
            for (int cb = 0; cb < Cfg::KCHUNKS; ++cb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (leader) {
                    if (a.debug & 2) {
                        mbar_arrive(&full[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full[stage], a.stage_bytes);
                        // tensor map dims: (ci, cblk, cross axis, conv axis, n)
                        tma_load_5d(&tmA, &full[stage], sA + (size_t)stage * a.stage_bytes, 0, cb, b0, a0, n);
                    }
                }
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp converged, one elected lane issues) =================
        const bool leader = elect_one();
        mbar_wait(bfull, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int mt = cta_m; mt < a.total_m_tiles; mt += m_stride, ++it) {
            const int buf = it & 1;
            const uint32_t use_parity = (it >> 1) & 1;
            mbar_wait(&tempty[buf], use_parity ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * TC_BN;
            for (int cb = 0; cb < Cfg::KCHUNKS; ++cb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t slab = smem_u32(sA + (size_t)stage * a.stage_bytes);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    // tap t = the slab shifted by tap_row[t] rows of 128 B (a multiple of 8 rows)
                    const uint64_t adesc = umma_desc_sw128(slab + a.tap_row[t] * 128);
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (t * Cfg::KCHUNKS + cb) * TC_B_ATOM_BYTES));
#pragma unroll
                    for (int k8 = 0; k8 < TC_KCH / 8; ++k8) {  // 8 tf32 = 32 bytes = 2 x 16B per MMA
                        // NOTE: keep this loop free of run-time switches — two extra predicated MMA variants
                        // here (an ablation experiment) cost 50 ns per MMA = 16 us per launch.
                        if (leader) umma_tf32(d_tmem, adesc + 2 * k8, bdesc + 2 * k8, TC_IDESC, (cb | t | k8) != 0 ? 1u : 0u);
                    }
                }
                if (leader) umma_commit(&empty[stage]);  // frees the slab when these MMAs have read it
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (leader) umma_commit(&tfull[buf]);  // accumulator complete -> epilogue
        }
    } else {
        // ================= epilogue (warps 2..): conv_tc_common.cuh =================
        tc_epilogue<Cfg::EPI_GROUPS, (AHEAD ? 2 : 0), 64, 2, TC_BN, false>(a, stg, tmem_base, tfull, tempty, n_half, cta_m, m_stride);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * TC_BN);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TcPlan {
    int vertical, dil, fwd_order;  // fwd_order: weight slot t reads offset (t-1)*d (1) or -(t-1)*d (0)
    int TA, TB, tiles_a, tiles_b, stages, stage_bytes, smem_bytes, m_ctas;
};

// Derive the plan from the public arguments; false = shape / tap pattern not served by this kernel.
static bool tc_make_plan(int N, int H, int W, int C, const int* dy, const int* dx, TcPlan* p, int epi_groups = 0) {
    if (epi_groups == 0) epi_groups = (C == 64) ? 2 : 1;
    if (!(C == 64 || C == 128) || N <= 0) return false;
    const bool vert = dy[0] != 0 || dy[2] != 0;
    const int* o = vert ? dy : dx;
    const int* z = vert ? dx : dy;
    if (z[0] || z[1] || z[2] || o[1] != 0 || o[0] != -o[2] || o[0] == 0) return false;
    p->vertical = vert ? 1 : 0;
    p->dil = o[0] < 0 ? -o[0] : o[0];
    p->fwd_order = o[0] < 0 ? 1 : 0;
    const int ext_a = vert ? H : W, ext_b = vert ? W : H;
    if (ext_a % 16 == 0 && ext_b % 8 == 0) {
        p->TA = 16;
        p->TB = 8;
    } else if (ext_a % 8 == 0 && ext_b % 16 == 0) {
        p->TA = 8;
        p->TB = 16;
    } else {
        return false;
    }
    if (p->TA + 2 * p->dil > 256) return false;  // TMA box limit
    p->tiles_a = ext_a / p->TA;
    p->tiles_b = ext_b / p->TB;
    p->stage_bytes = (p->TA + 2 * p->dil) * p->TB * 128;
    const int b_bytes = 3 * (C / 32) * TC_B_ATOM_BYTES;
    const int fixed = 1024 + b_bytes + epi_groups * TC_STG_BYTES + 512;  // alignment slack + B + epilogue staging + barriers
    int stages = (TC_SMEM_LIMIT - fixed) / p->stage_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return false;
    p->stages = stages;
    p->smem_bytes = fixed + stages * p->stage_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_halves = C / TC_BN;
    const long long tiles = (long long)N * p->tiles_a * p->tiles_b;
    long long m_ctas = sms / n_halves;
    if (m_ctas > tiles) m_ctas = tiles;
    p->m_ctas = (int)(m_ctas < 1 ? 1 : m_ctas);
    return tc_get_encode_fn() != nullptr;
}

}  // namespace lf

using namespace lf;

extern "C" void lf_conv1d_tc_set_variant(int v) { g_tc_variant = (v == 1) ? 1 : 2; }
extern "C" void lf_conv1d_tc_set_debug(int bits) { g_tc_debug = bits; }

// returns 0 if the shape is unsupported, else the number of rows of the optional colsum_partial output.
// Every shape the slab kernel takes is also taken by the per-tap kernel (and both use the same number of
// CTAs), which serves as the per-call fallback when a slab does not fit (large dilation on small maps).
extern "C" int lf_conv1d_tc_supported(int N, int H, int W, int C) { return lf_conv1d_tc_supported_v1(N, H, W, C); }

// 1 if a call with these taps runs on the slab kernel (and may therefore request stats_partial)
extern "C" int lf_conv1d_tc_slab_ok(int N, int H, int W, int C, int vertical, int dil) {
    if (g_tc_variant == 1 || dil < 1) return 0;
    TcPlan pl;
    const int zero[3] = {0, 0, 0}, off[3] = {-dil, 0, dil};
    return tc_make_plan(N, H, W, C, vertical ? off : zero, vertical ? zero : off, &pl) ? 1 : 0;
}

extern "C" int lf_conv1d_tc(const LfConvTcArgs* args, lf_stream_t stream_) {
    if (g_tc_variant == 1) return lf_conv1d_tc_v1(args, stream_);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!args) return LF_ERR_INVALID_ARGUMENT;
    const LfConvTcArgs& p = *args;
    LF_REQUIRE(p.in && p.wpack && p.out);
    TcPlan pl;
    if (!tc_make_plan(p.N, p.H, p.W, p.C, p.dy, p.dx, &pl)) return lf_conv1d_tc_v1(args, stream_);
    // residual-add launches: operands one tile ahead + two epilogue groups (if the extra staging tile still leaves
    // two slab stages at C=128)
    bool ahead = p.add_src && !p.mask_src;
    if (ahead && p.C == 128) {
        TcPlan pl2;
        if (tc_make_plan(p.N, p.H, p.W, p.C, p.dy, p.dx, &pl2, 2)) pl = pl2; else ahead = false;
    }
    TcEncodeTiledFn enc = tc_get_encode_fn();
    TcArgs a{};
    a.out = p.out; a.bias = p.bias; a.mask_src = p.mask_src; a.add_src = p.add_src; a.add_mask = p.add_mask;
    a.colsum_partial = p.colsum_partial;
    a.stats_partial = p.stats_partial;
    a.mask_scale = p.mask_scale;
    a.mask_shift = p.mask_shift;
    LF_REQUIRE(!p.mask_scale || (p.mask_shift && p.mask_src && p.stats_partial));
    a.N = p.N; a.H = p.H; a.W = p.W; a.Ctot = p.C; a.relu = p.relu;
    a.vertical = pl.vertical; a.TA = pl.TA; a.TB = pl.TB; a.dil = pl.dil;
    a.tb_shift = (pl.TB == 8) ? 3 : 4;
    a.slab_rows = (pl.TA + 2 * pl.dil) * pl.TB;
    a.tiles_a = pl.tiles_a; a.tiles_b = pl.tiles_b;
    a.stages = pl.stages; a.stage_bytes = pl.stage_bytes;
    a.debug = g_tc_debug;
    for (int t = 0; t < 3; ++t) a.tap_row[t] = (pl.fwd_order ? t : 2 - t) * pl.dil * pl.TB;
    a.n_halves = p.C / TC_BN;
    a.total_m_tiles = p.N * pl.tiles_a * pl.tiles_b;

    CUtensorMap tmA, tmB;
    {
        // activations [N,H,W,C] viewed as (ci:32, cblk:C/32, cross axis, conv axis, n); one box = the slab
        const cuuint64_t sx = (cuuint64_t)p.C * 4, sy = (cuuint64_t)p.W * p.C * 4;
        cuuint64_t dims[5] = {32, (cuuint64_t)(p.C / 32), (cuuint64_t)(pl.vertical ? p.W : p.H),
                              (cuuint64_t)(pl.vertical ? p.H : p.W), (cuuint64_t)p.N};
        cuuint64_t strides[4] = {128, pl.vertical ? sx : sy, pl.vertical ? sy : sx, (cuuint64_t)p.H * p.W * p.C * 4};
        cuuint32_t box[5] = {32, 1, (cuuint32_t)pl.TB, (cuuint32_t)(pl.TA + 2 * pl.dil), 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(p.in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return LF_ERR_CUDA;
    }
    {
        // packed weights [Cout][3*C] (K contiguous)
        cuuint64_t dims[2] = {(cuuint64_t)(3 * p.C), (cuuint64_t)p.C};
        cuuint64_t strides[1] = {(cuuint64_t)(3 * p.C) * 4};
        cuuint32_t box[2] = {32, TC_BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.wpack), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return LF_ERR_CUDA;
    }
    const int grid = pl.m_ctas * a.n_halves;
    cudaError_t e;
    if (p.C == 128 && ahead) {
        e = cudaFuncSetAttribute(conv1d_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_kernel<128, true>, grid, TcCfg<128, true>::THREADS, pl.smem_bytes, stream, tmA, tmB, a);
    } else if (p.C == 128) {
        e = cudaFuncSetAttribute(conv1d_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_kernel<128, false>, grid, TcCfg<128>::THREADS, pl.smem_bytes, stream, tmA, tmB, a);
    } else if (ahead) {
        e = cudaFuncSetAttribute(conv1d_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_kernel<64, true>, grid, TcCfg<64, true>::THREADS, pl.smem_bytes, stream, tmA, tmB, a);
    } else {
        e = cudaFuncSetAttribute(conv1d_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_kernel<64, false>, grid, TcCfg<64>::THREADS, pl.smem_bytes, stream, tmA, tmB, a);
    }
    return check_launch();
}

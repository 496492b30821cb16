// 3xTF32 ("fp32-grade") tcgen05 kernel for ERFNet's factorised 3-tap convolutions and their input gradients
// (non_bottleneck_1d, BP/Networks/ERFNet.py:29-37,44-53), C in {64, 128}, NHWC fp32 activations.
//
// The tensor core multiplies TF32 values (10 mantissa bits).  Every fp32 operand is split a = a_hi + a_lo with
// a_hi = the TF32 value the tensor core itself derives from the fp32 bits and a_lo = tf32(a - a_hi); then
//     a*w = a_hi*w_hi + (a_hi*w_lo + a_lo*w_hi) + O(2^-22 |a||w|)
// with fp32 accumulation: the products are as accurate as an fp32 FFMA chain (the dropped a_lo*w_lo term is below the
// fp32 rounding of the sum), so this mode meets the 1e-4 parity gates that the single-pass TF32 kernel (conv_tc.cu,
// 2^-11 per operand) misses by an order of magnitude.
//
// Same GEMM view, halo slab, persistent CTAs and epilogue as conv_tc.cu (read its header first).  What differs:
//  * Weights arrive pre-split from the packer ([2][Cout][3C]: hi, lo; both exactly representable in TF32).  In shared
//    memory the hi and lo rows of a (tap, 32-channel block) sit in ONE 128-row operand atom, so
//        MMA 1:  A_hi[128 x 8] * [W_hi | W_lo]^T   (N = 128)  ->  TMEM columns [0,64) = hi*hi, [64,128) = hi*lo
//    is a single instruction (the N=128 form runs at 69 cycles against 2 x 51 for two N=64 ones: the A tile is
//    fetched from shared memory once).
//  * Activations: the raw fp32 slab is the hi operand as it is (the tensor core ignores the 13 low mantissa bits).
//    When the hi MMAs of a slab have completed (tcgen05.commit -> mbarrier), the split warps rewrite the slab IN PLACE
//    with a_lo (no second buffer: shared memory goes to pipeline stages), fence the generic->async proxy, and
//        MMA 2:  A_lo * W_hi^T   (N = 64)   ->  accumulates into columns [64,128)
//    is issued one slab later than MMA 1 so the tensor pipe never waits for the split.
//    The epilogue adds the two column ranges (small terms first).
//  * hi|lo weights of all three taps take 96 KB per 64 input channels.  C = 64: resident.  C = 128: a double-buffered
//    48 KB slot per 32-channel block; the CTA keeps the accumulators of 4 tiles in TMEM (4 x 128 columns) and walks
//    block-major over them (block 0 of tiles 0..3, block 1 of tiles 0..3, ...), so a slot is refilled a whole
//    sub-phase (4 slabs) before it is needed again and is re-read from L2 once per 4 tiles.
// Warp roles: 0 = TMA producer (activation slabs + weight slot refills), 1 = MMA issuer (+TMEM alloc), 2.. = split warps,
// then 8 epilogue warps; 12 warps in all (3 per SM sub-partition keeps the register cap at 168).
#include <stdlib.h>

#include "conv_tc_common.cuh"

namespace lf {

constexpr int X3_NBUF = 4;                       // TMEM accumulator buffers
constexpr int X3_BUF_COLS = 128;                 // columns per buffer: [hi*hi | hi*lo + lo*hi]
constexpr int X3_B_ATOM_BYTES = 128 * 128;       // one (tap, 32-channel block): rows 0-63 hi, 64-127 lo   (16 KB)
constexpr int X3_B_SLOT_BYTES = 3 * X3_B_ATOM_BYTES;   // 48 KB: three taps of one 32-channel block
constexpr int X3_B_BYTES = 2 * X3_B_SLOT_BYTES;        // two slots
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 <<17, M>>4 <<24
constexpr uint32_t X3_IDESC_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t X3_IDESC_N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// The TF32 value tcgen05.mma.kind::tf32 derives from fp32 bits in shared memory: the low 13 mantissa bits are ignored
// (tools/micro/tf32_rounding.cu, profiles/r02/tf32_rounding.jsonl).
__device__ __forceinline__ float x3_hw_tf32(float a) { return __uint_as_float(__float_as_uint(a) & 0xffffe000u); }
// round-to-nearest (ties away) to TF32 with the dropped bits cleared: exact under any hardware conversion
__device__ __forceinline__ float x3_rna_tf32(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float x3_lo(float a) { return x3_rna_tf32(a - x3_hw_tf32(a)); }

template <int C, int AHEAD = 0, int EG = 2>
struct X3Cfg {
    static constexpr int NCB = C / TC_KCH;          // 32-channel blocks = slabs per tile
    // Two epilogue groups (one 32-channel half each) for every shape: at C = 128 the tiles of a group complete in a burst
    // (block-major walk) and the next group needs their TMEM buffers back at once -- a single group (5k cycles per
    // tile) stalled the MMA issuer for ~9 us per launch (tools/x3_ablate.py, profiles/r02).
    // EG = 1 (one group walks both halves, 18 KB less staging) only where two groups would leave fewer than two slab
    // stages: the 48 KB slabs of the large dilations on 40 x 80 maps (320 x 640 inputs).
    static constexpr int EPI_GROUPS = EG;
    static constexpr int SPLIT_WARPS = 2;
    static constexpr int EPI_T0 = 32 * (2 + SPLIT_WARPS);
    static constexpr int THREADS = EPI_T0 + 128 * EPI_GROUPS;
    static constexpr int STG_BYTES = EPI_GROUPS * TC_STG_BYTES;
    static constexpr bool BLOCK_MAJOR = C > 64;     // walk order inside a group of X3_NBUF tiles
};

template <int C, int AHEAD, int EG>
__global__ void __launch_bounds__(X3Cfg<C, AHEAD, EG>::THREADS, 1)
conv1d_tc_x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    pdl_trigger();
    using Cfg = X3Cfg<C, AHEAD, EG>;
    constexpr int NCB = Cfg::NCB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    uint8_t* sA = smem + X3_B_BYTES;
    float* stg = reinterpret_cast<float*>(sA + (size_t)a.stages * a.stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + Cfg::STG_BYTES);
    uint64_t* full = bars;                        // [S] TMA slab landed
    uint64_t* empty = full + TC_MAX_STAGES;       // [S] lo MMAs done: slab may be overwritten by TMA
    uint64_t* hdone = empty + TC_MAX_STAGES;      // [S] hi MMAs done: slab may be rewritten with a_lo
    uint64_t* lordy = hdone + TC_MAX_STAGES;      // [S] a_lo written
    uint64_t* bfull = lordy + TC_MAX_STAGES;      // [2] weight slot loaded
    uint64_t* bempty = bfull + 2;                 // [2] weight slot may be refilled
    uint64_t* tfull = bempty + 2;                 // [NBUF]
    uint64_t* tempty = tfull + X3_NBUF;           // [NBUF]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + X3_NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_half = blockIdx.x % a.n_halves;
    const int cta_m = blockIdx.x / a.n_halves;
    const int m_stride = gridDim.x / a.n_halves;
    const int my_tiles = cta_m < a.total_m_tiles ? (a.total_m_tiles - cta_m + m_stride - 1) / m_stride : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&hdone[s], 1);
            mbar_init(&lordy[s], Cfg::SPLIT_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bfull[b], 1);
            mbar_init(&bempty[b], 1);
        }
        for (int b = 0; b < X3_NBUF; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4 * Cfg::EPI_GROUPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, X3_NBUF * X3_BUF_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    // Step s of a group of gn <= X3_NBUF tiles touches (tile j of the group, 32-channel block cb):
    // tile-major when the weights are resident (C = 64), block-major when they stream through the two slots.
    auto step_of = [&](int s, int gn, int& j, int& cb) {
        if (Cfg::BLOCK_MAJOR) { cb = s / gn; j = s - cb * gn; } else { j = s / NCB; cb = s - j * NCB; }
    };

    if (warp == 0) {
        // ================= TMA producer (whole warp converged, one elected lane issues) =================
        const bool leader = elect_one();
        // Weight fills: sub-phase q = group * NCB + block goes to slot q & 1.  Fill q >= 2 may be issued once the lo MMAs of
        // the last slab of sub-phase q - 2 have completed (bempty); that slab's `empty` barrier completes at the same
        // moment, so the refill is issued from the slab loop at the step whose `empty` wait implies it -- no second
        // producer warp, and the slab ring keeps running ahead while a slot waits for its release.
        const int nq = Cfg::BLOCK_MAJOR ? ((my_tiles + X3_NBUF - 1) / X3_NBUF) * NCB : NCB;
        auto fill = [&](int q) {
            const int slot = q & 1, cb = q % NCB;
            if (q >= 2) mbar_wait(&bempty[slot], ((q >> 1) - 1) & 1);
            if (leader) {
                mbar_arrive_expect_tx(&bfull[slot], X3_B_SLOT_BYTES);
                for (int t = 0; t < 3; ++t) {
                    uint8_t* dst = sB + slot * X3_B_SLOT_BYTES + t * X3_B_ATOM_BYTES;
                    // packed weights [2][Cout][3*C] viewed as [2*Cout][3*C]: lo rows start at row Cout
                    tma_load_2d(&tmB, &bfull[slot], dst, t * C + cb * TC_KCH, n_half * TC_BN);
                    tma_load_2d(&tmB, &bfull[slot], dst + TC_BN * 128, t * C + cb * TC_KCH, a.Ctot + n_half * TC_BN);
                }
            }
        };
        // global step index of the last slab of sub-phase q (block-major order)
        auto release_step = [&](int q) {
            const int g = q / NCB, cb = q - g * NCB;
            const int gn = min(X3_NBUF, my_tiles - g * X3_NBUF);
            return g * X3_NBUF * NCB + cb * gn + gn - 1;
        };
        if (my_tiles > 0) {
            fill(0);
            fill(1);
        }
        int next_fill = 2;
        int stage = 0, gs = 0;
        uint32_t phase = 0;
        for (int g0 = 0; g0 < my_tiles; g0 += X3_NBUF) {
            const int gn = min(X3_NBUF, my_tiles - g0);
            for (int s = 0; s < gn * NCB; ++s, ++gs) {
                int j, cb;
                step_of(s, gn, j, cb);
                const int mt = cta_m + (g0 + j) * m_stride;
                const int ta = mt % a.tiles_a;
                const int tb = (mt / a.tiles_a) % a.tiles_b;
                const int n = mt / (a.tiles_a * a.tiles_b);
                mbar_wait(&empty[stage], phase ^ 1);   // lo MMAs of step gs - stages have completed
                while (next_fill < nq && gs - a.stages >= release_step(next_fill - 2)) fill(next_fill++);
                if (leader) {
                    mbar_arrive_expect_tx(&full[stage], a.stage_bytes);
                    // tensor map dims: (ci, cblk, cross axis, conv axis, n)
                    tma_load_5d(&tmA, &full[stage], sA + (size_t)stage * a.stage_bytes, 0, cb, tb * a.TB, ta * a.TA - a.dil, n);
                }
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        while (next_fill < nq) fill(next_fill++);   // refills whose release comes after the last slab was issued
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp converged, one elected lane issues) =================
        const bool leader = elect_one();
        const uint32_t sB_u32 = smem_u32(sB);
        int stage = 0;
        uint32_t phase = 0;
        // The lo MMAs of a slab are issued LAG slabs late (2 when the ring has >= 4 stages, else 1): by then its hi MMAs have
        // long completed and the split warps have rewritten it, so the issuer never blocks on `lordy` with an empty
        // tensor-pipe queue (with LAG = 1 the pipe was 50 % active: ncu profiles/r02).
        struct Pend {
            int stage, buf, slot;
            uint32_t phase, dtmem;
            bool last_of_tile, release_b;
        };
        Pend older{}, newer{};
        int npend = 0;
        const int LAG = a.stages >= 4 ? 2 : 1;
        auto issue_lo = [&](const Pend& p) {
            mbar_wait(&lordy[p.stage], p.phase);
            tc_fence_after();
            const uint32_t slab = smem_u32(sA + (size_t)p.stage * a.stage_bytes);
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const uint64_t adesc = umma_desc_sw128(slab + a.tap_row[t] * 128);
                const uint64_t bdesc = umma_desc_sw128(sB_u32 + p.slot * X3_B_SLOT_BYTES + t * X3_B_ATOM_BYTES);
#pragma unroll
                for (int k8 = 0; k8 < TC_KCH / 8; ++k8)
                    if (leader && !(a.debug & 8)) umma_tf32(p.dtmem + TC_BN, adesc + 2 * k8, bdesc + 2 * k8, X3_IDESC_N64, 1u);
            }
            if (leader) {
                umma_commit(&empty[p.stage]);                       // slab free
                if (p.release_b) umma_commit(&bempty[p.slot]);      // last use of this weight block in the group
                if (p.last_of_tile) umma_commit(&tfull[p.buf]);     // accumulator complete -> epilogue
            }
        };
        for (int g0 = 0; g0 < my_tiles; g0 += X3_NBUF) {
            const int gn = min(X3_NBUF, my_tiles - g0);
            for (int s = 0; s < gn * NCB; ++s) {
                int j, cb;
                step_of(s, gn, j, cb);
                const int it = g0 + j, buf = j;
                const int slot = cb & 1;
                if (Cfg::BLOCK_MAJOR ? (j == 0) : (g0 == 0 && j == 0)) {   // first use of a (re)filled weight slot
                    const int q = Cfg::BLOCK_MAJOR ? (g0 / X3_NBUF) * NCB + cb : cb;   // fill number (see the producer)
                    // Fill q >= 2 is released by the lo MMAs of the last slab of sub-phase q - 2.  In a group of fewer than
                    // LAG tiles those are still pending here (they would be issued AFTER this step's hi MMAs, which wait for
                    // the fill: a cycle -- the B=2 test shapes trapped on it): issue them first.
                    while (npend > 0 && ((older.release_b && older.slot == slot) ||
                                         (npend == 2 && newer.release_b && newer.slot == slot))) {
                        issue_lo(older);
                        older = newer;
                        --npend;
                    }
                    mbar_wait(&bfull[slot], (q >> 1) & 1);
                    tc_fence_after();
                }
                if (cb == 0) {   // first touch of this accumulator buffer in the group: the epilogue must have drained it
                    mbar_wait(&tempty[buf], ((it / X3_NBUF) & 1) ^ 1);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + buf * X3_BUF_COLS;
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t slab = smem_u32(sA + (size_t)stage * a.stage_bytes);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    // tap t = the slab shifted by tap_row[t] rows of 128 B (a multiple of 8 rows)
                    const uint64_t adesc = umma_desc_sw128(slab + a.tap_row[t] * 128);
                    const uint64_t bdesc = umma_desc_sw128(sB_u32 + slot * X3_B_SLOT_BYTES + t * X3_B_ATOM_BYTES);
#pragma unroll
                    for (int k8 = 0; k8 < TC_KCH / 8; ++k8)
                        if (leader) umma_tf32(d_tmem, adesc + 2 * k8, bdesc + 2 * k8, X3_IDESC_N128, (cb | t | k8) != 0 ? 1u : 0u);
                }
                if (leader) umma_commit(&hdone[stage]);   // the split warps may now rewrite the slab with a_lo
                if (npend == LAG) {
                    issue_lo(older);
                    older = newer;
                    --npend;
                }
                const Pend cur{stage, buf, slot, phase, d_tmem, cb == NCB - 1, Cfg::BLOCK_MAJOR && (j == gn - 1)};
                if (npend == 0) older = cur; else newer = cur;
                ++npend;
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        if (npend >= 1) issue_lo(older);
        if (npend == 2) issue_lo(newer);
    } else if (warp < 2 + Cfg::SPLIT_WARPS) {
        // ================= split warps: slab <- a_lo, in place, once its hi MMAs have completed =================
        const int tid = threadIdx.x - 64;
        constexpr int NT = 32 * Cfg::SPLIT_WARPS;
        const int nvec = a.slab_rows * 8;   // float4 per slab
        int stage = 0;
        uint32_t phase = 0;
        const int total_steps = my_tiles * NCB;
        for (int s = 0; s < total_steps; ++s) {
            mbar_wait(&full[stage], phase);    // TMA writes visible to this thread
            mbar_wait(&hdone[stage], phase);   // the tensor core has finished reading the fp32 values
            float4* p = reinterpret_cast<float4*>(sA + (size_t)stage * a.stage_bytes);
            int i = (a.debug & 4) ? nvec : tid;   // timing experiment: skip the rewrite
            for (; i + 3 * NT < nvec; i += 4 * NT) {
                float4 v0 = p[i], v1 = p[i + NT], v2 = p[i + 2 * NT], v3 = p[i + 3 * NT];
                v0.x = x3_lo(v0.x); v0.y = x3_lo(v0.y); v0.z = x3_lo(v0.z); v0.w = x3_lo(v0.w);
                v1.x = x3_lo(v1.x); v1.y = x3_lo(v1.y); v1.z = x3_lo(v1.z); v1.w = x3_lo(v1.w);
                v2.x = x3_lo(v2.x); v2.y = x3_lo(v2.y); v2.z = x3_lo(v2.z); v2.w = x3_lo(v2.w);
                v3.x = x3_lo(v3.x); v3.y = x3_lo(v3.y); v3.z = x3_lo(v3.z); v3.w = x3_lo(v3.w);
                p[i] = v0; p[i + NT] = v1; p[i + 2 * NT] = v2; p[i + 3 * NT] = v3;
            }
            for (; i < nvec; i += NT) {
                float4 v = p[i];
                v.x = x3_lo(v.x); v.y = x3_lo(v.y); v.z = x3_lo(v.z); v.w = x3_lo(v.w);
                p[i] = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&lordy[stage]);
            if (++stage == a.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else {
        // ================= epilogue warps: conv_tc_common.cuh =================
        tc_epilogue<Cfg::EPI_GROUPS, AHEAD, Cfg::EPI_T0, X3_NBUF, X3_BUF_COLS, true>(a, stg, tmem_base, tfull, tempty, n_half, cta_m,
                                                                                     m_stride);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, X3_NBUF * X3_BUF_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct X3Plan {
    int vertical, dil, fwd_order;  // fwd_order: weight slot t reads offset (t-1)*d (1) or -(t-1)*d (0)
    int TA, TB, tb_shift, tiles_a, tiles_b, stages, stage_bytes, smem_bytes, m_ctas, epi_groups;
};

// Derive the plan from the public arguments; false = shape / tap pattern not served by this kernel.
// Tile shapes: TA x TB = 128 pixels with TB a power of two such that the tap shift d*TB is a multiple of the 8-row
// swizzle period; the slab with the fewest rows wins (tall tiles for the large dilations: halo 2d / TA).
static bool x3_make_plan(int N, int H, int W, int C, const int* dy, const int* dx, X3Plan* p) {
    if (!(C == 64 || C == 128) || N <= 0) return false;
    const bool vert = dy[0] != 0 || dy[2] != 0;
    const int* o = vert ? dy : dx;
    const int* z = vert ? dx : dy;
    if (z[0] || z[1] || z[2] || o[1] != 0 || o[0] != -o[2] || o[0] == 0) return false;
    p->vertical = vert ? 1 : 0;
    p->dil = o[0] < 0 ? -o[0] : o[0];
    p->fwd_order = o[0] < 0 ? 1 : 0;
    const int ext_a = vert ? H : W, ext_b = vert ? W : H;
    int best_rows = 1 << 30;
    for (int sh = 0; sh <= 4; ++sh) {
        const int TB = 1 << sh, TA = 128 >> sh;
        if (ext_a % TA || ext_b % TB || (p->dil * TB) % 8 || TA + 2 * p->dil > 256) continue;
        const int rows = (TA + 2 * p->dil) * TB;
        if (rows < best_rows) {
            best_rows = rows;
            p->TA = TA;
            p->TB = TB;
            p->tb_shift = sh;
        }
    }
    if (best_rows == (1 << 30)) return false;
    p->tiles_a = ext_a / p->TA;
    p->tiles_b = ext_b / p->TB;
    p->stage_bytes = best_rows * 128;
    // alignment slack + weights + epilogue staging + barriers; two epilogue groups unless that leaves a single stage
    int epi_groups = 2;
    int fixed = 1024 + X3_B_BYTES + epi_groups * TC_STG_BYTES + 512;
    int stages = (TC_SMEM_LIMIT - fixed) / p->stage_bytes;
    if (stages < 2) {
        epi_groups = 1;
        fixed = 1024 + X3_B_BYTES + epi_groups * TC_STG_BYTES + 512;
        stages = (TC_SMEM_LIMIT - fixed) / p->stage_bytes;
    }
    p->epi_groups = epi_groups;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return false;   // the lo MMAs of slab i are issued after the hi MMAs of slab i+1: two slabs in flight at least
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

template <int C, int AHEAD, int EG>
static cudaError_t x3_launch(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const TcArgs& a) {
    cudaError_t e = cudaFuncSetAttribute(conv1d_tc_x3_kernel<C, AHEAD, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    lf_launch(conv1d_tc_x3_kernel<C, AHEAD, EG>, grid, X3Cfg<C, AHEAD, EG>::THREADS, smem_bytes, stream, tmA, tmB, a);
    return cudaSuccess;
}

}  // namespace lf

using namespace lf;

// rows of colsum_partial / stats_partial (= CTAs per 64-channel half), 0 if a call with these taps is not served
extern "C" int lf_conv1d_tc_x3_rows(int N, int H, int W, int C, int vertical, int dil) {
    if (dil < 1) return 0;
    X3Plan pl;
    const int zero[3] = {0, 0, 0}, off[3] = {-dil, 0, dil};
    if (!x3_make_plan(N, H, W, C, vertical ? off : zero, vertical ? zero : off, &pl)) return 0;
    return pl.m_ctas;
}

static int g_x3_debug = 0;
// timing experiments only (tools/x3_ablate.py): bit0 skips the epilogue body, bit2 the a_lo rewrite, bit3 the lo MMAs;
// outputs are meaningless while bits are set.  0 = normal operation.
extern "C" void lf_conv1d_tc_x3_set_debug(int bits) { g_x3_debug = bits; }

extern "C" int lf_conv1d_tc_x3(const LfConvTcArgs* args, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!args) return LF_ERR_INVALID_ARGUMENT;
    const LfConvTcArgs& p = *args;
    LF_REQUIRE(p.in && p.wpack && p.out);
    X3Plan pl;
    if (!x3_make_plan(p.N, p.H, p.W, p.C, p.dy, p.dx, &pl)) return LF_ERR_UNSUPPORTED;
    // Epilogue operands one tile ahead (two epilogue groups only: one half per thread): 2 = add_src + add_mask, 1 = a single
    // operand -- the ReLU mask of a masked input gradient or a pre-masked residual gradient (round 2: the mask loads issued
    // just before the accumulator wait were the top stall line of the masked launches in the ncu source view)
    int ahead = 0;
    if (pl.epi_groups == 2) {
        if (p.add_src && p.add_mask && !p.mask_src) ahead = 2;
        else if (p.mask_src || p.add_src) ahead = 1;
    }
    static const bool no_ahead1 = getenv("LANEFIT_X3_NOAHEAD1") != nullptr;   // A/B switch for the measurement of AHEAD = 1
    if (no_ahead1 && ahead == 1 && !(p.add_src && !p.mask_src)) ahead = 0;
    TcEncodeTiledFn enc = tc_get_encode_fn();
    TcArgs a{};
    a.out = p.out; a.bias = p.bias; a.mask_src = p.mask_src; a.add_src = p.add_src; a.add_mask = p.add_mask;
    a.colsum_partial = p.colsum_partial;
    a.stats_partial = p.stats_partial;
    a.mask_scale = p.mask_scale;
    a.mask_shift = p.mask_shift;
    LF_REQUIRE(!p.mask_scale || (p.mask_shift && p.mask_src && p.stats_partial));
    a.N = p.N; a.H = p.H; a.W = p.W; a.Ctot = p.C; a.relu = p.relu;
    a.vertical = pl.vertical; a.TA = pl.TA; a.TB = pl.TB; a.tb_shift = pl.tb_shift; a.dil = pl.dil;
    a.tiles_a = pl.tiles_a; a.tiles_b = pl.tiles_b;
    a.stages = pl.stages; a.stage_bytes = pl.stage_bytes;
    a.slab_rows = (pl.TA + 2 * pl.dil) * pl.TB;
    a.debug = g_x3_debug;
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
        // packed weights [2][Cout][3*C] (hi, lo; K contiguous) as one [2*Cout][3*C] matrix
        cuuint64_t dims[2] = {(cuuint64_t)(3 * p.C), (cuuint64_t)(2 * p.C)};
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
    if (pl.epi_groups == 1) {
        e = p.C == 128 ? x3_launch<128, 0, 1>(grid, pl.smem_bytes, stream, tmA, tmB, a)
                       : x3_launch<64, 0, 1>(grid, pl.smem_bytes, stream, tmA, tmB, a);
    } else if (p.C == 128) {
        e = ahead == 2 ? x3_launch<128, 2, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a)
          : ahead == 1 ? x3_launch<128, 1, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a)
                       : x3_launch<128, 0, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a);
    } else {
        e = ahead == 2 ? x3_launch<64, 2, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a)
          : ahead == 1 ? x3_launch<64, 1, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a)
                       : x3_launch<64, 0, 2>(grid, pl.smem_bytes, stream, tmA, tmB, a);
    }
    if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
    return check_launch();
}

// VARIANT 1 (one TMA box per tap; kept as the validated fallback of conv_tc.cu, selected with
// lf_conv1d_tc_set_variant(1)).
// tcgen05 (5th-gen tensor core) implicit-GEMM kernel for ERFNet's factorised 3-tap convolutions
// (non_bottleneck_1d: conv3x1 / conv1x3 with dilation, BP/Networks/ERFNet.py:29-37,44-53) and
// their input gradients, C in {64, 128}, NHWC fp32 activations, TF32 multiply / fp32 accumulate
// (the arithmetic cuDNN uses for the reference's fp32 convs on Ampere+).
//
//   out[n,y,x,co] = epi( sum_{t<3} sum_{ci} in[n, y+dy[t], x+dx[t], ci] * Wp[co][t*C + ci] )
//
// GEMM view per CTA tile: M = 128 pixels (a bx x by patch), N = 64 output channels, K = 3*C.
//  * B (all 3 taps of this CTA's 64 output channels, 48/96 KB) is TMA-loaded ONCE per CTA and
//    stays in shared memory (persistent CTA, one per SM); tiles stream through it.
//  * A: one TMA box per (tap, 32-channel chunk): [128 px x 32 ch] = 16 KB, 128B-swizzled; the tap
//    shift is a coordinate offset of the box, out-of-image pixels are zero-filled by TMA (this is
//    the conv padding, for any dilation).  6-8 stage mbarrier ring.
//  * MMA: tcgen05.mma.cta_group::1.kind::tf32, M=128 N=64 K=8, issued by one elected thread;
//    accumulators in TMEM, double-buffered (2 x 64 columns) so the epilogue of tile i overlaps
//    the MMAs of tile i+1.
//  * Epilogue: 4 warps, tcgen05.ld 32x32b (thread = pixel row), + bias, ReLU, ReLU-backward mask,
//    residual-gradient add, 128-bit stores.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue.
#include <cuda.h>

#include "lf_common.cuh"
#include "tc_ptx.cuh"

namespace lf {

constexpr int TC_THREADS_V1 = 192;
constexpr int TC_BM_V1 = 128;
constexpr int TC_BN_V1 = 64;
constexpr int TC_KCH_V1 = 32;                        // fp32 elements per 128-byte swizzle row
constexpr int TC_A_STAGE_BYTES_V1 = TC_BM_V1 * 128;     // 16 KB
constexpr int TC_B_ATOM_BYTES_V1 = TC_BN_V1 * 128;      // 8 KB

struct TcArgsV1 {
    float* out;
    const float* bias;
    const float* mask_src;
    const float* add_src;
    const float* add_mask;
    float* colsum_partial;  // [gridDim.x / n_halves][Ctot] per-CTA column sums of the output, or NULL
    int N, H, W, Ctot;
    int bx, by;
    int dy[3], dx[3];
    int relu;
    int n_halves;
    int total_m_tiles;
};

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | SWIZZLE_128B(2) <<61)
__device__ __forceinline__ uint64_t umma_desc_sw128_v1(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 <<17, M>>4 <<24
constexpr uint32_t TC_IDESC_V1 = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_BN_V1 >> 3) << 17) | ((TC_BM_V1 >> 4) << 24);

template <int C>
struct TcCfgV1 {
    static constexpr int KCHUNKS = C / TC_KCH_V1;         // 32-channel chunks per tap
    static constexpr int KSTEPS = 3 * KCHUNKS;         // pipeline steps per tile
    static constexpr int B_BYTES = KSTEPS * TC_B_ATOM_BYTES_V1;
    static constexpr int STAGES = (C == 128) ? 6 : 8;
    static constexpr int SMEM_BYTES = 1024 + B_BYTES + STAGES * TC_A_STAGE_BYTES_V1 + 256;
};

template <int C>
__global__ void __launch_bounds__(TC_THREADS_V1, 1)
conv1d_tc_v1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgsV1 a) {
    pdl_trigger();
    using Cfg = TcCfgV1<C>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    uint8_t* sA = smem + Cfg::B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + Cfg::STAGES * TC_A_STAGE_BYTES_V1);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
    uint64_t* bfull = bars + 2 * Cfg::STAGES;    // [1]
    uint64_t* tfull = bfull + 1;                 // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_half = blockIdx.x % a.n_halves;
    const int cta_m = blockIdx.x / a.n_halves;
    const int m_stride = gridDim.x / a.n_halves;
    const int tiles_x = a.W / a.bx, tiles_y = a.H / a.by;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(bfull, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * TC_BN_V1);
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
            for (int kb = 0; kb < Cfg::KSTEPS; ++kb)
                tma_load_2d(&tmB, bfull, sB + kb * TC_B_ATOM_BYTES_V1, kb * TC_KCH_V1, n_half * TC_BN_V1);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = cta_m; mt < a.total_m_tiles; mt += m_stride) {
            const int tx = mt % tiles_x;
            const int ty = (mt / tiles_x) % tiles_y;
            const int n = mt / (tiles_x * tiles_y);
            for (int t = 0; t < 3; ++t) {
                const int x0 = tx * a.bx + a.dx[t], y0 = ty * a.by + a.dy[t];
                for (int cb = 0; cb < Cfg::KCHUNKS; ++cb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) {
                        mbar_arrive_expect_tx(&full[stage], TC_A_STAGE_BYTES_V1);
                        tma_load_5d(&tmA, &full[stage], sA + stage * TC_A_STAGE_BYTES_V1, 0, cb, x0, y0, n);
                    }
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
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
            const uint32_t d_tmem = tmem_base + buf * TC_BN_V1;
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128_v1(smem_u32(sA + stage * TC_A_STAGE_BYTES_V1));
                const uint64_t bdesc = umma_desc_sw128_v1(smem_u32(sB + ks * TC_B_ATOM_BYTES_V1));
#pragma unroll
                for (int k8 = 0; k8 < TC_KCH_V1 / 8; ++k8)  // 8 tf32 = 32 bytes = 2 x 16B per MMA
                    if (leader) umma_tf32(d_tmem, adesc + 2 * k8, bdesc + 2 * k8, TC_IDESC_V1, (ks | k8) != 0 ? 1u : 0u);
                if (leader) umma_commit(&empty[stage]);  // frees the A stage when these MMAs have read it
                if (++stage == Cfg::STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (leader) umma_commit(&tfull[buf]);  // accumulator complete -> epilogue
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may access
        const int m = lane_base + lane;
        const int yy = m / a.bx, xx = m - yy * a.bx;
        float csum[TC_BN_V1];  // running column sums of this thread's pixel row over all tiles (bias gradient)
#pragma unroll
        for (int c = 0; c < TC_BN_V1; ++c) csum[c] = 0.f;
        int it = 0;
        for (int mt = cta_m; mt < a.total_m_tiles; mt += m_stride, ++it) {
            const int buf = it & 1;
            const uint32_t use_parity = (it >> 1) & 1;
            const int tx = mt % tiles_x;
            const int ty = (mt / tiles_x) % tiles_y;
            const int n = mt / (tiles_x * tiles_y);
            const size_t off = ((size_t)(n * a.H + ty * a.by + yy) * a.W + tx * a.bx + xx) * a.Ctot + n_half * TC_BN_V1;
            mbar_wait(&tfull[buf], use_parity);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < TC_BN_V1; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + buf * TC_BN_V1 + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    const int c = c0 + 4 * q;
                    if (a.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n_half * TC_BN_V1 + c));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (a.relu & 1) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    if (a.mask_src) {
                        const float4 mk = __ldg(reinterpret_cast<const float4*>(a.mask_src + off + c));
                        o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
                        o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
                    }
                    if (a.add_src) {
                        float4 ad = __ldg(reinterpret_cast<const float4*>(a.add_src + off + c));
                        if (a.add_mask) {
                            const float4 mk = __ldg(reinterpret_cast<const float4*>(a.add_mask + off + c));
                            ad.x = mk.x > 0.f ? ad.x : 0.f; ad.y = mk.y > 0.f ? ad.y : 0.f;
                            ad.z = mk.z > 0.f ? ad.z : 0.f; ad.w = mk.w > 0.f ? ad.w : 0.f;
                        }
                        o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
                    }
                    if (a.relu & 2) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(a.out + off + c) = o;
                    csum[c] += o.x; csum[c + 1] += o.y; csum[c + 2] += o.z; csum[c + 3] += o.w;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        if (a.colsum_partial) {
            __shared__ float cs[4][TC_BN_V1];
#pragma unroll
            for (int c = 0; c < TC_BN_V1; ++c) {
                const float v = warp_sum(csum[c]);
                if (lane == 0) cs[warp & 3][c] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only
            const int t = threadIdx.x - 64;
            if (t < TC_BN_V1)
                a.colsum_partial[(size_t)cta_m * a.Ctot + n_half * TC_BN_V1 + t] = (cs[0][t] + cs[1][t]) + (cs[2][t] + cs[3][t]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * TC_BN_V1);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool pick_patch_v1(int H, int W, int* bx, int* by) {
    // bx * by = 128 pixels, bx | W, by | H; prefer wide patches (longer contiguous runs)
    for (int x = 128; x >= 1; x >>= 1) {
        const int y = 128 / x;
        if (W % x == 0 && H % y == 0) {
            *bx = x;
            *by = y;
            return true;
        }
    }
    return false;
}

}  // namespace lf

using namespace lf;

static int tc_m_ctas_v1(int N, int H, int W, int C, int bx, int by) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_halves = C / TC_BN_V1;
    const long long tiles = (long long)N * (H / by) * (W / bx);
    long long m_ctas = sms / n_halves;
    if (m_ctas > tiles) m_ctas = tiles;
    return (int)(m_ctas < 1 ? 1 : m_ctas);
}

// returns 0 if the shape is unsupported, else the number of rows of the optional colsum_partial output
int lf_conv1d_tc_supported_v1(int N, int H, int W, int C) {
    int bx, by;
    if (!(C == 64 || C == 128) || N <= 0) return 0;
    if (!pick_patch_v1(H, W, &bx, &by)) return 0;
    if (tc_get_encode_fn() == nullptr) return 0;
    return tc_m_ctas_v1(N, H, W, C, bx, by);
}

int lf_conv1d_tc_v1(const LfConvTcArgs* args, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!args) return LF_ERR_INVALID_ARGUMENT;
    const LfConvTcArgs& p = *args;
    LF_REQUIRE(p.in && p.wpack && p.out);
    if (p.stats_partial || p.mask_scale) return LF_ERR_UNSUPPORTED;  // BN-statistics epilogue exists in the slab kernel only
    if (!(p.C == 64 || p.C == 128)) return LF_ERR_UNSUPPORTED;
    TcArgsV1 a{};
    if (!pick_patch_v1(p.H, p.W, &a.bx, &a.by)) return LF_ERR_UNSUPPORTED;
    TcEncodeTiledFn enc = tc_get_encode_fn();
    if (!enc) return LF_ERR_UNSUPPORTED;
    a.out = p.out; a.bias = p.bias; a.mask_src = p.mask_src; a.add_src = p.add_src; a.add_mask = p.add_mask;
    a.colsum_partial = p.colsum_partial;
    a.N = p.N; a.H = p.H; a.W = p.W; a.Ctot = p.C; a.relu = p.relu;
    for (int t = 0; t < 3; ++t) {
        a.dy[t] = p.dy[t];
        a.dx[t] = p.dx[t];
    }
    a.n_halves = p.C / TC_BN_V1;
    a.total_m_tiles = p.N * (p.H / a.by) * (p.W / a.bx);

    CUtensorMap tmA, tmB;
    if (!tc_encode_nhwc_map(enc, &tmA, p.in, p.N, p.H, p.W, p.C, a.bx, a.by)) return LF_ERR_CUDA;
    {
        // packed weights [Cout][3*C] (K contiguous)
        cuuint64_t dims[2] = {(cuuint64_t)(3 * p.C), (cuuint64_t)p.C};
        cuuint64_t strides[1] = {(cuuint64_t)(3 * p.C) * 4};
        cuuint32_t box[2] = {32, TC_BN_V1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.wpack), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return LF_ERR_CUDA;
    }
    const int m_ctas = tc_m_ctas_v1(p.N, p.H, p.W, p.C, a.bx, a.by);
    const int grid = m_ctas * a.n_halves;
    cudaError_t e;
    if (p.C == 128) {
        e = cudaFuncSetAttribute(conv1d_tc_v1_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgV1<128>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_v1_kernel<128>, grid, TC_THREADS_V1, TcCfgV1<128>::SMEM_BYTES, stream, tmA, tmB, a);
    } else {
        e = cudaFuncSetAttribute(conv1d_tc_v1_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgV1<64>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(conv1d_tc_v1_kernel<64>, grid, TC_THREADS_V1, TcCfgV1<64>::SMEM_BYTES, stream, tmA, tmB, a);
    }
    return check_launch();
}

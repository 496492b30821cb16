// tcgen05 "gather-GEMM" convolution for ERFNet's resolution-changing layers: the 3x3 stride-2 Conv2d of
// DownsamplerBlock (BP/Networks/ERFNet.py:15,19-22), the 3x3 stride-2 ConvTranspose2d of UpsamplerBlock
// (:66-73) and the input gradients of both.  TF32 multiply / fp32 accumulate, fp32 NHWC storage.
//
//   out[n, oy*oy_mul + oy0, ox, c] = bias[c] + sum_{t<ntaps} sum_{k<Kc} A_{map[t]}[n, oy+dy[t], ox+dx[t], k] * Wg[c][t*Kc + k]
//
// over the iteration domain (n, oy, ox) in [N] x [Hs] x [Ws].  The host (ops_net.py) turns the four layer
// forms into this one by viewing two horizontally adjacent pixels as one "pair pixel" with twice the
// channels (a free view of NHWC memory) and the even / odd input rows as two strided views (A_0, A_1), so
// that a stride-2 access pattern becomes unit-stride taps over pair pixels:
//   * stride-2 conv (Down forward, Up input-gradient): 6 taps = 3 kernel rows x pair offsets {-1, 0},
//     Kc = 2*Cin, weights zero where a (pair offset, half) combination is not a kernel column;
//   * stride-2 transposed conv (Up forward, Down input-gradient): one launch per output row parity
//     (2 or 4 taps over the un-strided input), Ng = 2*Cout = both output columns of the pair.
// Structure = conv_tc_v1.cu (one TMA box [128 px x 32 ch] per tap and 32-channel chunk, OOB zero fill = conv
// padding, mbarrier ring, warp-specialised, accumulators double-buffered in TMEM) except that the weight
// chunk [Ng x 32] streams through the ring next to its activation box (the weight matrices of these
// layers, up to 393 KB, do not fit in shared memory) and Ng is a run-time multiple of 16 (<= 128).
//
// precision 1 (3xTF32, template X3; see conv_tc_x3.cu for the arithmetic): the weight chunk is the pre-split pair
// [W_hi rows | W_lo rows] (one TMA box of 2*Ng rows = ONE UMMA operand of N = 2*Ng), pass 1 = A_hi * [W_hi | W_lo]^T
// into TMEM columns [0,Ng) | [Ng,2Ng); when it has completed two split warps rewrite the activation box in place with
// A_lo, and pass 2 = A_lo * W_hi^T accumulates into [Ng,2Ng) one stage later.  The epilogue adds the two ranges.
#include <cuda.h>

#include "lf_common.cuh"
#include "tc_ptx.cuh"

namespace lf {

constexpr int TG_THREADS = 192;       // producer, MMA issuer, 4 epilogue warps
constexpr int TG_THREADS_X3 = 256;    // + 2 split warps (warps 2, 3)
constexpr int TG_BM = 128;
constexpr int TG_A_BYTES = TG_BM * 128;  // 16 KB
constexpr int TG_MAX_STAGES = 8;
constexpr int TG_MAXT = LF_TCG_MAX_TAPS;
constexpr int TG_SMEM_LIMIT = 226 * 1024;

struct TgArgs {
    float* out;
    const float* bias;
    long long osn, osy, osx;
    int oy_mul, oy0;
    int relu;
    int N, Hs, Ws;
    int bx, by;
    int kchunks, Ng, ntaps;
    int map[TG_MAXT], dy[TG_MAXT], dx[TG_MAXT];
    int stages, stage_bytes;
    uint32_t idesc;       // pass 1: N = Ng (TF32) or 2*Ng (3xTF32)
    uint32_t idesc_lo;    // 3xTF32 pass 2: N = Ng
    int total_tiles;
};

__device__ __forceinline__ uint64_t tg_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float tg_lo(float a) {   // tf32(a - trunc_tf32(a)), see conv_tc_x3.cu
    const float r = a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}

template <bool X3>
__global__ void __launch_bounds__(X3 ? TG_THREADS_X3 : TG_THREADS, 1)
conv_tcg_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const TgArgs a) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * a.stage_bytes);
    uint64_t* full = bars;                       // [TG_MAX_STAGES]
    uint64_t* empty = bars + TG_MAX_STAGES;      // [TG_MAX_STAGES]
    uint64_t* tfull = bars + 2 * TG_MAX_STAGES;  // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint64_t* hdone = tempty + 2;                // [TG_MAX_STAGES] X3: pass 1 done, the A box may be rewritten
    uint64_t* lordy = hdone + TG_MAX_STAGES;     // [TG_MAX_STAGES] X3: A_lo written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lordy + TG_MAX_STAGES);
    constexpr int EPI_W0 = X3 ? 4 : 2;           // first epilogue warp
    constexpr int BUF_COLS = X3 ? 256 : 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = a.Ws / a.bx, tiles_y = a.Hs / a.by;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&hdone[s], 1);
            mbar_init(&lordy[s], 2);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BUF_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0) {
        // ================= TMA producer (converged warp, elected lane issues) =================
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = blockIdx.x; mt < a.total_tiles; mt += gridDim.x) {
            const int tx = mt % tiles_x;
            const int ty = (mt / tiles_x) % tiles_y;
            const int n = mt / (tiles_x * tiles_y);
            for (int t = 0; t < a.ntaps; ++t) {
                const CUtensorMap* tm = a.map[t] ? &tmA1 : &tmA0;
                const int x0 = tx * a.bx + a.dx[t], y0 = ty * a.by + a.dy[t];
                for (int cb = 0; cb < a.kchunks; ++cb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) {
                        uint8_t* st = smem + (size_t)stage * a.stage_bytes;
                        mbar_arrive_expect_tx(&full[stage], a.stage_bytes);
                        tma_load_5d(tm, &full[stage], st, 0, cb, x0, y0, n);
                        tma_load_2d(&tmB, &full[stage], st + TG_A_BYTES, (t * a.kchunks + cb) * 32, 0);   // X3: 2*Ng rows (hi | lo)
                    }
                    if (++stage == a.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (converged warp, elected lane issues) =================
        const bool leader = elect_one();
        const int ksteps = a.ntaps * a.kchunks;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        // X3: pass 2 (A_lo * W_hi) of the previous stage, issued one stage late
        bool p_valid = false, p_last = false;
        int p_stage = 0, p_buf = 0;
        uint32_t p_phase = 0, p_dtmem = 0;
        auto pass2 = [&]() {
            mbar_wait(&lordy[p_stage], p_phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + (size_t)p_stage * a.stage_bytes);
            const uint64_t adesc = tg_desc_sw128(st);
            const uint64_t bdesc = tg_desc_sw128(st + TG_A_BYTES);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
                if (leader) umma_tf32(p_dtmem + a.Ng, adesc + 2 * k8, bdesc + 2 * k8, a.idesc_lo, 1u);
            if (leader) {
                umma_commit(&empty[p_stage]);
                if (p_last) umma_commit(&tfull[p_buf]);
            }
        };
        for (int mt = blockIdx.x; mt < a.total_tiles; mt += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t use_parity = (it >> 1) & 1;
            mbar_wait(&tempty[buf], use_parity ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BUF_COLS;
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + (size_t)stage * a.stage_bytes);
                const uint64_t adesc = tg_desc_sw128(st);
                const uint64_t bdesc = tg_desc_sw128(st + TG_A_BYTES);
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)  // 8 tf32 = 32 bytes per MMA
                    if (leader) umma_tf32(d_tmem, adesc + 2 * k8, bdesc + 2 * k8, a.idesc, (ks | k8) != 0 ? 1u : 0u);
                if (X3) {
                    if (leader) umma_commit(&hdone[stage]);
                    if (p_valid) pass2();
                    p_valid = true;
                    p_stage = stage; p_phase = phase; p_buf = buf; p_dtmem = d_tmem;
                    p_last = (ks == ksteps - 1);
                } else {
                    if (leader) umma_commit(&empty[stage]);
                }
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (!X3 && leader) umma_commit(&tfull[buf]);
        }
        if (X3 && p_valid) pass2();
    } else if (X3 && warp < EPI_W0) {
        // ================= split warps (X3): A box <- A_lo in place once pass 1 has read it =================
        const int tid = threadIdx.x - 64;
        const int ksteps = a.ntaps * a.kchunks;
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = blockIdx.x; mt < a.total_tiles; mt += gridDim.x) {
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(&full[stage], phase);
                mbar_wait(&hdone[stage], phase);
                float4* p = reinterpret_cast<float4*>(smem + (size_t)stage * a.stage_bytes);
#pragma unroll 4
                for (int i = tid; i < TG_A_BYTES / 16; i += 64) {
                    float4 v = p[i];
                    v.x = tg_lo(v.x); v.y = tg_lo(v.y); v.z = tg_lo(v.z); v.w = tg_lo(v.w);
                    p[i] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&lordy[stage]);
                if (++stage == a.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ================= epilogue (warps 2..5): thread = pixel row of the tile =================
        const int lane_base = (warp & 3) * 32;
        const int m = lane_base + lane;
        const int yy = m / a.bx, xx = m - yy * a.bx;
        int it = 0;
        for (int mt = blockIdx.x; mt < a.total_tiles; mt += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t use_parity = (it >> 1) & 1;
            const int tx = mt % tiles_x;
            const int ty = (mt / tiles_x) % tiles_y;
            const int n = mt / (tiles_x * tiles_y);
            float* orow = a.out + (size_t)n * a.osn + (size_t)((ty * a.by + yy) * a.oy_mul + a.oy0) * a.osy +
                          (size_t)(tx * a.bx + xx) * a.osx;
            mbar_wait(&tfull[buf], use_parity);
            tc_fence_after();
            for (int c0 = 0; c0 < a.Ng; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + buf * BUF_COLS + c0;
                tmem_ld16(taddr, v);
                if (X3) {
                    uint32_t sm[16];
                    tmem_ld16(taddr + a.Ng, sm);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(sm[q]) + __uint_as_float(v[q]));
                } else {
                    tmem_ld_wait();
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    if (a.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + 4 * q));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (a.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(orow + c0 + 4 * q) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BUF_COLS);
    }
}

static bool tg_pick_patch(int Hs, int Ws, int* bx, int* by) {
    for (int x = 128; x >= 1; x >>= 1) {
        const int y = 128 / x;
        if (x <= 256 && y <= 256 && Ws % x == 0 && Hs % y == 0) {
            *bx = x;
            *by = y;
            return true;
        }
    }
    return false;
}

static bool tg_encode_view(TcEncodeTiledFn enc, CUtensorMap* tm, const LfTcgView& v, int N, int Kc, int bx, int by) {
    // (ci:32, cblk:Kc/32, x, y, n) with the view's element strides
    cuuint64_t dims[5] = {32, (cuuint64_t)(Kc / 32), (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)N};
    cuuint64_t strides[4] = {128, (cuuint64_t)v.sx * 4, (cuuint64_t)v.sy * 4, (cuuint64_t)v.sn * 4};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(v.ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace lf

using namespace lf;

extern "C" int lf_conv_tcg_supported(int N, int Hs, int Ws, int Kc, int Ng) {
    int bx, by;
    if (N <= 0 || Kc < 32 || Kc % 32 != 0 || Kc > 256 || Ng < 16 || Ng % 16 != 0 || Ng > 128) return 0;
    if (!tg_pick_patch(Hs, Ws, &bx, &by)) return 0;
    return tc_get_encode_fn() != nullptr ? 1 : 0;
}

extern "C" int lf_conv_tcg(const LfConvTcgArgs* args, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!args) return LF_ERR_INVALID_ARGUMENT;
    const LfConvTcgArgs& p = *args;
    LF_REQUIRE(p.a[0].ptr && p.wg && p.out && p.ntaps >= 1 && p.ntaps <= TG_MAXT);
    if (!lf_conv_tcg_supported(p.N, p.Hs, p.Ws, p.Kc, p.Ng)) return LF_ERR_UNSUPPORTED;
    LF_REQUIRE(p.osx % 4 == 0 && p.osy % 4 == 0 && p.osn % 4 == 0);
    TcEncodeTiledFn enc = tc_get_encode_fn();
    TgArgs a{};
    tg_pick_patch(p.Hs, p.Ws, &a.bx, &a.by);
    a.out = p.out; a.bias = p.bias; a.relu = p.relu;
    a.osn = p.osn; a.osy = p.osy; a.osx = p.osx; a.oy_mul = p.oy_mul; a.oy0 = p.oy0;
    a.N = p.N; a.Hs = p.Hs; a.Ws = p.Ws;
    a.kchunks = p.Kc / 32; a.Ng = p.Ng; a.ntaps = p.ntaps;
    bool two = false;
    for (int t = 0; t < p.ntaps; ++t) {
        LF_REQUIRE(p.map[t] == 0 || p.map[t] == 1);
        a.map[t] = p.map[t]; a.dy[t] = p.dy[t]; a.dx[t] = p.dx[t];
        two = two || p.map[t] == 1;
    }
    LF_REQUIRE(!two || p.a[1].ptr);
    LF_REQUIRE(p.precision == 0 || p.precision == 1);
    const bool x3 = p.precision == 1;
    const int nb = x3 ? 2 * p.Ng : p.Ng;   // rows of the weight chunk: W, or W_hi | W_lo
    a.stage_bytes = TG_A_BYTES + nb * 128;
    int stages = (TG_SMEM_LIMIT - 1024 - 512) / a.stage_bytes;
    if (stages > TG_MAX_STAGES) stages = TG_MAX_STAGES;
    a.stages = stages;
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nb >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
    a.idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Ng >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
    a.total_tiles = p.N * (p.Hs / a.by) * (p.Ws / a.bx);

    CUtensorMap tmA0, tmA1, tmB;
    if (!tg_encode_view(enc, &tmA0, p.a[0], p.N, p.Kc, a.bx, a.by)) return LF_ERR_CUDA;
    if (!tg_encode_view(enc, &tmA1, two ? p.a[1] : p.a[0], p.N, p.Kc, a.bx, a.by)) return LF_ERR_CUDA;
    {
        // weights [Ng][ntaps*Kc] (K contiguous); 3xTF32: [2][Ng][ntaps*Kc] = hi rows, then lo rows, one box of 2*Ng rows
        const int K = p.ntaps * p.Kc;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)nb};
        cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)nb};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.wg), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return LF_ERR_CUDA;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.total_tiles < sms ? a.total_tiles : sms;
    const int smem_bytes = 1024 + a.stages * a.stage_bytes + 512;
    cudaError_t e = x3 ? cudaFuncSetAttribute(conv_tcg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_LIMIT)
                       : cudaFuncSetAttribute(conv_tcg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_LIMIT);
    if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
    if (x3) lf_launch(conv_tcg_kernel<true>, grid, TG_THREADS_X3, smem_bytes, stream, tmA0, tmA1, tmB, a);
    else lf_launch(conv_tcg_kernel<false>, grid, TG_THREADS, smem_bytes, stream, tmA0, tmA1, tmB, a);
    return check_launch();
}

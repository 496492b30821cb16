// tcgen05 weight-gradient kernel for ERFNet's factorised 3-tap convolutions (C in {64,128}):
//     dW[t][ci][co] = sum_{n,y,x} X[n, y+dy[t], x+dx[t], ci] * dY[n, y, x, co]
// = three GEMMs D_t[ci][co] = A_t[ci][px] * B[co][px]^T whose K dimension is the pixel index.
// Both operands are MN-major for the tensor core: a TMA box [KP pixels x 32 channels] lands in shared
// memory as KP rows of 128 bytes = a column of UMMA MN-major swizzle atoms (32 channels contiguous,
// pixels along K); channel blocks of 32 sit KP*128 bytes apart (the descriptor's LBO).  For 32-bit
// MN-major operands the only legal shared-memory layout is the 128B swizzle with 32-byte atoms
// (UMMA LayoutType SWIZZLE_128B_BASE32B <-> CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: 32B chunk index ^=
// row & 3, 4 K-rows per atom); the plain 16-byte-atom swizzle silently yields zeros.
// The SAME fp32 NHWC tensors feed it that feed the forward kernel -- TF32 multiply, fp32 accumulate.
//
// One persistent CTA owns a contiguous range of pixel patches (split-K), streams
// {dY, X shifted by each tap} through an mbarrier ring (TMA zero-fills out-of-image pixels = conv
// padding), accumulates all three taps in TMEM (3 x C columns) and writes ONE partial
// [3][C][C] at the end; lf_wgrad_reduce sums the partials in fixed order into the reference
// weight layout.   Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue.
//   C = 128: per tap one MMA chain M=128 (ci) x N=128 (co).
//   C =  64: taps 0 and 1 share one M=128 chain (rows 0-63 / 64-127: their X boxes are adjacent in
//            shared memory, so one descriptor spans both); tap 2 runs a second M=128 chain whose upper
//            64 rows read the following (unrelated) boxes and are discarded.
#include <cuda.h>

#include "lf_common.cuh"
#include "tc_ptx.cuh"

namespace lf {

constexpr int WT_THREADS = 192;

struct WtArgs {
    float* partial;  // [nCTA][3][C][C]
    int N, H, W;
    int bx, by;
    int dy[3], dx[3];
    int total_patches;
};

template <int C>
struct WtCfg {
    static constexpr int CB = C / 32;                      // 32-channel blocks
    static constexpr int KP = (C == 128) ? 32 : 64;        // pixels per stage
    static constexpr int BOX_BYTES = KP * 128;             // one [KP px x 32 ch] box
    static constexpr int STAGE_BYTES = 4 * CB * BOX_BYTES; // dY + 3 shifted X   (64 KB)
    static constexpr int STAGES = 3;
    static constexpr int SLACK = 2 * BOX_BYTES;            // C=64: the tap-2 chain over-reads 2 boxes
    static constexpr int NACC = (C == 128) ? 3 : 2;        // accumulators of C columns each
    static constexpr int TMEM_COLS = (C == 128) ? 512 : 128;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + SLACK + 256;
    // cute::UMMA::InstrDescriptor: c=F32, a=b=TF32, a_major=b_major=MN (bits 15,16), N>>3 <<17, M>>4 <<24
    static constexpr uint32_t IDESC =
        (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
};

// MN-major operand, SWIZZLE_128B_BASE32B: 32 MN-elements (128 B) per K-row, 4 K-rows per swizzle atom
// (SBO = 512 B between atoms), next 32-element MN block `lbo_bytes` further.
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout 1 <<61)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

template <int C>
__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WtArgs a) {
    pdl_trigger();
    using Cfg = WtCfg<C>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::SLACK);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* done = bars + 2 * Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (a.total_patches + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * per;
    const int p_end = min(a.total_patches, p_begin + per);
    const int tiles_x = a.W / a.bx, tiles_y = a.H / a.by;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0) {
        // TMA producer: whole warp converged (uniform coordinates / addresses), one elected lane issues
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int p = p_begin; p < p_end; ++p) {
            const int tx = p % tiles_x;
            const int ty = (p / tiles_x) % tiles_y;
            const int n = p / (tiles_x * tiles_y);
            const int x0 = tx * a.bx, y0 = ty * a.by;
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
            if (leader) {
                mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                for (int cb = 0; cb < Cfg::CB; ++cb) tma_load_5d(&tmDY, &full[stage], st + cb * Cfg::BOX_BYTES, 0, cb, x0, y0, n);
                for (int t = 0; t < 3; ++t)
                    for (int cb = 0; cb < Cfg::CB; ++cb)
                        tma_load_5d(&tmX, &full[stage], st + ((1 + t) * Cfg::CB + cb) * Cfg::BOX_BYTES, 0, cb, x0 + a.dx[t],
                                    y0 + a.dy[t], n);
            }
            if (++stage == Cfg::STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp converged so the descriptors stay in uniform registers (see elect_one())
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        bool first = true;
        for (int p = p_begin; p < p_end; ++p) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k8 = 0; k8 < Cfg::KP / 8; ++k8) {
                const uint32_t koff = k8 * 1024;  // 8 pixel rows
                const uint64_t bdesc = umma_desc_mn_sw128(st + koff, Cfg::BOX_BYTES);
                const uint32_t acc = (first && k8 == 0) ? 0u : 1u;
                if (C == 128) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const uint64_t adesc = umma_desc_mn_sw128(st + (1 + t) * Cfg::CB * Cfg::BOX_BYTES + koff, Cfg::BOX_BYTES);
                        if (leader) umma_tf32(tmem_base + t * C, adesc, bdesc, Cfg::IDESC, acc);
                    }
                } else {
                    // rows 0-63 = tap 0, rows 64-127 = tap 1 (adjacent boxes)
                    const uint64_t ad01 = umma_desc_mn_sw128(st + 1 * Cfg::CB * Cfg::BOX_BYTES + koff, Cfg::BOX_BYTES);
                    // rows 0-63 = tap 2, rows 64-127 = whatever follows (discarded)
                    const uint64_t ad2 = umma_desc_mn_sw128(st + 3 * Cfg::CB * Cfg::BOX_BYTES + koff, Cfg::BOX_BYTES);
                    if (leader) {
                        umma_tf32(tmem_base, ad01, bdesc, Cfg::IDESC, acc);
                        umma_tf32(tmem_base + C, ad2, bdesc, Cfg::IDESC, acc);
                    }
                }
            }
            first = false;
            if (leader) umma_commit(&empty[stage]);
            if (++stage == Cfg::STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
        if (leader) umma_commit(done);
    } else {
        // epilogue: TMEM lane = GEMM row
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        float* dst = a.partial + (size_t)blockIdx.x * 3 * C * C;
        if (p_begin < p_end) {
            mbar_wait(done, 0);
            tc_fence_after();
        }
        for (int acc = 0; acc < Cfg::NACC; ++acc) {
            // which (tap, ci) does this accumulator row hold?
            int t, ci;
            bool valid = true;
            if (C == 128) {
                t = acc;
                ci = row;
            } else {
                t = (acc == 0) ? (row >> 6) : 2;
                ci = row & 63;
                valid = (acc == 0) || row < 64;
            }
            float* drow = dst + ((size_t)t * C + ci) * C;
            for (int c0 = 0; c0 < C; c0 += 16) {
                uint32_t v[16];
                if (p_begin < p_end) {
                    tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + acc * C + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = 0u;
                }
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(drow + c0 + 4 * q) =
                            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                        __uint_as_float(v[4 * q + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

static bool pick_patch_kp(int H, int W, int kp, int* bx, int* by) {
    for (int x = kp; x >= 1; x >>= 1) {
        const int y = kp / x;
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

extern "C" int lf_wgrad3_tc_ctas(int N, int H, int W, int C) {
    int bx, by;
    if (!(C == 64 || C == 128) || N <= 0) return 0;
    const int kp = (C == 128) ? 32 : 64;
    if (!pick_patch_kp(H, W, kp, &bx, &by)) return 0;
    if (!tc_get_encode_fn()) return 0;
    const long long patches = (long long)N * (H / by) * (W / bx);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // at least 4 patches per CTA so the split-K partials stay small next to the streamed operands
    long long ctas = patches / 4;
    if (ctas > sms) ctas = sms;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

extern "C" int lf_wgrad3_tc(const float* x, const float* dy, int N, int H, int W, int C, const int* tap_dy, const int* tap_dx,
                            float* partial, int nctas, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LF_REQUIRE(x && dy && partial && tap_dy && tap_dx && nctas >= 1);
    if (!(C == 64 || C == 128)) return LF_ERR_UNSUPPORTED;
    WtArgs a{};
    const int kp = (C == 128) ? 32 : 64;
    if (!pick_patch_kp(H, W, kp, &a.bx, &a.by)) return LF_ERR_UNSUPPORTED;
    TcEncodeTiledFn enc = tc_get_encode_fn();
    if (!enc) return LF_ERR_UNSUPPORTED;
    a.partial = partial; a.N = N; a.H = H; a.W = W;
    for (int t = 0; t < 3; ++t) {
        a.dy[t] = tap_dy[t];
        a.dx[t] = tap_dx[t];
    }
    a.total_patches = N * (H / a.by) * (W / a.bx);
    CUtensorMap tmX, tmDY;
    if (!tc_encode_nhwc_map(enc, &tmX, x, N, H, W, C, a.bx, a.by, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return LF_ERR_CUDA;
    if (!tc_encode_nhwc_map(enc, &tmDY, dy, N, H, W, C, a.bx, a.by, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return LF_ERR_CUDA;
    cudaError_t e;
    if (C == 128) {
        e = cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WtCfg<128>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(wgrad_tc_kernel<128>, nctas, WT_THREADS, WtCfg<128>::SMEM_BYTES, stream, tmX, tmDY, a);
    } else {
        e = cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WtCfg<64>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(wgrad_tc_kernel<64>, nctas, WT_THREADS, WtCfg<64>::SMEM_BYTES, stream, tmX, tmDY, a);
    }
    return check_launch();
}

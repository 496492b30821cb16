// 3xTF32 ("fp32-grade") tcgen05 weight-gradient kernel for ERFNet's factorised 3-tap convolutions (C in {64,128}):
//     dW[t][ci][co] = sum_{n,y,x} X[n, y+dy[t], x+dx[t], ci] * dY[n, y, x, co]
// Same GEMM view, MN-major operands (128B swizzle, 32-byte atoms), split-K over pixels and partial / reduce protocol as
// wgrad_tc.cu (read its header first).  Both operands are activations here, so both are split on the fly
// (x = x_hi + x_lo, the tensor core truncates fp32 to TF32; see conv_tc_x3.cu) and every product becomes
//     X_hi*G_hi + (X_hi*G_lo + X_lo*G_hi),   fp32 accumulate in TMEM:
//  * stage = [dY box][room for dY_lo][X shifted by tap 0, 1, 2]; TMA fills dY and the X boxes;
//  * split warps write dY_lo = tf32(dY - trunc(dY)) next to dY (channel blocks stay LBO apart, so [dY | dY_lo] is ONE
//    MN-major operand of twice the N extent);
//  * MMA pass 1 on the raw X boxes:  C = 64: X_hi * [G_hi | G_lo]^T (N = 128: columns [0,64) large term, [64,128)
//    correction);  C = 128: X_hi*G_hi and X_hi*G_lo as two N = 128 instructions into the same columns (3 taps x 256
//    columns would not fit TMEM);
//  * when pass 1 of a stage has completed the split warps rewrite the X boxes in place with X_lo, and pass 2
//    (X_lo * G_hi^T) is issued one stage later, so the tensor pipe does not wait for the rewrite.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = split warps, 6..9 = epilogue.
#include <cuda.h>

#include "lf_common.cuh"
#include "tc_ptx.cuh"

namespace lf {

constexpr int WX_THREADS = 320;

struct WxArgs {
    float* partial;  // [nCTA][3][C][C]
    int N, H, W;
    int bx, by;
    int dy[3], dx[3];
    int total_patches;
};

template <int C>
struct WxCfg {
    static constexpr int CB = C / 32;                      // 32-channel blocks
    static constexpr int KP = (C == 128) ? 16 : 32;        // pixels per stage
    static constexpr int BOX_BYTES = KP * 128;             // one [KP px x 32 ch] box
    static constexpr int STAGE_BYTES = 5 * CB * BOX_BYTES; // dY, dY_lo, 3 shifted X   (40 KB)
    static constexpr int LOAD_BYTES = 4 * CB * BOX_BYTES;  // what TMA writes per stage
    static constexpr int STAGES = 5;
    static constexpr int SLACK = 2 * BOX_BYTES;            // C=64: the tap-2 chain over-reads 2 boxes
    static constexpr bool NCONCAT = (C == 64);
    static constexpr int NACC = (C == 128) ? 3 : 2;        // accumulators
    static constexpr int ACC_COLS = 128;                   // C=128: 128 co; C=64: [64 co large | 64 co correction]
    static constexpr int TMEM_COLS = (C == 128) ? 512 : 256;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + SLACK + 256;
    // cute::UMMA::InstrDescriptor: c=F32, a=b=TF32, a_major=b_major=MN (bits 15,16), N>>3 <<17, M>>4 <<24
    static constexpr uint32_t idesc(int n) {
        return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    }
};

// MN-major operand, SWIZZLE_128B_BASE32B (see wgrad_tc.cu)
__device__ __forceinline__ uint64_t wx_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ float wx_lo(float a) {
    const float r = a - __uint_as_float(__float_as_uint(a) & 0xffffe000u);       // a - (what the tensor core sees)
    return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);       // rounded to TF32
}
__device__ __forceinline__ float4 wx_lo4(float4 v) { return make_float4(wx_lo(v.x), wx_lo(v.y), wx_lo(v.z), wx_lo(v.w)); }

template <int C>
__global__ void __launch_bounds__(WX_THREADS, 1)
wgrad_tc_x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WxArgs a) {
    pdl_trigger();
    using Cfg = WxCfg<C>;
    constexpr int CB = Cfg::CB, BOX = Cfg::BOX_BYTES, S = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES + Cfg::SLACK);
    uint64_t* full = bars;           // TMA landed
    uint64_t* empty = bars + S;      // pass 2 done: stage may be refilled
    uint64_t* glo = bars + 2 * S;    // dY_lo written
    uint64_t* hdone = bars + 3 * S;  // pass 1 done: X may be rewritten
    uint64_t* lordy = bars + 4 * S;  // X_lo written
    uint64_t* done = bars + 5 * S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (a.total_patches + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * per;
    const int p_end = min(a.total_patches, p_begin + per);
    const int tiles_x = a.W / a.bx, tiles_y = a.H / a.by;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&glo[s], 4);
            mbar_init(&hdone[s], 1);
            mbar_init(&lordy[s], 4);
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
                mbar_arrive_expect_tx(&full[stage], Cfg::LOAD_BYTES);
                for (int cb = 0; cb < CB; ++cb) tma_load_5d(&tmDY, &full[stage], st + cb * BOX, 0, cb, x0, y0, n);
                for (int t = 0; t < 3; ++t)
                    for (int cb = 0; cb < CB; ++cb)
                        tma_load_5d(&tmX, &full[stage], st + ((2 + t) * CB + cb) * BOX, 0, cb, x0 + a.dx[t], y0 + a.dy[t], n);
            }
            if (++stage == S) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp converged so the descriptors stay in uniform registers (see elect_one())
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        bool first = true, p_valid = false;
        int p_stage = 0;
        uint32_t p_phase = 0;
        // one pass over a stage: A = the X boxes (raw or rewritten), B = `b_off` bytes into the stage, N = n columns,
        // accumulating into column offset `col` of every accumulator
        auto pass = [&](int stg, uint32_t b_off, uint32_t idesc, uint32_t col, bool fresh) {
            const uint32_t st = smem_u32(smem + stg * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k8 = 0; k8 < Cfg::KP / 8; ++k8) {
                const uint32_t koff = k8 * 1024;  // 8 pixel rows
                const uint64_t bdesc = wx_desc_mn(st + b_off + koff, BOX);
                const uint32_t acc = (fresh && k8 == 0) ? 0u : 1u;
                if (C == 128) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const uint64_t adesc = wx_desc_mn(st + (2 + t) * CB * BOX + koff, BOX);
                        if (leader) umma_tf32(tmem_base + t * Cfg::ACC_COLS + col, adesc, bdesc, idesc, acc);
                    }
                } else {
                    // rows 0-63 = tap 0, rows 64-127 = tap 1 (adjacent boxes); second chain: rows 0-63 = tap 2,
                    // rows 64-127 = whatever follows (discarded)
                    const uint64_t ad01 = wx_desc_mn(st + 2 * CB * BOX + koff, BOX);
                    const uint64_t ad2 = wx_desc_mn(st + 4 * CB * BOX + koff, BOX);
                    if (leader) {
                        umma_tf32(tmem_base + col, ad01, bdesc, idesc, acc);
                        umma_tf32(tmem_base + Cfg::ACC_COLS + col, ad2, bdesc, idesc, acc);
                    }
                }
            }
        };
        auto pass2 = [&]() {   // X_lo * G_hi of the previous stage
            mbar_wait(&lordy[p_stage], p_phase);
            tc_fence_after();
            pass(p_stage, 0, Cfg::idesc(C), Cfg::NCONCAT ? 64 : 0, false);
            if (leader) umma_commit(&empty[p_stage]);
        };
        for (int p = p_begin; p < p_end; ++p) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (Cfg::NCONCAT) {
                mbar_wait(&glo[stage], phase);
                tc_fence_after();
                pass(stage, 0, Cfg::idesc(2 * C), 0, first);          // X_hi * [G_hi | G_lo]
            } else {
                pass(stage, 0, Cfg::idesc(C), 0, first);              // X_hi * G_hi
                mbar_wait(&glo[stage], phase);
                tc_fence_after();
                pass(stage, CB * BOX, Cfg::idesc(C), 0, false);       // X_hi * G_lo
            }
            first = false;
            if (leader) umma_commit(&hdone[stage]);
            if (p_valid) pass2();
            p_valid = true;
            p_stage = stage;
            p_phase = phase;
            if (++stage == S) {
                stage = 0;
                phase ^= 1;
            }
        }
        if (p_valid) pass2();
        if (leader) umma_commit(done);
    } else if (warp < 6) {
        // split warps.  dY_lo of stage p+1 is produced BEFORE waiting for pass 1 of stage p (its TMA landed long ago), so
        // the MMA issuer never waits for it behind the X rewrite of the previous stage.
        const int tid = threadIdx.x - 64;
        int stage = 0;
        uint32_t phase = 0;
        constexpr int GV = CB * BOX / 16, XV = 3 * CB * BOX / 16;   // float4 counts
        auto make_glo = [&](int stg, uint32_t ph) {
            uint8_t* st = smem + stg * Cfg::STAGE_BYTES;
            mbar_wait(&full[stg], ph);
            const float4* g = reinterpret_cast<const float4*>(st);
            float4* gl = reinterpret_cast<float4*>(st + CB * BOX);
#pragma unroll
            for (int i = tid; i < GV; i += 128) gl[i] = wx_lo4(g[i]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&glo[stg]);
        };
        if (p_begin < p_end) make_glo(0, 0);
        for (int p = p_begin; p < p_end; ++p) {
            const int nstage = (stage + 1 == S) ? 0 : stage + 1;
            const uint32_t nphase = (stage + 1 == S) ? phase ^ 1 : phase;
            if (p + 1 < p_end) make_glo(nstage, nphase);
            uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
            mbar_wait(&hdone[stage], phase);
            {
                float4* x = reinterpret_cast<float4*>(st + 2 * CB * BOX);
#pragma unroll 4
                for (int i = tid; i < XV; i += 128) x[i] = wx_lo4(x[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&lordy[stage]);
            stage = nstage;
            phase = nphase;
        }
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
                    const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + acc * Cfg::ACC_COLS + c0;
                    tmem_ld16(taddr, v);
                    if (Cfg::NCONCAT) {
                        uint32_t s[16];
                        tmem_ld16(taddr + 64, s);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(s[q]) + __uint_as_float(v[q]));
                    } else {
                        tmem_ld_wait();
                    }
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

static bool wx_pick_patch(int H, int W, int kp, int* bx, int* by) {
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

extern "C" int lf_wgrad3_tc_x3_ctas(int N, int H, int W, int C) {
    int bx, by;
    if (!(C == 64 || C == 128) || N <= 0) return 0;
    const int kp = (C == 128) ? WxCfg<128>::KP : WxCfg<64>::KP;
    if (!wx_pick_patch(H, W, kp, &bx, &by)) return 0;
    if (!tc_get_encode_fn()) return 0;
    const long long patches = (long long)N * (H / by) * (W / bx);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // at least 8 patches per CTA so the split-K partials stay small next to the streamed operands
    long long ctas = patches / 8;
    if (ctas > sms) ctas = sms;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

extern "C" int lf_wgrad3_tc_x3(const float* x, const float* dy, int N, int H, int W, int C, const int* tap_dy, const int* tap_dx,
                               float* partial, int nctas, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LF_REQUIRE(x && dy && partial && tap_dy && tap_dx && nctas >= 1);
    if (!(C == 64 || C == 128)) return LF_ERR_UNSUPPORTED;
    WxArgs a{};
    const int kp = (C == 128) ? WxCfg<128>::KP : WxCfg<64>::KP;
    if (!wx_pick_patch(H, W, kp, &a.bx, &a.by)) return LF_ERR_UNSUPPORTED;
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
        e = cudaFuncSetAttribute(wgrad_tc_x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WxCfg<128>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(wgrad_tc_x3_kernel<128>, nctas, WX_THREADS, WxCfg<128>::SMEM_BYTES, stream, tmX, tmDY, a);
    } else {
        e = cudaFuncSetAttribute(wgrad_tc_x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WxCfg<64>::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_cuda_error(e); return LF_ERR_CUDA; }
        lf_launch(wgrad_tc_x3_kernel<64>, nctas, WX_THREADS, WxCfg<64>::SMEM_BYTES, stream, tmX, tmDY, a);
    }
    return check_launch();
}

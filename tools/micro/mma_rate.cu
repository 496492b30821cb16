// Micro-benchmark: issue rate of tcgen05.mma (SS mode, K-major SW128 operands in shared memory) on sm_100a for
// kind::tf32 and kind::f16 (bf16), N in {64,128,256}, 1/2/4 rotating accumulators.  No global loads: operands are
// whatever is in shared memory (zeros).  Prints cycles per MMA and the implied MAC/cycle/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lanedetection_end2end_b200/csrc/tc_ptx.cuh"
using namespace lf;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// kind: 0 = tf32, 1 = bf16, 2 = tf32 with A in tensor memory (TS); M: 128 or 64
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int kind, int M, int N, int nacc, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    if (threadIdx.x < 32) {  // converged warp, elected lane issues (operands stay in uniform registers)
        const bool leader = elect_one();
        const uint32_t fmt = kind == 1 ? 1u : 2u;  // BF16 : TF32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
        const uint32_t a_tm = tm + 256;            // TS mode: A lives in columns 256.. (accumulators stay below 256)
        const uint64_t a = desc_sw128(smem_u32(smem)), b = desc_sw128(smem_u32(smem + 16 * 1024));
        // fully unrolled groups of 8 (no per-MMA integer work): accumulator = (j % nacc) * N precomputed
        uint32_t dsel[8];
        for (int j = 0; j < 8; ++j) dsel[j] = tm + (uint32_t)(((j % nacc) * N) % 512);
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (!leader) continue;
                if (kind == 0)
                    umma_tf32(dsel[j], a + 2 * (j & 3), b + 2 * (j & 3), idesc, 1u);
                else if (kind == 2)
                    umma_tf32_ts(dsel[j], a_tm + 8 * (j & 3), b + 2 * (j & 3), idesc, 1u);
                else
                    umma_f16(dsel[j], a + 2 * (j & 3), b + 2 * (j & 3), idesc, 1u);
            }
        }
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (leader && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tm, 512);
    }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 4096;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device clock attr %d kHz\n", clk_khz);
    for (int grid : {1, 148}) {
        for (int kind = 0; kind < 3; ++kind)
          for (int M : {128, 64})
            for (int N : {64, 128, 256})
                for (int nacc : {1, 2, 4}) {
                    if (nacc * N > (kind == 2 ? 256 : 512)) continue;
                    if (M == 64 && nacc != 2) continue;
                    mma_rate_kernel<<<grid, 128, 64 * 1024>>>(kind, M, N, nacc, reps, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("error %s\n", cudaGetErrorString(e));
                        return 1;
                    }
                    long long cyc;
                    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
                    const double per = (double)cyc / reps;
                    const int K = kind == 1 ? 16 : 8;
                    printf("{\"grid\": %d, \"kind\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"accumulators\": %d, \"cycles_per_mma\": %.1f, "
                           "\"mac_per_cycle_per_sm\": %.0f}\n",
                           grid, kind == 0 ? "tf32" : kind == 1 ? "bf16" : "tf32_ts", M, N, K, nacc, per, (double)M * N * K / per);
                }
    }
    return 0;
}

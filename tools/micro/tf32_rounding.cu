// Micro-experiment (for the round-2 "3xTF32" accurate mode): how does tcgen05.mma.kind::tf32 turn an fp32 operand read
// from shared memory into TF32 -- truncation, round-to-nearest-even, or round-to-nearest-away?  The split
// a = a_hi + a_lo must use the SAME conversion for a_hi that the tensor core applies to the raw fp32 value.
// One MMA: A[0][0] = test value, B[0][0] = 1, everything else 0  ->  D[0][0] = tf32(A[0][0]).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_rounding tf32_rounding.cu && ./tf32_rounding
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../lanedetection_end2end_b200/csrc/tc_ptx.cuh"
using namespace lf;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) tf32_round_kernel(const float* vals, int n, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* sA = reinterpret_cast<float*>(smem);              // [128 rows][32 floats] K-major SW128 (row 0: no permutation)
    float* sB = reinterpret_cast<float*>(smem + 16 * 1024);  // [64 rows][32 floats]
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int i = 0; i < n; ++i) {
        for (int j = threadIdx.x; j < 24 * 1024 / 4; j += blockDim.x) reinterpret_cast<uint32_t*>(smem)[j] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            sA[0] = vals[i];   // A[m=0][k=0]
            sB[0] = 1.0f;      // B[n=0][k=0]
        }
        // make the generic-proxy writes visible to the async (tensor core) proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) {
            const bool leader = elect_one();
            if (leader) {
                umma_tf32(tm, desc_sw128(smem_u32(sA)), desc_sw128(smem_u32(sB)), idesc, 0u);
                umma_commit(&bar);
            }
            mbar_wait(&bar, i & 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tm, v);
            tmem_ld_wait();
            if (threadIdx.x == 0) out[i] = __uint_as_float(v[0]);
        }
        tc_fence_before();
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tm, 64);
    }
}

static float from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t to_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

int main() {
    // 1.0 = 0x3f800000; tf32 keeps 10 mantissa bits -> the low 13 bits are dropped
    const uint32_t cases[] = {
        0x3f800000u | 0x0fffu,   // just below half an ulp      : trunc 1.0 | RN 1.0
        0x3f800000u | 0x1000u,   // exactly half an ulp (tie)    : trunc 1.0 | RNE 1.0 | RNA 1+ulp
        0x3f800000u | 0x1001u,   // just above half              : trunc 1.0 | RN 1+ulp
        0x3f800000u | 0x3000u,   // 1.5 ulp (tie, odd)           : trunc 1+ulp | RNE 1+2ulp | RNA 1+2ulp
        0x3f800000u | 0x1fffu,   // almost one ulp               : trunc 1.0 | RN 1+ulp
        0xbf800000u | 0x1001u,   // negative, just above half
    };
    const int n = sizeof(cases) / sizeof(cases[0]);
    float h[n], *d_in, *d_out, r[n];
    for (int i = 0; i < n; ++i) h[i] = from_bits(cases[i]);
    cudaMalloc(&d_in, n * 4);
    cudaMalloc(&d_out, n * 4);
    cudaMemcpy(d_in, h, n * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(tf32_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    tf32_round_kernel<<<1, 128, 32 * 1024>>>(d_in, n, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(r, d_out, n * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) {
        const uint32_t in = cases[i], out = to_bits(r[i]);
        const uint32_t trunc = in & ~0x1fffu, rna = (in + 0x1000u) & ~0x1fffu;
        const uint32_t rne = (in + 0x0fffu + ((in >> 13) & 1u)) & ~0x1fffu;
        printf("{\"in\": \"0x%08x\", \"out\": \"0x%08x\", \"is_trunc\": %d, \"is_rne\": %d, \"is_rna\": %d}\n", in, out, out == trunc,
               out == rne, out == rna);
    }
    return 0;
}

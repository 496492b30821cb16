// Shared pieces of the two tcgen05 3-tap convolution kernels (conv_tc.cu: TF32 multiply; conv_tc_x3.cu: 3xTF32 =
// fp32-grade products): kernel arguments, the K-major UMMA descriptor and the epilogue warps
// (TMEM -> registers -> staging tile -> coalesced stores with the fused ReLU / mask / residual-add / column statistics).
#pragma once
#include <cuda.h>

#include "lf_common.cuh"
#include "tc_ptx.cuh"

namespace lf {

constexpr int TC_BM = 128;
constexpr int TC_BN = 64;
constexpr int TC_KCH = 32;                    // fp32 elements per 128-byte swizzle row
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_STG_LD = 32 + 4;                         // staging row stride in floats: a 32-channel half + bank spread
constexpr int TC_STG_BYTES = TC_BM * TC_STG_LD * 4;       // 18 KB
constexpr int TC_SMEM_LIMIT = 226 * 1024;  // 227 KB opt-in maximum minus the 1 KB static epilogue scratch

struct TcArgs {
    float* out;
    const float* bias;
    const float* mask_src;
    const float* add_src;
    const float* add_mask;
    float* colsum_partial;  // [gridDim.x / n_halves][Ctot] per-CTA column sums of the output, or NULL
    double* stats_partial;  // [gridDim.x / n_halves][2][Ctot] per-CTA sum and sum of squares (BatchNorm), or NULL
    const float* mask_scale;  // non-NULL (with mask_shift): mask_src is a PRE-activation x; the ReLU mask bit is
    const float* mask_shift;  // fma(x, scale[c], shift[c]) > 0 and the second statistic = sum out*x  (BatchNorm backward)
    int N, H, W, Ctot;
    int vertical;           // conv axis: 1 = y (3x1), 0 = x (1x3)
    int TA, TB;             // tile extent along / across the conv axis (TA*TB = 128)
    int tb_shift;           // log2(TB)
    int dil;                // tap spacing d
    int tap_row[3];         // first slab row (128-byte rows) of the view used by weight slot t
    int tiles_a, tiles_b;
    int relu;               // bit 0: ReLU on conv + bias (before mask / add); bit 1: ReLU after the residual add
    int n_halves;
    int total_m_tiles;
    int stages, stage_bytes;
    int slab_rows;          // (TA + 2*dil) * TB
    int debug;              // timing experiments (lf_conv1d_tc_set_debug): results are garbage when != 0
};

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | SWIZZLE_128B(2) <<61)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Epilogue warps of one CTA (4 * EPI_GROUPS warps starting at thread EPI_T0).  Tiles mt = cta_m, cta_m + m_stride, ...;
// tile i lives in TMEM accumulator buffer i % NBUF (BUF_COLS columns apart).  X3: the buffer holds the large term in
// columns [0,64) and the sum of the two small correction terms in [64,128); the result is small + large.
// Group g of 4 warps (one warp per TMEM lane quadrant) walks HALVES 32-channel halves of every tile:
//  (1) tcgen05.ld 32x32b (thread = pixel row), staged into a padded shared-memory half-tile (18 KB);
//  (2) the half-tile is walked row-major so that 8 consecutive threads cover one pixel's 32 channels (one 128-byte
//      line): bias (this thread's four columns, held in registers for the whole kernel -- loading it per tile in (1) was
//      the top stall line of the round-2 ncu source view), ReLU, ReLU-backward mask, residual-gradient add, column sums /
//      sums of squares (bias gradient, BatchNorm statistics) and the output stores are fully coalesced 128-bit accesses.
// The mask / residual operands of a tile are prefetched into registers BEFORE waiting for its accumulator (AHEAD: one
// whole tile ahead), so their latency hides behind the MMAs.
template <int EPI_GROUPS, int AHEAD, int EPI_T0, int NBUF, int BUF_COLS, bool X3>
__device__ __forceinline__ void tc_epilogue(const TcArgs& a, float* stg, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                            int n_half, int cta_m, int m_stride) {
    constexpr int NH = 2 / EPI_GROUPS;      // 32-channel halves each group walks per tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may access
    const int m = lane_base + lane;         // phase 1: this thread's pixel row of the tile
    const int grp = (threadIdx.x - EPI_T0) >> 7;  // epilogue group (0 when there is only one)
    const int et = (threadIdx.x - EPI_T0) & 127;  // 0..127 inside the group
    const int c4 = et & 7;                  // phase 2: float4 column inside a 32-channel half (fixed)
    const int r0 = et >> 3;                 // phase 2: first of this thread's 8 rows (r0, r0+16, ...)
    const int tb_shift = a.tb_shift;
    const int h_first = (EPI_GROUPS == 2) ? grp : 0;  // first (only) channel half of this group
    float* stg_g = stg + grp * (TC_STG_BYTES / 4);
    const int bar_id = 1 + grp;
    const bool pre_mask = a.mask_src != nullptr;
    const bool pre_add = (a.add_src != nullptr) && !pre_mask;  // both given: add_src is read in the loop
    float4 csum[NH], csq[NH];               // running column sums / sums of squares, per channel half
    float4 msc[NH], msh[NH];                // BatchNorm scale / shift of this thread's columns (mask_scale mode)
    float4 bsv[NH];                         // bias of this thread's (fixed) phase-2 columns: loaded once, not per tile
    const bool affine_mask = a.mask_scale != nullptr;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
        csum[hh] = csq[hh] = msc[hh] = msh[hh] = bsv[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias) bsv[hh] = __ldg(reinterpret_cast<const float4*>(a.bias + n_half * TC_BN + 32 * (h_first + hh) + 4 * c4));
        if (affine_mask) {
            msc[hh] = __ldg(reinterpret_cast<const float4*>(a.mask_scale + n_half * TC_BN + 32 * (h_first + hh) + 4 * c4));
            msh[hh] = __ldg(reinterpret_cast<const float4*>(a.mask_shift + n_half * TC_BN + 32 * (h_first + hh) + 4 * c4));
        }
    }
    // AHEAD: 0 = operands fetched just before the accumulator wait; 1 = ONE operand (the ReLU mask, or a pre-masked residual
    // gradient) a whole tile ahead; 2 = add_src and add_mask both a tile ahead (32 more registers)
    float4 nxt_a[AHEAD ? 8 : 1], nxt_m[AHEAD == 2 ? 8 : 1];  // operands of the next tile in flight
    int it = 0;
    for (int mt = cta_m; mt < a.total_m_tiles; mt += m_stride, ++it) {
        const int buf = it % NBUF;
        const uint32_t use_parity = (it / NBUF) & 1;
        const int ta = mt % a.tiles_a;
        const int tb = (mt / a.tiles_a) % a.tiles_b;
        const int n = mt / (a.tiles_a * a.tiles_b);
        // global offsets of this thread's 8 phase-2 rows (channel 4*c4 of the group's first half)
        size_t roff[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = r0 + 16 * j;
            const int ap = r >> tb_shift, bp = r & (a.TB - 1);  // slab order: cross axis fastest
            const int pa = ta * a.TA + ap, pb = tb * a.TB + bp;
            const int y = a.vertical ? pa : pb, x = a.vertical ? pb : pa;
            roff[j] = ((size_t)(n * a.H + y) * a.W + x) * a.Ctot + n_half * TC_BN + 32 * h_first + 4 * c4;
        }
        // Prefetch the ReLU mask (or the gated residual gradient).  One half per thread (two groups): ONE TILE AHEAD —
        // the loads for tile i+1 are issued before tile i is processed, so their latency hides behind a whole
        // tile of epilogue work (the epilogue, not the tensor pipe, bounds these layers).  Two halves per
        // thread (one group): no registers for that; prefetch this tile's operands before waiting on the accumulator.
        float4 pre[NH][8];
        if constexpr (NH == 1 && AHEAD != 0) {
            if (pre_mask || pre_add) {
                auto issue = [&](int mtn) {
                    const int tan = mtn % a.tiles_a;
                    const int tbn = (mtn / a.tiles_a) % a.tiles_b;
                    const int nn = mtn / (a.tiles_a * a.tiles_b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = r0 + 16 * j;
                        const int ap = r >> tb_shift, bp = r & (a.TB - 1);
                        const int pa = tan * a.TA + ap, pb = tbn * a.TB + bp;
                        const int y = a.vertical ? pa : pb, x = a.vertical ? pb : pa;
                        const size_t off = ((size_t)(nn * a.H + y) * a.W + x) * a.Ctot + n_half * TC_BN + 32 * h_first + 4 * c4;
                        nxt_a[j] = __ldg(reinterpret_cast<const float4*>((pre_mask ? a.mask_src : a.add_src) + off));
                        if constexpr (AHEAD == 2) {
                            if (!pre_mask && a.add_mask) nxt_m[j] = __ldg(reinterpret_cast<const float4*>(a.add_mask + off));
                        }
                    }
                };
                if (it == 0) issue(mt);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = nxt_a[j];
                    if constexpr (AHEAD == 2) {
                        if (!pre_mask && a.add_mask) {
                            const float4 mk = nxt_m[j];
                            v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
                            v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
                        }
                    }
                    pre[0][j] = v;
                }
                if (mt + m_stride < a.total_m_tiles) issue(mt + m_stride);
            }
        } else if (pre_mask || pre_add) {
#pragma unroll
            for (int hh = 0; hh < NH; ++hh)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (pre_mask) {
                        pre[hh][j] = __ldg(reinterpret_cast<const float4*>(a.mask_src + roff[j] + 32 * hh));
                    } else {
                        float4 ad = __ldg(reinterpret_cast<const float4*>(a.add_src + roff[j] + 32 * hh));
                        if (a.add_mask) {
                            const float4 mk = __ldg(reinterpret_cast<const float4*>(a.add_mask + roff[j] + 32 * hh));
                            ad.x = mk.x > 0.f ? ad.x : 0.f; ad.y = mk.y > 0.f ? ad.y : 0.f;
                            ad.z = mk.z > 0.f ? ad.z : 0.f; ad.w = mk.w > 0.f ? ad.w : 0.f;
                        }
                        pre[hh][j] = ad;
                    }
                }
        }
        mbar_wait(&tfull[buf], use_parity);
        tc_fence_after();
        if (a.debug & 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
            continue;
        }
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            const int h = h_first + hh;
            // ---- phase 1: TMEM -> registers -> (+bias, ReLU) -> staging half-tile, one pixel row per thread
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + buf * BUF_COLS + 32 * h + c0;
                tmem_ld16(taddr, v);
                if constexpr (X3) {
                    uint32_t s[16];
                    tmem_ld16(taddr + TC_BN, s);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(s[q]) + __uint_as_float(v[q]));
                } else {
                    tmem_ld_wait();
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)   // raw accumulator values: bias / ReLU are applied in phase 2 (bias in registers)
                    *reinterpret_cast<float4*>(&stg_g[m * TC_STG_LD + c0 + 4 * q]) =
                        make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                    __uint_as_float(v[4 * q + 3]));
            }
            if (hh == NH - 1) tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // staging complete (this group only)
            if (hh == NH - 1 && lane == 0) mbar_arrive(&tempty[buf]);  // this warp's TMEM reads are done
            // ---- phase 2: row-major walk, 8 consecutive threads = one pixel's 32 channels (128 B)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r0 + 16 * j;
                const size_t off = roff[j] + 32 * hh;
                float4 o = *reinterpret_cast<const float4*>(&stg_g[r * TC_STG_LD + 4 * c4]);
                o.x += bsv[hh].x; o.y += bsv[hh].y; o.z += bsv[hh].z; o.w += bsv[hh].w;
                if (a.relu & 1) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                if (pre_mask) {
                    float4 mk = pre[hh][j];
                    if (affine_mask) {   // the same fma the forward BatchNorm-apply performed before its ReLU: bit-identical mask
                        mk.x = fmaf(mk.x, msc[hh].x, msh[hh].x); mk.y = fmaf(mk.y, msc[hh].y, msh[hh].y);
                        mk.z = fmaf(mk.z, msc[hh].z, msh[hh].z); mk.w = fmaf(mk.w, msc[hh].w, msh[hh].w);
                    }
                    o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
                    o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
                }
                if (pre_add) {
                    o.x += pre[hh][j].x; o.y += pre[hh][j].y; o.z += pre[hh][j].z; o.w += pre[hh][j].w;
                } else if (a.add_src) {
                    float4 ad = __ldg(reinterpret_cast<const float4*>(a.add_src + off));
                    if (a.add_mask) {
                        const float4 mk = __ldg(reinterpret_cast<const float4*>(a.add_mask + off));
                        ad.x = mk.x > 0.f ? ad.x : 0.f; ad.y = mk.y > 0.f ? ad.y : 0.f;
                        ad.z = mk.z > 0.f ? ad.z : 0.f; ad.w = mk.w > 0.f ? ad.w : 0.f;
                    }
                    o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
                }
                if (a.relu & 2) {   // ReLU AFTER the residual add: the closing relu(conv + x) of an eval-mode (BN-folded) block
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                *reinterpret_cast<float4*>(a.out + off) = o;
                csum[hh].x += o.x; csum[hh].y += o.y; csum[hh].z += o.z; csum[hh].w += o.w;
                if (affine_mask) {   // pre_mask holds: pre[hh][j] is the BatchNorm input x
                    const float4 mk = pre[hh][j];
                    csq[hh].x = fmaf(o.x, mk.x, csq[hh].x); csq[hh].y = fmaf(o.y, mk.y, csq[hh].y);
                    csq[hh].z = fmaf(o.z, mk.z, csq[hh].z); csq[hh].w = fmaf(o.w, mk.w, csq[hh].w);
                } else {
                    csq[hh].x = fmaf(o.x, o.x, csq[hh].x); csq[hh].y = fmaf(o.y, o.y, csq[hh].y);
                    csq[hh].z = fmaf(o.z, o.z, csq[hh].z); csq[hh].w = fmaf(o.w, o.w, csq[hh].w);
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // staging half-tile may be overwritten
        }
    }
    if (a.colsum_partial || a.stats_partial) {
        // 16 threads share each (half, float4 column): combine them through the (now free) staging tile
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            *reinterpret_cast<float4*>(&stg_g[(hh * 16 + r0) * TC_STG_LD + 4 * c4]) = csum[hh];
            *reinterpret_cast<float4*>(&stg_g[(32 + hh * 16 + r0) * TC_STG_LD + 4 * c4]) = csq[hh];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (et < 32 * NH) {
            const int hh = et >> 5, c = et & 31;
            const int ch = n_half * TC_BN + 32 * (h_first + hh) + c;
            double tot = 0.0, tsq = 0.0;
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                tot += (double)stg_g[(hh * 16 + g) * TC_STG_LD + c];
                tsq += (double)stg_g[(32 + hh * 16 + g) * TC_STG_LD + c];
            }
            if (a.colsum_partial) a.colsum_partial[(size_t)cta_m * a.Ctot + ch] = (float)tot;
            if (a.stats_partial) {
                double* sp = a.stats_partial + (size_t)cta_m * 2 * a.Ctot + ch;
                sp[0] = tot;
                sp[a.Ctot] = tsq;
            }
        }
    }
}

}  // namespace lf

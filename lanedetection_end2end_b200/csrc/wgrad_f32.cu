// Weight gradients as a split-K GEMM over pixels (fp32 CUDA cores, parity mode) plus the
// deterministic split reductions.  See LfWgradArgs in include/lanefit_b200.h.
// Replaces autograd's convolution_backward weight/bias paths for every Conv2d /
// ConvTranspose2d of BP/Networks/ERFNet.py (:15,29-37,101).
#include "lf_common.cuh"
#include "lf_net.h"

namespace lf {

constexpr int WG_TILE = 64;
constexpr int WG_PX = 16;
constexpr int WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS) wgrad_f32_kernel(const WgradArgs a) {
    __shared__ __align__(16) float Ps[2][WG_PX][WG_TILE];
    __shared__ __align__(16) float Qs[2][WG_PX][WG_TILE];
    const int tid = threadIdx.x;
    const int tilesQ = a.CqPad / WG_TILE, tilesP = a.CpPad / WG_TILE;
    const int tq = blockIdx.x % tilesQ;
    const int tp = (blockIdx.x / tilesQ) % tilesP;
    const int t = blockIdx.x / (tilesQ * tilesP);
    const int split = blockIdx.y;
    const long long M = (long long)a.N * a.Hs * a.Ws;
    long long per = (M + a.nsplit - 1) / a.nsplit;
    per = ((per + WG_PX - 1) / WG_PX) * WG_PX;
    const long long m_begin = (long long)split * per;
    const long long m_end = min(M, m_begin + per);

    const int lp = tid >> 4;   // pixel within the chunk
    const int c4 = tid & 15;   // float4 column
    const int tx = tid & 15, ty = tid >> 4;
    const int cp0 = tp * WG_TILE + 4 * c4;
    const int cq0 = tq * WG_TILE + 4 * c4;
    const bool p_ch_ok = cp0 < a.Cp, q_ch_ok = cq0 < a.Cq;
    const int pdy = a.pdy[t], pdx = a.pdx[t], qdy = a.qdy[t], qdx = a.qdx[t];
    const bool do_qsum = (a.qsum_partial != nullptr) && t == 0 && tp == 0 && ty == 0;

    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    float qs[4] = {0.f, 0.f, 0.f, 0.f};

    float4 rp, rq;
    auto load = [&](long long m0) {
        const long long m = m0 + lp;
        rp = make_float4(0.f, 0.f, 0.f, 0.f);
        rq = rp;
        if (m < m_end) {
            const int n = (int)(m / (a.Hs * a.Ws));
            const int rem = (int)(m - (long long)n * (a.Hs * a.Ws));
            const int j = rem / a.Ws;
            const int i = rem - j * a.Ws;
            const int py = j * a.psy + pdy, px = i * a.psx + pdx;
            const int qy = j * a.qsy + qdy, qx = i * a.qsx + qdx;
            const bool pin = py >= 0 && py < a.Hp && px >= 0 && px < a.Wp;
            const bool qin = qy >= 0 && qy < a.Hq && qx >= 0 && qx < a.Wq;
            // out-of-tensor operands are zero, so a product term exists only where both are inside
            if (pin && p_ch_ok)
                rp = __ldg(reinterpret_cast<const float4*>(
                    a.P + ((size_t)(n * a.Hp + py) * a.Wp + px) * a.p_cstride + a.p_coff + cp0));
            if (qin && q_ch_ok)
                rq = __ldg(reinterpret_cast<const float4*>(
                    a.Q + ((size_t)(n * a.Hq + qy) * a.Wq + qx) * a.q_cstride + a.q_coff + cq0));
        }
    };
    auto store = [&](int buf) {
        *reinterpret_cast<float4*>(&Ps[buf][lp][4 * c4]) = rp;
        *reinterpret_cast<float4*>(&Qs[buf][lp][4 * c4]) = rq;
    };

    if (m_begin < m_end) {
        load(m_begin);
        store(0);
    }
    __syncthreads();
    int it = 0;
    for (long long m0 = m_begin; m0 < m_end; m0 += WG_PX, ++it) {
        const int buf = it & 1;
        const bool more = (m0 + WG_PX) < m_end;
        if (more) load(m0 + WG_PX);
#pragma unroll
        for (int p = 0; p < WG_PX; ++p) {
            const float4 av = *reinterpret_cast<const float4*>(&Ps[buf][p][4 * ty]);
            const float4 bv = *reinterpret_cast<const float4*>(&Qs[buf][p][4 * tx]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(aa[r], bb[c], acc[r][c]);
            if (do_qsum) {
                qs[0] += bb[0]; qs[1] += bb[1]; qs[2] += bb[2]; qs[3] += bb[3];
            }
        }
        if (more) store(buf ^ 1);
        __syncthreads();
    }
    float* dst = a.partial + (((size_t)split * a.ntaps + t) * a.CpPad + tp * WG_TILE + 4 * ty) * a.CqPad +
                 tq * WG_TILE + 4 * tx;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(dst + (size_t)r * a.CqPad) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (do_qsum)
        *reinterpret_cast<float4*>(a.qsum_partial + (size_t)split * a.CqPad + tq * WG_TILE + 4 * tx) =
            make_float4(qs[0], qs[1], qs[2], qs[3]);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int ntaps, int Cp, int Cq, int CpPad,
                                    int CqPad, float* __restrict__ dst, int st, int sp, int sq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = ntaps * Cp * Cq;
    if (idx >= total) return;
    const int cq = idx % Cq;
    const int cp = (idx / Cq) % Cp;
    const int t = idx / (Cq * Cp);
    const size_t stride = (size_t)ntaps * CpPad * CqPad;
    const float* p = partial + ((size_t)t * CpPad + cp) * CqPad + cq;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += p[(size_t)k * stride];
    dst[(size_t)t * st + (size_t)cp * sp + (size_t)cq * sq] = s;
}

__global__ void vec_reduce_kernel(const float* __restrict__ partial, int nsplit, int C, int Cpad, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(size_t)k * Cpad + c];
    dst[c] = s;
}

constexpr int COLSUM_THREADS = 256;
constexpr int COLSUM_PIX_PER_BLOCK = 2048;

__global__ void __launch_bounds__(COLSUM_THREADS) colsum_kernel(const float* __restrict__ src, long long npix, int C,
                                                                 int cstride, int coff, float* __restrict__ partial, int Cpad) {
    // thread -> (pixel lane, channel); channels fastest so global reads are contiguous
    __shared__ float red[COLSUM_THREADS];
    const int tid = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * COLSUM_PIX_PER_BLOCK;
    const long long p1 = min(npix, p0 + COLSUM_PIX_PER_BLOCK);
    for (int cbase = 0; cbase < C; cbase += COLSUM_THREADS) {
        const int cw = min(C - cbase, COLSUM_THREADS);   // channels handled this pass
        const int lanes = COLSUM_THREADS / cw > 0 ? COLSUM_THREADS / cw : 1;
        const int c = tid % cw, pl = tid / cw;
        float s = 0.f;
        if (pl < lanes)
            for (long long p = p0 + pl; p < p1; p += lanes) s += src[(size_t)p * cstride + coff + cbase + c];
        red[tid] = (pl < lanes) ? s : 0.f;
        __syncthreads();
        if (tid < cw) {
            float tot = 0.f;
            for (int l = 0; l < lanes; ++l) tot += red[l * cw + tid];
            partial[(size_t)blockIdx.x * Cpad + cbase + tid] = tot;
        }
        __syncthreads();
    }
}

}  // namespace lf

using namespace lf;

extern "C" int lf_wgrad_f32(const LfWgradArgs* args, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!args) return LF_ERR_INVALID_ARGUMENT;
    const WgradArgs& a = *args;
    LF_REQUIRE(a.P && a.Q && a.partial);
    LF_REQUIRE(a.N > 0 && a.Hs > 0 && a.Ws > 0 && a.ntaps >= 1 && a.ntaps <= LF_MAX_TAPS && a.nsplit >= 1);
    LF_REQUIRE(a.Cp % 4 == 0 && a.Cq % 4 == 0 && a.p_cstride % 4 == 0 && a.q_cstride % 4 == 0);
    LF_REQUIRE(a.p_coff % 4 == 0 && a.q_coff % 4 == 0);
    LF_REQUIRE(a.CpPad % WG_TILE == 0 && a.CqPad % WG_TILE == 0 && a.CpPad >= a.Cp && a.CqPad >= a.Cq);
    dim3 grid((a.CpPad / WG_TILE) * (a.CqPad / WG_TILE) * a.ntaps, a.nsplit);
    wgrad_f32_kernel<<<grid, WG_THREADS, 0, stream>>>(a);
    return check_launch();
}

extern "C" int lf_wgrad_reduce(const float* partial, int nsplit, int ntaps, int Cp, int Cq, int CpPad, int CqPad,
                               float* dst, int st, int sp, int sq, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LF_REQUIRE(partial && dst && nsplit >= 1 && ntaps >= 1 && Cp >= 1 && Cq >= 1);
    const int total = ntaps * Cp * Cq;
    wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(partial, nsplit, ntaps, Cp, Cq, CpPad, CqPad, dst, st,
                                                                 sp, sq);
    return check_launch();
}

extern "C" int lf_vec_reduce(const float* partial, int nsplit, int C, int Cpad, float* dst, lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LF_REQUIRE(partial && dst && nsplit >= 1 && C >= 1 && Cpad >= C);
    vec_reduce_kernel<<<(C + 127) / 128, 128, 0, stream>>>(partial, nsplit, C, Cpad, dst);
    return check_launch();
}

extern "C" int lf_colsum_blocks(long long npix) {
    return (int)((npix + COLSUM_PIX_PER_BLOCK - 1) / COLSUM_PIX_PER_BLOCK);
}

extern "C" int lf_colsum(const float* src, long long npix, int C, int cstride, int coff, float* partial, int Cpad,
                         lf_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LF_REQUIRE(src && partial && npix > 0 && C >= 1 && Cpad >= C);
    colsum_kernel<<<lf_colsum_blocks(npix), COLSUM_THREADS, 0, stream>>>(src, npix, C, cstride, coff, partial, Cpad);
    return check_launch();
}

// Second-generation memory-bound kernels: GroupNorm backward, the Dice sums and (until elementwise4.cuh replaced
// them) trilinear x2.
//
// The first versions (elementwise.cuh) were instruction-bound, not HBM-bound (ncu, profiles/r01_*):
// per-element shared-memory reads of the per-channel constants and scalar FP32 math cost ~125
// instructions per 16-byte vector.  Here a warp owns ONE 8-channel chunk, so the per-channel
// constants live in registers, and the arithmetic is packed FP32x2 (FFMA2/FMUL2/FADD2, new on
// sm_100): ~60 instructions per vector, which puts the kernels back under the HBM roofline.
//
// Replaces aten::native_group_norm_backward + leaky_relu_backward (model.py:95-96, 105-112, 338).
#pragma once
#include "elementwise.cuh"

namespace b200 {

constexpr int kRedLines = 128;      // lines one gn_bwd_reduce2 CTA may own
// Register budget of the two GroupNorm-backward kernels.  In the training step they run beside a weight-gradient CTA
// (12 K registers, ~217 KB of shared memory per SM): at 128 registers per thread only ONE of their CTAs fits next to it
// (8 warps per SM for a memory-bound kernel), at 64 three do.  kGnBwdMinCtas / kGnBwdUnroll trade per-thread loads in
// flight for resident warps.  Measured on the training step (min CTAs, unroll): (2,4) 5.66 ms, (3,2) 5.68, (3,4) 5.87,
// (4,2) 5.95, (4,1) 6.02 - loads in flight per thread beat resident warps; (2,4) stays.  Round 2b tried the other
// settings for the SMALL tensors only (levels 1-3, < 48 MB, where one (2,4) CTA fits beside a weight-gradient CTA):
// (4,2) 5.47 ms and (3,2) 5.38 ms against 5.35 ms - the same answer.
#ifndef B200_GNBWD_MINCTAS
#define B200_GNBWD_MINCTAS 2
#endif
#ifndef B200_GNBWD_UNROLL
#define B200_GNBWD_UNROLL 4
#endif
constexpr int kGnBwdMinCtas = B200_GNBWD_MINCTAS;
constexpr int kGnBwdUnroll = B200_GNBWD_UNROLL;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
// bf16x2 word -> two floats (low half first)
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ void unpack4(const uint4& q, float2* f) {
    f[0] = bf2_to_f2(q.x); f[1] = bf2_to_f2(q.y); f[2] = bf2_to_f2(q.z); f[3] = bf2_to_f2(q.w);
}
__device__ __forceinline__ uint4 pack4(const float2* f) {
    uint4 q;
    q.x = pack_bf16x2(f[0].x, f[0].y); q.y = pack_bf16x2(f[1].x, f[1].y);
    q.z = pack_bf16x2(f[2].x, f[2].y); q.w = pack_bf16x2(f[3].x, f[3].y);
    return q;
}
__device__ __forceinline__ float2 lrelu_mask(float2 z) { return make_float2(z.x > 0.f ? 1.f : 0.01f, z.y > 0.f ? 1.f : 0.01f); }
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp -> chunk assignment of a 256-thread CTA: with nch = C/8 chunks, nb = min(nch, 8) chunks are
// processed per pass, 8/nb warps share one chunk (splitting its voxel list), nch/8 passes.
struct WarpChunk {
    int nb, wpc, npass, cvb, ws;
};
__device__ __forceinline__ WarpChunk warp_chunk(int C) {
    WarpChunk m;
    const int nch = C >> 3, w = threadIdx.x >> 5;
    m.nb = nch < 8 ? nch : 8;
    m.wpc = 8 / m.nb;
    m.npass = (nch + 7) >> 3;
    m.cvb = w % m.nb;
    m.ws = w / m.nb;
    return m;
}

// ---------------------------------------------------------------------------------------
// GroupNorm (+LeakyReLU) backward, pass 1: per (n, c) sums  S1 = sum dz, S2 = sum dz * xhat
//   dz = dy * lrelu'(z), z = xhat*gamma + beta, xhat = (x - mean) * rstd.
// grid = (blocks, N); CTA b owns lines [b*lpb, b*lpb+lpb) of sample n; partial[n][blocks][C][2].
// ---------------------------------------------------------------------------------------
// fin.coef != nullptr: the last CTA of each sample also does pass 2 (below) for that sample, and the last of those the
// sum over samples for dgamma / dbeta - no finalize launch between the reduction and the apply kernel.
struct GnBwdFin {
    float* coef;              // [N][C][2]; nullptr: pass 2 is a separate launch (gn_bwd_finalize2_kernel)
    double* tot;              // [N][C][2] scratch: per-sample (S1, S2) in double, for the sum over samples
    float* dgamma;            // [C]
    float* dbeta;             // [C]
    unsigned int* tickets;    // [N + 1] zero-initialised words (one per sample + one for the samples), left zero
    double m;                 // elements per (sample, group)
};

__device__ __forceinline__ void gn_bwd_finalize_sample(const float* __restrict__ partial, int blocks, int n, int N, int C,
                                                       const float* __restrict__ gamma, const GnBwdFin& f,
                                                       unsigned int* s_ticket);

__global__ void __launch_bounds__(256, kGnBwdMinCtas)
gn_bwd_reduce2_kernel(ActRef x, ActRef dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ partial,
                      Vol v, int C, int do_lrelu, FastDiv by_W, int lpb, GnBwdFin fin) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ long long s_rows[kRedLines];
    __shared__ float s_red[8][16];
    const int n = blockIdx.y;
    const int nlines = v.D * v.H;
    const int line0 = blockIdx.x * lpb;
    const int nl = min(lpb, nlines - line0);
    for (int li = threadIdx.x; li < nl; li += blockDim.x) {
        const int line = line0 + li;
        const int d = line / v.H, h = line - d * v.H;
        s_rows[li] = v.row(n, d + 1, h + 1, 1);
    }
    __syncthreads();
    const WarpChunk wc = warp_chunk(C);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gs = C >> 3;
    const int total = nl * v.W;
    const int stride = wc.wpc * 32;
    for (int pass = 0; pass < wc.npass; ++pass) {
        const int cv = wc.cvb + 8 * pass;
        float2 a2[4], b2[4], p1[4], p2[4], s1[4], s2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ka[2], kb[2], k1[2], k2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cv * 8 + 2 * j + e;
                const int g = c / gs;
                const float r = rstd[n * 8 + g], m = mean[n * 8 + g];
                ka[e] = r; kb[e] = -m * r;
                k1[e] = r * gamma[c]; k2[e] = -m * r * gamma[c] + beta[c];
            }
            a2[j] = f2(ka[0], ka[1]); b2[j] = f2(kb[0], kb[1]);
            p1[j] = f2(k1[0], k1[1]); p2[j] = f2(k2[0], k2[1]);
            s1[j] = f2(0.f, 0.f); s2[j] = f2(0.f, 0.f);
        }
        constexpr int U = kGnBwdUnroll;
        for (int i0 = wc.ws * 32 + lane; i0 < total; i0 += stride * U) {
            uint4 qx[U], qd[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = i0 + u * stride;
                ok[u] = idx < total;
                if (ok[u]) {
                    const int li = by_W.div(idx);
                    const long long r = s_rows[li] + (idx - li * v.W);
                    qx[u] = ld16(x.at(cv, r));
                    qd[u] = ld16(dy.at(cv, r));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                float2 fx[4], fd[4];
                unpack4(qx[u], fx);
                unpack4(qd[u], fd);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 xh = __ffma2_rn(fx[j], a2[j], b2[j]);
                    float2 dz = fd[j];
                    if (do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
                    s1[j] = __fadd2_rn(s1[j], dz);
                    s2[j] = __ffma2_rn(dz, xh, s2[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s1[j].x = warp_sum(s1[j].x); s1[j].y = warp_sum(s1[j].y);
            s2[j].x = warp_sum(s2[j].x); s2[j].y = warp_sum(s2[j].y);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_red[warp][2 * j] = s1[j].x; s_red[warp][2 * j + 1] = s1[j].y;
                s_red[warp][8 + 2 * j] = s2[j].x; s_red[warp][8 + 2 * j + 1] = s2[j].y;
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < wc.nb * 16) {
            const int cb = threadIdx.x >> 4, k = threadIdx.x & 15;
            float acc = 0.f;
            for (int q = 0; q < wc.wpc; ++q) acc += s_red[cb + q * wc.nb][k];       // fixed order
            const int c = (cb + 8 * pass) * 8 + (k & 7);
            partial[(((size_t)n * gridDim.x + blockIdx.x) * C + c) * 2 + (k >> 3)] = acc;
        }
        __syncthreads();
    }
    __shared__ unsigned int s_ticket;
    if (fin.coef != nullptr && cta_draws_last_ticket(fin.tickets + n, gridDim.x, &s_ticket))
        gn_bwd_finalize_sample(partial, (int)gridDim.x, n, v.N, C, gamma, fin, &s_ticket);
}

// Pass 2 inside the reduction kernel: ONE CTA (the last of sample n to finish) sums that sample's partial[blocks][C][2]
// - a [blocks][2C] matrix whose rows are contiguous: thread = (column, row group), coalesced ld.cg loads, double
// accumulation, row groups combined in a fixed order (deterministic) -, writes coef[n] and the per-sample sums; the
// last sample to get there adds the samples up for dgamma / dbeta.
__device__ __forceinline__ void gn_bwd_finalize_sample(const float* __restrict__ partial, int blocks, int n, int N, int C,
                                                       const float* __restrict__ gamma, const GnBwdFin& f,
                                                       unsigned int* s_ticket) {
    __shared__ double s_acc[256];
    __shared__ double s_col[256];       // [c * 2 + which], C <= 128
    const int t = threadIdx.x;
    const int cols = 2 * C;                                 // 32 .. 256
    const int R = 256 / cols;                               // row groups
    const int col = t % cols, rg = t / cols;
    {
        const float* src = partial + (size_t)n * blocks * cols + col;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int b = rg;
        for (; b + 3 * R < blocks; b += 4 * R) {
            const float v0 = __ldcg(src + (size_t)b * cols), v1 = __ldcg(src + (size_t)(b + R) * cols);
            const float v2 = __ldcg(src + (size_t)(b + 2 * R) * cols), v3 = __ldcg(src + (size_t)(b + 3 * R) * cols);
            a0 += (double)v0; a1 += (double)v1; a2 += (double)v2; a3 += (double)v3;
        }
        for (; b < blocks; b += R) a0 += (double)__ldcg(src + (size_t)b * cols);
        s_acc[t] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (t < cols) {
        double a = 0.0;
        for (int q = 0; q < R; ++q) a += s_acc[q * cols + t];
        s_col[t] = a;
        f.tot[(size_t)n * cols + t] = a;
    }
    __syncthreads();
    const int gs = C >> 3;
    if (t < C) {
        const int g = t / gs;
        double A = 0.0, B = 0.0;
        for (int k = 0; k < gs; ++k) {
            const double gm = (double)gamma[g * gs + k];
            A += s_col[2 * (g * gs + k)] * gm;
            B += s_col[2 * (g * gs + k) + 1] * gm;
        }
        f.coef[((size_t)n * C + t) * 2 + 0] = (float)(A / f.m);
        f.coef[((size_t)n * C + t) * 2 + 1] = (float)(B / f.m);
    }
    if (t == 0) f.tickets[n] = 0u;
    if (cta_draws_last_ticket(f.tickets + N, (unsigned)N, s_ticket)) {
        if (t < C) {
            double s1 = 0.0, s2 = 0.0;
            for (int q = 0; q < N; ++q) {                    // fixed order over the samples
                s1 += __ldcg(f.tot + ((size_t)q * C + t) * 2);
                s2 += __ldcg(f.tot + ((size_t)q * C + t) * 2 + 1);
            }
            f.dbeta[t] = (float)s1;
            f.dgamma[t] = (float)s2;
        }
        if (t == 0) f.tickets[N] = 0u;
    }
}

// pass 2: partial[n][blocks][C][2] -> coef[n][C][2] = (A_g, B_g)/m of the channel's group, and
// dgamma[c] = sum_n S2, dbeta[c] = sum_n S1.   grid = 8 (one CTA per group), 256-1024 threads.
// Warp o sums one (channel, which) over the blocks (lanes stride the blocks, double accumulation,
// fixed shuffle tree): deterministic.
__global__ void __launch_bounds__(1024)
gn_bwd_finalize2_kernel(const float* __restrict__ partial, int blocks, int N, int C, double m,
                        const float* __restrict__ gamma, float* __restrict__ coef, float* __restrict__ dgamma,
                        float* __restrict__ dbeta) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ double s_sum[8][64];     // [n % 8][k*2 + which], k < gs <= 32
    __shared__ double s_tot[64];
    const int g = blockIdx.x, gs = C >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nout = gs * 2;
    const int nwarps = blockDim.x >> 5;     // the (sample, output) pairs are spread over all warps: the kernel is a chain
                                            // of dependent-load latencies, so its time is the number of rounds per warp
    if ((int)threadIdx.x < nout) s_tot[threadIdx.x] = 0.0;
    for (int n0 = 0; n0 < N; n0 += 8) {
        const int nn = min(8, N - n0);
        // (sample, output) pairs over the 8 warps; 4 independent loads in flight per lane
        for (int po = warp; po < nn * nout; po += nwarps) {
            const int nl = po / nout, o = po - nl * nout;
            const int c = g * gs + (o >> 1);
            const float* src = partial + ((size_t)(n0 + nl) * blocks * C + c) * 2 + (o & 1);
            const size_t bs = (size_t)C * 2;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int b = lane;
            for (; b + 96 < blocks; b += 128) {
                const float v0 = src[(size_t)b * bs], v1 = src[(size_t)(b + 32) * bs];
                const float v2 = src[(size_t)(b + 64) * bs], v3 = src[(size_t)(b + 96) * bs];
                a0 += (double)v0; a1 += (double)v1; a2 += (double)v2; a3 += (double)v3;
            }
            for (; b < blocks; b += 32) a0 += (double)src[(size_t)b * bs];
            const double a = warp_sum_d((a0 + a1) + (a2 + a3));
            if (lane == 0) s_sum[nl][o] = a;
        }
        __syncthreads();
        if ((int)threadIdx.x < nout)
            for (int nl = 0; nl < nn; ++nl) s_tot[threadIdx.x] += s_sum[nl][threadIdx.x];
        for (int pc = threadIdx.x; pc < nn * gs; pc += blockDim.x) {
            const int nl = pc / gs, k0 = pc - nl * gs;
            double A = 0.0, B = 0.0;
            for (int k = 0; k < gs; ++k) {
                A += s_sum[nl][2 * k] * (double)gamma[g * gs + k];
                B += s_sum[nl][2 * k + 1] * (double)gamma[g * gs + k];
            }
            const int c = g * gs + k0;
            coef[((size_t)(n0 + nl) * C + c) * 2 + 0] = (float)(A / m);
            coef[((size_t)(n0 + nl) * C + c) * 2 + 1] = (float)(B / m);
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < gs) {
        dbeta[g * gs + threadIdx.x] = (float)s_tot[2 * threadIdx.x];
        dgamma[g * gs + threadIdx.x] = (float)s_tot[2 * threadIdx.x + 1];
    }
}

// pass 3: dx = rstd * (dz*gamma - A_g - xhat*B_g) = dz*p1 + x*c1 + c0 with per-channel constants
//   p1 = rstd*gamma, c1 = -rstd^2 B, c0 = -rstd (A + b B), b = -mean*rstd;  z = x*p1 + (b*gamma + beta).
// aff_partial (optional): this CTA's per-channel sums S1 = sum dz, S2 = sum dz * xhat over its voxels, layout
// [n][blocks per sample][C][2] like gn_bwd_reduce2's - the input of gn_bwd_finalize2 for dgamma / dbeta when the
// group sums that dx needs came from the producing conv's epilogue (common.cuh gnb_accumulate) and no reduce pass ran.
template <bool AFF>
__global__ void __launch_bounds__(256, AFF ? 2 : kGnBwdMinCtas)
gn_bwd_apply2_kernel(ActRef x, ActRef dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ coef,
                     ActRef dx, Vol v, int C, int do_lrelu, FastDiv by_W, int lpb, float* __restrict__ aff_partial) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ long long s_rows[16];
    __shared__ float s_red[8][16];
    const int line0 = blockIdx.x * lpb;
    fill_line_rows(v, line0, lpb, s_rows);
    __syncthreads();
    const int n = line0 / (v.D * v.H);
    const WarpChunk wc = warp_chunk(C);
    const int lane = threadIdx.x & 31;
    const int gs = C >> 3;
    const int total = lpb * v.W;
    const int stride = wc.wpc * 32;
    for (int pass = 0; pass < wc.npass; ++pass) {
        const int cv = wc.cvb + 8 * pass;
        float2 p1[4], p2[4], c1[4], c0[4], a2[4], b2[4], s1[4], s2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float k1[2], k2[2], k3[2], k4[2], ka[2], kb[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cv * 8 + 2 * j + e;
                const int g = c / gs;
                const float r = rstd[n * 8 + g], b = -mean[n * 8 + g] * r;
                const float A = coef[((size_t)n * C + c) * 2 + 0], B = coef[((size_t)n * C + c) * 2 + 1];
                k1[e] = r * gamma[c]; k2[e] = b * gamma[c] + beta[c];
                k3[e] = -r * r * B; k4[e] = -r * (A + b * B);
                ka[e] = r; kb[e] = b;
            }
            p1[j] = f2(k1[0], k1[1]); p2[j] = f2(k2[0], k2[1]);
            c1[j] = f2(k3[0], k3[1]); c0[j] = f2(k4[0], k4[1]);
            a2[j] = f2(ka[0], ka[1]); b2[j] = f2(kb[0], kb[1]);
            s1[j] = f2(0.f, 0.f); s2[j] = f2(0.f, 0.f);
        }
        constexpr int U = kGnBwdUnroll;
        for (int i0 = wc.ws * 32 + lane; i0 < total; i0 += stride * U) {
            uint4 qx[U], qd[U];
            long long rr[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = i0 + u * stride;
                ok[u] = idx < total;
                if (ok[u]) {
                    const int li = by_W.div(idx);
                    rr[u] = s_rows[li] + (idx - li * v.W);
                    qx[u] = ld16(x.at(cv, rr[u]));
                    qd[u] = ld16(dy.at(cv, rr[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                float2 fx[4], fd[4];
                unpack4(qx[u], fx);
                unpack4(qd[u], fd);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 dz = fd[j];
                    if (do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
                    if (AFF) {
                        s1[j] = __fadd2_rn(s1[j], dz);
                        s2[j] = __ffma2_rn(dz, __ffma2_rn(fx[j], a2[j], b2[j]), s2[j]);
                    }
                    fx[j] = __ffma2_rn(dz, p1[j], __ffma2_rn(fx[j], c1[j], c0[j]));
                }
                st16(dx.at(cv, rr[u]), pack4(fx));
            }
        }
        if (AFF) {          // same fixed-order combination as gn_bwd_reduce2_kernel
            const int warp = threadIdx.x >> 5;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s1[j].x = warp_sum(s1[j].x); s1[j].y = warp_sum(s1[j].y);
                s2[j].x = warp_sum(s2[j].x); s2[j].y = warp_sum(s2[j].y);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s_red[warp][2 * j] = s1[j].x; s_red[warp][2 * j + 1] = s1[j].y;
                    s_red[warp][8 + 2 * j] = s2[j].x; s_red[warp][8 + 2 * j + 1] = s2[j].y;
                }
            }
            __syncthreads();
            if ((int)threadIdx.x < wc.nb * 16) {
                const int cb = threadIdx.x >> 4, k = threadIdx.x & 15;
                float acc = 0.f;
                for (int q = 0; q < wc.wpc; ++q) acc += s_red[cb + q * wc.nb][k];       // fixed order
                const int c = (cb + 8 * pass) * 8 + (k & 7);
                const int bps = v.D * v.H / lpb;                                         // CTAs per sample
                const int b = blockIdx.x - n * bps;
                aff_partial[(((size_t)n * bps + b) * C + c) * 2 + (k >> 3)] = acc;
            }
            __syncthreads();
        }
    }
}

// Group sums from the producing conv's epilogue (common.cuh gnb_accumulate: per CTA a_g = sum gamma dz, u_g = sum gamma dz c)
// -> coef[n][C][2] = (A_g, B_g) / m of the channel's group, the two constants gn_bwd_apply2 needs:
//   A = sum_g gamma S1 = sum a,   B = sum_g gamma S2 = rstd * sum u + (-mean * rstd) * sum a      (xhat = c * rstd - mean * rstd)
// grid = N, block = 256: each of the 16 sums is split over 16 threads (strided over CTAs, fixed order) - deterministic.
__global__ void gn_bwd_fold_finalize_kernel(const float* __restrict__ gpart, int ctas, int N, int C, double m,
                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                            float* __restrict__ coef) {
    __shared__ double s_part[256];
    __shared__ double s_sum[16];
    const int n = blockIdx.x, t = threadIdx.x;
    const int o = t >> 4, j = t & 15;
    double a = 0.0;
    for (int c = j; c < ctas; c += 16) a += (double)gpart[((size_t)c * N + n) * 16 + o];
    s_part[t] = a;
    __syncthreads();
    if (t < 16) {
        double acc = 0.0;
        for (int q = 0; q < 16; ++q) acc += s_part[t * 16 + q];
        s_sum[t] = acc;
    }
    __syncthreads();
    const int gs = C >> 3;
    for (int c = t; c < C; c += blockDim.x) {
        const int g = c / gs;
        const double r = (double)rstd[n * 8 + g], b = -(double)mean[n * 8 + g] * r;
        const double A = s_sum[g], U = s_sum[8 + g];
        coef[((size_t)n * C + c) * 2 + 0] = (float)(A / m);
        coef[((size_t)n * C + c) * 2 + 1] = (float)((r * U + b * A) / m);
    }
}


// ---------------------------------------------------------------------------------------
// sigmoid backward + layout (model.py:431 backward): a CTA owns `lpb` lines; thread = voxel.
//   dlogit = gp * p * (1-p) -> chunk 0 of an act tensor; per-CTA bias-grad partials [blocks][4]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sigmoid_bwd_pack2_kernel(const float* __restrict__ gp, const float* __restrict__ probs, ActRef dlogit,
                         float* __restrict__ bias_partial, Vol v, int Creal, int lpb, FastDiv by_W) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ float s_red[8][4];
    const int line0 = blockIdx.x * lpb;
    const int nlines = v.N * v.D * v.H;
    const int nl = min(lpb, nlines - line0);
    const size_t plane = (size_t)v.D * v.H * v.W;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int idx = threadIdx.x; idx < nl * v.W; idx += blockDim.x) {
        const int li = by_W.div(idx), w = idx - li * v.W;
        int n, d, h;
        line_coords(v, line0 + li, n, d, h);
        const size_t src = (size_t)n * Creal * plane + ((size_t)d * v.H + h) * v.W + w;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            f[i] = 0.f;
            if (i < Creal) {
                const float p = probs[src + (size_t)i * plane];
                f[i] = gp[src + (size_t)i * plane] * p * (1.f - p);
                if (i < 4) bsum[i] += f[i];
            }
        }
        st16(dlogit.at(0, v.row(n, d + 1, h + 1, 1 + w)), pack_bf16x8(f));
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) bsum[c] = warp_sum(bsum[c]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int c = 0; c < 4; ++c) s_red[warp][c] = bsum[c];
    __syncthreads();
    if (threadIdx.x < 4) {
        float a = 0.f;
        for (int k = 0; k < 8; ++k) a += s_red[k][threadIdx.x];
        bias_partial[(size_t)blockIdx.x * 4 + threadIdx.x] = a;
    }
}

// partial[B*C][blocks][2] -> sums[8] (I_c at [c], U_c at [4+c]): one CTA, thread (c, which, j) strides
// the (sample, block) pairs, fixed-order combination through shared memory.
__global__ void __launch_bounds__(256)
dice_sums2_kernel(const float* __restrict__ partial, int B, int C, int blocks, float* __restrict__ sums) {
    __shared__ double s_p[256];
    const int o = threadIdx.x >> 5, j = threadIdx.x & 31;      // o = c*2 + which (< 8), 32 threads per output
    const int c = o >> 1, which = o & 1;
    double a = 0.0;
    if (c < C)
        for (int i = j; i < B * blocks; i += 32) {
            const int b = i / blocks, k = i - b * blocks;
            a += (double)partial[((size_t)(b * C + c) * blocks + k) * 2 + which];
        }
    s_p[threadIdx.x] = a;
    __syncthreads();
    if (j == 0 && c < C) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += s_p[o * 32 + q];
        sums[which * 4 + c] = (float)t;
    }
}

}  // namespace b200

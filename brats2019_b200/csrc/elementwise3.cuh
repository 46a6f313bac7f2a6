// GroupNorm (+LeakyReLU) backward in ONE launch for the deep, small levels of the U-Net
// (aten::native_group_norm_backward + leaky_relu_backward; model.py:95-96, 105-112).
//
// The three-kernel form of elementwise2.cuh (reduce -> finalize -> apply) reads x and dy twice and, at the
// 32^3 and 16^3 levels (2-8 MB tensors), spends most of its ~25 us in three launch latencies.  Here a
// thread-block cluster of 8 CTAs owns one "unit" - the chunks that hold whole groups: one group of C/8 >= 8
// channels, or the one chunk that holds 8/(C/8) small groups - keeps its share of x and dy in shared memory
// (read from HBM/L2 once), exchanges the per-channel sums over distributed shared memory behind one hardware
// cluster barrier per sample, and applies the result straight from shared memory.  Samples are walked in
// order by the same cluster, so dgamma/dbeta need no cross-CTA atomics and every sum has a fixed order
// (deterministic).  Used when the unit's slice of one sample fits the cluster's shared memory; larger tensors
// take the three-kernel path.
#pragma once
#include <cooperative_groups.h>

#include "elementwise2.cuh"

namespace b200 {

constexpr int kGnClusterSize = 8;
constexpr int kGnClusterThreads = 256;
constexpr int kGnClusterMaxLines = 256;     // lines of one sample a CTA may own
constexpr int kGnClusterMaxUcs = 4;         // chunks per unit (C <= 256)

struct GnClusterParams {
    ActRef x, dy, dx;
    const float *mean, *rstd, *gamma, *beta;
    float *dgamma, *dbeta;
    Vol v;
    int C, do_lrelu;
    FastDiv by_W;
    int lpc;            // lines per CTA (ceil(D*H / cluster size))
    int nv;             // voxel capacity of a CTA's shared-memory tile: lpc * W
    int ucs;            // chunks per unit
    double m;           // elements per (sample, group)
    // register-resident form only (gn_bwd_creg_kernel): CTAs per unit, and the global-memory exchange between them
    int csize;
    double* xchg;             // [units][csize][2][64] per-CTA partial sums, double-buffered by sample
    unsigned int* counters;   // [units][2] arrive / exit counters: zero before the first launch, left zero
};

__global__ void __launch_bounds__(kGnClusterThreads, 1)
gn_bwd_cluster_kernel(const __grid_constant__ GnClusterParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) uint8_t gsm[];
    uint4* sx = reinterpret_cast<uint4*>(gsm);                       // [ucs][nv]
    uint4* sd = sx + (size_t)p.ucs * p.nv;                           // [ucs][nv]
    __shared__ long long s_rows[kGnClusterMaxLines];
    __shared__ float s_red[kGnClusterThreads / 32][16];
    __shared__ double s_part[2][kGnClusterMaxUcs * 16];              // this CTA's sums, double-buffered by sample
    __shared__ double s_tot[kGnClusterMaxUcs * 16];                  // cluster totals: [chunk][which*8 + ch]
    __shared__ float s_coef[kGnClusterMaxUcs * 8][4];                // p1, p2, c1, c0 per channel of the unit

    const Vol& v = p.v;
    const int rank = (int)cluster.block_rank();
    const int unit = blockIdx.x / kGnClusterSize;
    const int cv0 = unit * p.ucs;                                    // first chunk of the unit
    const int lines = v.D * v.H;
    const int line0 = rank * p.lpc;
    const int nl = max(0, min(p.lpc, lines - line0));
    const int total = nl * v.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gs = p.C >> 3;
    const int nch_unit = p.ucs * 8;

    for (int li = threadIdx.x; li < nl; li += blockDim.x) {
        const int line = line0 + li;
        const int d = line / v.H, h = line - d * v.H;
        s_rows[li] = v.row(0, d + 1, h + 1, 1);
    }
    double acc_db = 0.0, acc_dg = 0.0;      // thread k < nch_unit of rank 0: dbeta / dgamma of channel k
    __syncthreads();

    for (int n = 0; n < v.N; ++n) {
        const long long nrow = (long long)n * v.sample_rows();
        const int buf = n & 1;
        // ---------------- phase 1: load the tile, per-channel sums S1 = sum dz, S2 = sum dz*xhat ----------------
        for (int c = 0; c < p.ucs; ++c) {
            const int cv = cv0 + c;
            float2 a2[4], b2[4], p1[4], p2[4], s1[4], s2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float ka[2], kb[2], k1[2], k2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ch = cv * 8 + 2 * j + e;
                    const int g = ch / gs;
                    const float r = p.rstd[n * 8 + g], mu = p.mean[n * 8 + g];
                    ka[e] = r; kb[e] = -mu * r;
                    k1[e] = r * p.gamma[ch]; k2[e] = -mu * r * p.gamma[ch] + p.beta[ch];
                }
                a2[j] = f2(ka[0], ka[1]); b2[j] = f2(kb[0], kb[1]);
                p1[j] = f2(k1[0], k1[1]); p2[j] = f2(k2[0], k2[1]);
                s1[j] = f2(0.f, 0.f); s2[j] = f2(0.f, 0.f);
            }
            uint4* tx = sx + (size_t)c * p.nv;
            uint4* td = sd + (size_t)c * p.nv;
            constexpr int U = 4;
            for (int i0 = threadIdx.x; i0 < total; i0 += kGnClusterThreads * U) {
                uint4 qx[U], qd[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = i0 + u * kGnClusterThreads;
                    ok[u] = idx < total;
                    if (ok[u]) {
                        const int li = p.by_W.div(idx);
                        const long long r = nrow + s_rows[li] + (idx - li * v.W);
                        qx[u] = ld16(p.x.at(cv, r));
                        qd[u] = ld16(p.dy.at(cv, r));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!ok[u]) continue;
                    const int idx = i0 + u * kGnClusterThreads;
                    tx[idx] = qx[u];
                    td[idx] = qd[u];
                    float2 fx[4], fd[4];
                    unpack4(qx[u], fx);
                    unpack4(qd[u], fd);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 xh = __ffma2_rn(fx[j], a2[j], b2[j]);
                        float2 dz = fd[j];
                        if (p.do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
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
            if (threadIdx.x < 16) {
                double a = 0.0;
                for (int w = 0; w < kGnClusterThreads / 32; ++w) a += (double)s_red[w][threadIdx.x];     // fixed order
                s_part[buf][c * 16 + threadIdx.x] = a;
            }
            __syncthreads();
        }
        // ---------------- phase 2: cluster totals over distributed shared memory ----------------
        cluster.sync();          // every CTA's s_part[buf] is written (and its reads of sample n-2 are long done)
        if ((int)threadIdx.x < p.ucs * 16) {
            double a = 0.0;
            for (int r = 0; r < kGnClusterSize; ++r) {
                const double* remote = cluster.map_shared_rank(&s_part[buf][0], r);
                a += remote[threadIdx.x];                                           // fixed rank order
            }
            s_tot[threadIdx.x] = a;
        }
        __syncthreads();
        if ((int)threadIdx.x < nch_unit) {
            const int k = threadIdx.x;                      // channel of the unit
            const int ch = cv0 * 8 + k;
            const int g = ch / gs;
            const int k_first = g * gs - cv0 * 8;           // first channel of the group, unit-relative
            double A = 0.0, B = 0.0;
            for (int q = 0; q < gs; ++q) {
                const int kk = k_first + q;
                const double gam = (double)p.gamma[cv0 * 8 + kk];
                A += s_tot[(kk >> 3) * 16 + (kk & 7)] * gam;
                B += s_tot[(kk >> 3) * 16 + 8 + (kk & 7)] * gam;
            }
            const float Af = (float)(A / p.m), Bf = (float)(B / p.m);
            const float r = p.rstd[n * 8 + g], b = -p.mean[n * 8 + g] * r;
            s_coef[k][0] = r * p.gamma[ch];
            s_coef[k][1] = b * p.gamma[ch] + p.beta[ch];
            s_coef[k][2] = -r * r * Bf;
            s_coef[k][3] = -r * (Af + b * Bf);
            acc_db += s_tot[(k >> 3) * 16 + (k & 7)];
            acc_dg += s_tot[(k >> 3) * 16 + 8 + (k & 7)];
        }
        __syncthreads();
        // ---------------- phase 3: dx = dz*p1 + x*c1 + c0 from the shared-memory tile ----------------
        for (int c = 0; c < p.ucs; ++c) {
            const int cv = cv0 + c;
            float2 p1[4], p2[4], c1[4], c0[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* e0 = s_coef[c * 8 + 2 * j];
                const float* e1 = s_coef[c * 8 + 2 * j + 1];
                p1[j] = f2(e0[0], e1[0]); p2[j] = f2(e0[1], e1[1]);
                c1[j] = f2(e0[2], e1[2]); c0[j] = f2(e0[3], e1[3]);
            }
            const uint4* tx = sx + (size_t)c * p.nv;
            const uint4* td = sd + (size_t)c * p.nv;
            for (int idx = threadIdx.x; idx < total; idx += kGnClusterThreads) {
                const int li = p.by_W.div(idx);
                const long long r = nrow + s_rows[li] + (idx - li * v.W);
                float2 fx[4], fd[4];
                unpack4(tx[idx], fx);
                unpack4(td[idx], fd);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 dz = fd[j];
                    if (p.do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
                    fx[j] = __ffma2_rn(dz, p1[j], __ffma2_rn(fx[j], c1[j], c0[j]));
                }
                st16(p.dx.at(cv, r), pack4(fx));
            }
        }
        __syncthreads();         // the tile and s_coef are rewritten by the next sample
    }
    if (rank == 0 && (int)threadIdx.x < nch_unit) {
        p.dbeta[cv0 * 8 + threadIdx.x] = (float)acc_db;
        p.dgamma[cv0 * 8 + threadIdx.x] = (float)acc_dg;
    }
    cluster.sync();              // no CTA may exit while a peer can still read its shared memory
}

// ---------------------------------------------------------------------------------------
// Register-resident one-launch form (round 2).  The CTA's share of x and dy stays in REGISTERS between the reduction
// and the apply pass (UCS chunks x VPT 16-byte vectors per thread and tensor, UCS * VPT <= 8), and the `csize` CTAs of a
// unit exchange their partial sums through global memory behind a software barrier (arrive counter + spin) instead of a
// hardware cluster barrier.  Both choices are about co-scheduling, not about the kernel's own speed: inside the training
// step a weight-gradient GEMM of the side stream sits on every SM; it leaves ~20 KB of shared memory and ~50 K registers,
// and - measured, profiles/r02_ab_gn_creg.txt - the block scheduler does not place a thread-block CLUSTER beside it at
// all (the cluster launch waited for the GEMM to drain, shared-memory tiles or not), while plain CTAs slip in at once.
// Plain CTAs + spin barrier need every CTA of a unit resident at the same time: grid <= SM count (planner), the kernels
// already on the machine never wait on this one, and a watchdog turns a protocol bug into a trap instead of a hang.
// Same arithmetic and the same fixed summation order as the cluster kernel (thread -> warp shuffle -> warps in order ->
// ranks in order -> samples in order): deterministic.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int UCS, int VPT>
__global__ void __maxnreg__(192)        // 256 threads x 192 registers fit beside a weight-gradient CTA (256 x 56)
gn_bwd_creg_kernel(const __grid_constant__ GnClusterParams p) {
    __shared__ long long s_rows[kGnClusterMaxLines];
    __shared__ float s_red[kGnClusterThreads / 32][16];
    __shared__ double s_tot[kGnClusterMaxUcs * 16];
    __shared__ float s_coef[kGnClusterMaxUcs * 8][4];

    const Vol& v = p.v;
    const int csize = p.csize;
    const int unit = blockIdx.x / csize;
    const int rank = blockIdx.x - unit * csize;
    const int cv0 = unit * UCS;
    const int lines = v.D * v.H;
    const int line0 = rank * p.lpc;
    const int nl = max(0, min(p.lpc, lines - line0));
    const int total = nl * v.W;                                      // <= kGnClusterThreads * VPT (planner)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gs = p.C >> 3;
    constexpr int nch_unit = UCS * 8;
    double* my_xchg = p.xchg + ((size_t)unit * csize + rank) * 128;          // [2][64]
    unsigned int* arrive = p.counters + 2 * unit;

    for (int li = threadIdx.x; li < nl; li += blockDim.x) {
        const int line = line0 + li;
        const int d = line / v.H, h = line - d * v.H;
        s_rows[li] = v.row(0, d + 1, h + 1, 1);
    }
    double acc_db = 0.0, acc_dg = 0.0;
    __syncthreads();
    // this thread's voxels: idx = threadIdx.x + u * 256 (the same for every chunk and sample)
    long long roff[VPT];
    bool ok[VPT];
#pragma unroll
    for (int u = 0; u < VPT; ++u) {
        const int idx = threadIdx.x + u * kGnClusterThreads;
        ok[u] = idx < total;
        roff[u] = 0;
        if (ok[u]) {
            const int li = p.by_W.div(idx);
            roff[u] = s_rows[li] + (idx - li * v.W);
        }
    }

    for (int n = 0; n < v.N; ++n) {
        const long long nrow = (long long)n * v.sample_rows();
        const int buf = n & 1;
        uint4 rx[UCS][VPT], rd[UCS][VPT];
        // ---------------- phase 1: every load of the tile in flight at once, then the per-channel sums ----------------
#pragma unroll
        for (int c = 0; c < UCS; ++c)
#pragma unroll
            for (int u = 0; u < VPT; ++u)
                if (ok[u]) {
                    rx[c][u] = ld16(p.x.at(cv0 + c, nrow + roff[u]));
                    rd[c][u] = ld16(p.dy.at(cv0 + c, nrow + roff[u]));
                }
#pragma unroll
        for (int c = 0; c < UCS; ++c) {
            const int cv = cv0 + c;
            float2 a2[4], b2[4], p1[4], p2[4], s1[4], s2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float ka[2], kb[2], k1[2], k2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ch = cv * 8 + 2 * j + e;
                    const int g = ch / gs;
                    const float r = p.rstd[n * 8 + g], mu = p.mean[n * 8 + g];
                    ka[e] = r; kb[e] = -mu * r;
                    k1[e] = r * p.gamma[ch]; k2[e] = -mu * r * p.gamma[ch] + p.beta[ch];
                }
                a2[j] = f2(ka[0], ka[1]); b2[j] = f2(kb[0], kb[1]);
                p1[j] = f2(k1[0], k1[1]); p2[j] = f2(k2[0], k2[1]);
                s1[j] = f2(0.f, 0.f); s2[j] = f2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                if (!ok[u]) continue;
                float2 fx[4], fd[4];
                unpack4(rx[c][u], fx);
                unpack4(rd[c][u], fd);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 xh = __ffma2_rn(fx[j], a2[j], b2[j]);
                    float2 dz = fd[j];
                    if (p.do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
                    s1[j] = __fadd2_rn(s1[j], dz);
                    s2[j] = __ffma2_rn(dz, xh, s2[j]);
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
            if (threadIdx.x < 16) {
                double a = 0.0;
                for (int w = 0; w < kGnClusterThreads / 32; ++w) a += (double)s_red[w][threadIdx.x];     // fixed order
                my_xchg[buf * 64 + c * 16 + threadIdx.x] = a;
            }
            __syncthreads();
        }
        // ---------------- phase 2: unit totals through global memory behind the arrive counter ----------------
        if (threadIdx.x == 0) {
            __threadfence();                                   // this CTA's sums (ordered by the barrier above) before the arrival
            atomicAdd(arrive, 1u);
            const unsigned int target = (unsigned int)csize * (unsigned int)(n + 1);
            unsigned int spins = 0;
            while (ld_volatile_u32(arrive) < target) {
                if (++spins == (1u << 24)) {
                    printf("gn_bwd_creg: barrier timeout unit %d rank %d sample %d\n", unit, rank, n);
                    __trap();
                }
            }
            __threadfence();
        }
        __syncthreads();
        if ((int)threadIdx.x < UCS * 16) {
            double a = 0.0;
            const double* src = p.xchg + (size_t)unit * csize * 128 + buf * 64 + threadIdx.x;
            for (int r = 0; r < csize; ++r) a += __ldcg(src + (size_t)r * 128);                  // fixed rank order
            s_tot[threadIdx.x] = a;
        }
        __syncthreads();
        if ((int)threadIdx.x < nch_unit) {
            const int k = threadIdx.x;
            const int ch = cv0 * 8 + k;
            const int g = ch / gs;
            const int k_first = g * gs - cv0 * 8;
            double A = 0.0, B = 0.0;
            for (int q = 0; q < gs; ++q) {
                const int kk = k_first + q;
                const double gam = (double)p.gamma[cv0 * 8 + kk];
                A += s_tot[(kk >> 3) * 16 + (kk & 7)] * gam;
                B += s_tot[(kk >> 3) * 16 + 8 + (kk & 7)] * gam;
            }
            const float Af = (float)(A / p.m), Bf = (float)(B / p.m);
            const float r = p.rstd[n * 8 + g], b = -p.mean[n * 8 + g] * r;
            s_coef[k][0] = r * p.gamma[ch];
            s_coef[k][1] = b * p.gamma[ch] + p.beta[ch];
            s_coef[k][2] = -r * r * Bf;
            s_coef[k][3] = -r * (Af + b * Bf);
            acc_db += s_tot[(k >> 3) * 16 + (k & 7)];
            acc_dg += s_tot[(k >> 3) * 16 + 8 + (k & 7)];
        }
        __syncthreads();
        // ---------------- phase 3: dx = dz*p1 + x*c1 + c0 from the registers ----------------
#pragma unroll
        for (int c = 0; c < UCS; ++c) {
            float2 p1[4], p2[4], c1[4], c0[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* e0 = s_coef[c * 8 + 2 * j];
                const float* e1 = s_coef[c * 8 + 2 * j + 1];
                p1[j] = f2(e0[0], e1[0]); p2[j] = f2(e0[1], e1[1]);
                c1[j] = f2(e0[2], e1[2]); c0[j] = f2(e0[3], e1[3]);
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                if (!ok[u]) continue;
                float2 fx[4], fd[4];
                unpack4(rx[c][u], fx);
                unpack4(rd[c][u], fd);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 dz = fd[j];
                    if (p.do_lrelu) dz = __fmul2_rn(dz, lrelu_mask(__ffma2_rn(fx[j], p1[j], p2[j])));
                    fx[j] = __ffma2_rn(dz, p1[j], __ffma2_rn(fx[j], c1[j], c0[j]));
                }
                st16(p.dx.at(cv0 + c, nrow + roff[u]), pack4(fx));
            }
        }
        __syncthreads();         // s_coef / s_red are rewritten by the next sample
    }
    if (rank == 0 && (int)threadIdx.x < nch_unit) {
        p.dbeta[cv0 * 8 + threadIdx.x] = (float)acc_db;
        p.dgamma[cv0 * 8 + threadIdx.x] = (float)acc_dg;
    }
    // the last CTA of the unit to leave puts both counters back to zero (every CTA has passed its last spin by then)
    if (threadIdx.x == 0) {
        const unsigned int left = atomicAdd(arrive + 1, 1u);
        if (left == (unsigned int)csize - 1u) {
            arrive[0] = 0u;
            arrive[1] = 0u;
        }
    }
}

}  // namespace b200

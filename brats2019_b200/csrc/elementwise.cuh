// Memory-bound kernels of the ResUNet hot path: layout conversion, GroupNorm apply/backward,
// LeakyReLU, residual add, trilinear x2 up-sampling (forward + adjoint), space-to-depth,
// sigmoid/Dice, weight packing and the deterministic partial reductions.
//
// All activation tensors are chunk-planar zero-halo bf16 (common.cuh).  Kernels that walk a
// tensor use one CTA per (n, d, h) line; a thread handles 16-byte vectors (8 channels of one
// voxel) with the voxel index fastest across threads, so a warp touches 512 contiguous bytes of
// one chunk plane per access.  Halo and guard rows are never touched (they stay zero).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kEwThreads = 256;
constexpr int kMaxRedLines = 512;    // lines one gn_bwd_reduce CTA may own

__device__ __forceinline__ void line_coords(const Vol& v, int line, int& n, int& d, int& h) {
    h = line % v.H;
    int t = line / v.H;
    d = t % v.D;
    n = t / v.D;
}
// A CTA owns `lpb` consecutive lines (all of one sample: lpb divides D*H).  Vector j of the CTA
// -> (line, 8-channel chunk cv, voxel w); voxel index fastest so a warp reads 512 contiguous bytes.
// The first row of each line is computed once per CTA into shared memory (line_rows).
struct LineGeom {
    FastDiv by_nvec, by_W;
    int nvec, W, lpb;
};
struct LineVec {
    int cv;
    long long row;      // row of the voxel in the padded volume
    bool ok;
};
__device__ __forceinline__ void fill_line_rows(const Vol& v, int line0, int lpb, long long* s_rows) {
    for (int ll = threadIdx.x; ll < lpb; ll += blockDim.x) {
        const int line = line0 + ll;
        const int h = line % v.H;
        const int t = line / v.H;
        s_rows[ll] = v.row(t / v.D, (t % v.D) + 1, h + 1, 1);
    }
}
__device__ __forceinline__ LineVec line_vec(const LineGeom& g, const long long* s_rows, int j) {
    LineVec r;
    const int ll = g.by_nvec.div(j);
    const int i = j - ll * g.nvec;
    r.ok = ll < g.lpb;
    r.cv = g.by_W.div(i);
    r.row = s_rows[r.ok ? ll : 0] + (i - r.cv * g.W);
    return r;
}
constexpr int kUnroll = 4;

__device__ __forceinline__ uint4 ld16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(__nv_bfloat16* p, const uint4& q) { *reinterpret_cast<uint4*>(p) = q; }

// ---------------------------------------------------------------------------------------
// (B,Creal,D,H,W) fp32 NCDHW  ->  act; only chunk 0 is written (Creal <= 8), higher chunks of the
// destination stay zero.  Entry layout conversion for model.py:407-412 (`input = x[0]`).
// ---------------------------------------------------------------------------------------
// element loads of the model's external tensors: the reference's fp32, or the compact staging formats a host pipeline
// may use to cut host->device traffic (bf16 images, uint8 targets); converted to fp32 in registers
__device__ __forceinline__ float ld_f(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ld_f(const uint8_t* p, size_t i) { return (float)p[i]; }
__device__ __forceinline__ float4 ld_f4(const float* p, size_t i4) { return reinterpret_cast<const float4*>(p)[i4]; }
__device__ __forceinline__ float4 ld_f4(const __nv_bfloat16* p, size_t i4) {
    const uint2 q = reinterpret_cast<const uint2*>(p)[i4];
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
}
__device__ __forceinline__ float4 ld_f4(const uint8_t* p, size_t i4) {
    const uchar4 q = reinterpret_cast<const uchar4*>(p)[i4];
    return make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
}

template <typename T>
__global__ void pack_input_kernel(const T* __restrict__ x, ActRef out, Vol v, int Creal) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    int n, d, h;
    line_coords(v, blockIdx.x, n, d, h);
    const size_t plane = (size_t)v.D * v.H * v.W;
    const size_t src0 = (size_t)n * Creal * plane + ((size_t)d * v.H + h) * v.W;
    const long long row0 = v.row(n, d + 1, h + 1, 1);
    for (int w = threadIdx.x; w < v.W; w += blockDim.x) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (i < Creal) ? ld_f(x, src0 + (size_t)i * plane + w) : 0.f;
        st16(out.at(0, row0 + w), pack_bf16x8(f));
    }
}

// ---------------------------------------------------------------------------------------
// GroupNorm statistics: conv epilogue partials [ctas][N][16] -> mean/rstd [N][8]
// (aten::native_group_norm statistics, model.py:95-96/338; eps 1e-5, biased variance)
// ---------------------------------------------------------------------------------------
// grid = N, block = 256: each of the 16 sums is split over 16 threads (strided over CTAs, fixed
// order) and combined in a fixed order -> deterministic, ~ctas/16 dependent loads of latency.
// coef (optional, with gamma/beta): [N][3][C] = (p1 | p2 | gamma), z = x * p1 + p2 the normalised value whose sign the
// LeakyReLU backward needs - read by the data-gradient conv that folds this GroupNorm's backward sums into its epilogue
// (common.cuh gnb_accumulate).  do_lrelu == 0 (norm_input): p1 = 0, p2 = 1, i.e. the mask is always 1.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int ctas, int N, double count, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ coef, int C, int do_lrelu) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ double s_part[256];
    __shared__ double s_sum[16];
    const int n = blockIdx.x, t = threadIdx.x;
    const int o = t >> 4, j = t & 15;
    double a = 0.0;
    for (int c = j; c < ctas; c += 16) a += (double)partial[((size_t)c * N + n) * 16 + o];
    s_part[t] = a;
    __syncthreads();
    if (t < 16) {
        double acc = 0.0;
        for (int q = 0; q < 16; ++q) acc += s_part[t * 16 + q];
        s_sum[t] = acc;
    }
    __syncthreads();
    __shared__ float s_m[8], s_r[8];
    if (t < 8) {
        const double m = s_sum[t] / count;
        double var = s_sum[8 + t] / count - m * m;
        if (var < 0.0) var = 0.0;
        mean[n * 8 + t] = (float)m;
        rstd[n * 8 + t] = (float)(1.0 / sqrt(var + (double)eps));
        s_m[t] = (float)m;
        s_r[t] = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (coef == nullptr) return;
    __syncthreads();
    const int gs = C >> 3;
    for (int c = t; c < C; c += blockDim.x) {
        const int g = c / gs;
        const float sc = s_r[g] * gamma[c];                       // the same scale / shift gn_apply_kernel forms
        float* o = coef + (size_t)n * 3 * C;
        o[c] = do_lrelu ? sc : 0.f;
        o[C + c] = do_lrelu ? beta[c] - s_m[g] * sc : 1.f;
        o[2 * C + c] = gamma[c];
    }
}

// ---------------------------------------------------------------------------------------
// y = act( (x - mean) * rstd * gamma + beta ) [+ residual]      (model.py:105-115, 413)
//   residual is added AFTER the activation (out = x + lrelu(gn(conv))), model.py:115.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
gn_apply_kernel(ActRef x, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, ActRef residual, ActRef out, Vol v,
                int C, int do_lrelu, LineGeom lg) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ float s_scale[256], s_shift[256];
    __shared__ long long s_rows[16];
    const int lpb = lg.lpb;
    const int line0 = blockIdx.x * lpb;
    fill_line_rows(v, line0, lpb, s_rows);
    const int n = line0 / (v.D * v.H);
    const int gs = C / 8;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / gs;
        const float sc = rstd[n * 8 + g] * gamma[c];
        s_scale[c] = sc;
        s_shift[c] = beta[c] - mean[n * 8 + g] * sc;
    }
    __syncthreads();
    const int nvec = v.W * (C / 8);
    const int total = lpb * nvec;
    for (int j0 = threadIdx.x; j0 < total; j0 += blockDim.x * kUnroll) {
        LineVec lv[kUnroll];
        uint4 qx[kUnroll], qr[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            lv[u] = line_vec(lg, s_rows, j0 + u * blockDim.x);
            if (lv[u].ok) {
                qx[u] = ld16(x.at(lv[u].cv, lv[u].row));
                if (residual.base) qr[u] = ld16(residual.at(lv[u].cv, lv[u].row));
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!lv[u].ok) continue;
            const int cb = lv[u].cv * 8;
            float f[8];
            unpack_bf16x8(qx[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float z = f[k] * s_scale[cb + k] + s_shift[cb + k];
                f[k] = do_lrelu ? lrelu(z) : z;
            }
            if (residual.base) {
                float r[8];
                unpack_bf16x8(qr[u], r);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] += r[k];
            }
            st16(out.at(lv[u].cv, lv[u].row), pack_bf16x8(f));
        }
    }
}

// (Trilinear x2 and its adjoint: elementwise4.cuh.)

// ---------------------------------------------------------------------------------------
// space-to-depth / depth-to-space for the k2 s2 conv (model.py:360-363):
//   s2d channel ((kd*2+kh)*2+kw)*C + c of coarse voxel (d,h,w) = fine[2d+kd][2h+kh][2w+kw][c]
// i.e. coarse chunk (tap8 * C/8 + cf) <- fine chunk cf.  A thread moves the two kw neighbours
// (32 contiguous bytes of a fine plane).  `vc` = coarse volume; C = fine channels.
// d2s optionally adds a residual (fine layout) - the skip-gradient add.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
s2d_kernel(ActRef fine, ActRef coarse, Vol vc, int C) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    Vol vf{vc.N, vc.D * 2, vc.H * 2, vc.W * 2};
    int n, d, h;
    line_coords(vc, blockIdx.x, n, d, h);
    const int ccf = C / 8;
    const int items = ccf * 4 * vc.W;
    const long long orow0 = vc.row(n, d + 1, h + 1, 1);
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int w = i % vc.W;
        const int t = i / vc.W;
        const int dh = t & 3, cf = t >> 2;
        const long long fr = vf.row(n, 2 * d + (dh >> 1) + 1, 2 * h + (dh & 1) + 1, 2 * w + 1);
        const __nv_bfloat16* src = fine.at(cf, fr);
        st16(coarse.at((dh * 2 + 0) * ccf + cf, orow0 + w), ld16(src));
        st16(coarse.at((dh * 2 + 1) * ccf + cf, orow0 + w), ld16(src + 8));
    }
}

__global__ void __launch_bounds__(kEwThreads)
d2s_kernel(ActRef coarse, ActRef residual, ActRef fine, Vol vc, int C) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    Vol vf{vc.N, vc.D * 2, vc.H * 2, vc.W * 2};
    int n, d, h;
    line_coords(vc, blockIdx.x, n, d, h);
    const int ccf = C / 8;
    const int items = ccf * 4 * vc.W;
    const long long irow0 = vc.row(n, d + 1, h + 1, 1);
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int w = i % vc.W;
        const int t = i / vc.W;
        const int dh = t & 3, cf = t >> 2;
        const long long fr = vf.row(n, 2 * d + (dh >> 1) + 1, 2 * h + (dh & 1) + 1, 2 * w + 1);
#pragma unroll
        for (int kw = 0; kw < 2; ++kw) {
            uint4 q = ld16(coarse.at((dh * 2 + kw) * ccf + cf, irow0 + w));
            if (residual.base) {
                float a[8], b[8];
                unpack_bf16x8(q, a);
                unpack_bf16x8(ld16(residual.at(cf, fr + kw)), b);
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] += b[k];
                q = pack_bf16x8(a);
            }
            st16(fine.at(cf, fr + kw), q);
        }
    }
}

// out = a + b over the interior (gradient accumulation where two paths meet)
__global__ void __launch_bounds__(kEwThreads)
add_kernel(ActRef a, ActRef b, ActRef out, Vol v, int C, LineGeom lg) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ long long s_rows[16];
    const int lpb = lg.lpb;
    const int line0 = blockIdx.x * lpb;
    fill_line_rows(v, line0, lpb, s_rows);
    __syncthreads();
    const int nvec = v.W * (C / 8);
    const int total = lpb * nvec;
    for (int j0 = threadIdx.x; j0 < total; j0 += blockDim.x * kUnroll) {
        LineVec lv[kUnroll];
        uint4 qa[kUnroll], qb[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            lv[u] = line_vec(lg, s_rows, j0 + u * blockDim.x);
            if (lv[u].ok) {
                qa[u] = ld16(a.at(lv[u].cv, lv[u].row));
                qb[u] = ld16(b.at(lv[u].cv, lv[u].row));
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!lv[u].ok) continue;
            float fa[8], fb[8];
            unpack_bf16x8(qa[u], fa);
            unpack_bf16x8(qb[u], fb);
#pragma unroll
            for (int k = 0; k < 8; ++k) fa[k] += fb[k];
            st16(out.at(lv[u].cv, lv[u].row), pack_bf16x8(fa));
        }
    }
}

// out[j] = sum_i partial[i*stride + j], j < n_out, in double, fixed order.
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int count, int stride, int n_out,
                                       float* __restrict__ out) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    const int j = blockIdx.x;
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) a += (double)partial[(size_t)i * stride + j];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && j < n_out) out[j] = (float)sh[0];
}

// ---------------------------------------------------------------------------------------
// Dice_loss_joint (loss.py:98-122).  probs/target fp32 (B,C,S) contiguous, C <= 4.
//   sums[c] = sum p*g, sums[4+c] = sum (p^2 + g)          (per-CTA partials [blocks][8])
// ---------------------------------------------------------------------------------------
template <typename TG>
__global__ void __launch_bounds__(kEwThreads)
dice_partial_kernel(const float* __restrict__ p, const TG* __restrict__ g, float* __restrict__ partial, int B,
                    int C, long long S) {
    // grid = (blocks_x, B*C)
    const int bc = blockIdx.y;
    const int c = bc % C;
    const float* pp = p + (size_t)bc * S;
    const TG* gg = g + (size_t)bc * S;
    float si = 0.f, su = 0.f;
    // vector loads only when both rows really are aligned for them (a sliced batch is contiguous but may start anywhere)
    const bool vec = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(pp) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(gg) & (4 * sizeof(TG) - 1)) == 0);
    const long long nv = vec ? S / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(pp)[i];
        const float4 b = ld_f4(gg, (size_t)i);
        si += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        su += (a.x * a.x + b.x) + (a.y * a.y + b.y) + (a.z * a.z + b.z) + (a.w * a.w + b.w);
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
        const float gv = ld_f(gg, (size_t)i);
        si += pp[i] * gv;
        su += pp[i] * pp[i] + gv;
    }
    __shared__ float s_i[kEwThreads / 32], s_u[kEwThreads / 32];
    si = warp_sum(si);
    su = warp_sum(su);
    if ((threadIdx.x & 31) == 0) { s_i[threadIdx.x >> 5] = si; s_u[threadIdx.x >> 5] = su; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < kEwThreads / 32; ++k) { a += s_i[k]; b += s_u[k]; }
        float* o = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
        o[0] = a;
        o[1] = b;
        (void)c;
    }
}

// loss = priority * (1 - mean_c 2 (I_c + 1e-6) / (U_c + 2e-6))
__global__ void dice_loss_kernel(const float* __restrict__ sums, int C, float priority, float* __restrict__ loss) {
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc += 2.f * (sums[c] + 1e-6f) / (sums[4 + c] + 2e-6f);
        loss[0] = priority * (1.f - acc / C);
    }
}

// dL/dp = gout * priority * -(2/C) (g U_c - 2 p I_c) / U_c^2     
template <typename TG>
__global__ void __launch_bounds__(kEwThreads)
dice_bwd_kernel(const float* __restrict__ p, const TG* __restrict__ g, const float* __restrict__ sums,
                const float* __restrict__ gout, float priority, float* __restrict__ dp, int B, int C, long long S) {
    const int bc = blockIdx.y;
    const int c = bc % C;
    const float I = sums[c] + 1e-6f, U = sums[4 + c] + 2e-6f;
    const float k = -gout[0] * priority * (2.f / C) / (U * U);
    const float* pp = p + (size_t)bc * S;
    const TG* gg = g + (size_t)bc * S;
    float* dd = dp + (size_t)bc * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
        dd[i] = k * (ld_f(gg, (size_t)i) * U - 2.f * pp[i] * I);
}

// ---------------------------------------------------------------------------------------
// BCE_Loss (loss.py:64-79): -mean( g log(p+1e-6) + w (1-g) log((1+1e-6) - p) ), w = bg_weight.
//   forward: per-CTA partial sums of the bracket  (partial[blocks])
//   backward: dL/dp = -gout/numel * ( g/(p+1e-6) - w (1-g)/((1+1e-6) - p) )
// ---------------------------------------------------------------------------------------
template <typename TG>
__global__ void __launch_bounds__(kEwThreads)
bce_partial_kernel(const float* __restrict__ p, const TG* __restrict__ g, float w, float* __restrict__ partial,
                   long long n) {
    float acc = 0.f;
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(g) & (4 * sizeof(TG) - 1)) == 0);
    const long long nv = vec ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(p)[i];
        const float4 b = ld_f4(g, (size_t)i);
        acc += b.x * logf(a.x + 1e-6f) + w * (1.f - b.x) * logf((1.f + 1e-6f) - a.x);
        acc += b.y * logf(a.y + 1e-6f) + w * (1.f - b.y) * logf((1.f + 1e-6f) - a.y);
        acc += b.z * logf(a.z + 1e-6f) + w * (1.f - b.z) * logf((1.f + 1e-6f) - a.z);
        acc += b.w * logf(a.w + 1e-6f) + w * (1.f - b.w) * logf((1.f + 1e-6f) - a.w);
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gv = ld_f(g, (size_t)i);
        acc += gv * logf(p[i] + 1e-6f) + w * (1.f - gv) * logf((1.f + 1e-6f) - p[i]);
    }
    __shared__ float s_red[kEwThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < kEwThreads / 32; ++k) a += s_red[k];
        partial[blockIdx.x] = a;
    }
}
// loss = -sum / numel   (sum[0] may have been all-reduced; numel is the global element count)
__global__ void bce_loss_kernel(const float* __restrict__ sum, double numel, float* __restrict__ loss) {
    if (threadIdx.x == 0) loss[0] = (float)(-(double)sum[0] / numel);
}
template <typename TG>
__global__ void __launch_bounds__(kEwThreads)
bce_bwd_kernel(const float* __restrict__ p, const TG* __restrict__ g, const float* __restrict__ gout, float w,
               float inv_numel, float* __restrict__ dp, long long n) {
    const float k = -gout[0] * inv_numel;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gv = ld_f(g, (size_t)i);
        dp[i] = k * (gv / (p[i] + 1e-6f) - w * (1.f - gv) / ((1.f + 1e-6f) - p[i]));
    }
}

// ---------------------------------------------------------------------------------------
// Weight packing: PyTorch fp32 (Cout, Cin, kd, kh, kw) -> bf16 UMMA B-operand stage images
//   packed[job][g][tg][tl][kchunk][n][8]  (one contiguous w_stage per (job, g, tg))
// GEMM view: Wg[tap][k][n].  `kind` selects how (tap,k,n) map into the PyTorch tensor:
//   0 forward          k = ci, n = co, tap as is
//   1 data-gradient k3 k = co, n = ci, tap mirrored (26 - tap); k1: tap 0
//   2 forward k2s2 over space-to-depth: k = tap8*Cin + ci, n = co
//   3 data-gradient k2s2 (to space-to-depth layout): k = co, n = tap8*Cin + ci
// ci_off: first input channel of the PyTorch tensor this GEMM's k (kind 0) or n (kind 1) refers to.
// ---------------------------------------------------------------------------------------
struct PackParams {
    int kind, Cout_w, Cin_w, taps_w;   // PyTorch tensor dims (taps_w = kd*kh*kw)
    int ci_off;
    int K_real, N_real;                // logical GEMM extents before padding
    int n_jobs, KG, NTG, TG, KC, Nmma;
    int fold;                          // kw-fold: GEMM tap = kd*3+kh (9 taps), n = kw*(Nmma/3) + channel
};

__device__ __forceinline__ float fetch_weight(const float* w, const PackParams& q, int tap, int k, int n) {
    if (q.fold) {
        const int cg = q.Nmma / 3;
        const int kw = n / cg, c = n - kw * cg;
        if (k >= q.K_real || c >= q.N_real) return 0.f;
        const int tp = tap * 3 + kw;
        if (q.kind == 0) return w[((size_t)c * q.Cin_w + (k + q.ci_off)) * q.taps_w + tp];
        return w[((size_t)k * q.Cin_w + (c + q.ci_off)) * q.taps_w + (q.taps_w - 1 - tp)];
    }
    if (k >= q.K_real || n >= q.N_real) return 0.f;
    int co, ci, tp;
    switch (q.kind) {
        case 0: co = n; ci = k + q.ci_off; tp = tap; break;
        case 1: co = k; ci = n + q.ci_off; tp = q.taps_w - 1 - tap; break;
        case 2: co = n; ci = k % q.Cin_w; tp = k / q.Cin_w; break;
        default: co = k; ci = n % q.Cin_w; tp = n / q.Cin_w; break;
    }
    return w[((size_t)co * q.Cin_w + ci) * q.taps_w + tp];
}

__device__ __forceinline__ void pack_weight_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                 const PackParams& q, size_t i) {
    size_t t = i;
    const int nrow = t % q.Nmma; t /= q.Nmma;
    const int kch = t % (q.KC / 8); t /= (q.KC / 8);
    const int tl = t % q.TG; t /= q.TG;
    const int tg = t % q.NTG; t /= q.NTG;
    const int g = t % q.KG; t /= q.KG;
    const int job = (int)t;
    const int tap = tg * q.TG + tl;
    const int n = job * q.Nmma + nrow;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fetch_weight(w, q, tap, g * q.KC + kch * 8 + e, n);
    *reinterpret_cast<uint4*>(packed + i * 8) = pack_bf16x8(f);
}
__host__ __device__ inline size_t pack_weight_total(const PackParams& q) {
    return (size_t)q.n_jobs * q.KG * q.NTG * q.TG * (q.KC / 8) * q.Nmma;
}
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, PackParams q) {
    const size_t total = pack_weight_total(q);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        pack_weight_elem(w, packed, q, i);
}

// Marching conv (conv_march.cuh): packed[ks][t = kh*3+kw][k-chunk 2][band*CO + c][8], bands = kd 2,1,0,2,1
// so that a start offset of 0/1/2 bands rotates which kd tap lands in which accumulator slot.
struct MarchPackParams {
    int kind, Cout_w, Cin_w, ci_off, K_real, N_real, KS, CO;
};
__device__ __forceinline__ void pack_weight_march_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                       const MarchPackParams& q, int i) {
    int t = i;
    const int row = t % (5 * q.CO); t /= 5 * q.CO;
    const int kch = t % 2; t /= 2;
    const int tap9 = t % 9; t /= 9;
    const int ks = t;
    const int band = row / q.CO, c = row - band * q.CO;
    const int kd = (band == 0 || band == 3) ? 2 : ((band == 1 || band == 4) ? 1 : 0);
    const int tp = kd * 9 + tap9;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = ks * 16 + kch * 8 + e;
        float v = 0.f;
        if (k < q.K_real && c < q.N_real) {
            if (q.kind == 0) v = w[((size_t)c * q.Cin_w + (k + q.ci_off)) * 27 + tp];
            else v = w[((size_t)k * q.Cin_w + (c + q.ci_off)) * 27 + (26 - tp)];
        }
        f[e] = v;
    }
    *reinterpret_cast<uint4*>(packed + (size_t)i * 8) = pack_bf16x8(f);
}
__global__ void pack_weight_march_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, MarchPackParams q) {
    const int total = q.KS * 9 * 2 * 5 * q.CO;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) pack_weight_march_elem(w, packed, q, i);
}

// Band-marching conv (conv_band.cuh): packed[rot 3][kw 3][k-chunk 2][kh idx 3 (kh = 2,1,0)][kd band 3][16][8],
// kd of band position b under rotation r = (2,1,0,2,1)[r + b].
__device__ __forceinline__ void pack_weight_band_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                      const MarchPackParams& q, int i) {
    int t = i;
    const int row = t % 144; t /= 144;
    const int kch = t % 2; t /= 2;
    const int kw = t % 3; t /= 3;
    const int rot = t;
    const int khidx = row / 48, bpos = (row % 48) / 16, c = row % 16;
    const int kh = 2 - khidx;
    const int bi = rot + bpos;
    const int kd = (bi == 0 || bi == 3) ? 2 : ((bi == 1 || bi == 4) ? 1 : 0);
    const int tp = kd * 9 + kh * 3 + kw;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kch * 8 + e;
        float v = 0.f;
        if (k < q.K_real && c < q.N_real) {
            if (q.kind == 0) v = w[((size_t)c * q.Cin_w + (k + q.ci_off)) * 27 + tp];
            else v = w[((size_t)k * q.Cin_w + (c + q.ci_off)) * 27 + (26 - tp)];
        }
        f[e] = v;
    }
    *reinterpret_cast<uint4*>(packed + (size_t)i * 8) = pack_bf16x8(f);
}
__global__ void pack_weight_band_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, MarchPackParams q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * 2 * 144) pack_weight_band_elem(w, packed, q, i);
}

// All weights of the model in ONE launch: a device-resident job table (built once on the host by
// b200_pack_table_build) maps CTA ranges to (weight tensor, packed image, layout parameters).
struct PackJobDev {
    const float* w;
    __nv_bfloat16* packed;
    int layout;            // 0: conv_gemm.cuh images, 1: conv_march.cuh image, 2: conv_band.cuh image
    int block0, nblocks;   // CTA range of this job (256 vectors per CTA)
    int pad_;
    PackParams q;
    MarchPackParams mq;
};
__global__ void pack_weights_batched_kernel(const PackJobDev* __restrict__ jobs, int n_jobs) {
    __shared__ int s_job;
    // every thread tests a few jobs: one round of independent loads instead of a dependent scan
    for (int j = threadIdx.x; j < n_jobs; j += blockDim.x) {
        const int b0 = jobs[j].block0;
        if ((int)blockIdx.x >= b0 && (int)blockIdx.x < b0 + jobs[j].nblocks) s_job = j;
    }
    __syncthreads();
    const PackJobDev& J = jobs[s_job];
    const size_t i = (size_t)((int)blockIdx.x - J.block0) * blockDim.x + threadIdx.x;
    if (J.layout == 0) {
        if (i < pack_weight_total(J.q)) pack_weight_elem(J.w, J.packed, J.q, i);
    } else if (J.layout == 1) {
        if (i < (size_t)(J.mq.KS * 9 * 2 * 5 * J.mq.CO)) pack_weight_march_elem(J.w, J.packed, J.mq, (int)i);
    } else {
        if (i < (size_t)(9 * 2 * 144)) pack_weight_band_elem(J.w, J.packed, J.mq, (int)i);
    }
}

// ---------------------------------------------------------------------------------------
// wgrad partial reduction: partial[job][split][nacc][M][Nmma] -> fp32 grad, PyTorch layout.
//   kind 0: k3 conv  dW[co][ci][kd][kh][kw]           (banded / folded / multi-acc per flags)
//   kind 1: k1 conv  dW[co][ci_off + ci]
//   kind 2: k2s2 via s2d: X channel k = tap8*Cin + ci -> dW[co][ci][tap8]
// ---------------------------------------------------------------------------------------
struct WgradReduceParams {
    int kind, Cout_w, Cin_w, taps_w, ci_off;
    int Cout_g, Cin_g;       // channel extents of dY / X as seen by the GEMM (padded)
    int banded, folded, accs, dual, kw_acc;
    int n_jobs, splits, nacc, M, Nmma;
    int accumulate;          // add into grad instead of overwriting
};

// A 256-thread CTA owns 256/SG quads of 4 consecutive accumulator elements; thread (sg, quad) sums the
// K splits s = sg, sg+SG, ... of its quad with float4 loads (4 independent loads in flight), the SG
// partial sums are combined in a fixed order through shared memory (deterministic), and the quad is
// scattered into the PyTorch-layout gradient.  The host picks SG (1..8) so that small outputs with
// many splits (fine levels) and large outputs with few splits (coarse levels) both fill the machine.
__device__ __forceinline__ void wgrad_scatter(float* __restrict__ grad, const WgradReduceParams& q, int job, int i, double a) {
    const int col = i % q.Nmma;
    int r = i / q.Nmma;
    const int row = r % q.M;
    r /= q.M;
    const int t = r % q.nacc;
    int co, ci_w, tp;
    if (q.kind == 0) {
        // jobs enumerate (kd, accumulator-axis tap, fold-axis tap) with the fold axis fastest
        int jj = job;
        const int jtf = q.folded ? 0 : jj % 3; if (!q.folded) jj /= 3;
        const int jta = q.accs ? 0 : jj % 3;   if (!q.accs) jj /= 3;
        const int jkd = q.banded ? 0 : jj;
        int kd, tf;
        if (q.dual) {
            // 8-row blocks (band, chunk, range) and 8-col blocks (fold, chunk, range); the caller only passes
            // range-0 elements (the range-1 twin at +8 rows, +8 cols has been added to `a`)
            const int rb = row >> 4, cb = col >> 4;
            kd = rb / (q.Cout_g >> 3);
            co = (rb % (q.Cout_g >> 3)) * 8 + (row & 7);
            tf = cb / (q.Cin_g >> 3);
            ci_w = (cb % (q.Cin_g >> 3)) * 8 + (col & 7);
        } else {
            kd = q.banded ? row / q.Cout_g : jkd;
            co = q.banded ? row % q.Cout_g : row;
            tf = q.folded ? col / q.Cin_g : jtf;
            ci_w = q.folded ? col % q.Cin_g : col;
        }
        const int ta = q.accs ? t : jta;
        const int kh = q.kw_acc ? tf : ta, kw = q.kw_acc ? ta : tf;
        if (kd > 2 || tf > 2) return;      // unused band / padding of the M or N extent
        tp = (kd * 3 + kh) * 3 + kw;
    } else if (q.kind == 1) {
        if (job * q.Nmma + col >= q.Cin_g) return;
        co = row; ci_w = q.ci_off + job * q.Nmma + col; tp = 0;
    } else {
        const int k = job * q.Nmma + col;
        co = row; tp = k / q.Cin_w; ci_w = k % q.Cin_w;
    }
    if (co >= q.Cout_w || ci_w >= q.Cin_w || tp >= q.taps_w) return;
    const size_t o = ((size_t)co * q.Cin_w + ci_w) * q.taps_w + tp;
    if (q.accumulate) grad[o] += (float)a;
    else grad[o] = (float)a;
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, WgradReduceParams q, int SG) {
    pdl_wait();         // (launched as a programmatic dependent of the weight-gradient GEMM, common.cuh)
    __shared__ double s_acc[256][4];
    const int per_job = q.nacc * q.M * q.Nmma;            // multiple of 4
    const int quads = q.n_jobs * (per_job >> 2);
    const int qpc = 256 / SG;                             // quads per CTA
    const int sg = threadIdx.x / qpc, ql = threadIdx.x - sg * qpc;
    const int quad = blockIdx.x * qpc + ql;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int job = 0, i = 0;
    bool active = quad < quads;
    if (active) {
        i = quad << 2;
        job = i / per_job;
        i -= job * per_job;
        if (q.dual) {      // only range-0 blocks own an output; their range-1 twin sits 8 rows and 8 columns further
            const int col = i % q.Nmma, row = (i / q.Nmma) % q.M;
            active = (((row >> 3) & 1) == 0) && (((col >> 3) & 1) == 0);
        }
    }
    if (active) {
        const float* src = partial + (size_t)job * q.splits * per_job + (size_t)i;
        if (q.dual) {
            const float* tw = src + 8 * q.Nmma + 8;
            for (int s2 = sg; s2 < q.splits; s2 += SG) {
                const float4 v = *reinterpret_cast<const float4*>(tw + (size_t)s2 * per_job);
                a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            }
        }
        int s = sg;
        for (; s + 3 * SG < q.splits; s += 4 * SG) {
            const float4 v0 = *reinterpret_cast<const float4*>(src + (size_t)s * per_job);
            const float4 v1 = *reinterpret_cast<const float4*>(src + (size_t)(s + SG) * per_job);
            const float4 v2 = *reinterpret_cast<const float4*>(src + (size_t)(s + 2 * SG) * per_job);
            const float4 v3 = *reinterpret_cast<const float4*>(src + (size_t)(s + 3 * SG) * per_job);
            a0 += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
            a1 += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
            a2 += ((double)v0.z + (double)v1.z) + ((double)v2.z + (double)v3.z);
            a3 += ((double)v0.w + (double)v1.w) + ((double)v2.w + (double)v3.w);
        }
        for (; s < q.splits; s += SG) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * per_job);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
    }
    if (SG > 1) {
        s_acc[threadIdx.x][0] = a0; s_acc[threadIdx.x][1] = a1; s_acc[threadIdx.x][2] = a2; s_acc[threadIdx.x][3] = a3;
        __syncthreads();
        if (sg != 0) return;
        for (int k = 1; k < SG; ++k) {
            a0 += s_acc[k * qpc + ql][0]; a1 += s_acc[k * qpc + ql][1];
            a2 += s_acc[k * qpc + ql][2]; a3 += s_acc[k * qpc + ql][3];
        }
    }
    if (!active) return;
    wgrad_scatter(grad, q, job, i + 0, a0);
    wgrad_scatter(grad, q, job, i + 1, a1);
    wgrad_scatter(grad, q, job, i + 2, a2);
    wgrad_scatter(grad, q, job, i + 3, a3);
}

}  // namespace b200

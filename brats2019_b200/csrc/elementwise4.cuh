// Third-generation trilinear x2 kernels (aten::upsample_trilinear3d(+_backward) + leaky_relu, model.py:7-14, 422).
//
// ncu on the second generation (elementwise2.cuh, profiles/r01_ncu_upsample_issue_bound.txt): 70-85 % issue-slot
// utilisation at 35-60 % of HBM bandwidth - the kernels were bound by instruction issue (245 instructions per
// 16-byte output vector forward, 526 and 1139 in the two adjoint passes), not by memory.  The stencils below do
// the same arithmetic with each loaded vector unpacked once and reused for every output it feeds, compile-time
// tap weights and pointer increments instead of per-load 64-bit index arithmetic:
//   forward   one CTA per COARSE line, thread = (chunk, coarse w): 27 loads -> the 2x2x2 block of fine voxels
//   adjoint A thread = (chunk, coarse w): its own two fine voxels loaded and masked once, the outer taps by warp shuffle
//   adjoint B 16 row offsets per CTA in shared memory, fixed 4x4 taps with clamped indices
// Tap rule (align_corners = False, scale 2): fine 2k = 0.25 in[k-1] + 0.75 in[k], fine 2k+1 = 0.75 in[k] +
// 0.25 in[k+1], indices clamped to the volume; the adjoint is the same 4-tap stencil (0.25, 0.75, 0.75, 0.25) over
// fine indices 2k-1 .. 2k+2 clamped to the volume (a clamped duplicate adds its weight, exactly as the forward
// clamp does).
#pragma once
#include "elementwise2.cuh"

namespace b200 {

__device__ __forceinline__ float2 lrelu2(float2 v) {
    const float2 s = __fmul2_rn(v, f2(0.01f, 0.01f));
    return f2(fmaxf(v.x, s.x), fmaxf(v.y, s.y));
}

__global__ void __launch_bounds__(128)
upsample2x_fwd3_kernel(ActRef in, ActRef out, Vol vc, int C, int do_lrelu, FastDiv by_Wc) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    Vol vf{vc.N, vc.D * 2, vc.H * 2, vc.W * 2};
    int n, d, h;
    line_coords(vc, blockIdx.x, n, d, h);
    const int dd[3] = {max(d - 1, 0), d, min(d + 1, vc.D - 1)};
    const int hh[3] = {max(h - 1, 0), h, min(h + 1, vc.H - 1)};
    const float2 q25 = f2(0.25f, 0.25f), q75 = f2(0.75f, 0.75f);
    const long long plane = in.plane_rows * 8;                       // elements between chunk planes
    const __nv_bfloat16* ibase = in.base + in.guard * 8;
    __nv_bfloat16* obase = out.base + out.guard * 8;
    const long long oplane = out.plane_rows * 8;
    const int items = vc.W * (C >> 3);
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int cv = by_Wc.div(i), w = i - cv * vc.W;
        const int wm = max(w - 1, 0), wp = min(w + 1, vc.W - 1);
        float2 o[2][2][2][4];                                        // [fine d parity][fine h parity][fine w parity][4 x 2 ch]
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float2 hb[2][2][4];                                      // blended over w and h for this coarse slice
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const __nv_bfloat16* r = ibase + cv * plane + vc.row(n, dd[a] + 1, hh[b] + 1, 1) * 8;
                float2 tm[4], tc[4], tp[4];
                unpack4(ld16(r + wm * 8), tm);
                unpack4(ld16(r + w * 8), tc);
                unpack4(ld16(r + wp * 8), tp);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 c75 = __fmul2_rn(tc[j], q75);
                    const float2 ev = __ffma2_rn(tm[j], q25, c75), od = __ffma2_rn(tp[j], q25, c75);
                    if (b == 0) { hb[0][0][j] = __fmul2_rn(ev, q25); hb[0][1][j] = __fmul2_rn(od, q25); }
                    if (b == 1) {
                        hb[0][0][j] = __ffma2_rn(ev, q75, hb[0][0][j]); hb[0][1][j] = __ffma2_rn(od, q75, hb[0][1][j]);
                        hb[1][0][j] = __fmul2_rn(ev, q75);              hb[1][1][j] = __fmul2_rn(od, q75);
                    }
                    if (b == 2) { hb[1][0][j] = __ffma2_rn(ev, q25, hb[1][0][j]); hb[1][1][j] = __ffma2_rn(od, q25, hb[1][1][j]); }
                }
            }
#pragma unroll
            for (int fh = 0; fh < 2; ++fh)
#pragma unroll
                for (int fw = 0; fw < 2; ++fw)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 v = hb[fh][fw][j];
                        if (a == 0) o[0][fh][fw][j] = __fmul2_rn(v, q25);
                        if (a == 1) { o[0][fh][fw][j] = __ffma2_rn(v, q75, o[0][fh][fw][j]); o[1][fh][fw][j] = __fmul2_rn(v, q75); }
                        if (a == 2) o[1][fh][fw][j] = __ffma2_rn(v, q25, o[1][fh][fw][j]);
                    }
        }
#pragma unroll
        for (int fd = 0; fd < 2; ++fd)
#pragma unroll
            for (int fh = 0; fh < 2; ++fh) {
                __nv_bfloat16* dst = obase + cv * oplane + (vf.row(n, 2 * d + fd + 1, 2 * h + fh + 1, 1) + 2 * w) * 8;
                if (do_lrelu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { o[fd][fh][0][j] = lrelu2(o[fd][fh][0][j]); o[fd][fh][1][j] = lrelu2(o[fd][fh][1][j]); }
                }
                st16(dst, pack4(o[fd][fh][0]));
                st16(dst + 8, pack4(o[fd][fh][1]));
            }
    }
}

// Adjoint pass A: T[n, fd, fh, w] = 0.25 g[2w-1] + 0.75 g[2w] + 0.75 g[2w+1] + 0.25 g[2w+2] (indices clamped) with
// g = dy * lrelu'(y).  One CTA per fine line; a thread owns (chunk, coarse w), loads and masks its OWN two fine
// voxels 2w, 2w+1 (a warp reads 1 KB contiguous) and takes the two outer taps from its neighbour lanes by shuffle;
// only the lanes at a warp or line edge load theirs.  (A first version gave each thread a run of four coarse voxels:
// fewer instructions, but lanes 128 B apart - 32 sectors per request - and slower than the second generation.)
__device__ __forceinline__ void up_masked(const __nv_bfloat16* gp, const __nv_bfloat16* yp, int f, int do_lrelu, float2* g) {
    unpack4(ld16(gp + (long long)f * 8), g);
    if (do_lrelu) {
        float2 yy[4];
        unpack4(ld16(yp + (long long)f * 8), yy);
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = __fmul2_rn(g[j], lrelu_mask(yy[j]));
    }
}

__global__ void __launch_bounds__(128)
upsample2x_bwd_w3_kernel(ActRef dy, ActRef y, ActRef T, Vol vc, int C, int do_lrelu, FastDiv by_Wc) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    Vol vf{vc.N, vc.D * 2, vc.H * 2, vc.W * 2};
    Vol vt{vc.N, vc.D * 2, vc.H * 2, vc.W};
    int n, d, h;
    line_coords(vf, blockIdx.x, n, d, h);
    const long long irow0 = vf.row(n, d + 1, h + 1, 1);
    const long long orow0 = vt.row(n, d + 1, h + 1, 1);
    const int items = vc.W * (C >> 3);
    const int lane = threadIdx.x & 31;
    const float2 q25 = f2(0.25f, 0.25f), q75 = f2(0.75f, 0.75f);
    for (int base = 0; base < items; base += blockDim.x) {          // uniform trip count: every lane takes part in the shuffles
        const int i = base + threadIdx.x;
        const bool active = i < items;
        const int cv = active ? by_Wc.div(i) : 0, w = active ? i - cv * vc.W : 0;
        const __nv_bfloat16* gp = dy.at(cv, irow0);
        const __nv_bfloat16* yp = y.at(cv, irow0);
        float2 ev[4], od[4];
        if (active) {
            up_masked(gp, yp, 2 * w, do_lrelu, ev);
            up_masked(gp, yp, 2 * w + 1, do_lrelu, od);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { ev[j] = f2(0.f, 0.f); od[j] = f2(0.f, 0.f); }
        }
        float2 pm[4], pn[4];                                         // g[2w-1] (the left lane's odd voxel), g[2w+2] (the right lane's even voxel)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pm[j].x = __shfl_up_sync(0xffffffffu, od[j].x, 1); pm[j].y = __shfl_up_sync(0xffffffffu, od[j].y, 1);
            pn[j].x = __shfl_down_sync(0xffffffffu, ev[j].x, 1); pn[j].y = __shfl_down_sync(0xffffffffu, ev[j].y, 1);
        }
        if (!active) continue;
        if (lane == 0 || w == 0) {
            if (w > 0) up_masked(gp, yp, 2 * w - 1, do_lrelu, pm);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) pm[j] = ev[j];           // clamp: fine index -1 -> 0
            }
        }
        if (lane == 31 || w == vc.W - 1) {
            if (w < vc.W - 1) up_masked(gp, yp, 2 * w + 2, do_lrelu, pn);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) pn[j] = od[j];           // clamp: fine index 2W -> 2W - 1
            }
        }
        float2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[j] = __ffma2_rn(__fadd2_rn(pm[j], pn[j]), q25, __fmul2_rn(__fadd2_rn(ev[j], od[j]), q75));
        st16(T.at(cv, orow0 + w), pack4(acc));
    }
}

// Adjoint pass B: dcoarse[n, d, h, w] = sum over fine (fd, fh) in {2d-1 .. 2d+2} x {2h-1 .. 2h+2} (clamped) of
// wd * wh * T[n, fd, fh, w].  One CTA per coarse line; the 16 row offsets are computed once per CTA.
__global__ void __launch_bounds__(128)
upsample2x_bwd_dh3_kernel(ActRef T, ActRef dcoarse, Vol vc, int C, FastDiv by_Wc) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    Vol vt{vc.N, vc.D * 2, vc.H * 2, vc.W};
    __shared__ long long s_off[16];
    int n, d, h;
    line_coords(vc, blockIdx.x, n, d, h);
    if (threadIdx.x < 16) {
        const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
        const int fd = min(max(2 * d - 1 + a, 0), 2 * vc.D - 1), fh = min(max(2 * h - 1 + b, 0), 2 * vc.H - 1);
        s_off[threadIdx.x] = (T.guard + vt.row(n, fd + 1, fh + 1, 1)) * 8;
    }
    __syncthreads();
    const long long orow0 = vc.row(n, d + 1, h + 1, 1);
    const long long plane = T.plane_rows * 8;
    const int items = vc.W * (C >> 3);
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int cv = by_Wc.div(i), w = i - cv * vc.W;
        const __nv_bfloat16* base = T.base + cv * plane + w * 8;
        float2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = f2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float2 row[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) row[j] = f2(0.f, 0.f);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float2 g[4];
                unpack4(ld16(base + s_off[a * 4 + b]), g);
                const float wb = (b == 0 || b == 3) ? 0.25f : 0.75f;
                const float2 wt = f2(wb, wb);
#pragma unroll
                for (int j = 0; j < 4; ++j) row[j] = __ffma2_rn(g[j], wt, row[j]);
            }
            const float wa = (a == 0 || a == 3) ? 0.25f : 0.75f;
            const float2 wt = f2(wa, wa);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(row[j], wt, acc[j]);
        }
        st16(dcoarse.at(cv, orow0 + w), pack4(acc));
    }
}

// ---------------------------------------------------------------------------------------
// Layout conversions at the two ends of the model, four voxels per thread with 16-byte fp32 loads (the
// one-voxel-per-thread forms spent ~260 instructions per voxel on index arithmetic and scalar loads: 70 % issue-slot
// utilisation at half of the HBM bandwidth, profiles/r01_ncu_upsample_issue_bound.txt).  A CTA owns `lpb` <= 16 lines
// whose source offsets and destination rows are computed once into shared memory.  W % 4 == 0 and 16-byte aligned
// fp32 tensors are required (the entry points fall back to the scalar kernels otherwise).
// ---------------------------------------------------------------------------------------
// (B,Creal,D,H,W) fp32 NCDHW -> chunk 0 of an act tensor (model.py:407-412), Creal <= 8.
template <typename T>
__global__ void __launch_bounds__(256)
pack_input4_kernel(const T* __restrict__ x, ActRef out, Vol v, int Creal, int lpb, FastDiv by_Q) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ long long s_src[16], s_row[16];
    const int line0 = blockIdx.x * lpb;
    const int nl = min(lpb, v.N * v.D * v.H - line0);
    const size_t plane = (size_t)v.D * v.H * v.W;
    if ((int)threadIdx.x < nl) {
        int n, d, h;
        line_coords(v, line0 + threadIdx.x, n, d, h);
        s_src[threadIdx.x] = (long long)((size_t)n * Creal * plane + ((size_t)d * v.H + h) * v.W);
        s_row[threadIdx.x] = v.row(n, d + 1, h + 1, 1);
    }
    __syncthreads();
    const int Q = v.W >> 2;
    for (int idx = threadIdx.x; idx < nl * Q; idx += blockDim.x) {
        const int li = by_Q.div(idx), q = idx - li * Q;
        const size_t src4 = (size_t)(s_src[li] >> 2) + q;        // in units of four elements (W % 4 == 0, aligned base)
        float4 c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            c[i] = (i < Creal) ? ld_f4(x, src4 + (size_t)i * (plane >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat16* dst = out.at(0, s_row[li] + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = k == 0 ? c[i].x : k == 1 ? c[i].y : k == 2 ? c[i].z : c[i].w;
            st16(dst + k * 8, pack_bf16x8(f));
        }
    }
}

// sigmoid backward + layout (model.py:431 backward): dlogit = gp * p * (1 - p) -> chunk 0 of an act tensor
// (channels >= Creal stay zero), per-CTA bias-gradient partials [blocks][4].  Creal <= 4.
__global__ void __launch_bounds__(256)
sigmoid_bwd_pack4_kernel(const float* __restrict__ gp, const float* __restrict__ probs, ActRef dlogit,
                         float* __restrict__ bias_partial, Vol v, int Creal, int lpb, FastDiv by_Q) {
    pdl_wait();         // (launched as a programmatic dependent: nothing above touches global memory, common.cuh)
    pdl_trigger();      // the next kernel may become resident while this grid drains - never more than one kernel ahead
    __shared__ long long s_src[16], s_row[16];
    __shared__ float s_red[8][4];
    const int line0 = blockIdx.x * lpb;
    const int nl = min(lpb, v.N * v.D * v.H - line0);
    const size_t plane = (size_t)v.D * v.H * v.W;
    if ((int)threadIdx.x < nl) {
        int n, d, h;
        line_coords(v, line0 + threadIdx.x, n, d, h);
        s_src[threadIdx.x] = (long long)((size_t)n * Creal * plane + ((size_t)d * v.H + h) * v.W);
        s_row[threadIdx.x] = v.row(n, d + 1, h + 1, 1);
    }
    __syncthreads();
    const int Q = v.W >> 2;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int idx = threadIdx.x; idx < nl * Q; idx += blockDim.x) {
        const int li = by_Q.div(idx), q = idx - li * Q;
        const long long src = s_src[li] + 4 * q;
        float4 g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < Creal) {
                const float4 pr = *reinterpret_cast<const float4*>(probs + src + (size_t)i * plane);
                const float4 gg = *reinterpret_cast<const float4*>(gp + src + (size_t)i * plane);
                g[i] = make_float4(gg.x * pr.x * (1.f - pr.x), gg.y * pr.y * (1.f - pr.y), gg.z * pr.z * (1.f - pr.z),
                                   gg.w * pr.w * (1.f - pr.w));
                bsum[i] += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            }
        }
        __nv_bfloat16* dst = dlogit.at(0, s_row[li] + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                f[i] = i < 4 ? (k == 0 ? g[i].x : k == 1 ? g[i].y : k == 2 ? g[i].z : g[i].w) : 0.f;
            st16(dst + k * 8, pack_bf16x8(f));
        }
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

}  // namespace b200

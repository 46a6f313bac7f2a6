// Fused Adam (amsgrad, L2 weight decay) over flat fp32 buffers: the optimizer step of the reference recipe
// (main.py:133-138 `optim.Adam(lr=2e-5, weight_decay=1e-6, amsgrad=True)`, stepped at train.py:220) as ONE
// memory-bound kernel instead of the nine multi_tensor_apply launches of torch's fused Adam (336 us per step for
// 4.5 M parameters; 5 reads + 4 writes of 18 MB is ~25 us of HBM time).
//
// Parameters, gradients and the three moment buffers live at the same offsets of five flat buffers; the ACTIVE
// element ranges (parameters that received a gradient: the reference's 7 dead parameters never do, and
// torch.optim.Adam skips them entirely - no weight decay either) are given as a short segment table.
// `step` (fp32 scalar, like torch's capturable Adam) and `lr` live in device memory so that the launch can be
// replayed inside a CUDA graph; the last CTA to finish advances `step`.
//
// Arithmetic follows torch/optim/adam.py (_single_tensor_adam) in fp32:
//   g += wd * p;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  vmax = max(vmax, v)
//   p -= (lr / (1 - b1^t)) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kAdamMaxSegs = 64;
constexpr int kAdamThreads = 256;

struct AdamSegs {
    int n;
    long long begin[kAdamMaxSegs];      // first float4 of the segment in the flat buffers
    long long cum[kAdamMaxSegs + 1];    // float4s before the segment in the compacted index space
};

struct AdamParams {
    float* p;
    const float* g;
    float* m;
    float* v;
    float* vmax;                // may be null (amsgrad off)
    const float* lr;            // device scalar
    float* step;                // device scalar: number of steps taken so far
    unsigned* ticket;           // device counter, zero between launches
    float beta1, beta2, eps, weight_decay;
    double beta1d, beta2d;                      // for the bias corrections 1 - beta^t (torch computes them in double)
    float one_minus_beta1, one_minus_beta2;     // computed in double on the host, like torch's `1 - beta2` Python floats
    int lr_step_size;           // > 0: lr *= lr_gamma ^ floor(step / lr_step_size)  (StepLR stepped once per batch)
    float lr_gamma;
    AdamSegs segs;
};

__global__ void __launch_bounds__(kAdamThreads)
adam_step_kernel(const __grid_constant__ AdamParams q) {
    __shared__ float s_c[4];
    if (threadIdx.x == 0) {
        const float t_old = *q.step;
        const double t = (double)t_old + 1.0;
        float lr = *q.lr;
        if (q.lr_step_size > 0) lr *= (float)pow((double)q.lr_gamma, floor((double)t_old / q.lr_step_size));
        const double bc1 = 1.0 - pow(q.beta1d, t);
        const double bc2 = 1.0 - pow(q.beta2d, t);
        s_c[0] = (float)((double)lr / bc1);          // step size
        s_c[1] = (float)sqrt(bc2);                   // sqrt of the second bias correction
    }
    __syncthreads();
    const float step_size = s_c[0], bc2_sqrt = s_c[1];
    const float b1 = q.beta1, b2 = q.beta2, eps = q.eps, wd = q.weight_decay;
    const float w1 = q.one_minus_beta1, w2 = q.one_minus_beta2;
    const long long total = q.segs.cum[q.segs.n];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int lo = 0, hi = q.segs.n - 1;              // segment of compact index i (a handful of entries)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (q.segs.cum[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const long long o = q.segs.begin[lo] + (i - q.segs.cum[lo]);
        float4 p4 = reinterpret_cast<float4*>(q.p)[o];
        const float4 g4 = reinterpret_cast<const float4*>(q.g)[o];
        float4 m4 = reinterpret_cast<float4*>(q.m)[o];
        float4 v4 = reinterpret_cast<float4*>(q.v)[o];
        float4 x4 = q.vmax ? reinterpret_cast<float4*>(q.vmax)[o] : make_float4(0.f, 0.f, 0.f, 0.f);
        float* pp = &p4.x; const float* gg = &g4.x; float* mm = &m4.x; float* vv = &v4.x; float* xx = &x4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float g = gg[e] + wd * pp[e];
            mm[e] = mm[e] + w1 * (g - mm[e]);
            vv[e] = b2 * vv[e] + w2 * g * g;
            float d = vv[e];
            if (q.vmax) { xx[e] = fmaxf(xx[e], vv[e]); d = xx[e]; }
            pp[e] -= step_size * (mm[e] / (sqrtf(d) / bc2_sqrt + eps));
        }
        reinterpret_cast<float4*>(q.p)[o] = p4;
        reinterpret_cast<float4*>(q.m)[o] = m4;
        reinterpret_cast<float4*>(q.v)[o] = v4;
        if (q.vmax) reinterpret_cast<float4*>(q.vmax)[o] = x4;
    }
    // the last CTA to finish advances the step counter: every CTA has read it by then
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(q.ticket, 1u) == gridDim.x - 1) {
            *q.step = *q.step + 1.f;
            *q.ticket = 0u;
        }
    }
}

}  // namespace b200

// Common device helpers for the sm_100a kernels: mbarrier, TMA, TMEM and tcgen05 wrappers.
// Everything here is inline PTX; there is no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Perf probes (skip-a-stage switches, in-kernel cycle accounting) exist only in the probes build
// (`python -m brats2019_b200.build --probes`, -DB200_PROBES): in the default library B200_DBG() is a compile-time
// false and every probe branch is dead code.
#ifdef B200_PROBES
#define B200_DBG(p, bits) (((p).debug & (bits)) != 0)
#else
#define B200_DBG(p, bits) false
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------
// Activation layout in HBM: chunk-planar, zero-halo ("NC/8 DHW 8c" with halos and guards)
//
//   act[chunk][guard + row + guard][8]     bf16, chunk = channel / 8
//   row = ((n*(D+2) + dp)*(H+2) + hp)*(W+2) + wp    linear index over the zero-padded volume
//
//   * interior voxels are dp,hp,wp >= 1; every halo voxel is 0 and is never written, so a
//     3x3x3 tap is a constant row shift and needs no boundary handling;
//   * `guard` zero rows before and after each plane let tiles run past either end of the
//     tensor (ragged last tile, halo of the first/last slice, wgrad K tails) without any
//     out-of-bounds logic;
//   * the 8 channels x R rows a tensor-core operand needs are ONE contiguous range of a plane,
//     fetched by a single cp.async.bulk, and land in shared memory already in the SWIZZLE_NONE
//     canonical UMMA layout (8 rows x 16 B core matrices).
// ---------------------------------------------------------------------------------------
struct Vol {
    int N, D, H, W;
    __host__ __device__ int Dp() const { return D + 2; }
    __host__ __device__ int Hp() const { return H + 2; }
    __host__ __device__ int Wp() const { return W + 2; }
    __host__ __device__ long long slice_rows() const { return (long long)Hp() * Wp(); }
    __host__ __device__ long long sample_rows() const { return (long long)Dp() * Hp() * Wp(); }
    __host__ __device__ long long total_rows() const { return (long long)N * sample_rows(); }
    __host__ __device__ long long row(int n, int dp, int hp, int wp) const {
        return (((long long)n * Dp() + dp) * Hp() + hp) * Wp() + wp;
    }
    // zero rows kept before and after every chunk plane
    __host__ __device__ long long guard_rows() const { return (slice_rows() + 2LL * Wp() + 1024 + 7) / 8 * 8; }
    __host__ __device__ long long plane_rows() const { return total_rows() + 2 * guard_rows(); }
};

// Division by a runtime constant without the ~25-instruction integer divide:
//   n / d == umulhi(n, mul) >> sh   with mul = ceil(2^(31+s)/d), sh = s-1, s = ceil(log2 d)
// exact for 0 <= n < 2^24 (tests/test_plan_cpu.py checks the formula for every divisor in use).
struct FastDiv {
    unsigned mul, sh, d;
    __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((unsigned)n, mul) >> sh); }
};
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    f.mul = 0;
    f.sh = 0;
    if (d <= 1) return f;
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    f.mul = (unsigned)(((1ull << (31 + s)) + d - 1) / d);
    f.sh = s - 1;
    return f;
}

// A view of one activation tensor: base pointer of the allocation + plane geometry.
struct ActRef {
    __nv_bfloat16* base;
    long long plane_rows, guard;
    __host__ __device__ __nv_bfloat16* at(int chunk, long long row) const {
        return base + ((size_t)chunk * plane_rows + guard + row) * 8;
    }
};
__host__ __device__ inline ActRef make_act(const void* p, const Vol& v) {
    ActRef a;
    a.base = (__nv_bfloat16*)p;
    a.plane_rows = v.plane_rows();
    a.guard = v.guard_rows();
    return a;
}

// ---------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may become resident while its stream predecessor is still draining;
// pdl_wait() blocks until that predecessor has completed and its memory is visible - everything that reads or writes a
// tensor of an earlier kernel goes after it, what comes before (mbarrier init, TMEM allocation, the first weight
// copies: weights were packed long before) overlaps the predecessor's tail.  pdl_trigger() in the PREDECESSOR lets the
// dependent grid start launching once every CTA of this grid has started (without it the trigger is the grid's end and
// nothing overlaps).  Both are no-ops in an ordinary launch.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin on the barrier phase.  A watchdog turns a protocol bug into a trap instead of a hang
// (a hung GPU box costs a strike); 2^31 polls is minutes, far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == 0x40000000u) {
            printf("mbar_wait timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------
// One 128-bit read-only load.  Written as asm because nvcc turns `*reinterpret_cast<const uint4*>(ptr)` in the conv
// epilogues into FOUR 32-bit LDG.E (cuobjdump: 64 scalar loads per 128-channel row, 0 vector loads): four times the L1
// requests, and at the 32-byte row stride of the depth-to-space epilogue 32 sectors per request - ncu showed that kernel
// bound by L1 sector throughput (profiles/r02c_ncu_d2s_summary.txt).  Not volatile: a pure function of the address (the
// operand is never written by the kernel that reads it), so the compiler may schedule it early.
__device__ __forceinline__ uint4 ld_nc_v4(const void* ptr) {
    uint4 q;
    asm("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(ptr));
    return q;
}

// 256-bit global accesses (sm_100: LDG / STG .ENL2.256): one full 32-byte sector per lane; `ptr` 32-byte aligned.
__device__ __forceinline__ void ld_nc_v8(const void* ptr, uint4& a, uint4& b) {
    asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(ptr));
}
__device__ __forceinline__ void st_v8(void* ptr, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// Request the 32-byte sector holding `ptr` from L2 (no register, no dependency: used to take DRAM latency off a chain of
// dependent epilogue loads).
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// 1D bulk copy global -> shared (UBLKCP); bytes % 16 == 0, both addresses 16B aligned.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// TMEM + tcgen05
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
    // no "memory" clobber: ordering against the mbarrier waits is provided by tc_fence_after(); a
    // clobber here makes the compiler re-load every kernel parameter between two MMA issues.
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8x16B core matrices).
//   K-major : 8 rows x 16B core matrix contiguous (128B); LBO = byte stride between the two
//             16B K-chunks of one K=16 slab; SBO = byte stride between 8-row groups.
//   MN-major: core matrix = 8 K-rows x 16B (8 MN elements) contiguous (128B); SBO = byte stride
//             between 8-element MN blocks; LBO = byte stride between 8-row K groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // D format: f32
           | (1u << 7)                     // A format: bf16
           | (1u << 10)                    // B format: bf16
           | ((uint32_t)a_mn_major << 15)  // A major: 0 = K, 1 = MN
           | ((uint32_t)b_mn_major << 16)  // B major
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Same load without the wait: issue several, then tmem_ld_wait() once, then read the registers.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// packing helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
    uint4 q;
    q.x = pack_bf16x2(f[0], f[1]);
    q.y = pack_bf16x2(f[2], f[3]);
    q.z = pack_bf16x2(f[4], f[5]);
    q.w = pack_bf16x2(f[6], f[7]);
    return q;
}
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

// ---------------------------------------------------------------------------------------
// GroupNorm-backward statistics folded into the epilogue of the data-gradient conv that PRODUCES the gradient
// (replaces the gn_bwd_reduce2 pass over two tensors; model.py:105-112 backward).  For the GroupNorm y = lrelu(z),
// z = c * p1 + p2 (p1 = rstd * gamma, p2 = beta - mean * rstd * gamma), whose input gradient is needed next, the
// closed form needs per (sample, group)   A = sum gamma * dz,   U = sum gamma * dz * c   with dz = dy * lrelu'(z):
// `dy` are the CO values this thread is about to store (already rounded to bf16, exactly what the apply kernel will
// read back), `cq` the CO/8 vectors of the saved conv output c at the same voxel, coef = [3][C] (p1 | p2 | gamma) of
// the sample.  a[g], u[g] accumulate the eight groups (CO / 8 channels each).
// ---------------------------------------------------------------------------------------
template <int CO>
__device__ __forceinline__ void gnb_accumulate(const uint4* dyq, const uint4* cq, const float* __restrict__ coef, int C,
                                               float* a, float* u) {
    constexpr int GS = CO / 8;
#pragma unroll
    for (int k = 0; k < CO / 8; ++k) {
        float dy[8], c[8];
        unpack_bf16x8(dyq[k], dy);
        unpack_bf16x8(cq[k], c);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(coef + k * 8 + h * 4));
            const float4 p2 = __ldg(reinterpret_cast<const float4*>(coef + C + k * 8 + h * 4));
            const float4 gm = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + k * 8 + h * 4));
            const float P1[4] = {p1.x, p1.y, p1.z, p1.w}, P2[4] = {p2.x, p2.y, p2.z, p2.w}, G[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = k * 8 + h * 4 + e;
                const float z = fmaf(c[h * 4 + e], P1[e], P2[e]);
                const float w = z > 0.f ? G[e] : 0.01f * G[e];
                const float t = dy[h * 4 + e] * w;
                a[i / GS] += t;
                u[i / GS] = fmaf(t, c[h * 4 + e], u[i / GS]);
            }
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// ---------------------------------------------------------------------------------------
// "The last CTA finishes the reduction": instead of a separate tiny kernel that sums per-CTA partial results (a
// 4-9 us link in a dependent chain of launches, 50 of them per training step), every CTA takes a ticket after its
// partials are written and the one that draws the last ticket does the finalize pass, reading the other CTAs' partials
// through L2 (ld.cg).  `ticket` is a zero-initialised word in global memory; the last CTA leaves it zero again, so
// the same word serves every launch on a stream (and every CUDA-graph replay).
// Call from ALL threads of the CTA, after the partials have been stored; returns true in every thread of the last CTA.
// Memory ordering: bar.sync orders the CTA's stores before thread 0's fence, the fence (cumulative, gpu scope) before
// the ticket; the last CTA fences again after reading the ticket before it loads the partials.
// ---------------------------------------------------------------------------------------
// s_ticket: one word of shared memory (the tensor-core kernels have no static shared memory to spare: they pass a
// word of their - by then idle - dynamic allocation).
__device__ __forceinline__ bool cta_draws_last_ticket(unsigned int* ticket, unsigned int total, unsigned int* s_ticket) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        *s_ticket = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    const bool last = (*s_ticket == total - 1u);
    if (last) __threadfence();
    return last;
}

// GroupNorm statistics finished by the last CTA of the conv whose epilogue produced the partial sums
// (replaces gn_finalize_kernel; same summation order, bit-identical mean / rstd).
struct GnFin {
    float* mean;              // [N][8]; nullptr: no fused finalize (the caller runs b200_gn_finalize)
    float* rstd;              // [N][8]
    unsigned int* ticket;     // one zero-initialised word
    double count;             // elements per (sample, group)
    float eps;
};
// partial: [ctas][N][16] (sum[8] | sumsq[8]); scratch: >= 256 doubles of shared memory; all threads of a CTA with
// blockDim.x >= 256.
__device__ __noinline__ void gn_stats_finalize_cta(const float* partial, int ctas, int N, const GnFin& f, double* scratch) {
    const int t = threadIdx.x;
    for (int n = 0; n < N; ++n) {
        if (t < 256) {
            const int o = t >> 4, j = t & 15;
            double a = 0.0;
#pragma unroll 4
            for (int c = j; c < ctas; c += 16) a += (double)__ldcg(partial + ((size_t)c * N + n) * 16 + o);
            scratch[t] = a;
        }
        __syncthreads();
        if (t < 8) {
            double s = 0.0, q = 0.0;
            for (int k = 0; k < 16; ++k) s += scratch[t * 16 + k];
            for (int k = 0; k < 16; ++k) q += scratch[(8 + t) * 16 + k];
            const double m = s / f.count;
            double var = q / f.count - m * m;
            if (var < 0.0) var = 0.0;
            f.mean[n * 8 + t] = (float)m;
            f.rstd[n * 8 + t] = (float)(1.0 / sqrt(var + (double)f.eps));
        }
        __syncthreads();
    }
    if (t == 0) *f.ticket = 0u;
}

}  // namespace b200

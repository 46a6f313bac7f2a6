// Marching weight-gradient GEMM for the 3x3x3 16 <-> 16 channel convs of level 0 (model.py:72-73, 336, 348;
// the weight half of aten::convolution_backward, train.py:210): six launches per training step, the largest
// single item of the backward pass.
//
// STATUS: correct (tests/gpu_opcheck.py checks it against torch autograd next to the linear-row kernel) but OPT-IN
// (B200_WGRAD_MARCH=1): it removes the L2 -> SMEM ingest bound, not the MMA-issue bound, and measured slower
// (profiles/r01_ab_wgrad_march.txt).  Two earlier variants of this file folded kw into M (three row-shifted copies
// of dY) and (kd, kh) into N = 144 - one MMA per 16 voxels, MMA time 68 us - but needed 75 KB of operands per
// 2400-cycle step, more than a CTA can keep in flight (155 us).
//
//   dW[kd][kh][kw][co][ci] = sum_{n,d,h,w} dY[d][h][w][co] * X[d+kd-1][h+kh-1][w+kw-1][ci]
//
// The linear-row form in wgrad_gemm.cuh stacks 3 slice-shifted copies of dY in M and 3 line-shifted copies of X in
// N: 12 planes cross L2 -> shared memory per row range, 841 MB per launch at 2 x 128^3, which is the bandwidth the
// bulk-copy path sustains (~6.5 TB/s) for the whole 124 us - the kernel is bound by that ingest and starves the
// memory-bound GroupNorm-backward kernels it is meant to overlap with.  Here a CTA owns a band of BH lines and
// marches along D, and every byte is loaded once (plus the band's two halo lines):
//   * dY slices live in a 3-deep ring in shared memory, [slot][chunk][BH lines x (W+2) rows]: the blocks
//     (slot, chunk) sit at one uniform stride, so ONE MN-major A operand (M = 48, issued as 64) covers the three
//     kd taps - the slices d-1, d, d+1 of the gradient that meet slice d of the input;
//   * the band's BH + 2 lines of X slice d are streamed line-interleaved, [line][chunk][(W+2) rows], so that the
//     blocks (kh, chunk) of three neighbouring lines also sit at one uniform stride: a B operand of N = 48;
//   * the kw taps are three accumulators whose B operand starts one row apart.
// 340 MB instead of 841 MB cross L2 -> SMEM; the MMA count is the same as before (3 per 16 voxels, minus the padded
// rows the linear form also multiplies).  The ring slot of slice d+1-kd rotates with d, so the accumulators come in
// three sets selected by the slot of the centre slice (3 sets x 3 kw x 48 = 432 TMEM columns); the reduce kernel
// undoes the rotation.  K runs over the W interior voxels of a line (W % 16 == 0; dY is zero on the halo).
// A 3-slot ring has no prefetch lead (slice d-1 is read until the last MMA of step d, slice d+2 is needed by the
// first MMA of step d+1), so the bulk copies land in a double-buffered staging area a step ahead and the four
// epilogue warps - idle until the end - copy each slice into its ring slot (17 KB, shared to shared) the moment
// the slot is released; the band is split in two half bands with a ring each, so that the copy for one half hides
// behind the MMAs of the other.
// Each CTA writes one fp32 partial [9 accumulators][64][48]; wgrad_march_reduce_kernel sums CTAs and sets in a
// fixed order (deterministic) into the PyTorch (Cout, Cin, 3, 3, 3) gradient.
#pragma once
#include "conv_march.cuh"

namespace b200 {

constexpr int kWgmThreads = 256;
constexpr int kWgmN = 48;           // (kh 3) x 16 ci
constexpr int kWgmM = 64;           // (ring slot 3) x 16 co, padded to the 64-row MMA
constexpr int kWgmAccs = 9;         // (set 3) x (kw 3)
constexpr int kWgmXStages = 2;

struct WgradMarchParams {
    int N, D, H, W, Wp;
    int BH, n_bands;
    long long units;                // N * n_bands * D, slice fastest
    int ksteps;                     // W / 16
    int BHs;                        // lines per half band (BH = 2 * BHs, or BHs = BH when BH is odd / 1)
    int nsub;                       // 1 or 2 half bands, each with its own ring: the refill of one hides behind the MMAs of the other
    unsigned y_plane_bytes;         // BH * Wp * 16: one chunk of the band's lines of one slice (staging layout [chunk][BH lines])
    unsigned y_sub_bytes;           // BHs * Wp * 16: one chunk of a half band (ring block stride)
    unsigned y_slot_bytes;          // 2 planes
    unsigned x_blk_bytes;           // Wp * 16: one (line, chunk) block
    unsigned x_stage_bytes;         // (BH + 2) * 2 blocks
    unsigned smem_y_off, smem_stg_off, smem_x_off, smem_bar_off;
    ActRef dy, x;
    float* partial;                 // [cta][9][64][48]
    int debug;
};

struct WgmSeg {
    int n, band, d0, d1;            // interior slice range [d0, d1] of sample n, band of lines
};
__device__ __forceinline__ int wgm_segment(const WgradMarchParams& p, long long u, long long u_end, WgmSeg& s) {
    const int d0 = (int)(u % p.D);
    long long t = u / p.D;
    s.band = (int)(t % p.n_bands);
    s.n = (int)(t / p.n_bands);
    s.d0 = d0;
    long long len = p.D - d0;
    if (len > u_end - u) len = u_end - u;
    s.d1 = d0 + (int)len - 1;
    return (int)len;
}

__global__ void __launch_bounds__(kWgmThreads, 1)
wgrad_march_kernel(const __grid_constant__ WgradMarchParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x, ctas = gridDim.x;
    const long long u_begin = p.units * cta / ctas, u_end = p.units * (cta + 1) / ctas;

    uint8_t* smem_y = smem + p.smem_y_off;               // dY rings: [half band][slot 3][chunk 2][BHs*Wp rows]
    uint8_t* smem_stg = smem + p.smem_stg_off;           // 2 staging buffers of one slot each (follow the ring: the
                                                         // 64-row A operand reads one slot past the ring - finite data)
    uint8_t* smem_x = smem + p.smem_x_off;               // X stages: [line][chunk][Wp rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* y_full = bars;                              // [sub 2][slot 3]  fillers -> MMA (128 arrivals)
    uint64_t* y_empty = bars + 6;                         // [sub 2][slot 3]  MMA commit -> fillers
    uint64_t* x_full = bars + 12;                         // [2]  bulk copies -> MMA
    uint64_t* x_empty = bars + 14;                        // [2]  MMA commit -> X producer
    uint64_t* done = bars + 16;
    uint64_t* stg_full = bars + 17;                       // [2]  bulk copies -> fillers
    uint64_t* stg_empty = bars + 19;                      // [2]  fillers (128 arrivals) -> dY producer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 6; ++i) { mbar_init(&y_full[i], 128); mbar_init(&y_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
            mbar_init(&stg_full[i], 1); mbar_init(&stg_empty[i], 128);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 4) {                 // every MMA accumulates: the nine accumulators start at zero
        const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (uint32_t c = 0; c < (uint32_t)(kWgmAccs * kWgmN); c += 16) tmem_st16_zero(tl + c);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long SS = (long long)(p.H + 2) * p.Wp;     // rows per padded slice

    if (warp == 0) {
        // ============ X producer: lines band*BH .. +BH+1 (padded) of the step's slice, line-interleaved ============
        long long i = 0;
        for (long long u = u_begin; u < u_end;) {
            WgmSeg sg;
            u += wgm_segment(p, u, u_end, sg);
            for (int dxp = sg.d0 + 1; dxp <= sg.d1 + 1; ++dxp, ++i) {
                const int st = (int)(i % kWgmXStages);
                mbar_wait(&x_empty[st], (uint32_t)(((i / kWgmXStages) & 1) ^ 1));
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[st], p.x_stage_bytes);
                    const long long row0 = ((long long)sg.n * (p.D + 2) + dxp) * SS + (long long)(sg.band * p.BH) * p.Wp;
                    const uint8_t* s0 = reinterpret_cast<const uint8_t*>(p.x.at(0, row0));
                    const uint8_t* s1 = reinterpret_cast<const uint8_t*>(p.x.at(1, row0));
                    uint8_t* dst = smem_x + (size_t)st * p.x_stage_bytes;
                    for (int ln = 0; ln < p.BH + 2; ++ln) {
                        bulk_load_1d(dst, s0, p.x_blk_bytes, &x_full[st]);
                        bulk_load_1d(dst + p.x_blk_bytes, s1, p.x_blk_bytes, &x_full[st]);
                        dst += 2 * p.x_blk_bytes; s0 += p.x_blk_bytes; s1 += p.x_blk_bytes;
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ============ dY producer: the band's BH interior lines of one padded slice into a staging buffer ============
        long long j = 0;
        for (long long u = u_begin; u < u_end;) {
            WgmSeg sg;
            u += wgm_segment(p, u, u_end, sg);
            // X slices d0+1 .. d1+1 (padded) meet dY slices d0 .. d1+2 (padded)
            for (int dyp = sg.d0; dyp <= sg.d1 + 2; ++dyp, ++j) {
                const int b = (int)(j & 1);
                mbar_wait(&stg_empty[b], (uint32_t)(((j >> 1) & 1) ^ 1));
                if (elect_one()) {
                    mbar_arrive_expect_tx(&stg_full[b], p.y_slot_bytes);
                    const long long row0 = ((long long)sg.n * (p.D + 2) + dyp) * SS + (long long)(sg.band * p.BH + 1) * p.Wp;
                    uint8_t* dst = smem_stg + (size_t)b * p.y_slot_bytes;
                    bulk_load_1d(dst, p.dy.at(0, row0), p.y_plane_bytes, &stg_full[b]);
                    bulk_load_1d(dst + p.y_plane_bytes, p.dy.at(1, row0), p.y_plane_bytes, &stg_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer ============
        constexpr uint32_t idesc = make_idesc(kWgmM, kWgmN, 1, 1);
        // MN-major SWIZZLE_NONE: LBO = 128 B between 8-row K groups, SBO = stride between 8-channel blocks
        // (A: blocks (slot, chunk) of one half band's ring; B: blocks (line, chunk) of the X stage)
        const uint64_t a_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((p.y_sub_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((p.x_blk_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
        const uint32_t ybase16 = smem_u32(smem_y) >> 4, xbase16 = smem_u32(smem_x) >> 4;
        const uint32_t xstage16 = p.x_stage_bytes >> 4;
        const uint32_t ysubring16 = (6u * p.y_sub_bytes) >> 4;      // one half band's ring: 3 slots x 2 chunks
        const uint32_t yline16 = (uint32_t)p.Wp;                    // one line = Wp rows of 16 B
        const uint32_t xline16 = (uint32_t)(2 * p.Wp);              // one line of an X stage = 2 blocks
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, tb = 0, w_wait = 0, t_issue = 0, nsteps = 0;
        MARCH_PROF_T(tb);
        long long j = 0, i = 0;                                     // dY refills / X stages consumed so far
        for (long long u = u_begin; u < u_end;) {
            WgmSeg sg;
            u += wgm_segment(p, u, u_end, sg);
            const int nl = min(p.BH, p.H - sg.band * p.BH);         // lines of this band inside the volume
            for (int dxp = sg.d0 + 1; dxp <= sg.d1 + 1; ++dxp, ++i) {
                const int st = (int)(i % kWgmXStages);
                const uint32_t rho = (uint32_t)((j + 1) % 3);       // ring slot of the centre slice
                const uint32_t dset = tmem_base + rho * 3u * (uint32_t)kWgmN;
                for (int sub = 0; sub < p.nsub; ++sub) {
                    MARCH_PROF_T(t0);
                    // dY slices dxp-1, dxp, dxp+1 are refills j, j+1, j+2; all but the last were awaited by earlier steps
                    uint64_t* yf = y_full + sub * 3;
                    if (dxp == sg.d0 + 1) {
                        mbar_wait(&yf[(int)(j % 3)], (uint32_t)((j / 3) & 1));
                        mbar_wait(&yf[(int)((j + 1) % 3)], (uint32_t)(((j + 1) / 3) & 1));
                    }
                    mbar_wait(&yf[(int)((j + 2) % 3)], (uint32_t)(((j + 2) / 3) & 1));
                    if (sub == 0) mbar_wait(&x_full[st], (uint32_t)((i / kWgmXStages) & 1));
                    MARCH_PROF_T(t1);
                    w_wait += t1 - t0;
                    tc_fence_after();
                    const int l0 = sub * p.BHs, l1 = min(nl, l0 + p.BHs);
                    if (elect_one()) {
                        uint32_t a16 = ybase16 + (uint32_t)sub * ysubring16 + 1;                       // dY row wp = 1 of the half band's line 0
                        uint32_t b16 = xbase16 + (uint32_t)st * xstage16 + (uint32_t)l0 * xline16;     // X row wp = 0 of stage line l0
                        for (int l = l0; l < l1; ++l) {
#pragma unroll
                            for (int t = 0; t < 3; ++t) {
                                uint32_t a = a16, b = b16 + (uint32_t)t;
                                const uint32_t dtm = dset + (uint32_t)(t * kWgmN);
                                for (int ks = 0; ks < p.ksteps; ++ks) {
                                    umma_bf16(dtm, a_hi | (uint64_t)(a & 0x3FFF), b_hi | (uint64_t)(b & 0x3FFF), idesc, 1u);
                                    a += 16; b += 16;                            // 16 rows of 16 B
                                }
                            }
                            a16 += yline16; b16 += xline16;
                        }
                        uint64_t* ye = y_empty + sub * 3;
                        umma_commit(&ye[(int)(j % 3)]);                          // slice dxp-1 is not needed again
                        if (dxp == sg.d1 + 1) {                                  // end of the segment: release the other two
                            umma_commit(&ye[(int)((j + 1) % 3)]);
                            umma_commit(&ye[(int)((j + 2) % 3)]);
                        }
                        if (sub == p.nsub - 1) umma_commit(&x_empty[st]);
                    }
                    __syncwarp();
                    MARCH_PROF_T(t2);
                    t_issue += t2 - t1;
                }
                ++nsteps;
                ++j;
            }
            j += 2;                                                 // the segment loaded two slices more than it has steps
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
        if (prof && lane == 0 && cta < 160) {
            g_march_prof[cta * 16 + 3] = (unsigned long long)(clock64() - tb);
            g_march_prof[cta * 16 + 4] = (unsigned long long)w_wait;
            g_march_prof[cta * 16 + 5] = (unsigned long long)t_issue;
            g_march_prof[cta * 16 + 6] = (unsigned long long)nsteps;
        }
    } else if (warp >= 4) {
        // ---- main loop: copy every staged dY slice into its ring slots as soon as they are released ----
        {
            const int f = threadIdx.x - 128;
            const int sub_vecs = (int)(p.y_sub_bytes >> 4);           // 16-byte vectors of one chunk of a half band
            const int plane_vecs = (int)(p.y_plane_bytes >> 4);
            long long j = 0;
            for (long long u = u_begin; u < u_end;) {
                WgmSeg sg;
                u += wgm_segment(p, u, u_end, sg);
                for (int dyp = sg.d0; dyp <= sg.d1 + 2; ++dyp, ++j) {
                    const int b = (int)(j & 1), slot = (int)(j % 3);
                    mbar_wait(&stg_full[b], (uint32_t)((j >> 1) & 1));
                    const uint4* src = reinterpret_cast<const uint4*>(smem_stg + (size_t)b * p.y_slot_bytes);
                    for (int sub = 0; sub < p.nsub; ++sub) {
                        mbar_wait(&y_empty[sub * 3 + slot], (uint32_t)(((j / 3) & 1) ^ 1));
                        // ring: [sub][slot][chunk][BHs lines]; staging: [chunk][BH lines]
                        uint4* dst = reinterpret_cast<uint4*>(smem_y) + (size_t)((sub * 3 + slot) * 2) * sub_vecs;
                        for (int idx = f; idx < 2 * sub_vecs; idx += 128) {
                            const int c = idx >= sub_vecs ? 1 : 0;
                            const int r = idx - c * sub_vecs;
                            dst[idx] = src[c * plane_vecs + sub * sub_vecs + r];
                        }
                        fence_proxy_async_smem();                         // generic writes -> visible to the tensor core
                        mbar_arrive(&y_full[sub * 3 + slot]);
                    }
                    mbar_arrive(&stg_empty[b]);
                }
            }
        }
        // ============ epilogue: TMEM -> fp32 partial [9][64][48] ============
        const int ew = warp - 4;
        float* dst = p.partial + (size_t)cta * kWgmAccs * kWgmM * kWgmN;
        const int row = ew * 16 + lane;                   // M = 64: rows 16*ew .. +15 live in lanes 0-15 of quadrant ew
        const bool row_ok = lane < 16;
        mbar_wait(done, 0);
        tc_fence_after();
        for (int acc = 0; acc < kWgmAccs; ++acc) {
#pragma unroll
            for (int c0 = 0; c0 < kWgmN; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kWgmN + c0), v);
                if (row_ok) {
                    float4* o = reinterpret_cast<float4*>(dst + ((size_t)acc * kWgmM + row) * kWgmN + c0);
                    o[0] = make_float4(v[0], v[1], v[2], v[3]);
                    o[1] = make_float4(v[4], v[5], v[6], v[7]);
                    o[2] = make_float4(v[8], v[9], v[10], v[11]);
                    o[3] = make_float4(v[12], v[13], v[14], v[15]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, 512);
}

// dW[kd][kh][kw][co][ci] = sum over CTAs and sets rho of P[cta][rho*3 + kw][s*16 + co][kh*16 + ci] with
// s = (rho + 4 - kd) % 3: the ring slot that held gradient slice d + 1 - kd when the centre slice sat in slot rho.
// One thread per (output float4 over ci, CTA group); groups are combined in a fixed order in shared memory.
struct WgmReduceParams {
    int ctas, Cout_w, Cin_w, accumulate;
};
constexpr int kWgmReduceGroups = 16;
__global__ void __launch_bounds__(256)
wgrad_march_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, WgmReduceParams q) {
    __shared__ double s_acc[256][4];
    constexpr int QPB = 256 / kWgmReduceGroups;           // output quads per CTA
    const int g = threadIdx.x / QPB, ql = threadIdx.x % QPB;
    const int quad = blockIdx.x * QPB + ql;               // (tap, co, ci/4): 27 * 16 * 4 quads
    const bool active = quad < 27 * 16 * 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int tap = 0, co = 0, ci0 = 0;
    if (active) {
        ci0 = (quad & 3) * 4; co = (quad >> 2) & 15; tap = quad >> 6;
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        for (int c = g; c < q.ctas; c += kWgmReduceGroups) {
#pragma unroll
            for (int rho = 0; rho < 3; ++rho) {
                const int s = (rho + 4 - kd) % 3;
                const float4 v = *reinterpret_cast<const float4*>(
                    partial + (((size_t)c * kWgmAccs + rho * 3 + kw) * kWgmM + s * 16 + co) * kWgmN + kh * 16 + ci0);
                a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            }
        }
    }
    s_acc[threadIdx.x][0] = a0; s_acc[threadIdx.x][1] = a1; s_acc[threadIdx.x][2] = a2; s_acc[threadIdx.x][3] = a3;
    __syncthreads();
    if (g != 0 || !active) return;
    for (int k = 1; k < kWgmReduceGroups; ++k) {
        a0 += s_acc[k * QPB + ql][0]; a1 += s_acc[k * QPB + ql][1];
        a2 += s_acc[k * QPB + ql][2]; a3 += s_acc[k * QPB + ql][3];
    }
    if (co >= q.Cout_w) return;
    const double a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int ci = ci0 + e;
        if (ci >= q.Cin_w) continue;
        const size_t o = ((size_t)co * q.Cin_w + ci) * 27 + tap;
        if (q.accumulate) grad[o] += (float)a[e];
        else grad[o] = (float)a[e];
    }
}

}  // namespace b200

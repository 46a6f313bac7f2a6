// "Marching" implicit-GEMM 3x3x3 convolution for small channel counts (Cin, Cout <= 32) on tcgen05.
//
// Same role as conv_gemm.cuh (aten::convolution k3 p1 of model.py:72-73, 336, 348 and the data half
// of its backward, train.py:210), specialised for the two fine levels of the U-Net where 75 % of the
// FLOPs live and where a 128 x Cout x 16 MMA leaves the tensor pipe idle (an MMA costs
// max(46, (A bytes + B bytes) / 128) cycles: operand fetch from shared memory, not math, is the limit).
//
// Two ideas:
//  * kd-fold with accumulation ACROSS input slices in TMEM.  For one input slice s the MMA computes,
//    for all three kd taps at once (GEMM N = 3*Cout),
//        Q_kd[s][r] = sum_{kh,kw,ci} X[s][r + (kh-1)Wp + (kw-1)][ci] * W[kd,kh,kw][ci][co]
//    and  out[o] = Q_0[o-1] + Q_1[o] + Q_2[o+1].  Each 128-row block owns three TMEM accumulator slots
//    of Cout columns; the slot of output slice o is (o+1) mod 3 (padded slice index mod 3).  At input
//    slice s the three column blocks of the MMA's D tile land in the slots of o = s+1, s, s-1 - always
//    all three slots, so D is one contiguous column range; which kd goes to which slot rotates with
//    s mod 3, and that rotation is a start-address offset into a 5-band weight image [kd2|kd1|kd0|kd2|kd1].
//    After input slice s the slot of o = s-1 is complete: the epilogue drains it, writes zeros back and
//    hands it to o = s+2.  No shuffles, no partial-sum exchange: the epilogue reads Cout columns per row.
//  * marching along D.  A CTA owns a strip of MB*128 consecutive in-slice rows and walks the slices in
//    order, so every input slice of the strip is loaded ONCE (a ring of shared-memory slots filled by
//    cp.async.bulk), instead of (BD+2)/BD times by slice-group tiles.  Work units (sample, strip, slice)
//    are split into equal contiguous ranges over the CTAs; a range that crosses into another strip
//    starts a new segment (two extra input slices).
//
// Warp roles: w0 activation producer (bulk copies), w1 weight loader (once), w2 MMA issuer,
// w3 TMEM allocator, w4.. epilogue groups of 4 warps (group g takes blocks b = g, g+G, ...).
#pragma once
#include "common.cuh"
#include "conv_gemm.cuh"

namespace b200 {

constexpr int kMarchEpiGroups = 2;
constexpr int kMarchThreads = 128 + 128 * kMarchEpiGroups;
constexpr int kMarchMaxMB = 10;
constexpr int kMarchMaxSlots = 6;
constexpr unsigned kMarchTailBytes = 2048;   // barriers + TMEM slot + GroupNorm partials

struct MarchParams {
    int N, D, H, W, Wp, SS;
    FastDiv by_Wp;
    // tiling
    int MB, TR;            // blocks per strip, rows per strip (128*MB)
    int Q0, QN;            // first interior in-slice row (Wp+1) and number of rows up to the last interior one
    int n_strips;          // strips per slice
    long long units;       // N * n_strips * D work units (sample, strip, output slice), slice fastest
    int KS;                // Cin / 16 (K steps per tap)
    int SRp;               // rows per shared-memory slot plane (TR + 2*Wp + 2, rounded up to 8)
    int nslots;
    unsigned plane_bytes, slot_bytes, w_bytes, wtile_bytes;
    unsigned smem_x_off, smem_w_off, smem_bar_off;
    unsigned tmem_cols;
    // operands
    ActRef src;
    const __nv_bfloat16* wpacked;     // [KS][9 taps (kh,kw)][2 k-chunks][5 bands * CO][8]
    // epilogue
    int lrelu_out;
    ActRef out, residual;
    float* stats_partial;             // optional [ctas][N][16]
    GnFin gn_fin;                     // mean != nullptr: the last CTA turns stats_partial into mean / rstd (common.cuh)
    ActRef gnb_x;                     // != null: stats_partial receives GroupNorm-BACKWARD sums (common.cuh gnb_accumulate) of
    const float* gnb_coef;            //          the GroupNorm whose conv output is gnb_x; coef = [N][3][CO]
    const float* bias;                // EPI_SIGMOID
    float* probs;
    float* logits;
    int n_out_real;
    int debug;                        // perf probes: 1 = skip activation loads, 2 = skip MMAs
};

// In-kernel cycle accounting for the perf probes (debug bit 256): [cta][16] counters, see tools/perf_probe.py.
__device__ unsigned long long g_march_prof[160 * 16];
#define MARCH_PROF_T(var) if (prof) var = clock64()

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One segment = a run of consecutive output slices [d0, d1] of one (sample, strip).
struct MarchSeg {
    int n, strip, d0, d1;
};
// Units [u, u_end) -> first segment starting at u.  Returns the number of units consumed.
__device__ __forceinline__ int march_segment(const MarchParams& p, long long u, long long u_end, MarchSeg& s) {
    const int d0 = (int)(u % p.D);
    const long long t = u / p.D;
    s.strip = (int)(t % p.n_strips);
    s.n = (int)(t / p.n_strips);
    s.d0 = d0;
    long long len = p.D - d0;
    if (len > u_end - u) len = u_end - u;
    s.d1 = d0 + (int)len - 1;
    return (int)len;
}

template <int CO, int EPI>
__global__ void __launch_bounds__(kMarchThreads, 1)
conv_march_kernel(const __grid_constant__ MarchParams p) {
    constexpr int NM = 3 * CO;                 // GEMM N: three kd taps side by side
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x, ctas = gridDim.x;
    const long long u_begin = p.units * cta / ctas, u_end = p.units * (cta + 1) / ctas;

    uint8_t* smem_x = smem + p.smem_x_off;
    uint8_t* smem_w = smem + p.smem_w_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* x_full = bars;                               // [kMarchMaxSlots]
    uint64_t* x_empty = x_full + kMarchMaxSlots;           // [kMarchMaxSlots]
    uint64_t* acc_done = x_empty + kMarchMaxSlots;         // [kMarchMaxMB]  MMA -> epilogue
    uint64_t* acc_free = acc_done + kMarchMaxMB;           // [kMarchMaxMB]  epilogue -> MMA
    uint64_t* w_full = acc_free + kMarchMaxMB;             // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);   // [4*kMarchEpiGroups warps][16]

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nslots; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < p.MB; ++i) { mbar_init(&acc_done[i], 1); mbar_init(&acc_free[i], 128); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // all accumulators start at zero (every MMA accumulates)
    if (warp >= 4 && warp < 8) {
        const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (uint32_t c = 0; c < p.tmem_cols; c += 16) tmem_st16_zero(tl + c);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();


    // Programmatic dependent launch: everything above (and the weight copies of warp 1, which never touch a tensor of an
    // earlier kernel) may overlap the tail of the stream predecessor; every other warp waits for it here.
    if (warp != 1) pdl_wait();
    // GroupNorm partial sums [ctas][N][16]: this CTA's row starts at zero (samples it never touches must contribute
    // nothing); the epilogue (the same warp among others) overwrites the entries of the samples it does touch.
    if (p.stats_partial && warp == 4) {
        for (int i = lane; i < p.N * 16; i += 32) p.stats_partial[(size_t)cta * p.N * 16 + i] = 0.f;
        __syncwarp();
    }
    if (warp == 0) {
        // ================= activation producer =================
        int xs = 0; uint32_t xph = 0;
        const uint32_t bytes = (uint32_t)p.SRp * 16;
        const int chunks = p.KS * 2;
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, w_empty = 0, tb = 0;
        MARCH_PROF_T(tb);
        for (long long u = u_begin; u < u_end;) {
            MarchSeg sg;
            u += march_segment(p, u, u_end, sg);
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi) {
                MARCH_PROF_T(t0);
                mbar_wait(&x_empty[xs], xph ^ 1);
                MARCH_PROF_T(t1);
                w_empty += t1 - t0;
                if (elect_one()) {
                    if (B200_DBG(p, 1)) {
                        mbar_arrive(&x_full[xs]);
                    } else {
                        mbar_arrive_expect_tx(&x_full[xs], bytes * chunks);
                        const long long row0 = ((long long)sg.n * (p.D + 2) + dpi) * p.SS + (long long)sg.strip * p.TR;
                        uint8_t* dst = smem_x + (size_t)xs * p.slot_bytes;
                        for (int c = 0; c < chunks; ++c)
                            bulk_load_1d(dst + (size_t)c * p.plane_bytes, p.src.at(c, row0), bytes, &x_full[xs]);
                    }
                }
                __syncwarp();
                if (++xs == p.nslots) { xs = 0; xph ^= 1; }
            }
        }
        if (prof && lane == 0 && cta < 160) {
            g_march_prof[cta * 16 + 0] = (unsigned long long)(clock64() - tb);
            g_march_prof[cta * 16 + 1] = (unsigned long long)w_empty;
        }
    } else if (warp == 1) {
        // ================= weights: resident for the whole kernel =================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, p.w_bytes);
            // <= 64 KB per bulk copy keeps each transaction well inside the tx-count range
            for (unsigned off = 0; off < p.w_bytes; off += 32768) {
                const unsigned n = p.w_bytes - off < 32768 ? p.w_bytes - off : 32768;
                bulk_load_1d(smem_w + off, reinterpret_cast<const uint8_t*>(p.wpacked) + off, n, w_full);
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc(128, NM, 0, 0);
        const uint32_t plane16 = p.plane_bytes >> 4;
        // K-major SWIZZLE_NONE: LBO = stride between the two 8-channel K chunks, SBO = 128 B per 8 rows
        const uint64_t a_hi = ((uint64_t)(plane16 & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)((5 * CO * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t xbase16 = smem_u32(smem_x) >> 4, wbase16 = smem_u32(smem_w) >> 4;
        const uint32_t slot16 = p.slot_bytes >> 4, wtile16 = p.wtile_bytes >> 4;
        const bool skip_mma = B200_DBG(p, 2);
        const int MB = p.MB, KS = p.KS;
        uint32_t tof[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) tof[t] = (uint32_t)((t / 3) * p.Wp + (t % 3));
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, w_x = 0, w_acc = 0, t_issue = 0, tb = 0, nsteps = 0;
        MARCH_PROF_T(tb);
        mbar_wait(w_full, 0);
        tc_fence_after();
        int xs = 0; uint32_t xph = 0;
        uint32_t k = 0;
        for (long long u = u_begin; u < u_end;) {
            MarchSeg sg;
            u += march_segment(p, u, u_end, sg);
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi, ++k) {
                MARCH_PROF_T(t0);
                mbar_wait(&x_full[xs], xph);
                MARCH_PROF_T(t1);
                w_x += t1 - t0;
                ++nsteps;
                tc_fence_after();
                const uint32_t rot = (uint32_t)((4 - dpi % 3) % 3);
                const uint32_t xst16 = xbase16 + xs * slot16;
                const uint32_t wst16 = wbase16 + rot * CO;           // band offset: CO rows * 16 B
                for (int b = 0; b < MB; ++b) {
                    MARCH_PROF_T(t2);
                    mbar_wait(&acc_free[b], (k & 1) ^ 1);
                    MARCH_PROF_T(t3);
                    w_acc += t3 - t2;
                    tc_fence_after();
                    if (!skip_mma && elect_one()) {
                        const uint32_t dtm = tmem_base + (uint32_t)(b * NM);
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint32_t a_blk16 = xst16 + (uint32_t)(ks * 2) * plane16 + (uint32_t)(b * 128);
                            const uint32_t w_ks16 = wst16 + (uint32_t)(ks * 9) * wtile16;
#pragma unroll
                            for (int t = 0; t < 9; ++t)
                                umma_bf16(dtm, a_hi | (uint64_t)((a_blk16 + tof[t]) & 0x3FFF),
                                          b_hi | (uint64_t)((w_ks16 + (uint32_t)t * wtile16) & 0x3FFF), idesc, 1u);
                        }
                    }
                    __syncwarp();
                    if (elect_one()) umma_commit(&acc_done[b]);
                    __syncwarp();
                    MARCH_PROF_T(t2);
                    t_issue += t2 - t3;
                }
                if (elect_one()) umma_commit(&x_empty[xs]);
                __syncwarp();
                if (++xs == p.nslots) { xs = 0; xph ^= 1; }
            }
        }
        if (prof && lane == 0 && cta < 160) {
            g_march_prof[cta * 16 + 2] = (unsigned long long)(clock64() - tb);
            g_march_prof[cta * 16 + 3] = (unsigned long long)w_x;
            g_march_prof[cta * 16 + 4] = (unsigned long long)w_acc;
            g_march_prof[cta * 16 + 5] = (unsigned long long)t_issue;
            g_march_prof[cta * 16 + 6] = (unsigned long long)nsteps;
        }
        // every MMA of this CTA is issued: what is left is its last epilogue - the dependent kernel may start launching
        // (it still waits for this whole grid in its own pdl_wait; triggering at the start would park its CTAs here for
        // the whole kernel)
        pdl_trigger();
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int grp = (warp - 4) >> 2;
        const int ew = warp & 3;             // TMEM lane quadrant of this warp
        const int m = ew * 32 + lane;
        constexpr int GS = CO / 8;           // channels per GroupNorm group
        float ssum[8], ssq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        int cur_n = -1;
        const bool do_stats = (EPI == EPI_BF16) && (p.stats_partial != nullptr);
        const bool do_gnb = do_stats && p.gnb_x.base != nullptr;
        const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);

        auto flush_stats = [&](int n) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = warp_sum(ssum[i]); ssq[i] = warp_sum(ssq[i]); }
            const int w8 = warp - 4;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { stat_smem[w8 * 16 + i] = ssum[i]; stat_smem[w8 * 16 + 8 + i] = ssq[i]; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kMarchEpiGroups) : "memory");
            if (w8 == 0 && lane < 16) {
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < 4 * kMarchEpiGroups; ++q) v += stat_smem[q * 16 + lane];
                p.stats_partial[((size_t)cta * p.N + n) * 16 + lane] = v;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kMarchEpiGroups) : "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        };

        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, w_done = 0, t_ld = 0, t_st = 0, t_rest = 0, tb = 0;
        MARCH_PROF_T(tb);
        uint32_t k = 0;
        for (long long u = u_begin; u < u_end;) {
            MarchSeg sg;
            u += march_segment(p, u, u_end, sg);
            if (do_stats && sg.n != cur_n) {
                if (cur_n >= 0) flush_stats(cur_n);
                cur_n = sg.n;
            }
            // which rows of this strip are interior voxels (independent of the slice)
            const int qbase = sg.strip * p.TR + m;        // in-slice row of block 0 (slot row 0 <-> in-slice row strip*TR)
            uint32_t rowmask = 0;
            for (int b = grp; b < p.MB; b += kMarchEpiGroups) {
                const int q = qbase + b * 128 + p.Q0;     // output row (in-slice) = slot row + Wp + 1
                const int hp = p.by_Wp.div(q);
                const int wp = q - hp * p.Wp;
                if (q < p.Q0 + p.QN && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W) rowmask |= 1u << b;
            }
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi, ++k) {
                const int dpo = dpi - 1;                              // padded index of the slice completed by this step
                const uint32_t slot = (uint32_t)((dpi + 2) % 3);      // == dpo mod 3
                const bool out_valid = dpi >= sg.d0 + 2;
                const bool last = dpi == sg.d1 + 2;
                const long long obase = ((long long)sg.n * (p.D + 2) + dpo) * p.SS + p.Q0 + qbase;
                for (int b = grp; b < p.MB; b += kMarchEpiGroups) {
                    MARCH_PROF_T(t0);
                    if (prof && t3) t_rest += t0 - t3;
                    mbar_wait(&acc_done[b], k & 1);
                    MARCH_PROF_T(t1);
                    w_done += t1 - t0;
                    tc_fence_after();
                    const uint32_t tblk = tlane + (uint32_t)(b * NM);
                    uint32_t r[CO];
#pragma unroll
                    for (int c0 = 0; c0 < CO; c0 += 16) tmem_ld16_nowait(tblk + slot * CO + c0, r + c0);
                    tmem_ld_wait();
                    MARCH_PROF_T(t2);
                    t_ld += t2 - t1;
                    // hand the slot(s) back zeroed: the completed one, or all three at the end of a segment
                    if (last) {
#pragma unroll
                        for (int c0 = 0; c0 < NM; c0 += 16) tmem_st16_zero(tblk + c0);
                    } else {
#pragma unroll
                        for (int c0 = 0; c0 < CO; c0 += 16) tmem_st16_zero(tblk + slot * CO + c0);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&acc_free[b]);
                    MARCH_PROF_T(t3);
                    t_st += t3 - t2;
                    if (!out_valid || !((rowmask >> b) & 1u)) continue;
                    const long long orow = obase + b * 128;
                    float v[CO];
#pragma unroll
                    for (int i = 0; i < CO; ++i) v[i] = __uint_as_float(r[i]);
                    if (EPI == EPI_BF16) {
                        if (do_stats && !do_gnb) {
#pragma unroll
                            for (int i = 0; i < CO; ++i) {
                                ssum[i / GS] += v[i];
                                ssq[i / GS] += v[i] * v[i];
                            }
                        }
                        uint4 gq[CO / 8];
                        if (do_gnb) {                    // the saved conv output of the GroupNorm differentiated next
#pragma unroll
                            for (int c = 0; c < CO / 8; ++c) {
                                const void* gp = p.gnb_x.at(c, orow);
                                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(gq[c].x), "=r"(gq[c].y), "=r"(gq[c].z), "=r"(gq[c].w) : "l"(gp));
                            }
                        }
                        if (p.residual.base) {
#pragma unroll
                            for (int c = 0; c < CO / 8; ++c) {
                                float f[8];
                                unpack_bf16x8(ld_nc_v4(p.residual.at(c, orow)), f);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[c * 8 + i] += f[i];
                            }
                        }
                        if (p.lrelu_out) {
#pragma unroll
                            for (int i = 0; i < CO; ++i) v[i] = lrelu(v[i]);
                        }
                        uint4 oq[CO / 8];
#pragma unroll
                        for (int c = 0; c < CO / 8; ++c) {
                            oq[c] = pack_bf16x8(v + c * 8);
                            *reinterpret_cast<uint4*>(p.out.at(c, orow)) = oq[c];
                        }
                        if (do_gnb) gnb_accumulate<CO>(oq, gq, p.gnb_coef + (size_t)sg.n * 3 * CO, CO, ssum, ssq);
                    } else {   // EPI_SIGMOID: first n_out_real columns are real; fp32 NCDHW output
                        const int q = qbase + b * 128 + p.Q0;
                        const int hp = p.by_Wp.div(q);
                        const int wp = q - hp * p.Wp;
                        const size_t plane = (size_t)p.D * p.H * p.W;
                        const size_t vox = ((size_t)(dpo - 1) * p.H + (hp - 1)) * p.W + (wp - 1);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < p.n_out_real) {
                                const float z = v[c] + p.bias[c];
                                const size_t o = ((size_t)sg.n * p.n_out_real + c) * plane + vox;
                                if (p.logits) p.logits[o] = z;
                                p.probs[o] = 1.f / (1.f + expf(-z));
                            }
                        }
                    }
                }
            }
        }
        if (do_stats && cur_n >= 0) flush_stats(cur_n);
        if (prof && lane == 0 && ew == 0 && cta < 160 && grp < 2) {
            const int o = cta * 16 + 7 + grp * 4;
            g_march_prof[o + 0] = (unsigned long long)w_done;
            g_march_prof[o + 1] = (unsigned long long)t_ld;
            g_march_prof[o + 2] = (unsigned long long)t_st;
            g_march_prof[o + 3] = (unsigned long long)t_rest;
            if (grp == 0) g_march_prof[cta * 16 + 15] = (unsigned long long)(clock64() - tb);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, p.tmem_cols);
    // GroupNorm statistics: the CTA that finishes last sums every CTA's partial row (the pipeline's shared memory is idle
    // by now: every bulk copy has been consumed) - no separate finalize launch between this conv and gn_apply.
    if (p.gn_fin.mean != nullptr &&
        cta_draws_last_ticket(p.gn_fin.ticket, gridDim.x, reinterpret_cast<unsigned int*>(smem + 2048)))
        gn_stats_finalize_cta(p.stats_partial, ctas, p.N, p.gn_fin, reinterpret_cast<double*>(smem));
}

}  // namespace b200

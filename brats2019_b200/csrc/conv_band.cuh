// "Band-marching" 3x3x3 convolution for 16 -> 16 channels (level 0 of the U-Net, 44 % of all FLOPs):
// the kd AND kh taps are folded into the GEMM N dimension (N = 9*16 = 144) and combined by accumulation
// in TMEM, so a 128-voxel line segment needs 3 MMAs (the kw taps) instead of 27 / 9, and every MMA runs
// at the full tensor rate (an M=128, K=16 MMA costs max(46, (A+B bytes)/128, N/2) cycles - with N = 48
// the A-operand fetch is the limit, with N = 144 the math is).
//
// Same role as conv_march.cuh (aten::convolution k3 p1, model.py:72-73, 336, 348 and its data gradient).
//
// Geometry.  The M dimension of a tile is a "window": 128 consecutive voxels of one line (padded column
// index wp = 1 + 128*c + m).  A CTA owns a band of BH = 8 output lines x one window column and marches
// along D (one input slice per step, loaded once into a shared-memory ring).  For input line ji of the
// band (ji = 0 .. BH+1, lines hp0-1 .. hp0+BH) and tap kw the MMA computes, for all (kh, kd) at once,
//     Q[kh][kd][m] = sum_ci X[d_in][line ji][m + kw][ci] * W[kd][kh][kw][ci][:]
// which belongs to output line jo = ji + 1 - kh and output slice d_in + 1 - kd.  TMEM columns are
// [line slot jo (BH+2)][kd slot (3)][16]: the three kh blocks of D land in line slots ji-1, ji, ji+1 -
// contiguous columns - and inside each line slot the kd blocks rotate over the three slice slots exactly
// as in conv_march.cuh (slice slot = padded output slice index mod 3).  Because the rotation now happens
// inside each kh block, the weight image is stored three times, once per rotation.  The first and the
// last input line of the band only feed one output line each and use N = 48 MMAs.
// After input slice d_in the slice slot of output slice d_in - 1 is complete in every line slot: the
// epilogue drains 16 columns per voxel, writes zeros back, and the slot is reused for slice d_in + 2.
//
// Warp roles: w0 activation producer, w1 weight loader, w2 MMA issuer, w3 TMEM allocator,
// w4.. three epilogue groups of 4 warps (group g takes line slots 1+g, 1+g+3, ...).
#pragma once
#include "conv_march.cuh"

namespace b200 {

constexpr int kBandEpiGroups = 3;            // measured round 2b: 4 groups (GroupNorm-backward fold compiled out, 84 registers) and
                                             // 5 groups change nothing (5.385-5.424 vs 5.408 ms per step; 0.062 ms alone either way)
constexpr int kBandThreads = 128 + 128 * kBandEpiGroups;
constexpr int kBandBH = 8;                        // output lines per band
constexpr int kBandLines = kBandBH + 2;           // input lines per step == line slots in TMEM
constexpr int kBandLineRows = 130;                // rows of one input line segment (128 + the two kw halos)
constexpr unsigned kBandTailBytes = 2048;

struct BandParams {
    int N, D, H, W, Wp, SS;
    int n_bands, n_cols;          // bands of BH lines per slice, 128-voxel window columns per line
    long long units;              // N * n_bands * n_cols * D, slice fastest
    int nslots;
    unsigned plane_bytes, slot_bytes, w_bytes, wimg_bytes;    // wimg = one (rot, kw) image: 2 k-chunks x 144 rows x 16 B
    unsigned smem_x_off, smem_w_off, smem_bar_off;
    ActRef src;
    const __nv_bfloat16* wpacked;     // [rot 3][kw 3][k-chunk 2][kh idx 3 (kh = 2,1,0)][kd band 3][16 co][8]
    int lrelu_out;
    ActRef out, residual;
    float* stats_partial;
    GnFin gn_fin;                     // mean != nullptr: the last CTA turns stats_partial into mean / rstd (common.cuh)
    ActRef gnb_x;                     // != null: stats_partial receives GroupNorm-BACKWARD sums (common.cuh gnb_accumulate) of
    const float* gnb_coef;            //          the GroupNorm whose conv output is gnb_x; coef = [N][3][16]
    const float* bias;
    float* probs;
    float* logits;
    int n_out_real;
    int debug;
};

struct BandSeg {
    int n, band, col, d0, d1;
};
__device__ __forceinline__ int band_segment(const BandParams& p, long long u, long long u_end, BandSeg& s) {
    const int d0 = (int)(u % p.D);
    long long t = u / p.D;
    s.col = (int)(t % p.n_cols); t /= p.n_cols;
    s.band = (int)(t % p.n_bands);
    s.n = (int)(t / p.n_bands);
    s.d0 = d0;
    long long len = p.D - d0;
    if (len > u_end - u) len = u_end - u;
    s.d1 = d0 + (int)len - 1;
    return (int)len;
}

template <int EPI>
__global__ void __launch_bounds__(kBandThreads, 1)
conv_band_kernel(const __grid_constant__ BandParams p) {
    constexpr int CO = 16;
    constexpr int LS = 3 * CO;                    // TMEM columns per line slot (three slice slots)
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
    uint64_t* acc_done = x_empty + kMarchMaxSlots;         // [kBandLines]  MMA -> epilogue, per line slot
    uint64_t* acc_free = acc_done + kBandLines;            // [kBandLines]  epilogue -> MMA
    uint64_t* w_full = acc_free + kBandLines;              // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);   // [4*kBandEpiGroups warps][16]

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nslots; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < kBandLines; ++i) { mbar_init(&acc_done[i], 1); mbar_init(&acc_free[i], 128 * (kBandBH / 2)); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 4 && warp < 8) {      // all accumulators start at zero (every MMA accumulates)
        const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (uint32_t c = 0; c < 512; c += 16) tmem_st16_zero(tl + c);
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
        // ================= activation producer: BH+2 line segments x 2 chunks per step =================
        int xs = 0; uint32_t xph = 0;
        constexpr uint32_t seg_bytes = kBandLineRows * 16;
        for (long long u = u_begin; u < u_end;) {
            BandSeg sg;
            u += band_segment(p, u, u_end, sg);
            const int hp_first = sg.band * kBandBH;            // padded index of input line ji = 0 (= hp0 - 1)
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi) {
                mbar_wait(&x_empty[xs], xph ^ 1);
                if (elect_one()) {
                    if (B200_DBG(p, 1)) {
                        mbar_arrive(&x_full[xs]);
                    } else {
                        mbar_arrive_expect_tx(&x_full[xs], seg_bytes * 2 * kBandLines);
                        const long long row0 = (((long long)sg.n * (p.D + 2) + dpi) * (p.H + 2) + hp_first) * p.Wp + sg.col * 128;
                        uint8_t* dst = smem_x + (size_t)xs * p.slot_bytes;
                        if (p.Wp == kBandLineRows) {       // the BH+2 line segments are one contiguous range
                            for (int c = 0; c < 2; ++c)
                                bulk_load_1d(dst + (size_t)c * p.plane_bytes, p.src.at(c, row0), seg_bytes * kBandLines, &x_full[xs]);
                        } else {
                            for (int c = 0; c < 2; ++c)
                                for (int ji = 0; ji < kBandLines; ++ji)
                                    bulk_load_1d(dst + (size_t)c * p.plane_bytes + (size_t)ji * seg_bytes,
                                                 p.src.at(c, row0 + (long long)ji * p.Wp), seg_bytes, &x_full[xs]);
                        }
                    }
                }
                __syncwarp();
                if (++xs == p.nslots) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= weights: resident for the whole kernel =================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, p.w_bytes);
            for (unsigned off = 0; off < p.w_bytes; off += 32768) {
                const unsigned n = p.w_bytes - off < 32768 ? p.w_bytes - off : 32768;
                bulk_load_1d(smem_w + off, reinterpret_cast<const uint8_t*>(p.wpacked) + off, n, w_full);
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc144 = make_idesc(128, 144, 0, 0);
        constexpr uint32_t idesc48 = make_idesc(128, 48, 0, 0);
        const uint32_t plane16 = p.plane_bytes >> 4;
        const uint64_t a_hi = ((uint64_t)(plane16 & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)144 << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);   // LBO = 144 rows * 16 B
        const uint32_t xbase16 = smem_u32(smem_x) >> 4, wbase16 = smem_u32(smem_w) >> 4;
        const uint32_t slot16 = p.slot_bytes >> 4, wimg16 = p.wimg_bytes >> 4;
        const bool skip_mma = B200_DBG(p, 2);
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, w_x = 0, w_acc = 0, t_issue = 0, tb = 0, nsteps = 0;
        unsigned long long gt0 = 0;
        MARCH_PROF_T(tb);
        if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
        mbar_wait(w_full, 0);
        tc_fence_after();
        int xs = 0; uint32_t xph = 0;
        uint32_t k = 0;
        for (long long u = u_begin; u < u_end;) {
            BandSeg sg;
            u += band_segment(p, u, u_end, sg);
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi, ++k) {
                MARCH_PROF_T(t0);
                mbar_wait(&x_full[xs], xph);
                MARCH_PROF_T(t1);
                w_x += t1 - t0;
                ++nsteps;
                tc_fence_after();
                const uint32_t rot = (uint32_t)((4 - dpi % 3) % 3);
                const uint32_t xst16 = xbase16 + xs * slot16;
                const uint32_t wrot16 = wbase16 + rot * 3 * wimg16;
                // The previous step's accumulators are handed back in two halves (line slots 1-4 and 5-8):
                // two barrier waits and two elected issue blocks per step instead of one per line (a barrier
                // wait costs ~100 cycles even when it has already completed; the 30 MMAs of a step take ~2000).
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    MARCH_PROF_T(t2);
                    mbar_wait(&acc_free[half], (k & 1) ^ 1);
                    MARCH_PROF_T(t3);
                    w_acc += t3 - t2;
                    tc_fence_after();
                    if (!skip_mma && elect_one()) {
#pragma unroll
                        for (int ji = (half == 0 ? 0 : 4); ji < (half == 0 ? 4 : kBandLines); ++ji) {
                            const uint32_t a16 = xst16 + (uint32_t)(ji * kBandLineRows);
                            if (ji == 0) {                 // line above the band: kh = 0 only -> line slot 1
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw)
                                    umma_bf16(tmem_base + LS, a_hi | (uint64_t)((a16 + kw) & 0x3FFF),
                                              b_hi | (uint64_t)((wrot16 + kw * wimg16 + 2 * LS) & 0x3FFF), idesc48, 1u);
                            } else if (ji == kBandLines - 1) {   // line below the band: kh = 2 only -> line slot BH
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw)
                                    umma_bf16(tmem_base + kBandBH * LS, a_hi | (uint64_t)((a16 + kw) & 0x3FFF),
                                              b_hi | (uint64_t)((wrot16 + kw * wimg16) & 0x3FFF), idesc48, 1u);
                            } else {
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw)
                                    umma_bf16(tmem_base + (uint32_t)((ji - 1) * LS), a_hi | (uint64_t)((a16 + kw) & 0x3FFF),
                                              b_hi | (uint64_t)((wrot16 + kw * wimg16) & 0x3FFF), idesc144, 1u);
                            }
                            if (ji >= 2) umma_commit(&acc_done[ji - 1]);   // line slot ji-1 has all its kh contributions
                        }
                    }
                    __syncwarp();
                    if (skip_mma && elect_one()) {
                        for (int ji = (half == 0 ? 2 : 4); ji < (half == 0 ? 4 : kBandLines); ++ji) umma_commit(&acc_done[ji - 1]);
                    }
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
            unsigned long long gt1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
            g_march_prof[cta * 16 + 0] = gt1 - gt0;          // nanoseconds (calibrates the SM clock)
            g_march_prof[cta * 16 + 1] = gt0;                // absolute start (spread over CTAs = launch skew)
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
        const int ew = warp & 3;
        const int m = ew * 32 + lane;
        constexpr int GS = CO / 8;
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
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kBandEpiGroups) : "memory");
            if (w8 == 0 && lane < 16) {
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < 4 * kBandEpiGroups; ++q) v += stat_smem[q * 16 + lane];
                p.stats_partial[((size_t)cta * p.N + n) * 16 + lane] = v;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kBandEpiGroups) : "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        };

        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, w_done = 0, t_ld = 0, t_st = 0, t_rest = 0, tb = 0;
        MARCH_PROF_T(tb);
        uint32_t k = 0;
        for (long long u = u_begin; u < u_end;) {
            BandSeg sg;
            u += band_segment(p, u, u_end, sg);
            if (do_stats && sg.n != cur_n) {
                if (cur_n >= 0) flush_stats(cur_n);
                cur_n = sg.n;
            }
            const int wp = 1 + sg.col * 128 + m;              // padded column of this thread's voxel
            const bool w_ok = wp <= p.W;
            const int hp0 = sg.band * kBandBH + 1;            // padded line index of line slot 1
            for (int dpi = sg.d0; dpi <= sg.d1 + 2; ++dpi, ++k) {
                const int dpo = dpi - 1;
                const uint32_t slot = (uint32_t)((dpi + 2) % 3);
                const bool out_valid = dpi >= sg.d0 + 2;
                const bool last = dpi == sg.d1 + 2;
                for (int jo = 1 + grp; jo <= kBandBH; jo += kBandEpiGroups) {
                    MARCH_PROF_T(t0);
                    if (prof && t3) t_rest += t0 - t3;
                    // The residual operand (the identity path of a Residual block's data gradient) is requested
                    // BEFORE the wait for the accumulator, so its L2/HBM latency hides behind the MMAs of this line;
                    // loaded after the TMEM read it serialised ~1 us per line and cost the kernel +50 us.
                    uint4 rq[CO / 8], gq[CO / 8];
                    bool have_res = false;
                    if (EPI == EPI_BF16 && (p.residual.base || do_gnb)) {
                        const int hp_r = hp0 + jo - 1;
                        const bool vox_ok = out_valid && w_ok && hp_r <= p.H;
                        have_res = vox_ok && p.residual.base != nullptr;
                        const long long rrow = (((long long)sg.n * (p.D + 2) + dpo) * (p.H + 2) + hp_r) * p.Wp + wp;
                        if (have_res) {
#pragma unroll
                            for (int c = 0; c < CO / 8; ++c) {
                                const void* rp = p.residual.at(c, rrow);
                                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(rq[c].x), "=r"(rq[c].y), "=r"(rq[c].z), "=r"(rq[c].w) : "l"(rp));
                            }
                        }
                        if (do_gnb && vox_ok) {           // the saved conv output of the GroupNorm differentiated next
#pragma unroll
                            for (int c = 0; c < CO / 8; ++c) {
                                const void* gp = p.gnb_x.at(c, rrow);
                                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(gq[c].x), "=r"(gq[c].y), "=r"(gq[c].z), "=r"(gq[c].w) : "l"(gp));
                            }
                        }
                    }
                    mbar_wait(&acc_done[jo], k & 1);
                    MARCH_PROF_T(t1);
                    w_done += t1 - t0;
                    tc_fence_after();
                    const uint32_t tl = tlane + (uint32_t)(jo * LS);
                    uint32_t r[CO];
                    tmem_ld16_nowait(tl + slot * CO, r);
                    tmem_ld_wait();
                    MARCH_PROF_T(t2);
                    t_ld += t2 - t1;
                    if (last) {
#pragma unroll
                        for (int c0 = 0; c0 < LS; c0 += 16) tmem_st16_zero(tl + c0);
                    } else {
                        tmem_st16_zero(tl + slot * CO);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&acc_free[(jo - 1) / (kBandBH / 2)]);
                    MARCH_PROF_T(t3);
                    t_st += t3 - t2;
                    const int hp = hp0 + jo - 1;
                    if (!out_valid || !w_ok || hp > p.H) continue;
                    const long long orow = (((long long)sg.n * (p.D + 2) + dpo) * (p.H + 2) + hp) * p.Wp + wp;
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
                        if (have_res) {
#pragma unroll
                            for (int c = 0; c < CO / 8; ++c) {
                                float f[8];
                                unpack_bf16x8(rq[c], f);
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
                    } else {   // EPI_SIGMOID
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
    if (warp == 3) tmem_dealloc(tmem_base, 512);
    // GroupNorm statistics: the CTA that finishes last sums every CTA's partial row (the pipeline's shared memory is idle
    // by now: every bulk copy has been consumed) - no separate finalize launch between this conv and gn_apply.
    if (p.gn_fin.mean != nullptr &&
        cta_draws_last_ticket(p.gn_fin.ticket, gridDim.x, reinterpret_cast<unsigned int*>(smem + 2048)))
        gn_stats_finalize_cta(p.stats_partial, ctas, p.N, p.gn_fin, reinterpret_cast<double*>(smem));
}

}  // namespace b200

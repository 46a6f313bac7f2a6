// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), bf16 in / fp32 accumulate.
//
// Replaces aten::convolution as reached from the reference at model.py:72-73 (k3 p1),
// model.py:336 (conv_input), model.py:348 (conv_output + sigmoid, model.py:431),
// model.py:362 (k2 s2, run as a 1x1x1 conv over a space-to-depth tensor), model.py:393/401 (k1),
// and, with flipped/transposed packed weights, the data-gradient half of its backward
// (train.py:210).
//
// Data layout: chunk-planar zero-halo activations (see common.cuh).  Because every activation
// tensor carries its zero halo in HBM, a 3x3x3 tap is a constant row shift in the linear row
// index: in[r + (kd-1)*SS + (kh-1)*Wp + (kw-1)].  A tile is a run of 128*MB consecutive rows
// in each of BD consecutive slices; its input window is loaded ONCE per 16-channel K group:
// each (8-channel chunk, slice) is one contiguous range of a chunk plane in HBM, fetched by a
// single cp.async.bulk (TMA bulk engine) straight into its [chunk][slice][row][8 ch] shared
// memory plane - exactly the SWIZZLE_NONE K-major canonical layout of a UMMA operand - so each
// of the 27 taps is just a different start address in the A descriptor.  Rows that land on
// halo positions are junk: computed, never stored.
//
// kw-fold (FOLD = 1, Cout <= 32): a K=16 MMA costs max(46, N/2) cycles regardless of N <= 92
// (measured, tools/umma_probe.cu), so with N = Cout = 16 the tensor pipe idles 5/6 of the time.
// The fold makes N = 3*Cout: column block kw holds  P_kw[r] = sum_{kd,kh,ci} X[r+(kd-1)SS+(kh-1)Wp] W[kd,kh,kw]
// (9 MMAs per run instead of 27) and the epilogue forms  out[r] = P_0[r-1] + P_1[r] + P_2[r+1]
// with warp shuffles (+ a 2-row exchange through shared memory at warp boundaries).  Blocks then
// advance by 126 rows: P rows 0 and 127 of a 128-row block only feed their neighbours.
//
// Warp roles (384 threads): w0 activation bulk-copy producer, w1 weight bulk-copy producer,
// w2 MMA issuer (one elected lane), w3 TMEM allocator, w4-7 and w8-11 two epilogue groups
// (TMEM -> regs -> HBM) that take alternate runs of a tile.
#pragma once
#include "common.cuh"

namespace b200 {

enum { MODE_K3 = 0, MODE_K1 = 1 };
enum { EPI_BF16 = 0, EPI_SIGMOID = 1, EPI_D2S = 2 };   // EPI_D2S: host-side name of the bf16 epilogue with ConvKParams::d2s set

constexpr int kMaxTaps = 27;
constexpr int kEpiGroups = 3;                       // epilogue warp groups (4 warps each)
constexpr int kConvThreads = 128 + 128 * kEpiGroups;
constexpr unsigned kConvTailBytes = 4096;   // barriers + TMEM slot + stats + fold exchange buffers

struct ConvKParams {
    // output volume (interior dims) and padded strides
    int N, D, H, W, Wp, SS;
    FastDiv by_SS, by_Wp;
    long long sample_rows, total_rows;
    int mode;
    // tiling
    int BD, MB, TR;        // slices per tile, 128-row blocks per slice-run, rows per run
    int Q0, QN;            // q-range (rows relative to the slice base) covered by the tiles
    int tiles_q, tiles_d, num_tiles;
    int whole;             // 1: q runs over the whole sample (BD == 1), 0: per slice
    int n_jobs;            // column blocks (each its own packed weights); grid = ctas * n_jobs
    int pair;              // 1: conv_pair.cuh (cta_group::2); the two column jobs are the two CTAs' halves of B, grid = 2 * pairs
    int Nm_pack;           // columns per packed-weight job (GEMM N of one CTA's B operand)
    // K loop
    int KG, KGa, KC, NTG, TG;
    // shared-memory plan
    int x_stages, w_stages;
    int w_resident;        // 1: all KG*NTG weight stages stay in shared memory for the whole kernel
    unsigned x_stage_bytes, w_stage_bytes, x_plane_bytes;
    int SRp, nslices, halo_rows;
    int tap_off[kMaxTaps];
    unsigned smem_x_off, smem_w_off, smem_bar_off;
    unsigned tmem_cols;
    // operands
    ActRef src_a, src_b;   // chunk-planar activations (src_b only when KG > KGa)
    const __nv_bfloat16* wpacked;
    // epilogue
    int Cout_total;        // channels of `out`
    int lrelu_out;         // apply LeakyReLU(0.01) before storing
    ActRef out;
    ActRef residual;       // optional (base == nullptr if unused), same layout as out, added before activation
    // depth-to-space store (MODE_K1, data gradient of the k2 s2 conv, model.py:360-363 backward): the GEMM's Cout = 8 * Cf
    // columns of coarse voxel (d, h, w) are the Cf channels of the eight fine voxels (2d+kd, 2h+kh, 2w+kw), column block
    // ((kd*2+kh)*2+kw) * Cf (elementwise.cuh s2d layout); `out` / `residual` are then FINE tensors (2D, 2H, 2W, Cf) and the
    // epilogue scatters its 16-byte vectors there (+ the skip gradient as residual) - no coarse tensor, no d2s pass.
    int prefetch_residual; // 1 (default): request a row block's residual sectors from L2 one block ahead (B200_RES_PREFETCH=0: off)
    int d2s_spread;        // d2s: (row block, column group) items dealt to all epilogue groups (default; B200_D2S_SPREAD=0: one group per row block)
    int d2s_v8;            // d2s with 32-byte accesses (opt-in B200_D2S_V8=1; W even, 32-byte aligned tensors)
    int d2s;               // 1: scatter
    int d2s_sh;            // log2(Cf / 8): coarse chunk q -> tap q >> sh, fine chunk q & ((1 << sh) - 1)
    float* stats_partial;            // optional [ctas][N][16] (sum[8], sumsq[8])
    GnFin gn_fin;                     // mean != nullptr: the last CTA turns stats_partial into mean / rstd (common.cuh)
    const float* bias;               // EPI_SIGMOID
    float* probs;                    // EPI_SIGMOID: fp32 NCDHW (unpadded)
    float* logits;                   // EPI_SIGMOID: optional fp32 NCDHW
    int n_out_real;                  // EPI_SIGMOID: real output channels (3)
    int debug;                       // perf probes only: 1 = skip activation loads, 2 = skip MMAs
};

struct TileCoord {
    int n, d0, q0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int t) {
    TileCoord c;
    int tq = t % p.tiles_q;
    int r = t / p.tiles_q;
    int td = r % p.tiles_d;
    c.n = r / p.tiles_d;
    c.d0 = td * p.BD;
    c.q0 = p.Q0 + tq * p.TR;
    return c;
}

template <int MODE, int EPI, int NMMA, int FOLD>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
    constexpr int RB = FOLD ? 126 : 128;            // output rows per 128-row MMA block
    constexpr int CO = FOLD ? NMMA / 3 : NMMA;      // output channels handled by this CTA job
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x % p.n_jobs;
    const int cta = blockIdx.x / p.n_jobs;
    const int ctas = gridDim.x / p.n_jobs;

    uint8_t* smem_x = smem + p.smem_x_off;
    uint8_t* smem_w = smem + p.smem_w_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + p.x_stages;
    uint64_t* w_full = x_empty + p.x_stages;
    uint64_t* w_empty = w_full + p.w_stages;
    uint64_t* t_full = w_empty + p.w_stages;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);  // [4*kEpiGroups warps][16]
    float* xch_smem = stat_smem + 4 * kEpiGroups * 16;           // [groups][2 bufs][4 warps][2][16]

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.x_stages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < p.w_stages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128 * kEpiGroups); }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int R = p.BD * p.MB;  // accumulators (runs) per tile

    // Role code is executed by the WHOLE warp with warp-uniform values; only the instruction
    // issue is predicated on elect_one().  (Putting the role under `lane == 0` makes ptxas wrap
    // every UTCHMMA/UTMALDG in an ELECT + R2UR waterfall: ~5x slower MMA issue, measured.)

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
        // ================= activation producer (TMA) =================
        int xs = 0; uint32_t xph = 0;
        for (int t = cta; t < p.num_tiles; t += ctas) {
            TileCoord tc = decode_tile(p, t);
            for (int g = 0; g < p.KG; ++g) {
                mbar_wait(&x_empty[xs], xph ^ 1);
                if (elect_one()) {
                    if (B200_DBG(p, 1)) {
                        mbar_arrive(&x_full[xs]);
                    } else {
                        mbar_arrive_expect_tx(&x_full[xs], p.x_stage_bytes);
                        // (the issue loop of this single thread must stay well below the MMA / HBM time of a stage:
                        // one pointer per slice, then a constant stride per chunk - no 64-bit multiplies inside)
                        const ActRef& src = (g < p.KGa) ? p.src_a : p.src_b;
                        const int chunk0 = (g < p.KGa ? g : g - p.KGa) * (p.KC / 8);
                        uint8_t* dst = smem_x + (size_t)xs * p.x_stage_bytes;
                        const uint32_t bytes = (uint32_t)p.SRp * 16;
                        const size_t plane_stride = (size_t)src.plane_rows * 8;
                        const int nch = p.KC / 8;
                        for (int s = 0; s < p.nslices; ++s) {
                            long long row0;
                            if (MODE == MODE_K3) {
                                const int dpi = (p.whole ? 0 : tc.d0 + 1) - 1 + s;
                                row0 = ((long long)tc.n * (p.D + 2) + dpi) * p.SS + tc.q0 - p.halo_rows;
                            } else {
                                row0 = (long long)t * p.TR;
                            }
                            const __nv_bfloat16* ps = src.at(chunk0, row0);
                            uint8_t* pd = dst + (size_t)s * bytes;
                            for (int c = 0; c < nch; ++c, ps += plane_stride, pd += p.x_plane_bytes)
                                bulk_load_1d(pd, ps, bytes, &x_full[xs]);
                        }
                    }
                }
                __syncwarp();
                if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= weight producer (bulk copy) =================
        int ws = 0; uint32_t wph = 0;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked) +
                              (size_t)job * p.KG * p.NTG * p.w_stage_bytes;
        for (int t = cta; t < p.num_tiles; t += ctas) {
            for (int g = 0; g < p.KG; ++g) {
                for (int tg = 0; tg < p.NTG; ++tg) {
                    if (!p.w_resident) mbar_wait(&w_empty[ws], wph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&w_full[ws], p.w_stage_bytes);
                        bulk_load_1d(smem_w + (size_t)ws * p.w_stage_bytes,
                                     wsrc + (size_t)(g * p.NTG + tg) * p.w_stage_bytes, p.w_stage_bytes, &w_full[ws]);
                    }
                    __syncwarp();
                    if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
                }
            }
            if (p.w_resident) break;      // one pass fills every slot; they are never released
        }
    } else if (warp == 2) {
        // ================= MMA issuer =================
        // Descriptor high words are loop invariant (LBO, SBO, version); only the 14-bit start
        // address field changes per MMA, so each issue costs a couple of integer adds.
        constexpr uint32_t idesc = make_idesc(128, NMMA, 0, 0);
        const uint32_t plane16 = p.x_plane_bytes >> 4;
        const uint64_t a_hi = ((uint64_t)(plane16 & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)((NMMA * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t xbase16 = smem_u32(smem_x) >> 4, wbase16 = smem_u32(smem_w) >> 4;
        const uint32_t xstage16 = p.x_stage_bytes >> 4, wstage16 = p.w_stage_bytes >> 4;
        const int ksteps = p.KC / 16;
        const uint32_t wtap16 = (uint32_t)(p.KC / 8) * NMMA;     // 16B units per tap in a weight stage
        const bool skip_mma = B200_DBG(p, 2);
        int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
        int it = 0;
        for (int t = cta; t < p.num_tiles; t += ctas, ++it) {
            const int as = it & 1;
            mbar_wait(&t_empty[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int g = 0; g < p.KG; ++g) {
                mbar_wait(&x_full[xs], xph);
                tc_fence_after();
                const uint32_t xst16 = xbase16 + xs * xstage16;
                for (int tg = 0; tg < p.NTG; ++tg) {
                    if (!p.w_resident || it == 0) {
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                    }
                    const uint32_t wst16 = wbase16 + ws * wstage16;
                    const int* toff = &p.tap_off[tg * p.TG];
                    uint32_t dtm = tmem_base + (uint32_t)(as * R * NMMA);
                    uint32_t first = (g | tg) == 0 ? 0u : 1u;
                    if (MODE == MODE_K3) {
                        // compile-time shape: KC = 16 (one K=16 MMA per tap), TGc taps per weight stage
                        constexpr int TGc = FOLD ? 3 : 9;
                        uint32_t tof[TGc];
#pragma unroll
                        for (int i = 0; i < TGc; ++i) tof[i] = (uint32_t)toff[i];
                        const int BD = p.BD, MB = p.MB, SRp = p.SRp;
                        if (!skip_mma && elect_one()) {
                            for (int dz = 0; dz < BD; ++dz) {
                                for (int mb = 0; mb < MB; ++mb, dtm += NMMA) {
                                    const uint32_t a_run16 = xst16 + (uint32_t)(dz * SRp + mb * RB);
#pragma unroll
                                    for (int tl = 0; tl < TGc; ++tl) {
                                        umma_bf16(dtm, a_hi | (uint64_t)(a_run16 + tof[tl]),
                                                  b_hi | (uint64_t)(wst16 + (uint32_t)(tl * 2 * NMMA)), idesc,
                                                  tl == 0 ? first : 1u);
                                    }
                                }
                            }
                        }
                    } else if (!skip_mma && elect_one()) {
                        for (int dz = 0; dz < p.BD; ++dz) {
                            for (int mb = 0; mb < p.MB; ++mb, dtm += NMMA) {
                                const uint32_t a_run16 = xst16 + (uint32_t)(dz * p.SRp + mb * 128);
                                uint32_t b16 = wst16;
                                uint32_t acc = first;
                                for (int tl = 0; tl < p.TG; ++tl) {
                                    uint32_t a16 = a_run16 + (uint32_t)toff[tl];
                                    for (int ks = 0; ks < ksteps; ++ks) {
                                        umma_bf16(dtm, a_hi | (uint64_t)(a16 & 0x3FFF), b_hi | (uint64_t)(b16 & 0x3FFF),
                                                  idesc, acc);
                                        acc = 1u;
                                        a16 += 2 * plane16;
                                        b16 += 2 * NMMA;
                                    }
                                    b16 += wtap16 - (uint32_t)ksteps * 2 * NMMA;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (!p.w_resident) {
                        if (elect_one()) umma_commit(&w_empty[ws]);
                        __syncwarp();
                    }
                    if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
                }
                if (elect_one()) umma_commit(&x_empty[xs]);
                __syncwarp();
                if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
            }
            if (elect_one()) umma_commit(&t_full[as]);
            __syncwarp();
        }
        // every MMA of this CTA is issued: what is left is its last epilogue - the dependent kernel may start launching
        // (it still waits for this whole grid in its own pdl_wait; triggering at the start would park its CTAs here for
        // the whole kernel)
        pdl_trigger();
    } else if (warp >= 4) {
        // ================= epilogue (kEpiGroups groups of 4 warps take runs round-robin) =================
        const int grp = (warp - 4) >> 2;
        const int ew = (warp - 4) & 3;      // TMEM lanes [32*ew, 32*ew+32)
        const int m = ew * 32 + lane;       // row within a 128-row block
        constexpr int GS = (CO >= 16 ? CO / 8 : 1);   // channels per GroupNorm group
        float ssum[8], ssq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        int cur_n = -1;
        const bool do_stats = (EPI == EPI_BF16) && (p.stats_partial != nullptr) && !B200_DBG(p, 64);
        float* xch = xch_smem + grp * (2 * 4 * 2 * 16);
        int xbuf = 0;

        auto flush_stats = [&](int n) {
            // all epilogue threads participate
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = warp_sum(ssum[i]); ssq[i] = warp_sum(ssq[i]); }
            const int w8 = warp - 4;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { stat_smem[w8 * 16 + i] = ssum[i]; stat_smem[w8 * 16 + 8 + i] = ssq[i]; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiGroups) : "memory");
            if (w8 == 0 && lane < 16) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < 4 * kEpiGroups; ++k) v += stat_smem[k * 16 + lane];
                p.stats_partial[((size_t)cta * p.N + n) * 16 + lane] = v;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiGroups) : "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        };

        // Residual operand (skip / residual-branch gradient of the data-gradient convs): the loads of a row sit behind
        // the tcgen05.wait::ld of their column group (a memory clobber), so a row block used to be a chain of CO/16
        // DRAM round trips - 8 x ~0.9 us per 128 rows in the depth-to-space data gradient, which ran at 2.9 TB/s with
        // nothing else on the GPU.  The whole block's residual sectors are now requested from L2 one block ahead: for
        // the first block of a tile BEFORE the wait for the tile's MMAs, for the next block before the current one is
        // processed.  (Same address arithmetic as the row decode below.)
        auto prefetch_residual = [&](int t, const TileCoord& tc, int r) {
            const int dz = r / p.MB, mb = r - dz * p.MB;
            bool valid;
            long long orow, frow0 = 0;
            if (MODE == MODE_K3) {
                const int q = tc.q0 + mb * RB + m - (FOLD ? 1 : 0);
                const int dpo = (p.whole ? 0 : tc.d0 + 1) + dz;
                const int dq = p.by_SS.div(q);
                const int r2 = q - dq * p.SS;
                const int hp = p.by_Wp.div(r2);
                const int wp = r2 - hp * p.Wp;
                const int dp = dpo + dq;
                valid = (q < p.Q0 + p.QN) && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W;
                if (FOLD) valid = valid && m >= 1 && m <= RB;
                orow = ((long long)tc.n * (p.D + 2) + dpo) * p.SS + q;
            } else {
                orow = (long long)t * p.TR + mb * 128 + m;
                valid = orow < p.total_rows;
                if (p.d2s) {
                    const long long dpf = orow / p.SS;
                    const int r2 = (int)(orow - dpf * p.SS);
                    const int hp = p.by_Wp.div(r2), wp = r2 - hp * p.Wp;
                    const int nn = (int)(dpf / (p.D + 2)), dp = (int)(dpf - (long long)nn * (p.D + 2));
                    valid = valid && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W;
                    frow0 = (((long long)nn * (2 * p.D + 2) + (2 * dp - 1)) * (2 * p.H + 2) + (2 * hp - 1)) * (2 * p.W + 2) +
                            (2 * wp - 1);
                }
            }
            if (!valid) return;
#pragma unroll
            for (int c0 = 0; c0 < CO; c0 += 16) {
                int ch = (job * CO + c0) >> 3;
                long long row = orow;
                if (MODE == MODE_K1 && p.d2s) {
                    const int tap8 = ch >> p.d2s_sh;
                    ch &= (1 << p.d2s_sh) - 1;
                    row = frow0 + (long long)(tap8 >> 2) * ((2 * p.H + 2) * (2 * p.W + 2)) +
                          ((tap8 >> 1) & 1) * (2 * p.W + 2) + (tap8 & 1);
                }
                prefetch_l2(p.residual.at(ch, row));
                prefetch_l2(p.residual.at(ch + 1, row));
            }
        };
        const bool pf_res = (EPI == EPI_BF16) && p.residual.base != nullptr && p.prefetch_residual != 0;

        int it = 0;
        for (int t = cta; t < p.num_tiles; t += ctas, ++it) {
            const int as = it & 1;
            TileCoord tc = decode_tile(p, t);
            if (do_stats && tc.n != cur_n) {
                if (cur_n >= 0) flush_stats(cur_n);
                cur_n = tc.n;
            }
            if (MODE == MODE_K1 && EPI == EPI_BF16 && !FOLD && (NMMA == 128 || NMMA == 256) && p.d2s && p.d2s_spread && !p.d2s_v8 && !do_stats && R <= 2) {
                // ---- depth-to-space scatter, work spread over ALL epilogue groups (round 2c) ----
                // A tile of this GEMM has R <= 2 row blocks (TMEM: 2 stages x R x 128 columns) but three epilogue groups, and a
                // row block is a chain of CO/16 dependent (TMEM read -> skip gradient -> store) steps: with one group per row
                // block a third of the epilogue warps idled and the kernel, which has the GPU to itself in the step, ran at
                // 0.46 of the HBM peak.  The (row block, column group) items are dealt round-robin to the three groups instead;
                // every group decodes both row blocks once and requests its own items' skip-gradient sectors from L2 before it
                // waits for the accumulators.  (The same dealing for the 3x3x3 convs with R = 1 / 2 / 4 was measured too: no gain
                // at 128 channels, 30 vs 26 us at 64 channels - profiles/r02c_ab_d2s.txt; not kept.)
                constexpr int NCG = CO / 16;
                bool val0 = false, val1 = false;
                long long fr0 = 0, fr1 = 0;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (r < R) {
                        const long long orow = (long long)t * p.TR + r * 128 + m;
                        bool valid = orow < p.total_rows;
                        const long long dpf = orow / p.SS;
                        const int r2 = (int)(orow - dpf * p.SS);
                        const int hp = p.by_Wp.div(r2), wp = r2 - hp * p.Wp;
                        const int nn = (int)(dpf / (p.D + 2)), dp = (int)(dpf - (long long)nn * (p.D + 2));
                        valid = valid && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W;
                        const long long f0 = (((long long)nn * (2 * p.D + 2) + (2 * dp - 1)) * (2 * p.H + 2) + (2 * hp - 1)) * (2 * p.W + 2) +
                                             (2 * wp - 1);
                        if (r == 0) { val0 = valid; fr0 = f0; } else { val1 = valid; fr1 = f0; }
                    }
                }
                const long long fslice = (long long)(2 * p.H + 2) * (2 * p.W + 2);
                const bool has_res = p.residual.base != nullptr;
                auto item_row = [&](int item, int& r, int& c0, int& ch, long long& row) {
                    r = item / NCG;
                    c0 = (item - r * NCG) * 16;
                    const int chg = (job * CO + c0) >> 3;
                    const int tap8 = chg >> p.d2s_sh;
                    ch = chg & ((1 << p.d2s_sh) - 1);
                    row = (r ? fr1 : fr0) + (long long)(tap8 >> 2) * fslice + ((tap8 >> 1) & 1) * (2 * p.W + 2) + (tap8 & 1);
                };
                if (has_res && p.prefetch_residual) {
                    for (int item = grp; item < R * NCG; item += kEpiGroups) {
                        int r, c0, ch; long long row;
                        item_row(item, r, c0, ch, row);
                        if (r ? val1 : val0) { prefetch_l2(p.residual.at(ch, row)); prefetch_l2(p.residual.at(ch + 1, row)); }
                    }
                }
                mbar_wait(&t_full[as], (it >> 1) & 1);
                tc_fence_after();
                for (int item = grp; item < R * NCG; item += kEpiGroups) {
                    int r, c0, ch; long long row;
                    item_row(item, r, c0, ch, row);
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((as * R + r) * NMMA) + (uint32_t)c0, v);
                    if (r ? val1 : val0) {
                        if (has_res) {
                            float f[8];
                            unpack_bf16x8(ld_nc_v4(p.residual.at(ch, row)), f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += f[i];
                            unpack_bf16x8(ld_nc_v4(p.residual.at(ch + 1, row)), f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
                        }
                        *reinterpret_cast<uint4*>(p.out.at(ch, row)) = pack_bf16x8(v);
                        *reinterpret_cast<uint4*>(p.out.at(ch + 1, row)) = pack_bf16x8(v + 8);
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[as]);
                continue;
            }
            if (pf_res && grp < R) prefetch_residual(t, tc, grp);
            mbar_wait(&t_full[as], (it >> 1) & 1);
            tc_fence_after();
            for (int r = grp; r < R; r += kEpiGroups) {
                if (pf_res && r + kEpiGroups < R) prefetch_residual(t, tc, r + kEpiGroups);
                const int dz = r / p.MB, mb = r - dz * p.MB;
                // ---- which voxel is this row? ----
                bool valid;
                long long orow;      // output row in the padded tensor
                long long frow0 = 0; // d2s: row of the fine voxel (2d, 2h, 2w)
                int vd = 0, vh = 0, vw = 0;
                if (MODE == MODE_K3) {
                    const int q = tc.q0 + mb * RB + m - (FOLD ? 1 : 0);
                    const int dpo = (p.whole ? 0 : tc.d0 + 1) + dz;
                    const int dq = p.by_SS.div(q);
                    const int r2 = q - dq * p.SS;
                    const int hp = p.by_Wp.div(r2);
                    const int wp = r2 - hp * p.Wp;
                    const int dp = dpo + dq;
                    valid = (q < p.Q0 + p.QN) && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W;
                    if (FOLD) valid = valid && m >= 1 && m <= RB;
                    orow = ((long long)tc.n * (p.D + 2) + dpo) * p.SS + q;
                    vd = dp - 1; vh = hp - 1; vw = wp - 1;
                } else {
                    orow = (long long)t * p.TR + mb * 128 + m;
                    valid = orow < p.total_rows;
                    if (p.d2s) {
                        // coarse padded row -> (n, dp, hp, wp); halo rows hold nothing; fine row of the voxel's (0,0,0) corner
                        const long long dpf = orow / p.SS;
                        const int r2 = (int)(orow - dpf * p.SS);
                        const int hp = p.by_Wp.div(r2), wp = r2 - hp * p.Wp;
                        const int nn = (int)(dpf / (p.D + 2)), dp = (int)(dpf - (long long)nn * (p.D + 2));
                        valid = valid && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W;
                        frow0 = (((long long)nn * (2 * p.D + 2) + (2 * dp - 1)) * (2 * p.H + 2) + (2 * hp - 1)) * (2 * p.W + 2) +
                                (2 * wp - 1);
                    }
                }
                const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((as * R + r) * NMMA);
                if (MODE == MODE_K1 && EPI == EPI_BF16 && !FOLD && NMMA == 128 && p.d2s && p.d2s_v8 && !do_stats) {
                    // ---- depth-to-space scatter with 32-byte accesses (round 2c; opt-in B200_D2S_V8=1: bit-identical to the
                    //      16-byte form below and measured slower - 177 vs 101 us - so the default stays the 16-byte form) ----
                    // A lane's two kw taps are the fine rows 2wp-1 and 2wp: 32 contiguous bytes that straddle a sector
                    // boundary, and one 16-byte access per lane at a 32-byte lane stride touches 32 half-used sectors
                    // per request (ncu: the kernel was bound by L1 sector throughput at 2.8 TB/s with the GPU to itself).
                    // Sector-aligned pairs are (own kw = 1, NEXT lane's kw = 0) = fine rows 2wp, 2wp+1: the kw = 0
                    // accumulators move one lane down by shuffle and every lane reads the skip gradient and writes its
                    // result with one 256-bit access per chunk.  Lanes whose neighbour is missing (warp edge, line edge,
                    // halo row) fall back to 16-byte accesses for the orphaned row.
                    const int Cf = 8 << p.d2s_sh;                                   // fine channels = columns per tap
                    const bool nvalid = (__shfl_down_sync(0xffffffffu, (int)valid, 1) != 0) && lane < 31;
                    const bool pvalid = (__shfl_up_sync(0xffffffffu, (int)valid, 1) != 0) && lane > 0;
                    const bool pair = valid && nvalid;                               // both valid <=> consecutive voxels of one line
                    const bool solo1 = valid && !nvalid;                             // my kw = 1 row on its own
                    const bool solo0 = valid && !pvalid;                             // my kw = 0 row: nobody above me takes it
                    const bool has_res = p.residual.base != nullptr;
                    const long long fslice = (long long)(2 * p.H + 2) * (2 * p.W + 2);
#pragma unroll
                    for (int c0 = 0; c0 < CO; c0 += 16) {
                        const int chg = (job * CO + c0) >> 3;                        // global chunk index of this column group
                        const int tap8 = chg >> p.d2s_sh;
                        if (tap8 & 1) continue;                                      // kw = 1 groups are handled with their kw = 0 partner
                        const int fch = chg & ((1 << p.d2s_sh) - 1);                 // first of the two fine chunks
                        uint32_t r0[16], r1[16];
                        tmem_ld16_nowait(trow + c0, r0);                             // kw = 0
                        tmem_ld16_nowait(trow + c0 + Cf, r1);                        // kw = 1 (the next tap: Cf columns further)
                        tmem_ld_wait();
                        float v0[16], v1[16], n0[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            v0[i] = __uint_as_float(r0[i]);
                            v1[i] = __uint_as_float(r1[i]);
                            n0[i] = __shfl_down_sync(0xffffffffu, v0[i], 1);         // the next voxel's kw = 0 values
                        }
                        const long long row0 = frow0 + (long long)(tap8 >> 2) * fslice + ((tap8 >> 1) & 1) * (2 * p.W + 2);
                        const long long row1 = row0 + 1;                             // even row: 32-byte aligned in every plane
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            if (pair) {
                                if (has_res) {
                                    uint4 qa, qb;
                                    ld_nc_v8(p.residual.at(fch + cc, row1), qa, qb);
                                    float f[8];
                                    unpack_bf16x8(qa, f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v1[8 * cc + i] += f[i];
                                    unpack_bf16x8(qb, f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) n0[8 * cc + i] += f[i];
                                }
                                st_v8(p.out.at(fch + cc, row1), pack_bf16x8(v1 + 8 * cc), pack_bf16x8(n0 + 8 * cc));
                            } else if (solo1) {
                                if (has_res) {
                                    float f[8];
                                    unpack_bf16x8(ld_nc_v4(p.residual.at(fch + cc, row1)), f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v1[8 * cc + i] += f[i];
                                }
                                *reinterpret_cast<uint4*>(p.out.at(fch + cc, row1)) = pack_bf16x8(v1 + 8 * cc);
                            }
                            if (solo0) {
                                if (has_res) {
                                    float f[8];
                                    unpack_bf16x8(ld_nc_v4(p.residual.at(fch + cc, row0)), f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v0[8 * cc + i] += f[i];
                                }
                                *reinterpret_cast<uint4*>(p.out.at(fch + cc, row0)) = pack_bf16x8(v0 + 8 * cc);
                            }
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int c0 = 0; c0 < CO; c0 += 16) {
                    float v[16];
                    if (FOLD) {
                        float a0[16], a2[16];
                        {
                            uint32_t r0[16], r1[16], r2[16];
                            if (!B200_DBG(p, 16)) {
                                tmem_ld16_nowait(trow + c0, r0);
                                tmem_ld16_nowait(trow + CO + c0, r1);
                                tmem_ld16_nowait(trow + 2 * CO + c0, r2);
                                tmem_ld_wait();
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) { r0[i] = m + i; r1[i] = m - i; r2[i] = m * i; }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                a0[i] = __uint_as_float(r0[i]);
                                v[i] = __uint_as_float(r1[i]);
                                a2[i] = __uint_as_float(r2[i]);
                            }
                        }
                        // rows m-1 / m+1 live in the neighbouring lanes; across a warp boundary they
                        // come through shared memory (lane 31's P_0 row and lane 0's P_2 row per warp).
                        float* xb = xch + xbuf * (4 * 2 * 16);
                        float bnd[16];      // read only by lanes 0/31; rows 0 and 127 of a block are never stored
                        if (!B200_DBG(p, 4)) {
                            if (lane == 31) {
                                float4* d4 = reinterpret_cast<float4*>(xb + (ew * 2 + 0) * 16);
#pragma unroll
                                for (int i = 0; i < 4; ++i) d4[i] = make_float4(a0[4 * i], a0[4 * i + 1], a0[4 * i + 2], a0[4 * i + 3]);
                            } else if (lane == 0) {
                                float4* d4 = reinterpret_cast<float4*>(xb + (ew * 2 + 1) * 16);
#pragma unroll
                                for (int i = 0; i < 4; ++i) d4[i] = make_float4(a2[4 * i], a2[4 * i + 1], a2[4 * i + 2], a2[4 * i + 3]);
                            }
                            asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                            if (lane == 0 || lane == 31) {
                                // (warp 0 / lane 0 and warp 3 / lane 31 read a neighbour slot that holds
                                // finite junk: their rows m = 0 and m = 127 are never stored)
                                const float4* s4 = reinterpret_cast<const float4*>(
                                    lane == 0 ? xb + (((ew + 3) & 3) * 2 + 0) * 16 : xb + (((ew + 1) & 3) * 2 + 1) * 16);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float4 t4 = s4[i];
                                    bnd[4 * i] = t4.x; bnd[4 * i + 1] = t4.y; bnd[4 * i + 2] = t4.z; bnd[4 * i + 3] = t4.w;
                                }
                            }
                        }
                        if (!B200_DBG(p, 8)) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float up = __shfl_up_sync(0xffffffffu, a0[i], 1);
                                const float dn = __shfl_down_sync(0xffffffffu, a2[i], 1);
                                v[i] += (lane == 0 ? bnd[i] : up) + (lane == 31 ? bnd[i] : dn);
                            }
                        }
                        xbuf ^= 1;
                    } else {
                        tmem_ld16(trow + c0, v);
                    }
                    if (EPI == EPI_BF16) {
                        if (valid) {
                            if (do_stats) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const int gi = (c0 + i) / GS;
                                    ssum[gi] += v[i];
                                    ssq[gi] += v[i] * v[i];
                                }
                            }
                            int ch = (job * CO + c0) >> 3;            // first of the two 8-channel chunks
                            if (MODE == MODE_K1 && p.d2s) {
                                const int tap8 = ch >> p.d2s_sh;
                                ch &= (1 << p.d2s_sh) - 1;
                                orow = frow0 + (long long)(tap8 >> 2) * ((2 * p.H + 2) * (2 * p.W + 2)) +
                                       ((tap8 >> 1) & 1) * (2 * p.W + 2) + (tap8 & 1);
                            }
                            if (p.residual.base) {
                                float f[8];
                                unpack_bf16x8(ld_nc_v4(p.residual.at(ch, orow)), f);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] += f[i];
                                unpack_bf16x8(ld_nc_v4(p.residual.at(ch + 1, orow)), f);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
                            }
                            if (p.lrelu_out) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i]);
                            }
                            if (!B200_DBG(p, 32) || v[0] == 123.456f) {
                                *reinterpret_cast<uint4*>(p.out.at(ch, orow)) = pack_bf16x8(v);
                                *reinterpret_cast<uint4*>(p.out.at(ch + 1, orow)) = pack_bf16x8(v + 8);
                            }
                        }
                    } else {  // EPI_SIGMOID: first n_out_real columns are real
                        if (valid && c0 == 0) {
                            const size_t plane = (size_t)p.D * p.H * p.W;
                            const size_t vox = ((size_t)vd * p.H + vh) * p.W + vw;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (c < p.n_out_real) {
                                    const float z = v[c] + p.bias[c];
                                    const size_t o = ((size_t)tc.n * p.n_out_real + c) * plane + vox;
                                    if (p.logits) p.logits[o] = z;
                                    p.probs[o] = 1.f / (1.f + expf(-z));
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&t_empty[as]);
        }
        if (do_stats && cur_n >= 0) flush_stats(cur_n);
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

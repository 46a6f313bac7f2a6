// Weight-gradient GEMM on tcgen05 (sm_100a): dW[tap][co][ci] = sum_rows dY[row][co] * X[row + shift(tap)][ci].
//
// Replaces the weight half of aten::convolution_backward reached from train.py:210 for every
// conv of model.py (k3: model.py:72-73, 336, 348; k1: 393, 401; k2s2 via space-to-depth: 362).
//
// The contraction index is the voxel row, so both operands are read MN-major straight out of
// the same [chunk][row][8 ch] shared-memory planes the forward kernel uses (SWIZZLE_NONE
// MN-major canonical layout: 8 rows x 16 B core matrices); each plane of a stage is one
// contiguous range of a chunk plane in HBM, fetched by a single cp.async.bulk.  Because HBM
// tensors carry zero halos and zero guard rows, dY is zero on every junk row and the sum may run
// over the plain linear row range, including past either end of the tensor.
//
// To fill the 64/128-row M dimension with small channel counts, M stacks `nband` copies of dY
// shifted by whole slices (one per kd tap) and N stacks `nfold` copies of X shifted by one line
// (one per kh tap); the three kw taps go to separate TMEM accumulators that read the same X planes
// at start addresses one row apart (a 2-row halo per stage; folding kw and giving kh to the
// accumulators instead needs a 2*(W+2)-row halo and doubles the shared-memory ingest, which is the
// bound of this kernel).  Each CTA owns a
// (job, K-split) pair, accumulates in TMEM over its row range and writes one fp32 partial;
// `wgrad_reduce_kernel` sums the partials in a fixed order (deterministic) into the fp32
// gradient in PyTorch (Cout, Cin, kd, kh, kw) layout.
#pragma once
#include "common.cuh"
#include "conv_march.cuh"     // g_march_prof / MARCH_PROF_T (in-kernel cycle accounting, debug only)

namespace b200 {

constexpr int kWgradThreads = 256;
constexpr int kMaxJobs = 32;

struct WgradKParams {
    long long total_rows;
    int Wp, SS;
    int KT;             // rows per plane of a pipeline stage (multiple of 16, <= 256)
    int nh;             // row ranges per stage: 1, or 2 ("dual": the stage covers rows [r0, r0+KT) and [r0+KT, r0+2KT),
                        // each with its own planes, stacked in M and N; only the diagonal blocks of D are meaningful)
    int XR;             // rows per X plane (KT + 2 when nacc == 3, else KT), rounded up to a multiple of 8
    int nband_loaded;   // dY bands actually loaded (3 when banded, else 1)
    int CoC, CiC;       // 8-channel chunks of dY / X per band / fold
    int nfold, nacc;
    int M, Nmma;
    int n_jobs, splits;
    int stages_per_split;       // pipeline stages each split walks
    int job_kd[kMaxJobs], job_kh[kMaxJobs], job_kw[kMaxJobs];   // fixed tap offsets in {-1,0,1}
    int job_xch[kMaxJobs];      // first X channel of this job (N-chunk jobs of wide 1x1 convs)
    int y_planes, x_planes;     // planes reserved in smem (>= loaded; M/8 and Nmma/8)
    unsigned y_plane_bytes, x_plane_bytes, stage_bytes, stage_tx_bytes;
    int stages;
    unsigned smem_bar_off;
    unsigned tmem_cols;
    ActRef dy, x;               // chunk-planar activations
    float* partial;             // [job][split][nacc][M][Nmma]
    int acc_shift;              // rows between the X start addresses of consecutive accumulators: 1 (accumulators = kw
                                // taps, N-folds = kh taps) or Wp (accumulators = kh taps, N-folds = kw taps)
    int fold_shift;             // rows between consecutive N-fold copies of X: Wp or 1 (the other of the two)
    int debug;                  // 256: per-role cycle counts into g_march_prof (tools/wgrad_prof.py)
};

__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ WgradKParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x % p.n_jobs;
    const int split = blockIdx.x / p.n_jobs;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* done = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }   // full: dY + X producer
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int stage_rows = p.KT * p.nh;
    const long long first_row = (long long)split * p.stages_per_split * stage_rows;
    int nst = p.stages_per_split;
    {
        long long remaining = p.total_rows - first_row;
        long long need = remaining <= 0 ? 0 : (remaining + stage_rows - 1) / stage_rows;
        if (need < nst) nst = (int)need;
    }
    const unsigned y_bytes = (unsigned)p.y_planes * p.y_plane_bytes;

    // Whole-warp role code with elect_one() around the issue (see conv_gemm.cuh for why).
    if (warp <= 1) {
        // ================= producers: warp 0 issues the dY planes of every stage, warp 1 the X planes =================
        // (one elected thread issues a cp.async.bulk every 42-84 cycles; with 24+ planes per stage a single issuer
        // was the bound of the 32-channel weight gradient, profiles/r01_wgrad_cycle_probe.txt)
        // Source offset of every plane for r0 = 0, computed once (one lane per plane): the per-stage issue
        // loop below must stay far below the MMA time of a stage (~2000 cycles), so it is one shared-memory
        // read and one add per bulk copy instead of the nested band/chunk/range address arithmetic.
        long long* s_yoff = reinterpret_cast<long long*>(smem + p.smem_bar_off + 256);     // [32] (tail region is 1 KB)
        long long* s_xoff = s_yoff + 32;                                                   // [32]
        const int ny = p.nband_loaded * p.CoC * p.nh, nx = p.nfold * p.CiC * p.nh;
        for (int pl = lane; pl < ny; pl += 32) {
            const int h = pl % p.nh, c = (pl / p.nh) % p.CoC, b = pl / (p.nh * p.CoC);
            const int kd = (p.nband_loaded > 1) ? (b - 1) : p.job_kd[job];
            s_yoff[pl] = ((long long)c * p.dy.plane_rows + p.dy.guard - (long long)kd * p.SS + (long long)h * p.KT) * 16;
        }
        for (int pl = lane; pl < nx; pl += 32) {
            const int h = pl % p.nh, c = (pl / p.nh) % p.CiC, f = pl / (p.nh * p.CiC);
            // tap of the fold (or the job's fixed tap) along the fold axis, and the accumulator axis start
            const int ft = (p.nfold > 1) ? (f - 1) : (p.fold_shift == 1 ? p.job_kw[job] : p.job_kh[job]);
            const int at = (p.nacc > 1) ? -1 : (p.acc_shift == 1 ? p.job_kw[job] : p.job_kh[job]);
            s_xoff[pl] = ((long long)(p.job_xch[job] / 8 + c) * p.x.plane_rows + p.x.guard + (long long)ft * p.fold_shift +
                          (long long)at * p.acc_shift + (long long)h * p.KT) * 16;
        }
        __syncwarp();
        const uint8_t* ysrc = reinterpret_cast<const uint8_t*>(p.dy.base);
        const uint8_t* xsrc = reinterpret_cast<const uint8_t*>(p.x.base);
        int s = 0; uint32_t ph = 0;
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, tb = 0, w_empty = 0, t_issue = 0;
        MARCH_PROF_T(tb);
        for (int i = 0; i < nst; ++i) {
            const long long r0b = (first_row + (long long)i * stage_rows) * 16;
            MARCH_PROF_T(t0);
            mbar_wait(&empty[s], ph ^ 1);
            MARCH_PROF_T(t1);
            w_empty += t1 - t0;
            if (elect_one()) {
                uint8_t* ybase = smem + (size_t)s * p.stage_bytes;
                uint8_t* xbase = ybase + y_bytes;
                if (warp == 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)ny * p.y_plane_bytes);
                    for (int pl = 0; pl < ny; ++pl)
                        bulk_load_1d(ybase + (size_t)pl * p.y_plane_bytes, ysrc + s_yoff[pl] + r0b, p.y_plane_bytes, &full[s]);
                } else {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)nx * p.x_plane_bytes);
                    for (int pl = 0; pl < nx; ++pl)
                        bulk_load_1d(xbase + (size_t)pl * p.x_plane_bytes, xsrc + s_xoff[pl] + r0b, p.x_plane_bytes, &full[s]);
                }
            }
            __syncwarp();
            MARCH_PROF_T(t2);
            t_issue += t2 - t1;
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (prof && lane == 0 && blockIdx.x < 160 && warp == 1) {
            g_march_prof[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - tb);
            g_march_prof[blockIdx.x * 16 + 1] = (unsigned long long)w_empty;
            g_march_prof[blockIdx.x * 16 + 2] = (unsigned long long)t_issue;
        }
    } else if (warp == 2) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc(p.M, p.Nmma, 1, 1);
        const uint32_t sbase16 = smem_u32(smem) >> 4;
        const uint32_t stage16 = p.stage_bytes >> 4, ybytes16 = y_bytes >> 4;
        // MN-major SWIZZLE_NONE: LBO = 128 B between 8-row K groups, SBO = plane stride between 8-channel blocks
        const uint64_t a_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((p.y_plane_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((p.x_plane_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
        const int ksteps = p.KT / 16;
        int s = 0; uint32_t ph = 0;
        const bool prof = B200_DBG(p, 256);
        long long t0 = 0, t1 = 0, t2 = 0, tb = 0, w_full = 0, t_issue = 0;
        MARCH_PROF_T(tb);
        for (int i = 0; i < nst; ++i) {
            MARCH_PROF_T(t0);
            mbar_wait(&full[s], ph);
            MARCH_PROF_T(t1);
            w_full += t1 - t0;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ya16 = sbase16 + s * stage16;
                const uint32_t xa16 = ya16 + ybytes16;
                uint32_t acc = i != 0;
                for (int t = 0; t < p.nacc; ++t) {
                    uint32_t a16 = ya16;
                    uint32_t b16 = xa16 + ((p.nacc > 1) ? (uint32_t)(t * p.acc_shift) : 0u);       // accumulator t = tap kw: X shifted by t rows
                    const uint32_t dtm = tmem_base + (uint32_t)(t * p.Nmma);
                    uint32_t acc_t = acc;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        umma_bf16(dtm, a_hi | (uint64_t)(a16 & 0x3FFF), b_hi | (uint64_t)(b16 & 0x3FFF), idesc, acc_t);
                        acc_t = 1u;
                        a16 += 16;      // 16 rows * 16 B, in 16-byte units
                        b16 += 16;
                    }
                }
            }
            __syncwarp();
            if (elect_one()) umma_commit(&empty[s]);
            __syncwarp();
            MARCH_PROF_T(t2);
            t_issue += t2 - t1;
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        pdl_trigger();      // every MMA is issued: the reduce kernel may become resident while the epilogues drain
        __syncwarp();
        if (prof && lane == 0 && blockIdx.x < 160) {
            g_march_prof[blockIdx.x * 16 + 3] = (unsigned long long)(clock64() - tb);
            g_march_prof[blockIdx.x * 16 + 4] = (unsigned long long)w_full;
            g_march_prof[blockIdx.x * 16 + 5] = (unsigned long long)t_issue;
            g_march_prof[blockIdx.x * 16 + 6] = (unsigned long long)nst;
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> fp32 partial =================
        const int ew = warp - 4;
        float* dst = p.partial + ((size_t)job * p.splits + split) * p.nacc * p.M * p.Nmma;
        int row; bool row_ok;
        if (p.M == 128) { row = ew * 32 + lane; row_ok = true; }
        else            { row = ew * 16 + lane; row_ok = lane < 16; }   // M=64: lanes 0-15 of each quadrant
        const bool prof = B200_DBG(p, 256);
        long long tb = 0, t1 = 0;
        MARCH_PROF_T(tb);
        if (nst > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
        }
        MARCH_PROF_T(t1);
        for (int t = 0; t < p.nacc; ++t) {
            for (int c0 = 0; c0 < p.Nmma; c0 += 16) {
                float v[16];
                if (nst > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(t * p.Nmma + c0), v);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (row_ok) {
                    float4* o = reinterpret_cast<float4*>(dst + ((size_t)t * p.M + row) * p.Nmma + c0);
                    // dual: 8-row / 8-column blocks alternate between the two row ranges; only blocks of the
                    // same range are products of matching rows, the others are never read
                    const bool lo = p.nh == 1 || ((row >> 3) & 1) == 0, hi = p.nh == 1 || ((row >> 3) & 1) == 1;
                    if (lo) {
                        o[0] = make_float4(v[0], v[1], v[2], v[3]);
                        o[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                    if (hi) {
                        o[2] = make_float4(v[8], v[9], v[10], v[11]);
                        o[3] = make_float4(v[12], v[13], v[14], v[15]);
                    }
                }
            }
        }
        if (prof && lane == 0 && ew == 0 && blockIdx.x < 160) {
            g_march_prof[blockIdx.x * 16 + 7] = (unsigned long long)(t1 - tb);          // wait for the last MMA
            g_march_prof[blockIdx.x * 16 + 8] = (unsigned long long)(clock64() - t1);   // TMEM -> global partial
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace b200

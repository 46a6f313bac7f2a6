// Line-marching weight-gradient GEMM for the 3x3x3 16 <-> 16 channel convs of level 0 (model.py:72-73, 336, 348;
// the weight half of aten::convolution_backward, train.py:210): six launches per training step, round 1's
// time-dominant kernel family.
//
//   dW[kd][kh][kw][co][ci] = sum_{n,d,h,w} dY[d][h][w][co] * X[d+kd-1][h+kh-1][w+kw-1][ci]
//
// ONE tcgen05.mma of M = 64 (48 used), N = 144 per 16 voxels covers all 27 taps (the linear-row kernel of
// wgrad_gemm.cuh issues three M=64/N=48 MMAs at the 46-cycle floor for the same work - 138 vs 72 cycles), and every
// operand byte crosses L2 -> shared memory about once instead of 3.1 times:
//
//   * A (M side) = dY line (d, h), three copies shifted by one voxel for the kw taps: blocks (kw, chunk) at one
//     uniform stride.  The three copies are three bulk copies of the same HBM line whose source / destination differ by
//     one 16-byte row (kw = 0: src + 16 B, kw = 2: dst + 16 B) into a ring of Ny expanded lines.  (Until round 2b the
//     line landed once in a raw ring and four warps made the shifted copies shared -> shared: no extra L2 traffic, but
//     24.6 KB of loads + stores per line on a shared-memory port that the MMA operand fetch already fills to 70 % - the
//     stage breakdown of profiles/r02_ab_wgrad_line_half.txt showed those copies, not the tensor pipe, bounding the
//     kernel.  The repeats hit L2.)
//   * B (N side) = the nine input lines (kd, kh) around (d, h): the band's lines of the slices d-1 .. d+1 live in a
//     shared-memory ring whose slot index is  q = 3*line + slice  (mod R), two chunk planes per slot, so that the nine
//     (kh, kd) lines of any output line are NINE CONSECUTIVE slots: blocks (kh, kd, chunk) at one uniform stride -
//     an N = 144 B operand straight out of the ring, no rotation between steps (the window slides by one slot per
//     slice and three per line).  Slot q of line L, slice s is the slot of line L+1, slice s-3: a line of slice d+2
//     is loaded as soon as line L+1 of slice d-1 has been used for the last time, which gives a prefetch lead of
//     about a slice.  The first 8 slots are mirrored behind the ring so that a 9-slot window never wraps.
//   * K runs over the W interior voxels of a line (W % 16 == 0); dY is zero on the halo voxels, X on its own.
//
// A CTA owns a band of LH lines of one sample and marches along D over a contiguous range of (sample, band, slice)
// units; all its MMAs accumulate into one 64 x 144 TMEM tile, written once as an fp32 partial [48][144];
// wgrad_line_reduce_kernel sums the CTAs in a fixed order (deterministic) into the PyTorch (Cout, Cin, 3, 3, 3)
// gradient.
//
// Warp roles: w0 X producer, w1 dY producer (cp.async.bulk + mbarrier expect_tx), w2 MMA issuer, w3 TMEM allocator,
// (then scout: waits for every step's operands and publishes a ready count the MMA warp polls), w4-7 epilogue.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kWglThreads = 256;
// Channel counts: dY has CY = 8 * NCHY channels (NCHY chunk planes per line), X has CX = 8 * NCHX:
//   <2, 2> (16 <-> 16, level 0): M = 64 (48 used), ONE MMA of N = 144 per 16 voxels (as described above);
//   <2, 1> (conv_input, model.py:336: 4 real input channels, the upper chunk of the packed input is zero by layout
//            contract): only chunk 0 of X is loaded, N = 72 - the MMA drops from 72 to the 46-cycle floor.  This launch is
//            the LAST kernel of the backward pass (nothing left to overlap it with): its time is step time;
//   <1, 2> (conv_output, model.py:348: 3 real output channels): only chunk 0 of dY is loaded and expanded (M = 24 of 64);
//   <4, 4> (32 <-> 32, level 1; opt-in): M = 128 (96 used), N = 288 exceeds one MMA, so the window is walked as THREE MMAs
//            of N = 96 = (kd 3) x 32 ci, one per kh (three consecutive ring slots each), into three column blocks of a
//            128 x 288 accumulator.  Same tensor time as the linear-row kernel (3 x 56 cycles per 16 voxels).
//   PAIR = 1 (round 2c; <2,2>, <2,1>, <1,2>): TWO dY lines (l, l+1) of a band per MMA.  M = (line 2) x (kw 3) x CY (96 of
//            128 rows at CY = 16), N = the window of the FOUR input lines l .. l+3 x (kd 3) x CX = 12 consecutive ring slots
//            (192 columns at CX = 16).  Row block j = 0 (line l) uses the column blocks of window lines 0..2, row block j = 1
//            (line l+1) those of window lines 1..3; the other quarter of the products lands in accumulator columns nobody
//            reads.  One N = 192 MMA costs 96 cycles for two lines (one N = 144 MMA per line: 72 each), and the
//            shared-memory port - the bound of the one-line form - carries A 4 KB + B 6 KB per two lines instead of
//            2 x (2 KB + 4.6 KB).  The epilogue writes the two row blocks as two partials [2][3 CY][9 CX] (columns
//            shifted back by 3 CX for j = 1), so the reduce kernel just sees twice as many CTAs.
template <int NCHY, int NCHX, int PAIR = 0> struct WglShape {
    static constexpr int CY = 8 * NCHY, CX = 8 * NCHX;
    static constexpr int Lines = PAIR ? 2 : 1;            // dY lines per MMA
    static constexpr int Win = PAIR ? 12 : 9;             // ring slots one B operand covers
    static constexpr int M = Lines * 3 * CY <= 64 ? 64 : 128;     // MMA rows: (line) x (kw 3) x CY, padded
    static constexpr int Rows = 3 * CY;                   // rows of one partial block
    static constexpr int Ntot = 9 * CX;                   // (kh 3) x (kd 3) x CX columns of one partial block
    static constexpr int Nacc = Win * CX;                 // accumulator columns
    static constexpr int Nmma = Nacc <= 256 ? Nacc : 3 * CX;   // columns of one MMA
    static constexpr int NM = Nacc / Nmma;                // MMAs per K step (1, or 3: one per kh)
    static constexpr int Slack = M / 8 - Lines * 3 * NCHY;     // planes past the last expanded dY line that an A operand touches
    static constexpr int Mirror = Win - 1;                // ring slots mirrored behind the ring: a window never wraps
    static constexpr int TmemCols = Nacc <= 128 ? 128 : (Nacc <= 256 ? 256 : 512);
    static_assert(!PAIR || (NCHY <= 2 && NCHX <= 2), "pair mode: 16-channel operands only");
};
constexpr int kWglM = 64;             // <2, 2> values, kept for the host code and the tests
constexpr int kWglN = 144;
constexpr int kWglRows = 48;
constexpr int kWglMirror = 8;         // ring slots mirrored behind the ring (window of 9 slots; pair mode: 11 for its window of 12)
constexpr int kWglNB = 64;            // X-load barriers (ring)
constexpr int kWglND = 16;            // step-done barriers (ring, power of two), >= LH + 2
constexpr int kWglMaxNy = 8;           // bound of both dY rings (raw lines, expanded lines)

struct WgradLineParams {
    int N, D, H, W, Wp;
    int LH, n_bands;
    long long units;                  // N * n_bands * D, slice fastest
    int ksteps;                       // W / 16
    int R;                            // ring slots (without the mirror): 3 * (LH + 2) + 1
    int NR;                           // (unused since round 2b: the raw dY ring is gone; kept for the plan-debug layout)
    int Ny;                           // expanded dY lines ((kw 3) x (chunk) planes each) in flight between L2 and the MMAs
    unsigned Lp;                      // bytes of one (line, chunk) plane: Wp * 16
    unsigned smem_x_off, smem_raw_off, smem_y_off, smem_bar_off;
    ActRef dy, x;
    float* partial;                   // [cta][48][144]
    int pair;                         // 1: two dY lines per MMA (WglShape<.., PAIR = 1>); every band then has an even line count
    int mirror;                       // ring slots mirrored behind the ring: 8, pair mode 11
    int debug;                        // B200_PROBES builds only: 1 no X copies, 2 no MMAs, 4 no shift copies, 8 no dY copies,
                                      // 256 cycle accounting of the MMA warp into g_wgl_prof
};

#ifdef B200_PROBES
__device__ unsigned long long g_wgl_prof[160 * 8];
#define WGL_DEBUG(bit) ((p.debug & (bit)) != 0)
#define WGL_CLOCK(var) var = clock64()
#else
#define WGL_DEBUG(bit) false
#define WGL_CLOCK(var)
#endif

struct WglSeg {
    int n, band, d0, len, nl;         // slices d0 .. d0+len-1 (interior, 0-based) of sample n; nl lines in the band
};
__device__ __forceinline__ int wgl_segment(const WgradLineParams& p, long long u, long long u_end, WglSeg& s) {
    const int d0 = (int)(u % p.D);
    long long t = u / p.D;
    s.band = (int)(t % p.n_bands);
    s.n = (int)(t / p.n_bands);
    s.d0 = d0;
    long long len = p.D - d0;
    if (len > u_end - u) len = u_end - u;
    s.len = (int)len;
    s.nl = min(p.LH, p.H - s.band * p.LH);
    return (int)len;
}

template <int NCHY, int NCHX, int PAIR = 0>
__global__ void __launch_bounds__(kWglThreads, 1)
wgrad_line_kernel(const __grid_constant__ WgradLineParams p) {
    using S = WglShape<NCHY, NCHX, PAIR>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x, ctas = gridDim.x;
    const long long u_begin = p.units * cta / ctas, u_end = p.units * (cta + 1) / ctas;

    uint8_t* smem_x = smem + p.smem_x_off;            // ring: [R + mirror slots][chunk 2][Wp rows][16 B]
    uint8_t* smem_y = smem + p.smem_y_off;            // expanded dY lines: [Ny][kw 3][chunk 2][Wp rows][16 B] (+ 2 planes of slack)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* x_full = bars;                          // [NB]  bulk copies of one X line pair -> MMA
    uint64_t* step_done = x_full + kWglNB;            // [ND]  MMA commit of one step -> both producers
    uint64_t* y_raw = step_done + kWglND;             // [MaxNy]  bulk copies of one dY line -> copy warps
    uint64_t* y_exp = y_raw + kWglMaxNy;              // [MaxNy]  copy warps (one arrival per warp) -> scout
    uint64_t* raw_free = y_exp + kWglMaxNy;           // [MaxNy]  copy warps (one arrival per warp) -> dY producer
    uint64_t* done = raw_free + kWglMaxNy;            // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    volatile int* ready = reinterpret_cast<volatile int*>(tmem_slot + 2);   // steps whose operands are in shared memory

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kWglNB; ++i) mbar_init(&x_full[i], 1);
        for (int i = 0; i < kWglND; ++i) mbar_init(&step_done[i], 1);
        for (int i = 0; i < kWglMaxNy; ++i) { mbar_init(&y_raw[i], 1); mbar_init(&y_exp[i], 1); mbar_init(&raw_free[i], 1); }
        mbar_init(done, 1);
        *ready = 0;
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, S::TmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long SS = (long long)(p.H + 2) * p.Wp;      // rows per padded slice
    const unsigned slot_bytes = (unsigned)NCHX * p.Lp;     // one X ring slot: NCHX chunk planes
    const unsigned yline_bytes = 3u * NCHY * p.Lp;         // one expanded dY line: (kw 3) x (chunk NCHY) planes

    if (warp == 0) {
        // ============ X producer: line pairs in (slice, line) order into ring slot (3 * line + slice) mod R ============
        const uint8_t* x0 = reinterpret_cast<const uint8_t*>(p.x.at(0, 0));
        const size_t xplane = (size_t)p.x.plane_rows * 16;     // bytes between the chunk planes of the tensor in HBM
        // All counters are 32-bit and advance incrementally (a 64-bit division costs ~100 cycles, and every cycle of this
        // loop delays the prefetch): kx = loads issued, t_base = steps before this segment, t_known = highest step known done.
        int kx = 0, t_base = 0, t_known = -1;
        for (long long u = u_begin; u < u_end;) {
            WglSeg sg;
            u += wgl_segment(p, u, u_end, sg);
            const int nl = sg.nl;
            if (t_base > 0 && t_known < t_base - 1) {      // the ring restarts: the previous segment must be consumed
                mbar_wait(&step_done[(t_base - 1) & (kWglND - 1)], (uint32_t)(((t_base - 1) / kWglND) & 1));
                t_known = t_base - 1;
            }
            // padded coordinates: slices d0 .. d0+len+1, lines band*LH .. band*LH + nl + 1
            const long long row00 = ((long long)sg.n * (p.D + 2) + sg.d0) * SS + (long long)(sg.band * p.LH) * p.Wp;
            int qs = 0;                                    // s mod R
            for (int s = 0; s < sg.len + 2; ++s) {
                const uint8_t* src0 = x0 + (row00 + (long long)s * SS) * 16;
                const int tn0 = t_base + (s - 3) * nl;
                int q = qs;                                // (3 * lam + s) mod R
                for (int lam = 0; lam < nl + 2; ++lam, ++kx) {
                    if (s >= 3) {
                        // the slot still holds line lam+1 of slice s-3, last read by step (centre slice s-2, line min(lam+1, nl-1))
                        const int tn = tn0 + min(lam + 1, nl - 1);
                        if (tn > t_known) {
                            mbar_wait(&step_done[tn & (kWglND - 1)], (uint32_t)((tn / kWglND) & 1));
                            t_known = tn;
                        }
                    }
                    if (elect_one()) {
                        const bool mirror = q < S::Mirror;
                        uint64_t* bar = &x_full[kx & (kWglNB - 1)];
                        if (WGL_DEBUG(1)) {
                            mbar_arrive(bar);
                        } else {
                            mbar_arrive_expect_tx(bar, (mirror ? 2u : 1u) * slot_bytes);
                            uint8_t* dst = smem_x + (size_t)q * slot_bytes;
#pragma unroll
                            for (int c = 0; c < NCHX; ++c) bulk_load_1d(dst + c * p.Lp, src0 + c * xplane, p.Lp, bar);
                            if (mirror) {
                                uint8_t* dm = dst + (size_t)p.R * slot_bytes;
#pragma unroll
                                for (int c = 0; c < NCHX; ++c) bulk_load_1d(dm + c * p.Lp, src0 + c * xplane, p.Lp, bar);
                            }
                        }
                    }
                    __syncwarp();
                    src0 += p.Lp;                          // next line of the slice: Wp rows of 16 B further
                    q += 3; if (q >= p.R) q -= p.R;
                }
                if (++qs == p.R) qs = 0;
            }
            t_base += sg.len * nl;
        }
    } else if (warp == 1) {
        // ============ dY producer: the three kw copies of one interior line per step, straight from L2 ============
        //   A_kw[r] = dY[r - kw + 1] for the rows r = 1 .. W the MMAs read: kw = 1 is the line itself, kw = 0 the line from
        //   its second row on, kw = 2 the line landed one row further (rows 0 / Wp-1 of those planes are never read)
        const uint8_t* y0 = reinterpret_cast<const uint8_t*>(p.dy.at(0, 0));
        const size_t yplane = (size_t)p.dy.plane_rows * 16;
        int t = 0, slot = 0;                               // lines issued; t mod Ny
        for (long long u = u_begin; u < u_end;) {
            WglSeg sg;
            u += wgl_segment(p, u, u_end, sg);
            for (int sd = 0; sd < sg.len; ++sd) {
                const long long row0 = ((long long)sg.n * (p.D + 2) + sg.d0 + 1 + sd) * SS + (long long)(sg.band * p.LH + 1) * p.Wp;
                const uint8_t* src0 = y0 + row0 * 16;
                for (int l = 0; l < sg.nl; ++l, ++t) {
                    if (t >= p.Ny) {                        // the slot is free once the MMAs of line t - Ny are done
                        const int tn = t - p.Ny;
                        mbar_wait(&step_done[tn & (kWglND - 1)], (uint32_t)((tn / kWglND) & 1));
                    }
                    if (elect_one()) {
                        uint64_t* bar = &y_exp[slot];
                        if (WGL_DEBUG(8)) {
                            mbar_arrive(bar);
                        } else {
                            const bool shifts = !WGL_DEBUG(4);
                            mbar_arrive_expect_tx(bar, (uint32_t)NCHY * (p.Lp + (shifts ? 2u * (p.Lp - 16u) : 0u)));
                            uint8_t* dst = smem_y + (size_t)slot * yline_bytes;
#pragma unroll
                            for (int c = 0; c < NCHY; ++c) {
                                const uint8_t* sc = src0 + c * yplane;
                                bulk_load_1d(dst + (size_t)(NCHY + c) * p.Lp, sc, p.Lp, bar);                      // kw = 1: dY[r]
                                if (shifts) {
                                    bulk_load_1d(dst + (size_t)c * p.Lp, sc + 16, p.Lp - 16u, bar);                // kw = 0: dY[r + 1]
                                    bulk_load_1d(dst + (size_t)(2 * NCHY + c) * p.Lp + 16, sc, p.Lp - 16u, bar);   // kw = 2: dY[r - 1]
                                }
                            }
                        }
                    }
                    __syncwarp();
                    src0 += p.Lp;
                    if (++slot == p.Ny) slot = 0;
                }
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer: W/16 x NM x (M, Nmma, K=16) per line ============
        constexpr uint32_t idesc = make_idesc(S::M, S::Nmma, 1, 1);
        // MN-major SWIZZLE_NONE: LBO = 128 B between 8-row K groups, SBO = Lp between consecutive 8-channel blocks
        const uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((p.Lp >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
        const uint32_t xbase16 = smem_u32(smem_x) >> 4, ybase16 = smem_u32(smem_y) >> 4;
        const uint32_t slot16 = slot_bytes >> 4, yline16 = yline_bytes >> 4;
        int t = 0, slot = 0, dslot = 0;                    // step count; t % Ny; t % ND
#ifdef B200_PROBES
        long long c0 = 0, c1 = 0, c3 = 0, w_x = 0, t_issue = 0, tb = clock64();
#endif
        for (long long u = u_begin; u < u_end;) {
            WglSeg sg;
            u += wgl_segment(p, u, u_end, sg);
            const int nl = sg.nl;
            int qs = 0;                                    // (sd - 1) mod R: ring slot of (line 0, slice sd-1)
            for (int sd = 1; sd <= sg.len; ++sd) {
                // Two lines per issue block: one poll, one elected region, 16 MMAs.  The scout warp (w3) has waited for
                // the operands of every step - the X lines up to line l+2 of slice sd+1 and the expanded dY line - and
                // published the count of ready steps: one shared-memory poll here instead of two to four mbarrier waits
                // (~100 cycles each even when complete).  The tcgen05 queue is shallow, so whatever this warp spends
                // between two MMAs the tensor pipe idles: no fences, divisions or 64-bit arithmetic in this loop.
                if constexpr (PAIR != 0) {
                    // pair mode: one MMA per K step covers the lines l, l+1 (A: two adjacent expanded dY lines, B: the 12-slot
                    // window of the input lines l .. l+3); two pairs per issue block.  nl is even (planner), the dY slot of a
                    // pair's first line is even and Ny is even, so the two lines are always adjacent in the ring.  Both
                    // lines' step_done barriers are committed behind the pair's last MMA: the producers' reuse rules (in
                    // units of lines) hold unchanged - a ring slot's last reader is the pair that contains the line the
                    // one-line rule names.
                    for (int l = 0; l < nl; l += 4) {
                        const int n4 = min(4, nl - l);
                        WGL_CLOCK(c0);
                        while (*ready < t + n4) __nanosleep(20);
                        WGL_CLOCK(c1);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (2 * j < n4) {
                                    int q0 = qs + 3 * (l + 2 * j);
                                    if (q0 >= p.R) q0 -= p.R;
                                    int sl = slot + 2 * j;
                                    if (sl >= p.Ny) sl -= p.Ny;
                                    const uint32_t a16 = ybase16 + (uint32_t)sl * yline16 + 1u;
                                    const uint32_t b16 = xbase16 + (uint32_t)q0 * slot16 + 1u;
                                    const uint32_t acc = (t + 2 * j) != 0;
                                    if (!WGL_DEBUG(2)) {
                                        if (p.ksteps == 8) {
#pragma unroll
                                            for (int ks = 0; ks < 8; ++ks)
                                                umma_bf16(tmem_base, hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                          hi | (uint64_t)((b16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        } else {
                                            for (int ks = 0; ks < p.ksteps; ++ks)
                                                umma_bf16(tmem_base, hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                          hi | (uint64_t)((b16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        }
                                    }
                                    umma_commit(&step_done[(dslot + 2 * j) & (kWglND - 1)]);
                                    umma_commit(&step_done[(dslot + 2 * j + 1) & (kWglND - 1)]);
                                }
                            }
                        }
                        __syncwarp();
                        t += n4;
                        slot += n4; if (slot >= p.Ny) slot -= p.Ny;
                        dslot = (dslot + n4) & (kWglND - 1);
#ifdef B200_PROBES
                        WGL_CLOCK(c3);
                        w_x += c1 - c0; t_issue += c3 - c1;
#endif
                    }
                } else {
                    for (int l = 0; l < nl; l += 2) {
                        const int n2 = min(2, nl - l);
                        WGL_CLOCK(c0);
                        while (*ready < t + n2) __nanosleep(20);      // not a hot spin: other kernels' warps share this SM
                        WGL_CLOCK(c1);
                        tc_fence_after();
                        if (WGL_DEBUG(2)) {
                            if (elect_one())
                                for (int j = 0; j < n2; ++j) umma_commit(&step_done[(dslot + j) & (kWglND - 1)]);
                        } else
                        if (elect_one()) {
    #pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (j < n2) {
                                    int q0 = qs + 3 * (l + j);
                                    if (q0 >= p.R) q0 -= p.R;
                                    int sl = slot + j;
                                    if (sl >= p.Ny) sl -= p.Ny;
                                    const uint32_t a16 = ybase16 + (uint32_t)sl * yline16 + 1u;     // row wp = 1 of the kw = 0 copy
                                    const uint32_t b16 = xbase16 + (uint32_t)q0 * slot16 + 1u;      // row wp = 1 of the window's first slot
                                    const uint32_t acc = (t + j) != 0;
                                    if (S::NM == 1) {
                                        if (p.ksteps == 8) {                              // W = 128: the level-0 lines of the benchmark
    #pragma unroll
                                            for (int ks = 0; ks < 8; ++ks)
                                                umma_bf16(tmem_base, hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                          hi | (uint64_t)((b16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        } else {
                                            for (int ks = 0; ks < p.ksteps; ++ks)             // 16 rows of 16 B per K step
                                                umma_bf16(tmem_base, hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                          hi | (uint64_t)((b16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        }
                                    } else {
                                        // one MMA per kh: the three slots (kd 0..2) of window row kh, column block kh
                                        const uint32_t kh16 = 3u * slot16;
                                        if (p.ksteps == 4) {                              // W = 64: the level-1 lines of the benchmark
    #pragma unroll
                                            for (int ks = 0; ks < 4; ++ks)
    #pragma unroll
                                                for (int kh = 0; kh < S::NM; ++kh)
                                                    umma_bf16(tmem_base + (uint32_t)(kh * S::Nmma), hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                              hi | (uint64_t)((b16 + kh * kh16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        } else {
                                            for (int ks = 0; ks < p.ksteps; ++ks)
    #pragma unroll
                                                for (int kh = 0; kh < S::NM; ++kh)
                                                    umma_bf16(tmem_base + (uint32_t)(kh * S::Nmma), hi | (uint64_t)((a16 + 16 * ks) & 0x3FFF),
                                                              hi | (uint64_t)((b16 + kh * kh16 + 16 * ks) & 0x3FFF), idesc, ks == 0 ? acc : 1u);
                                        }
                                    }
                                    umma_commit(&step_done[(dslot + j) & (kWglND - 1)]);
                                }
                            }
                        }
                        __syncwarp();
                        t += n2;
                        slot += n2; if (slot >= p.Ny) slot -= p.Ny;
                        dslot = (dslot + n2) & (kWglND - 1);
    #ifdef B200_PROBES
                        WGL_CLOCK(c3);
                        w_x += c1 - c0; t_issue += c3 - c1;
    #endif
                    }
                }
                if (++qs == p.R) qs = 0;
            }
        }
        if (elect_one()) umma_commit(done);
        pdl_trigger();      // every MMA is issued: the reduce kernel may become resident while the epilogues drain
        __syncwarp();
#ifdef B200_PROBES
        if (WGL_DEBUG(256) && lane == 0 && cta < 160) {
            g_wgl_prof[cta * 8 + 0] = (unsigned long long)(clock64() - tb);
            g_wgl_prof[cta * 8 + 1] = (unsigned long long)w_x;
            g_wgl_prof[cta * 8 + 2] = 0ull;
            g_wgl_prof[cta * 8 + 3] = (unsigned long long)t_issue;
            g_wgl_prof[cta * 8 + 4] = (unsigned long long)t;
        }
#endif
    } else if (warp == 3) {
        // ============ scout: waits for the operands of every step, publishes the count of ready steps ============
        // Plain volatile store of the count: the operands are in shared memory before their mbarrier completes, the
        // waits below observe that completion, and the store is issued after them.
        int t = 0, slot = 0, yph = 0, kx_waited = 0, seg_k0 = 0;
        for (long long u = u_begin; u < u_end;) {
            WglSeg sg;
            u += wgl_segment(p, u, u_end, sg);
            const int nl = sg.nl;
            for (int sd = 1; sd <= sg.len; ++sd) {
                for (int l = 0; l < nl; ++l) {
                    // X lines needed: everything up to line l+2 of slice sd+1 (loads arrive in (slice, line) order)
                    const int k_need = seg_k0 + (sd + 1) * (nl + 2) + l + 2;
                    while (kx_waited <= k_need) {
                        mbar_wait(&x_full[kx_waited & (kWglNB - 1)], (uint32_t)((kx_waited / kWglNB) & 1));
                        ++kx_waited;
                    }
                    mbar_wait(&y_exp[slot], (uint32_t)yph);
                    ++t;
                    if (lane == 0) *ready = t;
                    if (++slot == p.Ny) { slot = 0; yph ^= 1; }
                }
            }
            seg_k0 += (sg.len + 2) * (nl + 2);
        }
    } else if (warp >= 4) {
        // ============ epilogue: TMEM -> fp32 partial [Rows][Ntot] ============
        const int ew = warp - 4;
        // M = 64: rows 16*ew .. +15 live in lanes 0-15 of quadrant ew;  M = 128: row = TMEM lane
        const int row = S::M == 64 ? ew * 16 + lane : ew * 32 + lane;
        const bool row_ok = (S::M == 64 ? lane < 16 : true) && row < S::Lines * S::Rows;
        // pair mode: row block jb = 1 (the second line of a pair) owns the accumulator columns of window lines 1..3, i.e.
        // its partial's columns are the accumulator's shifted down by 3 CX; the two blocks are separate partials
        const int jb = row >= S::Rows ? 1 : 0;
        const int shift = jb * 3 * S::CX;
        float* dst = p.partial + (((size_t)cta * S::Lines + jb) * S::Rows + (row - jb * S::Rows)) * S::Ntot;
        const bool any = u_end > u_begin;
        if (any) {
            while (!mbar_try_wait(done, 0)) __nanosleep(256);     // idle for the whole main loop: do not take issue slots
            tc_fence_after();
        }
#pragma unroll
        for (int c0 = 0; c0 < S::Nacc; c0 += 16) {
            float v[16];
            if (any) {
                tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            if (row_ok) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                  // eight columns at a time (3 CX is a multiple of 8)
                    const int oc = c0 + 8 * hf - shift;
                    if (c0 + 8 * hf < S::Nacc && oc >= 0 && oc < S::Ntot) {
                        float4* o = reinterpret_cast<float4*>(dst + oc);
                        o[0] = make_float4(v[8 * hf + 0], v[8 * hf + 1], v[8 * hf + 2], v[8 * hf + 3]);
                        o[1] = make_float4(v[8 * hf + 4], v[8 * hf + 5], v[8 * hf + 6], v[8 * hf + 7]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem_base, S::TmemCols);
}

// dW[co][ci][kd][kh][kw] = sum over CTAs of P[cta][kw*CY + co][kh*3*CX + kd*CX + ci].
// One thread per (output float4 over ci, CTA group); groups are combined in a fixed order in shared memory.
struct WglReduceParams {
    int ctas, Cout_w, Cin_w, accumulate;
    int CY, CX;                                           // channels of the GEMM's two sides (8, 16 or 32 each)
};
constexpr int kWglReduceGroups = 32;
__global__ void __launch_bounds__(256)
wgrad_line_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, WglReduceParams q) {
    pdl_wait();         // (launched as a programmatic dependent of wgrad_line_kernel, common.cuh)
    __shared__ double s_acc[256][4];
    constexpr int QPB = 256 / kWglReduceGroups;           // output quads per CTA
    const int g = threadIdx.x / QPB, ql = threadIdx.x % QPB;
    const int quad = blockIdx.x * QPB + ql;               // (tap, co, ci/4): 27 * C * C/4 quads
    const int CYg = q.CY, CXg = q.CX, qpr = CXg / 4;      // quads per (tap, co) row
    const bool active = quad < 27 * CYg * qpr;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int tap = 0, co = 0, ci0 = 0;
    if (active) {
        ci0 = (quad % qpr) * 4; co = (quad / qpr) % CYg; tap = quad / (qpr * CYg);
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int Ntot = 9 * CXg;
        const float* src = partial + (size_t)(kw * CYg + co) * Ntot + kh * 3 * CXg + kd * CXg + ci0;
        const size_t cs = (size_t)(3 * CYg) * Ntot;
        // every thread's loads are issued together (the kernel is a chain of L2 latencies otherwise): groups of five
        int c = g;
        for (; c + 4 * kWglReduceGroups < q.ctas; c += 5 * kWglReduceGroups) {
            float4 v[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) v[i] = *reinterpret_cast<const float4*>(src + (size_t)(c + i * kWglReduceGroups) * cs);
#pragma unroll
            for (int i = 0; i < 5; ++i) { a0 += (double)v[i].x; a1 += (double)v[i].y; a2 += (double)v[i].z; a3 += (double)v[i].w; }
        }
        for (; c < q.ctas; c += kWglReduceGroups) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)c * cs);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
    }
    s_acc[threadIdx.x][0] = a0; s_acc[threadIdx.x][1] = a1; s_acc[threadIdx.x][2] = a2; s_acc[threadIdx.x][3] = a3;
    __syncthreads();
    if (g != 0 || !active) return;
    for (int k = 1; k < kWglReduceGroups; ++k) {
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

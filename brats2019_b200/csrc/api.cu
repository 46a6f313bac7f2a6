// Host side of libbrats_b200.so: launch planning, TMA tensor maps and the C ABI
// declared in include/brats_b200.h.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/brats_b200.h"
#include "conv_gemm.cuh"
#include "conv_pair.cuh"
#include "conv_march.cuh"
#include "conv_band.cuh"
#include "elementwise.cuh"
#include "elementwise2.cuh"
#include "wgrad_gemm.cuh"
#include "wgrad_march.cuh"
#include "wgrad_line.cuh"
#include "elementwise3.cuh"
#include "elementwise4.cuh"
#include "optimizer.cuh"

using namespace b200;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
#define CUDA_OK(x)                                                                        \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) return fail("%s failed: %s", #x, cudaGetErrorString(e_)); \
    } while (0)
#define LAUNCH_OK(name)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess) return fail("launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* b200_last_error(void) { return g_err; }

static const unsigned kMaxSmem = 227 * 1024;
// Per-device caches (one process may drive several GPUs: tests on cuda:1 after cuda:0, DataParallel-style use):
// the SM count sizes grids and partial-sum buffers, and cudaFuncAttributeMaxDynamicSharedMemorySize is a
// per-device function attribute.
static const int kMaxDevices = 64;
static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}
static int num_sms() {
    static int cache[kMaxDevices];
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return 148;   // B200; planning queries on a machine without a GPU
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cache[dev] = n;
        else { cudaGetLastError(); cache[dev] = 148; }
    }
    return cache[dev];
}
// Launch with a per-kernel scheduling priority (cudaLaunchAttributePriority; kept by CUDA-graph kernel nodes).  The
// block scheduler hands a freed SM slot to the pending CTA of the highest-priority kernel; running CTAs are never
// preempted.  In the training step's backward pass three classes of kernels are pending at once:
//   convs of the data-gradient chain  (critical path)                 -> kPrioConv   (highest)
//   weight-gradient GEMMs + their reduces (side stream)               -> kPrioWgrad
//   memory-bound kernels (GroupNorm backward, trilinear adjoint, ...) -> default (0, lowest)
// Without priorities a 2-CTA weight-gradient reduce queues behind the ~2000 not-yet-resident CTAs of a GroupNorm kernel
// launched a microsecond earlier (measured: 50 us of side-stream idle per level boundary).  B200_PRIO="conv wgrad"
// (e.g. "-2 -1") turns it on.  Measured: the priorities do move the weight-gradient kernels forward in the timeline, and
// the memory-bound kernels they then overlap slow down by the same amount - the backward pass is bound by the sum of the
// work, not by the launch order (5.504 vs 5.503 ms per step) - so the default is off.
static void launch_priorities(int& conv, int& wgrad) {
    static int c = 1, w = 1, init = 0;
    if (!init) {
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); least = greatest = 0; }
        c = 0; w = 0;            // measured (profiles/r02_ab_priorities.txt): no effect on the step, off by default
        if (const char* e = getenv("B200_PRIO")) sscanf(e, "%d %d", &c, &w);
        c = std::max(c, greatest); w = std::max(w, greatest);
        c = std::min(c, least); w = std::min(w, least);
        init = 1;
    }
    conv = c; wgrad = w;
}
static int prio_conv() { int c, w; launch_priorities(c, w); return c; }
static int prio_wgrad() { int c, w; launch_priorities(c, w); return w; }
// Programmatic dependent launch (common.cuh pdl_wait): B200_PDL=0 off, 1 (default) the conv kernels, 2 also the
// memory-bound kernels.  Measured (profiles/r02_ab_pdl.txt): 0 -> 1 gains 0.11 ms per step (the conv prologues - barrier
// init, TMEM allocation, first weight copies - run under the predecessor's tail); 1 -> 2 LOSES 0.28 ms: a memory-bound
// kernel has no prologue worth hiding, and its thousands of early-resident CTAs parked in griddepcontrol.wait sit on the
// SMs of the conv that is still running.
static int pdl_level() {
    int on = 1;
    if (const char* e = getenv("B200_PDL")) on = atoi(e);
    return on;
}
static bool pdl_enabled() { return pdl_level() >= 2; }      // memory-bound kernels
static bool pdl_conv() { return pdl_level() >= 1; }
// Memory-bound kernels with FEW CTAs (the deep levels: 32 - 600 CTAs) as dependents: there are no thousands of parked CTAs
// there, and each launch hides its ~2 us of launch latency under the predecessor's tail.  B200_PDL_EW_MAX = largest grid
// (in CTAs) that is launched this way at level 1 (0 = none).
static int pdl_ew_max() {
    int n = 0;
    if (const char* e = getenv("B200_PDL_EW_MAX")) n = atoi(e);
    return n;
}
static bool pdl_ew(long long grid_ctas) { return pdl_level() >= 2 || (pdl_level() >= 1 && grid_ctas <= pdl_ew_max()); }
// A conv launched as a programmatic dependent copies its weights BEFORE griddepcontrol.wait (they "were packed long
// before").  That is false for the first conv enqueued on a stream right after a weight-packing kernel: it is launched
// with full serialization instead.  (Host-side launch order per stream is what counts, also under stream capture.)
static thread_local void* g_pack_stream = reinterpret_cast<void*>(-1);
static void note_pack_launch(void* stream) { g_pack_stream = stream; }
static bool pdl_conv_on(cudaStream_t st) {
    if (g_pack_stream == (void*)st) {
        g_pack_stream = reinterpret_cast<void*>(-1);
        return false;
    }
    return pdl_conv();
}
// the reduce kernel behind a weight-gradient GEMM (side stream) as a dependent: measured neutral (5.405 / 5.437 vs
// 5.410 ms), off by default; B200_PDL_WGRAD=1 turns it on
static bool pdl_wgrad() {
    int on = 0;
    if (const char* e = getenv("B200_PDL_WGRAD")) on = atoi(e);
    return pdl_level() >= 1 && on;
}
// the few-CTA finalize kernels between a conv / reduction and its apply pass are dependents too (their handful of parked
// CTAs cost nothing): -0.02 ms per step; B200_PDL_SMALL=0 turns that off
static bool pdl_small() {
    int on = 1;
    if (const char* e = getenv("B200_PDL_SMALL")) on = atoi(e);
    return pdl_level() >= 2 || (pdl_level() >= 1 && on);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int prio, bool pdl,
                             Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (prio != 0) { attr[na].id = cudaLaunchAttributePriority; attr[na].val.priority = prio; ++na; }
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_prio(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int prio,
                               Args&&... args) {
    return launch_ex(kernel, grid, block, smem, st, prio, false, std::forward<Args>(args)...);
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device)
#define SET_MAX_SMEM_ONCE(...)                                                                                     \
    do {                                                                                                           \
        static bool done_[kMaxDevices];                                                                            \
        const int dev_ = current_device();                                                                         \
        if (dev_ < 0 || dev_ >= kMaxDevices || !done_[dev_]) {                                                     \
            CUDA_OK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem)); \
            if (dev_ >= 0 && dev_ < kMaxDevices) done_[dev_] = true;                                               \
        }                                                                                                          \
    } while (0)
extern "C" int b200_num_sms(void) { return num_sms(); }

extern "C" int b200_device_check(int dev) {
    int major = 0, minor = 0;
    CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) return fail("device %d is sm_%d%d; this library contains sm_100a code only", dev, major, minor);
    return 0;
}

// ---------------------------------------------------------------------------------------
// activation geometry (chunk-planar zero-halo layout, common.cuh)
// ---------------------------------------------------------------------------------------
extern "C" long long b200_act_guard_rows(int D, int H, int W) {
    Vol v{1, D, H, W};
    return v.guard_rows();
}
extern "C" long long b200_act_plane_rows(int N, int D, int H, int W) {
    Vol v{N, D, H, W};
    return v.plane_rows();
}
static int check_ptr16(const void* p, const char* what) {
    if (((uintptr_t)p & 15) != 0) return fail("%s pointer not 16-byte aligned", what);
    return 0;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline unsigned align_up(unsigned a, unsigned b) { return (a + b - 1) / b * b; }
static unsigned pow2_cols(unsigned c) {
    unsigned p = 32;
    while (p < c) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------
// conv planning
// ---------------------------------------------------------------------------------------
// kw-fold (conv_gemm.cuh): 3x3x3, bf16 epilogue, Cout 16 or 32 -> GEMM N = 3*Cout
static bool conv_fold(const b200_conv_desc* d) {
    static int no_fold = -1;
    if (no_fold < 0) { const char* e = getenv("B200_NO_FOLD"); no_fold = (e && atoi(e)) ? 1 : 0; }
    return !no_fold && d->mode == MODE_K3 && d->epi == EPI_BF16 && (d->Cout == 16 || d->Cout == 32);
}
static int conv_nmma(const b200_conv_desc* d) {
    if (conv_fold(d)) return 3 * d->Cout;
    return d->Cout > 256 ? 256 : d->Cout;
}
// CTA-pair kernel (conv_pair.cuh, tcgen05.mma.cta_group::2): 3x3x3, 128 output channels, bf16 epilogue.
// OPT-IN (B200_CONV_PAIR=1).  Bit-identical to conv_gemm and SLOWER (profiles/r02_conv_pair_ab.txt: 128 -> 128 at
// 2 x 16^3 23.6 vs 20.0 us, at 8 x 16^3 60.8 vs 49.0 us): the deep-level conv is not bound by the weight stream the pair
// halves.  tools/conv_deep_probe.py: weight stream alone 8.5 us, MMAs alone 7.2 us (64 cycles each, tools/umma_probe),
// together 20.4 us - the two do not overlap: at M = N = 128 the operand fetch of the MMA (8 KB per 64 cycles) takes the
// whole shared-memory bandwidth and the incoming bulk copies take their turn on the same port.
static bool conv_pair(const b200_conv_desc* d) {
    int on = 0;
    if (const char* e = getenv("B200_CONV_PAIR")) on = atoi(e);
    return on && d->mode == MODE_K3 && d->epi == EPI_BF16 && d->Cout == 128 && d->Cin_b == 0 && !conv_fold(d);
}
static int conv_kc(const b200_conv_desc* d) {
    if (d->mode == MODE_K3) return 16;
    for (int kc : {64, 32, 16})
        if (d->Cin_a % kc == 0 && d->Cin_b % kc == 0) return kc;
    return 0;
}

static int check_conv_desc(const b200_conv_desc* d) {
    if (!d) return fail("null conv desc");
    if (d->mode != MODE_K3 && d->mode != MODE_K1) return fail("conv mode %d unsupported", d->mode);
    if (d->N < 1 || d->D < 1 || d->H < 1 || d->W < 1) return fail("bad volume %dx%dx%dx%d", d->N, d->D, d->H, d->W);
    if (d->Cin_a < 16 || d->Cin_a % 16 || d->Cin_b < 0 || d->Cin_b % 16)
        return fail("Cin_a=%d / Cin_b=%d must be multiples of 16", d->Cin_a, d->Cin_b);
    const int n = d->Cout > 256 ? 256 : d->Cout;
    if (!(n == 16 || n == 32 || n == 64 || n == 128 || n == 256) || d->Cout % n)
        return fail("Cout=%d unsupported (16/32/64/128/256 or a multiple of 256)", d->Cout);
    if (d->mode == MODE_K3 && (d->Cout > 128 || d->Cin_b != 0)) return fail("k3 conv: Cout<=128 and one source only");
    if (d->epi == EPI_SIGMOID && (d->mode != MODE_K3 || d->Cout != 16)) return fail("sigmoid epilogue needs k3, Cout=16");
    if (d->epi == EPI_D2S && (d->mode != MODE_K1 || d->Cout % 128 || d->Cin_b != 0 ||
                              ((d->Cout / 64) & (d->Cout / 64 - 1)) != 0))
        return fail("depth-to-space epilogue needs a 1x1x1 GEMM with Cout = 8 * Cf, Cf a power of two >= 16");
    if (d->epi < EPI_BF16 || d->epi > EPI_D2S) return fail("conv epilogue %d unsupported", d->epi);
    if (d->W + 2 > 4000) return fail("W too large");
    return 0;
}

static int plan_conv(const b200_conv_desc* d, ConvKParams& p) {
    if (check_conv_desc(d)) return 1;
    memset(&p, 0, sizeof(p));
    const bool pair = conv_pair(d);
    const int Nfull = conv_nmma(d);                    // GEMM N of one MMA (accumulator columns per run)
    const int Nm = pair ? Nfull / 2 : Nfull;           // columns of one CTA's B operand = one packed-weight job
    p.pair = pair ? 1 : 0;
    p.Nm_pack = Nm;
    p.N = d->N; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Wp = d->W + 2;
    p.SS = (d->H + 2) * p.Wp;
    p.by_SS = make_fastdiv((unsigned)p.SS);
    p.by_Wp = make_fastdiv((unsigned)p.Wp);
    p.sample_rows = (long long)(d->D + 2) * p.SS;
    p.total_rows = p.sample_rows * d->N;
    p.mode = d->mode;
    const bool fold = conv_fold(d);
    const int RB = fold ? 126 : 128;
    p.n_jobs = fold ? 1 : d->Cout / Nm;
    p.KC = conv_kc(d);
    if (p.KC == 0) return fail("no K chunk size for Cin_a=%d Cin_b=%d", d->Cin_a, d->Cin_b);
    p.KGa = d->Cin_a / p.KC;
    p.KG = p.KGa + d->Cin_b / p.KC;
    p.Cout_total = d->Cout;
    const int sms = num_sms();
    const unsigned bar_bytes = kConvTailBytes + 128;

    if (d->mode == MODE_K3) {
        p.NTG = 3; p.TG = fold ? 3 : 9;
        p.w_stage_bytes = (unsigned)(p.TG * p.KC * Nm * 2);
        double best_cost = 1e300;
        int bBD = 0, bMB = 0, bWhole = 0;
        int fBD = 0, fMB = 0, fWhole = -1;                // diagnostics: B200_CONV_TILE="BD MB whole" forces a tile shape
        if (const char* e = getenv("B200_CONV_TILE")) sscanf(e, "%d %d %d", &fBD, &fMB, &fWhole);
        for (int whole = 0; whole < 2; ++whole)
            for (int BD : {1, 2, 4}) {
                if (whole && BD != 1) continue;
                if (BD > d->D || d->D % BD) continue;     // no ragged slice groups
                for (int MB : {1, 2, 4}) {
                    if (fBD && (BD != fBD || MB != fMB || (fWhole >= 0 && whole != fWhole))) continue;
                    const int R = BD * MB;
                    if (2 * R * Nfull > 512) continue;
                    const int TR = RB * MB;
                    const int SR = (MB - 1) * RB + 128 + 2 * p.Wp + (fold ? 0 : 2);
                    const int SRp = (SR + 7) / 8 * 8;
                    const unsigned xst = 2u * (BD + 2) * SRp * 16;
                    if (2 * xst + 2 * p.w_stage_bytes + bar_bytes > kMaxSmem) continue;
                    long long QN, tiles;
                    if (whole) {
                        QN = (long long)(d->D - 1) * p.SS + (long long)(d->H - 1) * p.Wp + d->W;
                        tiles = (long long)d->N * ceil_div(QN, TR);
                    } else {
                        QN = (long long)(d->H - 1) * p.Wp + d->W;
                        tiles = (long long)d->N * ceil_div(d->D, BD) * ceil_div(QN, TR);
                    }
                    const long long ctas = std::min<long long>(sms, tiles);
                    const double amp = (double)(BD + 2) * SRp / ((double)BD * TR);
                    const double cost = (double)ceil_div(tiles, ctas) * BD * TR * (1.0 + 0.04 * amp);
                    if (cost < best_cost) { best_cost = cost; bBD = BD; bMB = MB; bWhole = whole; }
                }
            }
        if (bBD == 0) return fail("k3 conv: no tile shape fits (W=%d, Cout=%d)", d->W, d->Cout);
        p.BD = bBD; p.MB = bMB; p.whole = bWhole;
        p.TR = RB * p.MB;
        const int SR = (p.MB - 1) * RB + 128 + 2 * p.Wp + (fold ? 0 : 2);
        p.SRp = (SR + 7) / 8 * 8;
        p.nslices = p.BD + 2;
        p.halo_rows = p.Wp + 1;
        p.x_plane_bytes = (unsigned)p.nslices * p.SRp * 16;
        p.x_stage_bytes = (unsigned)(p.KC / 8) * p.x_plane_bytes;
        if (p.whole) {
            p.Q0 = p.SS + p.Wp + 1;
            p.QN = (d->D - 1) * p.SS + (d->H - 1) * p.Wp + d->W;
            p.tiles_d = 1;
        } else {
            p.Q0 = p.Wp + 1;
            p.QN = (d->H - 1) * p.Wp + d->W;
            p.tiles_d = ceil_div(d->D, p.BD);
        }
        p.tiles_q = ceil_div(p.QN, p.TR);
        if (fold) {
            for (int t = 0; t < 9; ++t) p.tap_off[t] = (t / 3) * p.SRp + (t % 3) * p.Wp;     // (kd, kh); kw lives in N
        } else {
            for (int t = 0; t < 27; ++t) p.tap_off[t] = (t / 9) * p.SRp + ((t / 3) % 3) * p.Wp + (t % 3);
        }
    } else {
        p.NTG = 1; p.TG = 1;
        p.w_stage_bytes = (unsigned)(p.KC * Nm * 2);
        // rows per tile: as many 128-row blocks as TMEM allows (2 accumulator stages) while every SM still
        // gets >= 2 tiles: these GEMMs are HBM-bound and need bytes in flight, not tensor throughput
        p.MB = 1;
        for (int MB : {4, 2}) {
            if (2 * MB * Nm > 512) continue;
            if (ceil_div(p.total_rows, 128 * MB) >= 2 * sms) { p.MB = MB; break; }
        }
        p.BD = 1; p.whole = 0;
        p.TR = 128 * p.MB;
        p.SRp = p.TR;
        p.nslices = 1; p.halo_rows = 0;
        p.x_plane_bytes = (unsigned)p.TR * 16;
        p.x_stage_bytes = (unsigned)(p.KC / 8) * p.x_plane_bytes;
        p.Q0 = 0; p.QN = 0; p.tiles_d = 1;
        p.tiles_q = ceil_div(p.total_rows, p.TR);
        p.tap_off[0] = 0;
    }
    p.num_tiles = (p.mode == MODE_K3 ? p.N * p.tiles_d : 1) * p.tiles_q;
    // stages: as many as fit, up to 4; small filters stay resident (a weight stage round trip is
    // ~1.3 us, longer than the MMAs that consume it when Cout <= 32)
    unsigned budget = kMaxSmem - bar_bytes;
    p.x_stages = 2; p.w_stages = 2;
    const unsigned w_total = (unsigned)(p.KG * p.NTG) * p.w_stage_bytes;
    if (p.KG * p.NTG <= 8 && w_total <= 64 * 1024 && align_up(2 * p.x_stage_bytes, 128) + w_total <= budget) {
        p.w_resident = 1;
        p.w_stages = p.KG * p.NTG;
    }
    auto used = [&]() { return align_up(p.x_stages * p.x_stage_bytes, 128) + p.w_stages * p.w_stage_bytes; };
    if (used() > budget) return fail("conv: shared memory plan does not fit (%u bytes)", used());
    // k3: up to 4 stages; k1 (HBM-bound, small stages): up to 8 so that >= ~100 KB are in flight per SM
    // pair: half-size weight stages, up to 8 of them (the ring must cover the L2 round trip at the MMA rate)
    const int rounds = (d->mode == MODE_K1 || pair) ? 6 : 2;
    for (int round = 0; round < rounds; ++round) {
        if (!p.w_resident && (round < 2 || pair)) { ++p.w_stages; if (used() > budget) --p.w_stages; }
        if (pair && round >= 2) continue;              // activations: 4 stages
        ++p.x_stages; if (used() > budget) --p.x_stages;
    }
    p.smem_x_off = 0;
    p.smem_w_off = align_up(p.x_stages * p.x_stage_bytes, 128);
    p.smem_bar_off = align_up(p.smem_w_off + p.w_stages * p.w_stage_bytes, 16);
    p.tmem_cols = pow2_cols(2u * p.BD * p.MB * Nfull);
    if (p.tmem_cols > 512) return fail("conv: TMEM plan does not fit");
    return 0;
}


// ---------------------------------------------------------------------------------------
// marching conv (conv_march.cuh): 3x3x3, Cin and Cout in {16, 32}
// ---------------------------------------------------------------------------------------
static bool march_enabled() {
    static int off = -1;
    if (off < 0) { const char* e = getenv("B200_NO_MARCH"); off = (e && atoi(e)) ? 1 : 0; }
    return !off;
}
// returns 0 and fills p when the marching kernel applies to this conv, 1 otherwise (no error text)
static int plan_march(const b200_conv_desc* d, MarchParams& p) {
    if (!march_enabled() || d->mode != MODE_K3 || d->Cin_b != 0) return 1;
    if (!(d->Cin_a == 16 || d->Cin_a == 32) || !(d->Cout == 16 || d->Cout == 32)) return 1;
    if (d->epi == EPI_SIGMOID && d->Cout != 16) return 1;
    memset(&p, 0, sizeof(p));
    const int CO = d->Cout;
    p.N = d->N; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Wp = d->W + 2;
    p.SS = (d->H + 2) * p.Wp;
    p.by_Wp = make_fastdiv((unsigned)p.Wp);
    p.Q0 = p.Wp + 1;
    p.QN = (d->H - 1) * p.Wp + d->W;
    p.KS = d->Cin_a / 16;
    p.wtile_bytes = 2u * 5u * CO * 16u;
    p.w_bytes = (unsigned)p.KS * 9u * p.wtile_bytes;
    const int nb = ceil_div(p.QN, 128);
    const int sms = num_sms();
    const int maxMB = std::min(kMarchMaxMB, 512 / (3 * CO));
    double best = 1e300;
    int bMB = 0, bslots = 0;
    for (int MB = 1; MB <= maxMB; ++MB) {
        const int TR = 128 * MB;
        const int SR = TR + 2 * p.Wp + 2;
        const int SRp = (SR + 7) / 8 * 8;
        const unsigned slot = (unsigned)(d->Cin_a / 8) * SRp * 16u;
        if (TR > p.SS + 1000) continue;                       // guard rows cover at most SS + 2*Wp + 1024 of overrun
        if (slot >= (1u << 20)) continue;
        const unsigned avail = kMaxSmem - kMarchTailBytes - 1024 - p.w_bytes;
        int slots = (int)std::min<unsigned>(kMarchMaxSlots, avail / slot);
        if (slots < 3) continue;
        const long long strips = ceil_div(nb, MB);
        const long long units = (long long)d->N * strips * d->D;
        const double upc = (double)ceil_div(units, std::min<long long>(sms, units));
        const double steps = upc + 2.0 * (1.0 + upc / d->D);
        // cycles per step: MMA time of the MB blocks (operand-fetch bound: 46 cycles at N = 48, 56 at N = 96),
        // stretched by the shared-memory writes of the halo, plus the per-step hand-shake; with a single block
        // the next step's MMAs additionally wait for the epilogue to drain it (no interleaving)
        const double t_block = 9.0 * p.KS * (CO == 16 ? 46.0 : 56.0);
        const double cost = steps * (MB * t_block * (1.0 + 0.12 * ((double)SR / TR - 1.0)) + (MB == 1 ? 400.0 : 100.0));
        if (cost < best) { best = cost; bMB = MB; bslots = slots; }
    }
    if (bMB == 0) return 1;
    p.MB = bMB; p.TR = 128 * bMB;
    p.SRp = (p.TR + 2 * p.Wp + 2 + 7) / 8 * 8;
    p.plane_bytes = (unsigned)p.SRp * 16u;
    p.slot_bytes = (unsigned)(d->Cin_a / 8) * p.plane_bytes;
    p.nslots = std::min(bslots, 4);
    p.n_strips = ceil_div(nb, p.MB);
    p.units = (long long)d->N * p.n_strips * d->D;
    p.smem_x_off = 0;
    p.smem_w_off = align_up((unsigned)p.nslots * p.slot_bytes, 128);
    p.smem_bar_off = align_up(p.smem_w_off + p.w_bytes, 16);
    p.tmem_cols = pow2_cols((unsigned)(p.MB * 3 * CO));
    return 0;
}
static int march_ctas(const MarchParams& p) { return (int)std::min<long long>(num_sms(), p.units); }


// ---------------------------------------------------------------------------------------
// band-marching conv (conv_band.cuh): 3x3x3, 16 -> 16 channels, kd and kh folded (N = 144)
// ---------------------------------------------------------------------------------------
static bool band_enabled() {
    static int off = -1;
    if (off < 0) { const char* e = getenv("B200_NO_BAND"); off = (e && atoi(e)) ? 1 : 0; }
    return !off;
}
static int plan_band(const b200_conv_desc* d, BandParams& p) {
    if (!band_enabled() || !march_enabled() || d->mode != MODE_K3 || d->Cin_b != 0 || d->Cin_a != 16 || d->Cout != 16) return 1;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.D = d->D; p.H = d->H; p.W = d->W;
    p.Wp = d->W + 2;
    p.SS = (d->H + 2) * p.Wp;
    p.n_bands = ceil_div(d->H, kBandBH);
    p.n_cols = ceil_div(d->W, 128);
    // useful fraction of the MMA rows / lines; below ~0.6 the plain marching kernel is the better choice
    const double eff = ((double)d->W / (128.0 * p.n_cols)) * ((double)d->H / (kBandBH * p.n_bands));
    if (eff < 0.6) return 1;
    p.units = (long long)d->N * p.n_bands * p.n_cols * d->D;
    p.plane_bytes = (unsigned)kBandLines * kBandLineRows * 16u;
    p.slot_bytes = 2u * p.plane_bytes;
    p.wimg_bytes = 2u * 144u * 16u;
    p.w_bytes = 9u * p.wimg_bytes;
    const unsigned avail = kMaxSmem - kBandTailBytes - 1024 - p.w_bytes;
    p.nslots = (int)std::min<unsigned>(4, avail / p.slot_bytes);
    if (p.nslots < 3) return 1;
    p.smem_x_off = 0;
    p.smem_w_off = align_up((unsigned)p.nslots * p.slot_bytes, 128);
    p.smem_bar_off = align_up(p.smem_w_off + p.w_bytes, 16);
    return 0;
}
static int band_ctas(const BandParams& p) { return (int)std::min<long long>(num_sms(), p.units); }

static int conv_grid_ctas(const ConvKParams& p) {
    if (p.pair) return 2 * std::min(std::max(1, num_sms() / 2), (p.num_tiles + 1) / 2);    // whole CTA pairs
    int per_job = std::max(1, num_sms() / p.n_jobs);
    return std::min(per_job, p.num_tiles);
}

extern "C" size_t b200_conv_packed_weight_bytes(const b200_conv_desc* d) {
    BandParams bp;
    if (check_conv_desc(d) == 0 && plan_band(d, bp) == 0) return bp.w_bytes;
    MarchParams mp;
    if (check_conv_desc(d) == 0 && plan_march(d, mp) == 0) return mp.w_bytes;
    ConvKParams p;
    if (plan_conv(d, p)) return 0;
    return (size_t)p.n_jobs * p.KG * p.NTG * p.w_stage_bytes;
}
extern "C" int b200_conv_ctas(const b200_conv_desc* d) {
    BandParams bp;
    if (check_conv_desc(d) == 0 && plan_band(d, bp) == 0) return band_ctas(bp);
    MarchParams mp;
    if (check_conv_desc(d) == 0 && plan_march(d, mp) == 0) return march_ctas(mp);
    ConvKParams p;
    if (plan_conv(d, p)) return -1;
    return conv_grid_ctas(p);
}

// fills the layout parameters of one packing job (no launch); J.w / J.packed are set by the caller
static int make_pack_job(const b200_conv_desc* d, int kind, int Cout_w, int Cin_w, int taps_w, int ci_off, int K_real,
                         int N_real, PackJobDev& J, size_t& vectors) {
    if (check_conv_desc(d)) return 1;
    if (kind < 0 || kind > 3) return fail("bad weight kind %d", kind);
    memset(&J, 0, sizeof(J));
    {
        BandParams bp;
        if (plan_band(d, bp) == 0) {
            if (kind > B200_W_DGRAD || taps_w != 27) return fail("band conv: 3x3x3 forward / data-gradient weights only");
            if (K_real > 16 || N_real > 16) return fail("K_real/N_real exceed the GEMM extents");
            MarchPackParams& q = J.mq;
            q.kind = kind; q.Cout_w = Cout_w; q.Cin_w = Cin_w; q.ci_off = ci_off; q.K_real = K_real; q.N_real = N_real;
            q.KS = 1; q.CO = 16;
            J.layout = 2;
            vectors = 9u * 2u * 144u;
            return 0;
        }
    }
    {
        MarchParams mp;
        if (plan_march(d, mp) == 0) {
            if (kind > B200_W_DGRAD || taps_w != 27) return fail("marching conv: 3x3x3 forward / data-gradient weights only");
            if (K_real > d->Cin_a || N_real > d->Cout) return fail("K_real/N_real exceed the GEMM extents");
            MarchPackParams& q = J.mq;
            q.kind = kind; q.Cout_w = Cout_w; q.Cin_w = Cin_w; q.ci_off = ci_off; q.K_real = K_real; q.N_real = N_real;
            q.KS = mp.KS; q.CO = d->Cout;
            J.layout = 1;
            vectors = (size_t)q.KS * 9 * 2 * 5 * q.CO;
            return 0;
        }
    }
    ConvKParams p;
    if (plan_conv(d, p)) return 1;
    const int taps_g = (d->mode == MODE_K3) ? 27 : 1;
    if ((kind == B200_W_FWD || kind == B200_W_DGRAD) && taps_w != taps_g) return fail("taps mismatch %d vs %d", taps_w, taps_g);
    if ((kind == B200_W_FWD_S2D || kind == B200_W_DGRAD_S2D) && (taps_w != 8 || d->mode != MODE_K1))
        return fail("s2d weight kinds need a k1 GEMM and a 2x2x2 kernel");
    if (K_real > d->Cin_a + d->Cin_b || N_real > d->Cout) return fail("K_real/N_real exceed the GEMM extents");
    PackParams& q = J.q;
    q.kind = kind; q.Cout_w = Cout_w; q.Cin_w = Cin_w; q.taps_w = taps_w; q.ci_off = ci_off;
    q.K_real = K_real; q.N_real = N_real;
    q.n_jobs = p.n_jobs; q.KG = p.KG; q.NTG = p.NTG; q.TG = p.TG; q.KC = p.KC; q.Nmma = p.Nm_pack;
    q.fold = conv_fold(d) ? 1 : 0;
    if (q.fold && kind > B200_W_DGRAD) return fail("fold applies to 3x3x3 weights only");
    J.layout = 0;
    vectors = pack_weight_total(q);
    return 0;
}

extern "C" int b200_conv_pack_weight(const b200_conv_desc* d, int kind, const float* w, int Cout_w, int Cin_w,
                                     int taps_w, int ci_off, int K_real, int N_real, void* packed, void* stream) {
    PackJobDev J;
    size_t vectors = 0;
    if (make_pack_job(d, kind, Cout_w, Cin_w, taps_w, ci_off, K_real, N_real, J, vectors)) return 1;
    const int blocks = (int)((vectors + 255) / 256);
    if (J.layout == 2) {
        pack_weight_band_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, J.mq);
        LAUNCH_OK("pack_weight_band_kernel");
        note_pack_launch(stream);
    } else if (J.layout == 1) {
        pack_weight_march_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, J.mq);
        LAUNCH_OK("pack_weight_march_kernel");
        note_pack_launch(stream);
    } else {
        pack_weight_kernel<<<std::min(blocks, 1024), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, J.q);
        LAUNCH_OK("pack_weight_kernel");
        note_pack_launch(stream);
    }
    return 0;
}

// ---- all weights in one launch -----------------------------------------------------------------
extern "C" size_t b200_pack_table_entry_bytes(void) { return sizeof(PackJobDev); }
extern "C" int b200_pack_table_build(const b200_pack_job* jobs, int n_jobs, void* table_host, size_t table_bytes,
                                     int* total_blocks) {
    if (!jobs || n_jobs < 1 || !table_host || !total_blocks) return fail("pack_table_build: bad arguments");
    if (table_bytes < (size_t)n_jobs * sizeof(PackJobDev)) return fail("pack_table_build: table buffer too small");
    PackJobDev* T = (PackJobDev*)table_host;
    int block0 = 0;
    for (int j = 0; j < n_jobs; ++j) {
        size_t vectors = 0;
        if (make_pack_job(&jobs[j].desc, jobs[j].kind, jobs[j].Cout_w, jobs[j].Cin_w, jobs[j].taps_w, jobs[j].ci_off,
                          jobs[j].K_real, jobs[j].N_real, T[j], vectors))
            return 1;
        if (check_ptr16(jobs[j].packed, "packed weights")) return 1;
        T[j].w = jobs[j].w;
        T[j].packed = (__nv_bfloat16*)jobs[j].packed;
        T[j].block0 = block0;
        T[j].nblocks = (int)((vectors + 255) / 256);
        block0 += T[j].nblocks;
    }
    *total_blocks = block0;
    return 0;
}
extern "C" int b200_pack_table_run(const void* table_device, int n_jobs, int total_blocks, void* stream) {
    if (!table_device || n_jobs < 1 || total_blocks < 1) return fail("pack_table_run: bad arguments");
    pack_weights_batched_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>((const PackJobDev*)table_device, n_jobs);
    LAUNCH_OK("pack_weights_batched_kernel");
    note_pack_launch(stream);
    return 0;
}

template <int MODE, int EPI, int NM, int FOLD>
static int launch_conv(const ConvKParams& p, unsigned smem, int grid, cudaStream_t st) {
    SET_MAX_SMEM_ONCE(conv_gemm_kernel<MODE, EPI, NM, FOLD>);
    CUDA_OK(launch_ex(conv_gemm_kernel<MODE, EPI, NM, FOLD>, dim3(grid), dim3(kConvThreads), smem, st, prio_conv(), pdl_conv_on(st), p));
    LAUNCH_OK("conv_gemm_kernel");
    return 0;
}

template <int CO, int EPI>
static int launch_march(const MarchParams& p, unsigned smem, int grid, cudaStream_t st) {
    SET_MAX_SMEM_ONCE(conv_march_kernel<CO, EPI>);
    CUDA_OK(launch_ex(conv_march_kernel<CO, EPI>, dim3(grid), dim3(kMarchThreads), smem, st, prio_conv(), pdl_conv_on(st), p));
    LAUNCH_OK("conv_march_kernel");
    return 0;
}

static int run_march(const b200_conv_desc* d, MarchParams& p, const void* src_a, const void* packed, void* out,
                     const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                     float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef, const GnFin* fin,
                     cudaStream_t st) {
    if (!src_a || !packed) return fail("conv: null operand");
    if (d->epi == EPI_BF16 && !out) return fail("conv: null output");
    if (d->epi == EPI_SIGMOID && (!probs || !bias || n_out_real < 1 || n_out_real > 4))
        return fail("conv: sigmoid epilogue needs bias, probs and 1..4 real outputs");
    if (stats_partial && d->epi != EPI_BF16) return fail("conv: GroupNorm statistics only for the bf16 epilogue");
    if (check_ptr16(src_a, "src_a") || check_ptr16(out, "out") || check_ptr16(residual, "residual") ||
        check_ptr16(packed, "packed weights"))
        return 1;
    p.wpacked = (const __nv_bfloat16*)packed;
    p.lrelu_out = lrelu_out;
    p.stats_partial = stats_partial;
    p.bias = bias; p.probs = probs; p.logits = logits; p.n_out_real = n_out_real;
    {
        const char* dbg = getenv("B200_CONV_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }
    const int grid = march_ctas(p);
    Vol vol{d->N, d->D, d->H, d->W};
    p.src = make_act(src_a, vol);
    p.out = make_act(out, vol);
    p.residual = make_act(residual, vol);
    if (gnb_x && (!stats_partial || !gnb_coef || check_ptr16(gnb_x, "gnb_x") || check_ptr16(gnb_coef, "gnb_coef")))
        return fail("conv: GroupNorm-backward fold needs stats_partial, gnb_coef and 16-byte aligned pointers");
    p.gnb_x = make_act(gnb_x, vol);
    p.gnb_coef = gnb_coef;
    if (fin) {
        if (!stats_partial || gnb_x) return fail("conv: fused GroupNorm finalize needs forward statistics (stats_partial, no backward fold)");
        p.gn_fin = *fin;
    }
    const unsigned smem = p.smem_bar_off + kMarchTailBytes;
    if (d->epi == EPI_SIGMOID) return launch_march<16, EPI_SIGMOID>(p, smem, grid, st);
    if (d->Cout == 16) return launch_march<16, EPI_BF16>(p, smem, grid, st);
    return launch_march<32, EPI_BF16>(p, smem, grid, st);
}

template <int EPI>
static int launch_band(const BandParams& p, unsigned smem, int grid, cudaStream_t st) {
    SET_MAX_SMEM_ONCE(conv_band_kernel<EPI>);
    CUDA_OK(launch_ex(conv_band_kernel<EPI>, dim3(grid), dim3(kBandThreads), smem, st, prio_conv(), pdl_conv_on(st), p));
    LAUNCH_OK("conv_band_kernel");
    return 0;
}

static int run_band(const b200_conv_desc* d, BandParams& p, const void* src_a, const void* packed, void* out,
                    const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                    float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef, const GnFin* fin,
                    cudaStream_t st) {
    if (!src_a || !packed) return fail("conv: null operand");
    if (d->epi == EPI_BF16 && !out) return fail("conv: null output");
    if (d->epi == EPI_SIGMOID && (!probs || !bias || n_out_real < 1 || n_out_real > 4))
        return fail("conv: sigmoid epilogue needs bias, probs and 1..4 real outputs");
    if (stats_partial && d->epi != EPI_BF16) return fail("conv: GroupNorm statistics only for the bf16 epilogue");
    if (check_ptr16(src_a, "src_a") || check_ptr16(out, "out") || check_ptr16(residual, "residual") ||
        check_ptr16(packed, "packed weights"))
        return 1;
    p.wpacked = (const __nv_bfloat16*)packed;
    p.lrelu_out = lrelu_out;
    p.stats_partial = stats_partial;
    p.bias = bias; p.probs = probs; p.logits = logits; p.n_out_real = n_out_real;
    {
        const char* dbg = getenv("B200_CONV_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }
    const int grid = band_ctas(p);
    Vol vol{d->N, d->D, d->H, d->W};
    p.src = make_act(src_a, vol);
    p.out = make_act(out, vol);
    p.residual = make_act(residual, vol);
    if (gnb_x && (!stats_partial || !gnb_coef || check_ptr16(gnb_x, "gnb_x") || check_ptr16(gnb_coef, "gnb_coef")))
        return fail("conv: GroupNorm-backward fold needs stats_partial, gnb_coef and 16-byte aligned pointers");
    p.gnb_x = make_act(gnb_x, vol);
    p.gnb_coef = gnb_coef;
    if (fin) {
        if (!stats_partial || gnb_x) return fail("conv: fused GroupNorm finalize needs forward statistics (stats_partial, no backward fold)");
        p.gn_fin = *fin;
    }
    const unsigned smem = p.smem_bar_off + kBandTailBytes;
    if (d->epi == EPI_SIGMOID) return launch_band<EPI_SIGMOID>(p, smem, grid, st);
    return launch_band<EPI_BF16>(p, smem, grid, st);
}

extern "C" int b200_conv_supports_gnbwd(const b200_conv_desc* d) {
    if (check_conv_desc(d) || d->epi != EPI_BF16) return 0;
    BandParams bp;
    if (plan_band(d, bp) == 0) return 1;
    MarchParams mp;
    if (plan_march(d, mp) == 0) return 1;
    return 0;
}

static int conv_run_impl(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed, void* out,
                         const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                         float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef, const GnFin* fin,
                         void* stream);

extern "C" int b200_conv_run(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed,
                             void* out, const void* residual, int lrelu_out, float* stats_partial, const float* bias,
                             float* probs, float* logits, int n_out_real, void* stream) {
    return conv_run_impl(d, src_a, src_b, packed, out, residual, lrelu_out, stats_partial, bias, probs, logits, n_out_real,
                         nullptr, nullptr, nullptr, stream);
}

extern "C" int b200_conv_run_gnbwd(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed,
                                   void* out, const void* residual, int lrelu_out, float* stats_partial, const float* bias,
                                   float* probs, float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef,
                                   void* stream) {
    return conv_run_impl(d, src_a, src_b, packed, out, residual, lrelu_out, stats_partial, bias, probs, logits, n_out_real,
                         gnb_x, gnb_coef, nullptr, stream);
}

// 3x3x3 conv + GroupNorm statistics in ONE launch: the epilogue leaves per-CTA partial sums in stats_partial and the
// last CTA to finish reduces them to mean / rstd ([N][8] each) - what b200_conv_run(stats_partial) followed by
// b200_gn_finalize computes, bit for bit, without the second launch.  ticket: one zero-initialised 32-bit word of device
// memory, left zero again (reusable by the next launch on the same stream).
extern "C" int b200_conv_run_gn(const b200_conv_desc* d, const void* src_a, const void* packed, void* out,
                                float* stats_partial, float* mean, float* rstd, unsigned int* ticket, float eps,
                                void* stream) {
    if (check_conv_desc(d)) return 1;
    if (!stats_partial || !mean || !rstd || !ticket) return fail("conv_run_gn: null argument");
    if (d->Cout % 8) return fail("GroupNorm(8) needs C %% 8 == 0");
    GnFin fin;
    fin.mean = mean; fin.rstd = rstd; fin.ticket = ticket; fin.eps = eps;
    fin.count = (double)(d->Cout / 8) * d->D * d->H * d->W;
    return conv_run_impl(d, src_a, nullptr, packed, out, nullptr, 0, stats_partial, nullptr, nullptr, nullptr, 0, nullptr,
                         nullptr, &fin, stream);
}

static int conv_run_impl(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed, void* out,
                         const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                         float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef, const GnFin* fin,
                         void* stream) {
    if (check_conv_desc(d)) return 1;
    {
        BandParams bp;
        if (plan_band(d, bp) == 0)
            return run_band(d, bp, src_a, packed, out, residual, lrelu_out, stats_partial, bias, probs, logits,
                            n_out_real, gnb_x, gnb_coef, fin, (cudaStream_t)stream);
    }
    {
        MarchParams mp;
        if (plan_march(d, mp) == 0)
            return run_march(d, mp, src_a, packed, out, residual, lrelu_out, stats_partial, bias, probs, logits,
                             n_out_real, gnb_x, gnb_coef, fin, (cudaStream_t)stream);
    }
    if (gnb_x) return fail("conv: the GroupNorm-backward fold exists for the band / marching 3x3x3 kernels only (b200_conv_supports_gnbwd)");
    ConvKParams p;
    if (plan_conv(d, p)) return 1;
    if (!src_a || !packed) return fail("conv: null operand");
    if (d->Cin_b > 0 && !src_b) return fail("conv: second source missing");
    if (d->epi != EPI_SIGMOID && !out) return fail("conv: null output");
    if (d->epi == EPI_SIGMOID && (!probs || !bias || n_out_real < 1 || n_out_real > 4))
        return fail("conv: sigmoid epilogue needs bias, probs and 1..4 real outputs");
    if (stats_partial && (d->mode != MODE_K3 || d->epi != EPI_BF16 || (p.n_jobs != 1 && !p.pair)))
        return fail("conv: GroupNorm statistics only for the k3 bf16 path");
    cudaStream_t st = (cudaStream_t)stream;
    p.wpacked = (const __nv_bfloat16*)packed;
    p.lrelu_out = lrelu_out;
    p.stats_partial = stats_partial;
    if (fin) {
        if (!stats_partial) return fail("conv: fused GroupNorm finalize needs stats_partial");
        p.gn_fin = *fin;
    }
    p.bias = bias; p.probs = probs; p.logits = logits; p.n_out_real = n_out_real;
    {
        const char* dbg = getenv("B200_CONV_DEBUG");     // perf probes only (tools/perf_probe.py)
        p.debug = dbg ? atoi(dbg) : 0;
    }
    const int ctas = conv_grid_ctas(p);
    const int grid = p.pair ? ctas : ctas * p.n_jobs;
    if (check_ptr16(src_a, "src_a") || check_ptr16(src_b, "src_b") || check_ptr16(out, "out") ||
        check_ptr16(residual, "residual") || check_ptr16(packed, "packed weights"))
        return 1;
    Vol vol{d->N, d->D, d->H, d->W};
    p.src_a = make_act(src_a, vol);
    p.src_b = make_act(d->Cin_b > 0 ? src_b : src_a, vol);
    p.out = make_act(out, vol);
    p.residual = make_act(residual, vol);
    {
        const char* epf = getenv("B200_RES_PREFETCH");      // A/B: 0 = no L2 prefetch of the epilogue's residual operand
        p.prefetch_residual = (epf && !atoi(epf)) ? 0 : 1;
    }
    if (d->epi == EPI_D2S) {
        // the output and the residual are FINE tensors of Cout / 8 channels (conv_gemm.cuh ConvKParams::d2s)
        Vol fine{d->N, 2 * d->D, 2 * d->H, 2 * d->W};
        p.out = make_act(out, fine);
        p.residual = make_act(residual, fine);
        p.d2s = 1;
        {
            // 32-byte accesses (OPT-IN, B200_D2S_V8=1: bit-identical and measured SLOWER, 177 vs 101 us at 2 x 64^3 -> 128^3,
            // profiles/r02c_ab_d2s.txt): even rows of every chunk plane must be 32-byte aligned - guard rows and plane
            // rows are even for even W (common.cuh Vol), so it comes down to the base pointers
            const char* e8 = getenv("B200_D2S_V8");
            p.d2s_v8 = ((e8 && atoi(e8)) && d->W % 2 == 0 && ((uintptr_t)out & 31) == 0 && ((uintptr_t)residual & 31) == 0 &&
                        p.out.plane_rows % 2 == 0 && p.out.guard % 2 == 0) ? 1 : 0;
        }
        {
            const char* es = getenv("B200_D2S_SPREAD");
            p.d2s_spread = (es && !atoi(es)) ? 0 : 1;
        }
        p.d2s_sh = 0;
        while ((8 << p.d2s_sh) < d->Cout / 8) ++p.d2s_sh;
        if (lrelu_out) return fail("conv: no activation with the depth-to-space epilogue");
    }
    if (p.x_stage_bytes >= (1u << 20) || p.w_stage_bytes >= (1u << 20)) return fail("conv: stage exceeds the mbarrier tx-count range");
    const unsigned smem = p.smem_bar_off + kConvTailBytes;
    const int Nm = conv_nmma(d);
    if (p.pair) {
        SET_MAX_SMEM_ONCE(conv_pair_kernel<128>);
        CUDA_OK(launch_prio(conv_pair_kernel<128>, dim3(grid), dim3(kConvThreads), smem, st, prio_conv(), p));
        LAUNCH_OK("conv_pair_kernel");
        return 0;
    }
#define CONV_CASE(MODE, EPI, NM, FOLD) return launch_conv<MODE, EPI, NM, FOLD>(p, smem, grid, st)
    if (d->epi == EPI_SIGMOID) CONV_CASE(MODE_K3, EPI_SIGMOID, 16, 0);
    if (d->mode == MODE_K3) {
        if (conv_fold(d)) {
            if (Nm == 48) CONV_CASE(MODE_K3, EPI_BF16, 48, 1);
            if (Nm == 96) CONV_CASE(MODE_K3, EPI_BF16, 96, 1);
        }
        switch (Nm) {
            case 16: CONV_CASE(MODE_K3, EPI_BF16, 16, 0);
            case 32: CONV_CASE(MODE_K3, EPI_BF16, 32, 0);
            case 64: CONV_CASE(MODE_K3, EPI_BF16, 64, 0);
            case 128: CONV_CASE(MODE_K3, EPI_BF16, 128, 0);
        }
    } else {
        switch (Nm) {
            case 16: CONV_CASE(MODE_K1, EPI_BF16, 16, 0);
            case 32: CONV_CASE(MODE_K1, EPI_BF16, 32, 0);
            case 64: CONV_CASE(MODE_K1, EPI_BF16, 64, 0);
            case 128: CONV_CASE(MODE_K1, EPI_BF16, 128, 0);
            case 256: CONV_CASE(MODE_K1, EPI_BF16, 256, 0);
        }
    }
#undef CONV_CASE
    return fail("conv: no kernel for mode %d Cout %d", d->mode, d->Cout);
}

// ---------------------------------------------------------------------------------------
// wgrad planning
// ---------------------------------------------------------------------------------------
struct WgradPlan {
    WgradKParams k;
    int banded, folded, accs, dual;
    int kw_acc;     // 1: accumulators = kw taps (2-row halo), folds = kh taps;  0: accumulators = kh, folds = kw
    unsigned smem;
    int grid;
};

static int plan_wgrad(const b200_wgrad_desc* d, WgradPlan& P) {
    if (!d) return fail("null wgrad desc");
    if (d->mode != 0 && d->mode != 1) return fail("wgrad mode %d unsupported", d->mode);
    if (d->Cout < 16 || d->Cout % 16 || d->Cin < 16 || d->Cin % 16) return fail("wgrad: channels must be multiples of 16");
    memset(&P, 0, sizeof(P));
    WgradKParams& k = P.k;
    Vol v{d->N, d->D, d->H, d->W};
    k.total_rows = v.total_rows();
    k.Wp = v.Wp();
    k.SS = (int)v.slice_rows();
    if (d->mode == 0) {
        if (d->Cout > 128 || d->Cin > 128) return fail("wgrad k3: channels <= 128");
        if (d->Cout <= 32) { P.banded = 1; k.M = (d->Cout == 16) ? 64 : 128; }
        else { P.banded = 0; k.M = (d->Cout == 64) ? 64 : 128; }
        if (d->Cout == 48 || d->Cout == 80 || d->Cout == 96 || d->Cout == 112) return fail("wgrad k3: Cout %d unsupported", d->Cout);
        // kw in accumulators costs a 2-row halo, kh in N-folds costs 3 copies of X: take the accumulators first
        // and fold only while the three accumulators of the folded tile still fit TMEM (Cin <= 32)
        P.folded = (3 * d->Cin <= 256 && 9 * d->Cin <= 512) ? 1 : 0;
        k.Nmma = P.folded ? 3 * d->Cin : d->Cin;
        P.accs = (3 * k.Nmma <= 512) ? 1 : 0;
        // 16 x 16 channels: an MMA of M=64, N=48 still costs the 46-cycle operand-fetch floor; stacking a second
        // row range of the same stage in M and N (M=128, N=96, 56 cycles) does twice the work per MMA
        // (measured on B200: the kernel at 16 x 16 is then bound by the shared-memory ingest of the shifted operand
        // copies, ~7 TB/s aggregate, and the dual form is 10 % SLOWER than the plain one - it stays opt-in)
        static int want_dual = -1;
        if (want_dual < 0) { const char* e = getenv("B200_WGRAD_DUAL"); want_dual = (e && atoi(e)) ? 1 : 0; }
        P.dual = (want_dual && d->Cout == 16 && d->Cin == 16) ? 1 : 0;
        if (P.dual) { k.M = 128; k.Nmma = 96; }
        k.nfold = P.folded ? 3 : 1;
        k.nacc = P.accs ? 3 : 1;
        k.nband_loaded = P.banded ? 3 : 1;
        // tap roles: kd -> M bands (banded) or jobs; one of kh / kw -> N folds (shifted copies of X) and the
        // other -> TMEM accumulators (X start address + t * acc_shift rows: a halo of 2 * acc_shift rows)
        static int roles = -1;
        if (roles < 0) { const char* e = getenv("B200_WGRAD_KH_ACC"); roles = (e && atoi(e)) ? 0 : 1; }
        P.kw_acc = roles;
        k.acc_shift = P.kw_acc ? 1 : k.Wp;
        k.fold_shift = P.kw_acc ? k.Wp : 1;
        int j = 0;
        for (int kd = 0; kd < (P.banded ? 1 : 3); ++kd)
            for (int ta = 0; ta < (P.accs ? 1 : 3); ++ta)
                for (int tf = 0; tf < (P.folded ? 1 : 3); ++tf) {
                    const int kh = P.kw_acc ? tf : ta, kw = P.kw_acc ? ta : tf;
                    const bool kh_abs = P.kw_acc ? P.folded : P.accs, kw_abs = P.kw_acc ? P.accs : P.folded;
                    k.job_kd[j] = P.banded ? 0 : kd - 1;
                    k.job_kh[j] = kh_abs ? 0 : kh - 1;
                    k.job_kw[j] = kw_abs ? 0 : kw - 1;
                    k.job_xch[j] = 0;
                    ++j;
                }
        k.n_jobs = j;
    } else {
        if (d->Cout > 128) return fail("wgrad k1: Cout <= 128");
        P.banded = 0; P.folded = 0; P.accs = 0; P.dual = 0; P.kw_acc = 1;
        k.acc_shift = 1; k.fold_shift = k.Wp;
        k.M = d->Cout <= 64 ? 64 : 128;
        k.Nmma = d->Cin > 256 ? 256 : d->Cin;
        if (d->Cin % k.Nmma) return fail("wgrad k1: Cin %d unsupported", d->Cin);
        k.nfold = 1; k.nacc = 1; k.nband_loaded = 1;
        k.n_jobs = d->Cin / k.Nmma;
        for (int j = 0; j < k.n_jobs; ++j) { k.job_kd[j] = k.job_kh[j] = k.job_kw[j] = 0; k.job_xch[j] = j * k.Nmma; }
    }
    if (k.n_jobs > kMaxJobs) return fail("wgrad: too many jobs");
    k.nh = P.dual ? 2 : 1;
    k.CoC = d->Cout / 8;
    k.CiC = (d->mode == 0 ? d->Cin : k.Nmma) / 8;
    k.y_planes = k.M / 8;
    k.x_planes = k.Nmma / 8;
    if (k.nband_loaded * k.CoC * k.nh > k.y_planes) return fail("wgrad: internal plane count");
    // stage size: largest KT in {256,128,64} with >= 2 stages
    const unsigned bar_bytes = 1024;     // barriers + TMEM slot (first 256 B) + the producer's plane-offset tables
    k.stages = 0;
    for (int KT : {256, 128, 64}) {
        const int XR = (KT + (k.nacc > 1 ? 2 * k.acc_shift : 0) + 7) / 8 * 8;
        const unsigned ypl = (unsigned)KT * 16, xpl = (unsigned)XR * 16;
        const unsigned stage = k.y_planes * ypl + k.x_planes * xpl;
        int stages = (int)((kMaxSmem - bar_bytes) / stage);
        if (stages >= 2) {
            k.KT = KT; k.XR = XR;
            k.y_plane_bytes = ypl; k.x_plane_bytes = xpl; k.stage_bytes = stage;
            k.stages = std::min(stages, 4);
            break;
        }
    }
    if (k.stages == 0) return fail("wgrad: shared memory plan does not fit (Cout=%d Cin=%d W=%d)", d->Cout, d->Cin, d->W);
    k.stage_tx_bytes = (unsigned)(k.nband_loaded * k.CoC * k.nh) * k.y_plane_bytes +
                       (unsigned)(k.nfold * k.CiC * k.nh) * k.x_plane_bytes;
    k.smem_bar_off = k.stages * k.stage_bytes;
    P.smem = k.smem_bar_off + bar_bytes;
    k.tmem_cols = pow2_cols((unsigned)(k.nacc * k.Nmma));
    if (k.tmem_cols > 512) return fail("wgrad: TMEM plan does not fit");
    const int total_stages = ceil_div(k.total_rows, k.KT * k.nh);
    k.splits = std::max(1, std::min(num_sms() / k.n_jobs, total_stages));
    k.stages_per_split = ceil_div(total_stages, k.splits);
    k.splits = ceil_div(total_stages, k.stages_per_split);
    P.grid = k.splits * k.n_jobs;
    return 0;
}

// Marching form (wgrad_march.cuh) for the 3x3x3 16 <-> 16 convs: W a multiple of 16 and a band of lines that fits
// shared memory.  OPT-IN (B200_WGRAD_MARCH=1): measured on B200 (profiles/r01_ab_wgrad_march.txt) it moves 60 % fewer
// bytes from L2 but is bound by the same tcgen05.mma issue time as the linear-row kernel (3 MMAs of M=64/N=48 per 16
// voxels, ~35 cycles each) and pays ~20 us more prologue/epilogue: 137-145 us vs 118-124 us alone, 6.32 vs 6.05 ms per
// training step.  Kept, with its parity tests, as the starting point for a form that also cuts the MMA count.
struct WgradMarchPlan {
    WgradMarchParams k;
    unsigned smem;
    int grid;
};
static bool plan_wgrad_march(const b200_wgrad_desc* d, WgradMarchPlan& P, bool ignore_switch = false) {
    const char* e = getenv("B200_WGRAD_MARCH");         // read per call: tests switch forms in one process
    const int off = (ignore_switch || (e && atoi(e))) ? 0 : 1;
    if (off || d->mode != 0 || d->Cout != 16 || d->Cin != 16 || d->W % 16 || d->W < 16) return false;
    memset(&P, 0, sizeof(P));
    WgradMarchParams& k = P.k;
    k.N = d->N; k.D = d->D; k.H = d->H; k.W = d->W; k.Wp = d->W + 2;
    k.ksteps = d->W / 16;
    k.x_blk_bytes = (unsigned)k.Wp * 16u;
    const unsigned bar_bytes = 1024;
    // band height: the largest that fits shared memory (3 ring slots + 2 staging slots of dY, 2 stages of X),
    // preferring one that divides H (equal work per band)
    auto fits = [&](int BH) {
        if (BH > d->H || BH * k.Wp > 16383) return false;
        const unsigned ypl = (unsigned)BH * k.Wp * 16u;
        const unsigned need = 5u * 2u * ypl + (unsigned)kWgmXStages * (unsigned)(BH + 2) * 2u * k.x_blk_bytes + bar_bytes;
        return need <= kMaxSmem;
    };
    k.BH = 0;
    for (int BH : {8, 6, 4, 3, 2, 1}) if (fits(BH) && d->H % BH == 0) { k.BH = BH; break; }
    if (k.BH == 0) for (int BH : {8, 6, 4, 3, 2, 1}) if (fits(BH)) { k.BH = BH; break; }
    if (k.BH == 0) return false;
    k.n_bands = ceil_div(d->H, k.BH);
    k.units = (long long)d->N * k.n_bands * d->D;
    k.nsub = (k.BH % 2 == 0) ? 2 : 1;
    k.BHs = k.BH / k.nsub;
    k.y_plane_bytes = (unsigned)k.BH * k.Wp * 16u;
    k.y_sub_bytes = (unsigned)k.BHs * k.Wp * 16u;
    k.y_slot_bytes = 2u * k.y_plane_bytes;
    k.x_stage_bytes = (unsigned)(k.BH + 2) * 2u * k.x_blk_bytes;
    if (k.y_slot_bytes >= (1u << 20) || k.x_stage_bytes >= (1u << 20)) return false;
    k.smem_y_off = 0;
    k.smem_stg_off = 3u * k.y_slot_bytes;                 // directly after the ring (see the kernel)
    k.smem_x_off = k.smem_stg_off + 2u * k.y_slot_bytes;
    k.smem_bar_off = k.smem_x_off + (unsigned)kWgmXStages * k.x_stage_bytes;
    P.smem = k.smem_bar_off + bar_bytes;
    P.grid = (int)std::min<long long>(num_sms(), k.units);
    return true;
}

// Line-marching form (wgrad_line.cuh) for the 3x3x3 16 <-> 16 convs: W a multiple of 16 and a band of >= 2 lines whose
// ring fits shared memory.  Default; B200_NO_WGRAD_LINE=1 falls back to the linear-row kernel (A/B measurements).
struct WgradLinePlan {
    WgradLineParams k;
    unsigned smem;
    int grid;
    int nchy, nchx;     // chunk planes per line of dY / X: 2 (16 channels), 4 (32 channels), 1 (upper chunk known zero)
    int pair;           // two dY lines per MMA (wgrad_line.cuh, WglShape<.., PAIR = 1>)
};
// real_out / real_in: PyTorch channel counts of the gradient being computed (0 = unknown): with <= 8 of them the upper
// 8-channel chunk of the 16-channel tensor is zero by layout contract (b200_pack_input, b200_sigmoid_backward write zeros
// there) and is neither loaded nor multiplied.
static bool plan_wgrad_line(const b200_wgrad_desc* d, WgradLinePlan& P, bool ignore_switch = false, int real_out = 0,
                            int real_in = 0) {
    const char* e = getenv("B200_NO_WGRAD_LINE");        // read per call: tests switch forms in one process
    if (!ignore_switch && e && atoi(e)) return false;
    if (d->mode != 0 || d->Cout != d->Cin || (d->Cout != 16 && d->Cout != 32) || d->W % 16 || d->W < 16) return false;
    if (d->Cout == 32 && !ignore_switch) {
        // 32 <-> 32 channels: OPT-IN (B200_WGRAD_LINE32=1).  Correct (CPU replay + op checks) and slower than the linear-row
        // kernel (profiles/r02_ab_wgrad_line32.txt: 58.6 vs 48.3 us at 2 x 64^3): N = 288 needs three MMAs per K step, the
        // same tensor time as the linear form, and the shared->shared kw copies of dY plus twice as many (half as long)
        // bulk copies per line cost more shared-memory bandwidth and issue slots than the 3.4x L2 re-ingest they remove.
        const char* e32 = getenv("B200_WGRAD_LINE32");
        if (!e32 || !atoi(e32)) return false;
    }
    memset(&P, 0, sizeof(P));
    WgradLineParams& k = P.k;
    // chunk planes per line of dY / X (2: 16 channels, 4: 32 channels, 1: 16-channel tensor whose upper chunk is zero)
    unsigned nchy = (unsigned)d->Cout / 8u, nchx = (unsigned)d->Cin / 8u;
    {
        const char* eh = getenv("B200_WGL_NO_HALF");        // A/B: always load both chunks
        const bool half = !(eh && atoi(eh));
        if (half && d->Cout == 16 && real_out > 0 && real_out <= 8 && !(real_in > 0 && real_in <= 8)) nchy = 1;
        else if (half && d->Cin == 16 && real_in > 0 && real_in <= 8) nchx = 1;
    }
    // Pair mode (round 2c): two dY lines per MMA, window of 12 ring slots (wgrad_line.cuh).  Needs an even number of lines in
    // every band (H and LH even) and an even Ny >= 4.  OPT-IN: B200_WGL_PAIR is a mask over the channel variants (1: 16 x 16,
    // 2: conv_input, 4: conv_output).  Correct (CPU replay + op checks) and SLOWER at 2 x 128^3 (154 vs 108 us,
    // profiles/r02c_ab_wgrad_pair.txt): the 12-slot window + 11 mirrored slots leave room for bands of only 4 lines, and the
    // ring's prefetch lead is (band height - 4) lines - zero; DESIGN 4.2.
    bool pair = false;
    {
        const char* ep = getenv("B200_WGL_PAIR");
        const int mask = ep ? atoi(ep) : 0;
        const int bit = (nchy == 2 && nchx == 2) ? 1 : (nchy == 2 && nchx == 1) ? 2 : (nchy == 1 && nchx == 2) ? 4 : 0;
        pair = (mask & bit) != 0 && d->H % 2 == 0 && d->H >= 2;
    }
    const unsigned lines = pair ? 2u : 1u;
    const unsigned mirror = pair ? 11u : (unsigned)kWglMirror;
    const unsigned Mmma = lines * 3u * 8u * nchy <= 64u ? 64u : 128u;
    const unsigned slack = Mmma / 8u - lines * 3u * nchy;
    P.nchy = (int)nchy; P.nchx = (int)nchx; P.pair = pair ? 1 : 0;
    k.pair = P.pair; k.mirror = (int)mirror;
    k.N = d->N; k.D = d->D; k.H = d->H; k.W = d->W; k.Wp = d->W + 2;
    k.ksteps = d->W / 16;
    k.Lp = (unsigned)k.Wp * 16u;
    if (k.Lp >= (1u << 18)) return false;               // descriptor stride field
    const unsigned bar_bytes = 1024;
    // Shared memory: the X ring and Ny expanded dY lines (NR: the raw dY ring of the first version of the kernel, now 0).  ~12 KB per SM stay free on purpose: in the training
    // step this kernel runs beside the memory-bound GroupNorm-backward kernels (engine.py, side stream), whose CTAs need a
    // few KB each to be co-resident - with the whole 227 KB taken they queue behind this kernel and the overlap is lost
    // (measured: step 5.89 ms vs 5.81 ms).
    const unsigned budget = kMaxSmem - 12 * 1024;
    auto need = [&](int LH, int NR, int Ny) {
        const unsigned R = 3u * (LH + 2) + 1u;
        return (unsigned long long)(R + mirror) * nchx * k.Lp + 128u + (unsigned long long)NR * nchy * k.Lp +
               ((unsigned long long)Ny * 3u * nchy + slack) * k.Lp + bar_bytes;
    };
    k.LH = 0;
    {   // diagnostics: B200_WGL_LH / B200_WGL_NR / B200_WGL_NY force the plan (tools/wgrad_ab.py sweeps them)
        const char* el = getenv("B200_WGL_LH");
        const char* er = getenv("B200_WGL_NR");
        const char* en = getenv("B200_WGL_NY");
        if (el && er && en) {
            const int LH = atoi(el), NR = atoi(er), Ny = atoi(en);
            if (LH >= 1 && LH <= d->H && LH + 2 <= kWglND && NR >= 0 && NR <= kWglMaxNy && Ny >= 2 && Ny <= kWglMaxNy &&
                (!pair || (LH % 2 == 0 && Ny % 2 == 0 && Ny >= 4)) && need(LH, NR, Ny) <= kMaxSmem) { k.LH = LH; k.NR = NR; k.Ny = Ny; }
        }
    }
    // band height: prefer one that divides H (equal work per band), then the tallest; as many expanded dY lines in flight
    // as fit (>= 4: a slot's round trip is the MMAs of its line + ~2000 cycles of L2 latency, a line is ~600 cycles)
    const int lstep = pair ? 2 : 1;
    for (int pass = 0; pass < 2 && k.LH == 0; ++pass)
        for (int LH = 8; LH >= 2 && k.LH == 0; LH -= lstep) {
            if (LH > d->H || LH + 2 > kWglND) continue;     // a producer polls a step_done barrier at most LH steps late
            if (pass == 0 && d->H % LH) continue;
            for (int Ny = kWglMaxNy; Ny >= 4; Ny -= lstep)
                if (need(LH, 0, Ny) <= budget) { k.LH = LH; k.NR = 0; k.Ny = Ny; break; }
        }
    if (k.LH == 0) return false;
    k.R = 3 * (k.LH + 2) + 1;
    k.n_bands = ceil_div(d->H, k.LH);
    k.units = (long long)d->N * k.n_bands * d->D;
    k.smem_x_off = 0;
    k.smem_raw_off = align_up((unsigned)(k.R + mirror) * nchx * k.Lp, 128);
    k.smem_y_off = k.smem_raw_off + (unsigned)k.NR * nchy * k.Lp;
    k.smem_bar_off = align_up(k.smem_y_off + ((unsigned)k.Ny * 3u * nchy + slack) * k.Lp, 16);
    P.smem = k.smem_bar_off + bar_bytes;
    if (P.smem > kMaxSmem) return false;
    P.grid = (int)std::min<long long>(num_sms(), k.units);
    return true;
}

extern "C" size_t b200_wgrad_workspace_bytes(const b200_wgrad_desc* d) {
    WgradPlan P;
    if (plan_wgrad(d, P)) return 0;
    size_t bytes = (size_t)P.k.n_jobs * P.k.splits * P.k.nacc * P.k.M * P.k.Nmma * sizeof(float);
    WgradLinePlan LP;
    if (plan_wgrad_line(d, LP, true))
        bytes = std::max(bytes, (size_t)LP.grid * 2 * (3 * 8 * LP.nchy) * (9 * 8 * LP.nchx) * sizeof(float));   // (x 2: pair mode)
    WgradMarchPlan MP;
    if (plan_wgrad_march(d, MP))
        bytes = std::max(bytes, (size_t)MP.grid * kWgmAccs * kWgmM * kWgmN * sizeof(float));
    return bytes;
}

extern "C" int b200_wgrad_run(const b200_wgrad_desc* d, const void* dy, const void* x, void* workspace, float* grad,
                              int kind, int Cout_w, int Cin_w, int taps_w, int ci_off, int accumulate, void* stream) {
    WgradPlan P;
    if (plan_wgrad(d, P)) return 1;
    if (!dy || !x || !workspace || !grad) return fail("wgrad: null operand");
    if (kind == B200_G_K3 && (d->mode != 0 || taps_w != 27)) return fail("wgrad: kind/mode mismatch");
    if (kind == B200_G_K1 && (d->mode != 1 || taps_w != 1)) return fail("wgrad: kind/mode mismatch");
    if (kind == B200_G_S2D && (d->mode != 1 || taps_w != 8 || d->Cin != 8 * Cin_w)) return fail("wgrad: s2d kind mismatch");
    if (Cout_w > d->Cout) return fail("wgrad: Cout_w > Cout");
    cudaStream_t st = (cudaStream_t)stream;
    SET_MAX_SMEM_ONCE(wgrad_gemm_kernel);
    P.k.partial = (float*)workspace;
    if (check_ptr16(dy, "dy") || check_ptr16(x, "x")) return 1;
    Vol vol{d->N, d->D, d->H, d->W};
    P.k.dy = make_act(dy, vol);
    P.k.x = make_act(x, vol);
    static int wdbg = -1;
    if (wdbg < 0) { const char* e = getenv("B200_WGRAD_DEBUG"); wdbg = e ? atoi(e) : 0; }
    {
        WgradLinePlan LP;
        const char* em = getenv("B200_WGRAD_MARCH");
        if (kind == B200_G_K3 && !(em && atoi(em)) && plan_wgrad_line(d, LP, false, Cout_w, Cin_w + ci_off)) {
            LP.k.dy = P.k.dy; LP.k.x = P.k.x; LP.k.partial = (float*)workspace;
#ifdef B200_PROBES
            { const char* e2 = getenv("B200_WGL_DEBUG"); LP.k.debug = e2 ? atoi(e2) : 0; }
#endif
#define WGL_CASE(Y, X, PR)                                                                                                   \
    if (LP.nchy == Y && LP.nchx == X && LP.pair == PR) {                                                                     \
        SET_MAX_SMEM_ONCE(wgrad_line_kernel<Y, X, PR>);                                                                      \
        CUDA_OK(launch_prio(wgrad_line_kernel<Y, X, PR>, dim3(LP.grid), dim3(kWglThreads), LP.smem, st, prio_wgrad(), LP.k)); \
    }
            WGL_CASE(2, 2, 1) else WGL_CASE(2, 1, 1) else WGL_CASE(1, 2, 1) else
            WGL_CASE(2, 2, 0) else WGL_CASE(2, 1, 0) else WGL_CASE(1, 2, 0) else WGL_CASE(4, 4, 0) else return fail("wgrad_line: no kernel");
#undef WGL_CASE
            LAUNCH_OK("wgrad_line_kernel");
            WglReduceParams rq;
            rq.ctas = LP.grid * (LP.pair ? 2 : 1);      // pair mode: two partial blocks per CTA
            rq.Cout_w = Cout_w; rq.Cin_w = Cin_w; rq.accumulate = accumulate;
            rq.CY = 8 * LP.nchy; rq.CX = 8 * LP.nchx;
            constexpr int qpb = 256 / kWglReduceGroups;
            const int quads = 27 * rq.CY * rq.CX / 4;
            CUDA_OK(launch_ex(wgrad_line_reduce_kernel, dim3((quads + qpb - 1) / qpb), dim3(256), 0, st, prio_wgrad(), pdl_wgrad(),
                                (const float*)workspace, grad, rq));
            LAUNCH_OK("wgrad_line_reduce_kernel");
            return 0;
        }
    }
    {
        WgradMarchPlan MP;
        if (kind == B200_G_K3 && plan_wgrad_march(d, MP)) {
            SET_MAX_SMEM_ONCE(wgrad_march_kernel);
            MP.k.dy = P.k.dy; MP.k.x = P.k.x; MP.k.partial = (float*)workspace; MP.k.debug = wdbg;
            wgrad_march_kernel<<<MP.grid, kWgmThreads, MP.smem, st>>>(MP.k);
            LAUNCH_OK("wgrad_march_kernel");
            WgmReduceParams rq;
            rq.ctas = MP.grid; rq.Cout_w = Cout_w; rq.Cin_w = Cin_w; rq.accumulate = accumulate;
            constexpr int qpb = 256 / kWgmReduceGroups;
            wgrad_march_reduce_kernel<<<(27 * 16 * 4 + qpb - 1) / qpb, 256, 0, st>>>((const float*)workspace, grad, rq);
            LAUNCH_OK("wgrad_march_reduce_kernel");
            return 0;
        }
    }
    if (P.k.stage_tx_bytes >= (1u << 20)) return fail("wgrad: stage exceeds the mbarrier tx-count range");
    P.k.debug = wdbg;
    CUDA_OK(launch_prio(wgrad_gemm_kernel, dim3(P.grid), dim3(kWgradThreads), P.smem, st, prio_wgrad(), P.k));
    LAUNCH_OK("wgrad_gemm_kernel");
    WgradReduceParams q;
    q.kind = kind; q.Cout_w = Cout_w; q.Cin_w = Cin_w; q.taps_w = taps_w; q.ci_off = ci_off;
    q.Cout_g = d->Cout; q.Cin_g = d->Cin;
    q.banded = P.banded; q.folded = P.folded; q.accs = P.accs; q.dual = P.dual; q.kw_acc = P.kw_acc;
    q.n_jobs = P.k.n_jobs; q.splits = P.k.splits; q.nacc = P.k.nacc; q.M = P.k.M; q.Nmma = P.k.Nmma;
    q.accumulate = accumulate;
    const int quads = q.n_jobs * q.nacc * q.M * q.Nmma / 4;
    // split-groups per CTA: enough threads to fill the machine (~2 CTAs per SM), at most 32 and <= splits
    int SG = 1;
    while (SG < 32 && SG * 2 <= q.splits && (long long)quads * SG < 2LL * 256 * num_sms()) SG *= 2;
    const int qpc = 256 / SG;
    CUDA_OK(launch_ex(wgrad_reduce_kernel, dim3((quads + qpc - 1) / qpc), dim3(256), 0, st, prio_wgrad(), pdl_wgrad(),
                        (const float*)P.k.partial, grad, q, SG));
    LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------
// memory-bound ops
// ---------------------------------------------------------------------------------------
// lines per CTA for the line-walking kernels: a power of two <= 16 that divides D*H (so a CTA
// never straddles two samples) and still leaves >= ~8 CTAs per SM
static int lines_per_block(int N, int D, int H, int W = 0, int C = 0) {
    const long long lines = (long long)N * D * H;
    int lpb = 1;
    // CTAs per SM that must remain: 4 (measured on the training step: 8 -> 5.947 ms, 4 -> 5.904 ms, 2 -> 5.895 ms with a
    // slower forward; B200_EW_CTAS_PER_SM overrides)
    static int min_ctas = -1;
    if (min_ctas < 0) { const char* e = getenv("B200_EW_CTAS_PER_SM"); min_ctas = (e && atoi(e) > 0) ? atoi(e) : 4; }
    while (lpb < 16 && (D * H) % (lpb * 2) == 0 && lines / (lpb * 2) >= (long long)min_ctas * num_sms()) lpb *= 2;
    // small tensors (W, C given): at least ~16 KB of each operand per CTA, even if that leaves SMs idle
    if (W > 0 && C > 0)
        while (lpb < 16 && (D * H) % (lpb * 2) == 0 && (long long)lpb * W * C * 2 < 16384) lpb *= 2;
    return lpb;
}

static LineGeom make_line_geom(int W, int C, int lpb) {
    LineGeom g;
    g.nvec = W * (C / 8);
    g.W = W;
    g.lpb = lpb;
    g.by_nvec = make_fastdiv((unsigned)g.nvec);
    g.by_W = make_fastdiv((unsigned)W);
    return g;
}

static int check_act(int N, int D, int H, int W, int C) {
    if (N < 1 || D < 1 || H < 1 || W < 1) return fail("bad volume %dx%dx%dx%d", N, D, H, W);
    if (C < 8 || C % 8 || C > 1024) return fail("channel count %d unsupported", C);
    return 0;
}

// lines per CTA of the four-voxels-per-thread layout kernels: up to 16, while ~6 CTAs per SM remain
static int quad_lpb(int N, int D, int H) {
    const long long lines = (long long)N * D * H;
    int lpb = 1;
    while (lpb < 16 && lines / (lpb * 2) >= 6LL * num_sms()) lpb *= 2;
    return lpb;
}

template <typename T>
static int pack_input_t(const T* x, void* act_out, int N, int D, int H, int W, int Creal, cudaStream_t st) {
    Vol v{N, D, H, W};
    if (W % 4 == 0 && ((uintptr_t)x & (4 * sizeof(T) - 1)) == 0) {
        const int lpb = quad_lpb(N, D, H);
        CUDA_OK(launch_ex(pack_input4_kernel<T>, dim3((N * D * H + lpb - 1) / lpb), dim3(256), 0, st, 0, pdl_ew((long long)((N * D * H + lpb - 1) / lpb)), x, make_act(act_out, v), v, Creal, lpb,
                                                                          make_fastdiv((unsigned)(W / 4))));
        LAUNCH_OK("pack_input4_kernel");
        return 0;
    }
    CUDA_OK(launch_ex(pack_input_kernel<T>, dim3(N * D * H), dim3(128), 0, st, 0, pdl_ew((long long)(N * D * H)), x, make_act(act_out, v), v, Creal));
    LAUNCH_OK("pack_input_kernel");
    return 0;
}
extern "C" int b200_pack_input_t(const void* x, int x_dtype, void* act_out, int N, int D, int H, int W, int Creal, int Cpad,
                                 void* stream) {
    if (check_act(N, D, H, W, Cpad) || Creal > Cpad) return fail("pack_input: bad channels %d -> %d", Creal, Cpad);
    if (Creal > 8) return fail("pack_input: at most 8 real channels");
    if (x_dtype == B200_F32) return pack_input_t((const float*)x, act_out, N, D, H, W, Creal, (cudaStream_t)stream);
    if (x_dtype == B200_BF16) return pack_input_t((const __nv_bfloat16*)x, act_out, N, D, H, W, Creal, (cudaStream_t)stream);
    return fail("pack_input: input dtype %d unsupported (fp32 or bf16)", x_dtype);
}
extern "C" int b200_pack_input(const float* x, void* act_out, int N, int D, int H, int W, int Creal, int Cpad,
                               void* stream) {
    return b200_pack_input_t(x, B200_F32, act_out, N, D, H, W, Creal, Cpad, stream);
}

extern "C" int b200_gn_finalize(const float* stats_partial, int ctas, int N, int C, int D, int H, int W, float eps,
                                float* mean, float* rstd, void* stream) {
    if (C % 8) return fail("GroupNorm(8) needs C %% 8 == 0");
    const double count = (double)(C / 8) * D * H * W;
    CUDA_OK(launch_ex(gn_finalize_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, 0, pdl_small(), stats_partial, ctas, N, count, eps, mean, rstd, nullptr, nullptr,
                                                            nullptr, C, 1));
    LAUNCH_OK("gn_finalize_kernel");
    return 0;
}
extern "C" int b200_gn_finalize_coef(const float* stats_partial, int ctas, int N, int C, int D, int H, int W, float eps,
                                     const float* gamma, const float* beta, int do_lrelu, float* mean, float* rstd,
                                     float* coef, void* stream) {
    if (C % 8) return fail("GroupNorm(8) needs C %% 8 == 0");
    if (!gamma || !beta || !coef) return fail("gn_finalize_coef: null argument");
    if (check_ptr16(coef, "coef")) return 1;
    const double count = (double)(C / 8) * D * H * W;
    CUDA_OK(launch_ex(gn_finalize_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, 0, pdl_small(), stats_partial, ctas, N, count, eps, mean, rstd, gamma, beta, coef,
                                                            C, do_lrelu));
    LAUNCH_OK("gn_finalize_kernel");
    return 0;
}

extern "C" int b200_gn_apply(const void* x, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, const void* residual, void* out, int N, int D, int H, int W, int C,
                             int do_lrelu, void* stream) {
    if (check_act(N, D, H, W, C) || C > 256 || 256 % (C / 8)) return fail("gn_apply: C=%d unsupported", C);
    Vol v{N, D, H, W};
    const int lpb = lines_per_block(N, D, H);
    CUDA_OK(launch_ex(gn_apply_kernel, dim3(N * D * H / lpb), dim3(kEwThreads), 0, (cudaStream_t)stream, 0, pdl_ew((long long)(N * D * H / lpb)), make_act(x, v), mean, rstd, gamma, beta, make_act(residual, v), make_act(out, v), v, C, do_lrelu,
        make_line_geom(W, C, lpb)));
    LAUNCH_OK("gn_apply_kernel");
    return 0;
}

// reduction grid: ~4 CTAs per SM over the batch; each CTA owns `lpb` consecutive lines (<= kRedLines)
static void gn_bwd_grid(int N, int D, int H, int W, int C, int& blocks, int& lpb) {
    const int lines = D * H;
    int want = std::max(1, 4 * num_sms() / std::max(1, N));
    want = std::min(want, lines);
    lpb = (lines + want - 1) / want;
    // small tensors: at least ~64 KB of x+dy per CTA (a CTA costs a few microseconds of fixed latency, and
    // every extra CTA is another partial for the finalize kernel to add)
    const int line_bytes = 2 * W * C * 2;
    lpb = std::max(lpb, (65536 + line_bytes - 1) / line_bytes);
    lpb = std::min(std::min(lpb, kRedLines), lines);
    blocks = (lines + lpb - 1) / lpb;
}
static int gn_bwd_max_blocks() { return 4 * num_sms() + 8192; }

// Cluster form of GroupNorm backward (elementwise3.cuh): usable when one sample's slice of a unit, split over the
// 8 CTAs of a cluster, fits their shared memory.
static bool plan_gn_cluster(const Vol& v, int C, GnClusterParams& q, unsigned& smem) {
    // Opt-in (B200_GN_BWD_CLUSTER=1).  Measured on B200 (profiles/r01_ab_gn_cluster.txt): alone the kernel is faster
    // than reduce -> finalize -> apply at the 16^3 level (26 -> 18 us), but inside the training step it is SLOWER
    // (6.129 -> 6.157 ms with tiles <= 40 KB, 6.223 ms with tiles <= 160 KB): backward overlaps the data-gradient
    // chain with weight-gradient GEMMs on a side stream, those CTAs hold ~200 KB of shared memory on every SM, and a
    // cluster needs 8 SMs of one GPC with room for its tiles at the same time, so it waits where the small
    // register-only kernels of the three-kernel path slip in beside the GEMM.
    const char* e = getenv("B200_GN_BWD_CLUSTER");
    if (!e || atoi(e) == 0) return false;
    const int gs = C / 8;
    if (gs < 2 || (gs < 8 && 8 % gs) || (gs >= 8 && gs % 8)) return false;
    const int ucs = gs >= 8 ? gs / 8 : 1;
    if (ucs > kGnClusterMaxUcs) return false;
    const int lines = v.D * v.H;
    const int lpc = (lines + kGnClusterSize - 1) / kGnClusterSize;
    if (lpc > kGnClusterMaxLines) return false;
    const long long nv = (long long)lpc * v.W;
    const long long bytes = 2LL * ucs * nv * 16;
    // Tiles larger than ~24 KB cannot share an SM with a weight-gradient CTA of the side stream (200 KB of shared
    // memory), so the cluster would wait for that kernel to drain; B200_GN_BWD_CLUSTER_MAXKB moves the limit.
    const char* mk = getenv("B200_GN_BWD_CLUSTER_MAXKB");
    const long long max_kb = mk ? std::min(160, std::max(0, atoi(mk))) : 160;
    if (bytes > max_kb * 1024 || nv >= (1 << 24)) return false;
    static int attr_set = 0, clusters_ok = -1;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gn_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) {
            cudaGetLastError();
            clusters_ok = 0;
        }
        attr_set = 1;
    }
    if (clusters_ok < 0) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(kGnClusterSize); cfg.blockDim = dim3(kGnClusterThreads); cfg.dynamicSmemBytes = 160 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kGnClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, gn_bwd_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
        clusters_ok = nc > 0 ? 1 : 0;
    }
    if (!clusters_ok) return false;
    memset(&q, 0, sizeof(q));
    q.v = v; q.C = C;
    q.by_W = make_fastdiv((unsigned)v.W);
    q.lpc = lpc; q.nv = (int)nv; q.ucs = ucs;
    q.m = (double)gs * v.D * v.H * v.W;
    smem = (unsigned)bytes;
    return true;
}
// Register-resident one-launch form (elementwise3.cuh gn_bwd_creg_kernel): usable when 8 or 16 CTAs per unit can hold
// one sample's slice of the unit in registers (UCS chunks x VPT vectors per thread, UCS * VPT <= 8): the 16^3 and 32^3
// levels of the U-Net at any batch size.  B200_GN_BWD_CREG=0 turns it off.
typedef void (*GnCregKernel)(GnClusterParams);
static const int kGnCregMaxUnits = 32;
static GnCregKernel gn_creg_kernel(int ucs, int vpt) {
    switch (ucs * 16 + vpt) {
        case 1 * 16 + 1: return gn_bwd_creg_kernel<1, 1>;
        case 1 * 16 + 2: return gn_bwd_creg_kernel<1, 2>;
        case 1 * 16 + 4: return gn_bwd_creg_kernel<1, 4>;
        case 1 * 16 + 8: return gn_bwd_creg_kernel<1, 8>;
        case 2 * 16 + 1: return gn_bwd_creg_kernel<2, 1>;
        case 2 * 16 + 2: return gn_bwd_creg_kernel<2, 2>;
        case 2 * 16 + 4: return gn_bwd_creg_kernel<2, 4>;
        case 4 * 16 + 1: return gn_bwd_creg_kernel<4, 1>;
        case 4 * 16 + 2: return gn_bwd_creg_kernel<4, 2>;
    }
    return nullptr;
}
static bool plan_gn_creg(const Vol& v, int C, GnClusterParams& q, int& csize, int& vpt) {
    // OPT-IN (B200_GN_BWD_CREG=1): parity-green, 28 launches fewer, and no faster inside the training step
    // (profiles/r02_ab_gn_creg.txt: 5.575 vs 5.570 ms): beside a weight-gradient GEMM the one-launch kernel - a chain of
    // load -> shuffle -> barrier -> exchange -> apply latencies, twice (two samples) - stretches to the length of the three
    // small kernels it replaces.
    {
        const char* e = getenv("B200_GN_BWD_CREG");
        if (!e || atoi(e) == 0) return false;
    }
    const int gs = C / 8;
    if (gs < 2 || (gs < 8 && 8 % gs) || (gs >= 8 && gs % 8)) return false;
    const int ucs = gs >= 8 ? gs / 8 : 1;
    if (ucs > kGnClusterMaxUcs) return false;
    const int lines = v.D * v.H;
    for (int cs : {8, 16}) {
        const int lpc = (lines + cs - 1) / cs;
        if (lpc > kGnClusterMaxLines) continue;
        const long long nv = (long long)lpc * v.W;
        int vp = 1;
        while (vp < 8 && (long long)vp * kGnClusterThreads < nv) vp *= 2;
        if ((long long)vp * kGnClusterThreads < nv || ucs * vp > 8) continue;
        GnCregKernel k = gn_creg_kernel(ucs, vp);
        if (!k) continue;
        // every CTA of a unit spins on the others: all of them must be resident at once
        const int units = (C / 8) / ucs;
        if (units * cs > num_sms() || units > kGnCregMaxUnits) continue;
        memset(&q, 0, sizeof(q));
        q.v = v; q.C = C;
        q.by_W = make_fastdiv((unsigned)v.W);
        q.lpc = lpc; q.nv = (int)nv; q.ucs = ucs;
        q.m = (double)gs * v.D * v.H * v.W;
        csize = cs; vpt = vp;
        return nv < (1 << 24);
    }
    return false;
}
// workspace: tickets[N + 1] (32-bit words, padded to 64) | partial[N][blocks][C][2] | coef[N][C][2] | tot[N][C][2] (double)
// The ticket words must be ZERO when the workspace is first used (allocate it zero-filled); every launch leaves them zero.
// (words [0, N]: last-CTA tickets of the fused finalize; the last 64 words: arrive / exit counters of the
// register-resident form, two per unit)
static size_t gn_bwd_ticket_floats(int N) { return std::max<size_t>(256, ((size_t)N + 1 + 63) / 64 * 64 + 128); }
static size_t gn_bwd_xchg_floats() { return (size_t)kGnCregMaxUnits * 16 * 128 * 2; }     // doubles: [units][csize <= 16][2][64]
extern "C" size_t b200_gn_backward_workspace_floats(int N, int C) {
    return gn_bwd_ticket_floats(N) + std::max((size_t)N * gn_bwd_max_blocks() * C * 2 + (size_t)N * C * 2 + (size_t)N * C * 4 + 2,
                                              gn_bwd_xchg_floats());
}
// which form b200_gn_backward takes for this tensor: 0 = reduce -> finalize -> apply, 1 = shared-memory cluster kernel
// (opt-in), 2 = register-resident cluster kernel (one launch)
extern "C" int b200_gn_backward_form(int N, int D, int H, int W, int C) {
    if (check_act(N, D, H, W, C) || C > 256 || 256 % (C / 8) || C < 16) return -1;
    Vol v{N, D, H, W};
    GnClusterParams q;
    int csize = 0, vpt = 0;
    unsigned smem = 0;
    if (plan_gn_creg(v, C, q, csize, vpt)) return 2;
    if (plan_gn_cluster(v, C, q, smem)) return 1;
    return 0;
}
extern "C" int b200_gn_backward(const void* x, const void* dy, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta,
                                float* workspace, int N, int D, int H, int W, int C, int do_lrelu, void* stream) {
    if (check_act(N, D, H, W, C) || C > 256 || 256 % (C / 8) || C < 16) return fail("gn_backward: C=%d unsupported", C);
    Vol v{N, D, H, W};
    cudaStream_t st = (cudaStream_t)stream;
    {
        // small tensors: one cluster launch that keeps x and dy in registers between the two passes (elementwise3.cuh)
        GnClusterParams q;
        int csize = 0, vpt = 0;
        if (plan_gn_creg(v, C, q, csize, vpt)) {
            q.x = make_act(x, v); q.dy = make_act(dy, v); q.dx = make_act(dx, v);
            q.mean = mean; q.rstd = rstd; q.gamma = gamma; q.beta = beta; q.dgamma = dgamma; q.dbeta = dbeta;
            q.do_lrelu = do_lrelu;
            if (!workspace || ((uintptr_t)workspace & 7)) return fail("gn_backward: workspace must be 8-byte aligned");
            q.csize = csize;
            q.counters = reinterpret_cast<unsigned int*>(workspace) + gn_bwd_ticket_floats(N) - 64;
            q.xchg = reinterpret_cast<double*>(workspace + gn_bwd_ticket_floats(N));
            gn_creg_kernel(q.ucs, vpt)<<<(unsigned)((C / 8) / q.ucs * csize), kGnClusterThreads, 0, st>>>(q);
            LAUNCH_OK("gn_bwd_creg_kernel");
            return 0;
        }
    }
    {
        // (opt-in, B200_GN_BWD_CLUSTER=1) the shared-memory form of the same kernel
        GnClusterParams q;
        unsigned smem = 0;
        if (plan_gn_cluster(v, C, q, smem)) {
            q.x = make_act(x, v); q.dy = make_act(dy, v); q.dx = make_act(dx, v);
            q.mean = mean; q.rstd = rstd; q.gamma = gamma; q.beta = beta; q.dgamma = dgamma; q.dbeta = dbeta;
            q.do_lrelu = do_lrelu;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)((C / 8) / q.ucs * kGnClusterSize));
            cfg.blockDim = dim3(kGnClusterThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = kGnClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            CUDA_OK(cudaLaunchKernelEx(&cfg, gn_bwd_cluster_kernel, q));
            LAUNCH_OK("gn_bwd_cluster_kernel");
            return 0;
        }
    }
    int blocks, rlpb;
    gn_bwd_grid(N, D, H, W, C, blocks, rlpb);
    if (!workspace || ((uintptr_t)workspace & 7)) return fail("gn_backward: workspace must be 8-byte aligned");
    float* partial = workspace + gn_bwd_ticket_floats(N);
    if (blocks > gn_bwd_max_blocks()) return fail("gn_backward: volume too large");
    float* coef = partial + (size_t)N * gn_bwd_max_blocks() * C * 2;
    const FastDiv by_W = make_fastdiv((unsigned)W);
    const double m = (double)(C / 8) * D * H * W;
    // B200_GN_BWD_FUSED_FIN=1: pass 2 (partial -> coef, dgamma, dbeta) runs inside the reduction kernel, in its last CTAs
    // (C <= 128: one thread per (channel, S1|S2) column).  Opt-in: parity-green, one launch fewer, +0.06 ms per step
    // (profiles/r02_ab_last_cta_finalize.txt) - one CTA reading 75-150 KB of partials is slower than the 8-CTA kernel.
    int fused = 0;
    if (const char* e = getenv("B200_GN_BWD_FUSED_FIN")) fused = atoi(e) != 0;
    GnBwdFin fin;
    memset(&fin, 0, sizeof(fin));
    if (fused && C <= 128) {
        fin.coef = coef;
        fin.tot = reinterpret_cast<double*>(coef + (size_t)N * C * 2 + (((size_t)N * C * 2) & 1));
        fin.dgamma = dgamma; fin.dbeta = dbeta;
        fin.tickets = reinterpret_cast<unsigned int*>(workspace);
        fin.m = m;
    }
    CUDA_OK(launch_ex(gn_bwd_reduce2_kernel, dim3(blocks, N), dim3(256), 0, st, 0, pdl_ew((long long)(blocks) * (long long)(N)), make_act(x, v), make_act(dy, v), mean, rstd, gamma, beta,
                                                          partial, v, C, do_lrelu, by_W, rlpb, fin));
    LAUNCH_OK("gn_bwd_reduce2_kernel");
    if (fin.coef == nullptr) {
        // one warp per (sample, channel, S1|S2) sum, up to 32 warps: min(8, N) * (C/8) * 2 sums per CTA
        const int fin_threads = std::min(1024, std::max(256, std::min(8, N) * (C / 8) * 2 * 32));
        CUDA_OK(launch_ex(gn_bwd_finalize2_kernel, dim3(8), dim3(fin_threads), 0, st, 0, pdl_small(), partial, blocks, N, C, m, gamma, coef, dgamma, dbeta));
        LAUNCH_OK("gn_bwd_finalize2_kernel");
    }
    const int lpb = lines_per_block(N, D, H, W, C);
    CUDA_OK(launch_ex(gn_bwd_apply2_kernel<false>, dim3(N * D * H / lpb), dim3(256), 0, st, 0, pdl_ew((long long)(N * D * H / lpb)), make_act(x, v), make_act(dy, v), mean, rstd, gamma, beta,
                                                                coef, make_act(dx, v), v, C, do_lrelu, by_W, lpb, nullptr));
    LAUNCH_OK("gn_bwd_apply2_kernel");
    return 0;
}

// GroupNorm backward when the group sums came from the producing conv's epilogue (b200_conv_run_gnbwd): no reduce pass.
//   fold-finalize (gpart -> coef)  ->  apply (dx, + per-CTA per-channel sums)  ->  finalize2 (dgamma, dbeta)
// workspace: coef[N][C][2] | scratch coef[N][C][2] | aff_partial[N][blocks][C][2], blocks = D*H / lines-per-CTA of the apply
extern "C" size_t b200_gn_backward_folded_workspace_floats(int N, int D, int H, int W, int C) {
    const int lpb = lines_per_block(N, D, H, W, C);
    return gn_bwd_ticket_floats(N) + (size_t)N * C * 4 + (size_t)N * (D * H / lpb) * C * 2;
}
extern "C" int b200_gn_backward_folded(const void* x, const void* dy, const float* mean, const float* rstd,
                                       const float* gamma, const float* beta, const float* gpart, int ctas, void* dx,
                                       float* dgamma, float* dbeta, float* workspace, int N, int D, int H, int W, int C,
                                       int do_lrelu, void* stream) {
    if (check_act(N, D, H, W, C) || C > 256 || 256 % (C / 8) || C < 16) return fail("gn_backward_folded: C=%d unsupported", C);
    if (!gpart || ctas < 1 || !workspace) return fail("gn_backward_folded: null argument");
    Vol v{N, D, H, W};
    cudaStream_t st = (cudaStream_t)stream;
    const double m = (double)(C / 8) * D * H * W;
    workspace += gn_bwd_ticket_floats(N);      // the ticket words of b200_gn_backward (the two share one workspace)
    float* coef = workspace;
    float* coef_scratch = workspace + (size_t)N * C * 2;
    float* aff = workspace + (size_t)N * C * 4;
    gn_bwd_fold_finalize_kernel<<<N, 256, 0, st>>>(gpart, ctas, N, C, m, mean, rstd, coef);
    LAUNCH_OK("gn_bwd_fold_finalize_kernel");
    const FastDiv by_W = make_fastdiv((unsigned)W);
    const int lpb = lines_per_block(N, D, H, W, C);
    const int bps = D * H / lpb;
    CUDA_OK(launch_ex(gn_bwd_apply2_kernel<true>, dim3(N * bps), dim3(256), 0, st, 0, pdl_ew((long long)(N * bps)), make_act(x, v), make_act(dy, v), mean, rstd, gamma, beta, coef,
                                                       make_act(dx, v), v, C, do_lrelu, by_W, lpb, aff));
    LAUNCH_OK("gn_bwd_apply2_kernel");
    const int fin_threads = std::min(1024, std::max(256, std::min(8, N) * (C / 8) * 2 * 32));
    CUDA_OK(launch_ex(gn_bwd_finalize2_kernel, dim3(8), dim3(fin_threads), 0, st, 0, pdl_small(), aff, bps, N, C, m, gamma, coef_scratch, dgamma, dbeta));
    LAUNCH_OK("gn_bwd_finalize2_kernel");
    return 0;
}

extern "C" int b200_upsample2x(const void* coarse, void* fine, int N, int D, int H, int W, int C, int do_lrelu,
                               void* stream) {
    if (check_act(N, D, H, W, C)) return 1;
    Vol vc{N, D, H, W};
    Vol vf{N, 2 * D, 2 * H, 2 * W};
    CUDA_OK(launch_ex(upsample2x_fwd3_kernel, dim3(N * D * H), dim3(128), 0, (cudaStream_t)stream, 0, pdl_ew((long long)(N * D * H)), make_act(coarse, vc), make_act(fine, vf), vc, C,
                                                                       do_lrelu, make_fastdiv((unsigned)W)));
    LAUNCH_OK("upsample2x_fwd3_kernel");
    return 0;
}
// workspace: the w-reduced intermediate, act layout of volume (N, 2D, 2H, W) with C channels
extern "C" size_t b200_upsample2x_backward_workspace_bytes(int N, int D, int H, int W, int C) {
    Vol vt{N, 2 * D, 2 * H, W};
    return (size_t)(C / 8) * vt.plane_rows() * 16;
}
extern "C" int b200_upsample2x_backward(const void* dfine, const void* fine_out, void* dcoarse, void* workspace, int N,
                                        int D, int H, int W, int C, int do_lrelu, void* stream) {
    if (check_act(N, D, H, W, C)) return 1;
    if (!workspace || check_ptr16(workspace, "workspace")) return fail("upsample2x_backward: workspace missing or misaligned");
    Vol vc{N, D, H, W};
    Vol vf{N, 2 * D, 2 * H, 2 * W};
    Vol vt{N, 2 * D, 2 * H, W};
    cudaStream_t st = (cudaStream_t)stream;
    const FastDiv by_W = make_fastdiv((unsigned)W);
    CUDA_OK(launch_ex(upsample2x_bwd_w3_kernel, dim3(N * 2 * D * 2 * H), dim3(128), 0, st, 0, pdl_ew((long long)(N * 2 * D * 2 * H)), make_act(dfine, vf), make_act(fine_out, vf),
                                                               make_act(workspace, vt), vc, C, do_lrelu, by_W));
    LAUNCH_OK("upsample2x_bwd_w3_kernel");
    CUDA_OK(launch_ex(upsample2x_bwd_dh3_kernel, dim3(N * D * H), dim3(128), 0, st, 0, pdl_ew((long long)(N * D * H)), make_act(workspace, vt), make_act(dcoarse, vc), vc, C, by_W));
    LAUNCH_OK("upsample2x_bwd_dh3_kernel");
    return 0;
}

extern "C" int b200_space_to_depth(const void* fine, void* coarse, int N, int D, int H, int W, int C, void* stream) {
    if (check_act(N, D, H, W, C)) return 1;
    Vol vc{N, D, H, W};
    Vol vf{N, 2 * D, 2 * H, 2 * W};
    CUDA_OK(launch_ex(s2d_kernel, dim3(N * D * H), dim3(kEwThreads), 0, (cudaStream_t)stream, 0, pdl_ew((long long)(N * D * H)), make_act(fine, vf), make_act(coarse, vc), vc, C));
    LAUNCH_OK("s2d_kernel");
    return 0;
}
extern "C" int b200_depth_to_space(const void* coarse, const void* residual, void* fine, int N, int D, int H, int W,
                                   int C, void* stream) {
    if (check_act(N, D, H, W, C)) return 1;
    Vol vc{N, D, H, W};
    Vol vf{N, 2 * D, 2 * H, 2 * W};
    CUDA_OK(launch_ex(d2s_kernel, dim3(N * D * H), dim3(kEwThreads), 0, (cudaStream_t)stream, 0, pdl_ew((long long)(N * D * H)), make_act(coarse, vc), make_act(residual, vf),
                                                                  make_act(fine, vf), vc, C));
    LAUNCH_OK("d2s_kernel");
    return 0;
}
extern "C" int b200_add(const void* a, const void* b, void* out, int N, int D, int H, int W, int C, void* stream) {
    if (check_act(N, D, H, W, C)) return 1;
    Vol v{N, D, H, W};
    const int lpb = lines_per_block(N, D, H);
    CUDA_OK(launch_ex(add_kernel, dim3(N * D * H / lpb), dim3(kEwThreads), 0, (cudaStream_t)stream, 0, pdl_ew((long long)(N * D * H / lpb)), make_act(a, v), make_act(b, v), make_act(out, v),
                                                                        v, C, make_line_geom(W, C, lpb)));
    LAUNCH_OK("add_kernel");
    return 0;
}

static int sigmoid_lpb(int N, int D, int H) {
    const long long lines = (long long)N * D * H;
    int lpb = 1;
    while (lpb < 16 && lines / (lpb * 2) >= 6LL * num_sms()) lpb *= 2;
    return lpb;
}
extern "C" size_t b200_sigmoid_backward_workspace_floats(int N, int D, int H) { return (size_t)N * D * H * 4; }
extern "C" int b200_sigmoid_backward(const float* grad_probs, const float* probs, void* dlogit_act, float* dbias,
                                     float* workspace, int N, int D, int H, int W, int Creal, int Cpad, void* stream) {
    if (check_act(N, D, H, W, Cpad) || Creal > 4) return fail("sigmoid_backward: Creal=%d Cpad=%d unsupported", Creal, Cpad);
    (void)Cpad;     // only chunk 0 is written; the other chunks of the destination stay zero
    Vol v{N, D, H, W};
    cudaStream_t st = (cudaStream_t)stream;
    const int lpb = sigmoid_lpb(N, D, H);
    const int blocks = (N * D * H + lpb - 1) / lpb;
    if (W % 4 == 0 && (((uintptr_t)grad_probs | (uintptr_t)probs) & 15) == 0) {
        CUDA_OK(launch_ex(sigmoid_bwd_pack4_kernel, dim3(blocks), dim3(256), 0, st, 0, pdl_ew((long long)(blocks)), grad_probs, probs, make_act(dlogit_act, v), workspace, v, Creal, lpb,
                                                        make_fastdiv((unsigned)(W / 4))));
        LAUNCH_OK("sigmoid_bwd_pack4_kernel");
    } else {
        CUDA_OK(launch_ex(sigmoid_bwd_pack2_kernel, dim3(blocks), dim3(256), 0, st, 0, pdl_ew((long long)(blocks)), grad_probs, probs, make_act(dlogit_act, v), workspace, v, Creal, lpb,
                                                        make_fastdiv((unsigned)W)));
        LAUNCH_OK("sigmoid_bwd_pack2_kernel");
    }
    if (dbias) {
        CUDA_OK(launch_ex(reduce_partials_kernel, dim3(Creal), dim3(256), 0, st, 0, pdl_small(), workspace, blocks, 4, Creal, dbias));
        LAUNCH_OK("reduce_partials_kernel");
    }
    return 0;
}

static int dice_blocks() { return 2 * num_sms(); }
extern "C" size_t b200_dice_workspace_floats(int B, int C) { return (size_t)B * C * dice_blocks() * 2; }
// `target_dtype`: B200_F32 (the reference's float targets, loss.py:105-111) or B200_U8 (binary masks staged as bytes)
extern "C" int b200_dice_sums_t(const float* probs, const void* target, int target_dtype, float* sums, float* workspace,
                                int B, int C, long long S, void* stream) {
    if (C < 1 || C > 4) return fail("dice: C=%d unsupported (1..4)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int bx = dice_blocks();
    if (target_dtype == B200_F32)
        dice_partial_kernel<float><<<dim3(bx, B * C), kEwThreads, 0, st>>>(probs, (const float*)target, workspace, B, C, S);
    else if (target_dtype == B200_U8)
        dice_partial_kernel<uint8_t><<<dim3(bx, B * C), kEwThreads, 0, st>>>(probs, (const uint8_t*)target, workspace, B, C, S);
    else
        return fail("dice: target dtype %d unsupported (fp32 or uint8)", target_dtype);
    LAUNCH_OK("dice_partial_kernel");
    dice_sums2_kernel<<<1, 256, 0, st>>>(workspace, B, C, bx, sums);
    LAUNCH_OK("dice_sums2_kernel");
    return 0;
}
extern "C" int b200_dice_sums(const float* probs, const float* target, float* sums, float* workspace, int B, int C,
                              long long S, void* stream) {
    return b200_dice_sums_t(probs, target, B200_F32, sums, workspace, B, C, S, stream);
}
extern "C" int b200_dice_loss(const float* sums, int C, float priority, float* loss, void* stream) {
    dice_loss_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, C, priority, loss);
    LAUNCH_OK("dice_loss_kernel");
    return 0;
}
extern "C" int b200_dice_backward_t(const float* probs, const void* target, int target_dtype, const float* sums,
                                    const float* grad_out, float priority, float* grad_probs, int B, int C, long long S,
                                    void* stream) {
    if (C < 1 || C > 4) return fail("dice: C=%d unsupported (1..4)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(dice_blocks(), B * C);
    if (target_dtype == B200_F32)
        dice_bwd_kernel<float><<<grid, kEwThreads, 0, st>>>(probs, (const float*)target, sums, grad_out, priority, grad_probs, B, C, S);
    else if (target_dtype == B200_U8)
        dice_bwd_kernel<uint8_t><<<grid, kEwThreads, 0, st>>>(probs, (const uint8_t*)target, sums, grad_out, priority, grad_probs, B, C, S);
    else
        return fail("dice: target dtype %d unsupported (fp32 or uint8)", target_dtype);
    LAUNCH_OK("dice_bwd_kernel");
    return 0;
}
extern "C" int b200_dice_backward(const float* probs, const float* target, const float* sums, const float* grad_out,
                                  float priority, float* grad_probs, int B, int C, long long S, void* stream) {
    return b200_dice_backward_t(probs, target, B200_F32, sums, grad_out, priority, grad_probs, B, C, S, stream);
}

// BCE_Loss (loss.py:64-79) -------------------------------------------------------------------
extern "C" size_t b200_bce_workspace_floats(void) { return (size_t)dice_blocks(); }
extern "C" int b200_bce_sum_t(const float* probs, const void* target, int target_dtype, float bg_weight, float* sum,
                              float* workspace, long long numel, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int bx = dice_blocks();
    if (target_dtype == B200_F32)
        bce_partial_kernel<float><<<bx, kEwThreads, 0, st>>>(probs, (const float*)target, bg_weight, workspace, numel);
    else if (target_dtype == B200_U8)
        bce_partial_kernel<uint8_t><<<bx, kEwThreads, 0, st>>>(probs, (const uint8_t*)target, bg_weight, workspace, numel);
    else
        return fail("bce: target dtype %d unsupported (fp32 or uint8)", target_dtype);
    LAUNCH_OK("bce_partial_kernel");
    CUDA_OK(launch_ex(reduce_partials_kernel, dim3(1), dim3(256), 0, st, 0, pdl_small(), workspace, bx, 1, 1, sum));
    LAUNCH_OK("reduce_partials_kernel");
    return 0;
}
extern "C" int b200_bce_sum(const float* probs, const float* target, float bg_weight, float* sum, float* workspace,
                            long long numel, void* stream) {
    return b200_bce_sum_t(probs, target, B200_F32, bg_weight, sum, workspace, numel, stream);
}
extern "C" int b200_bce_loss(const float* sum, double global_numel, float* loss, void* stream) {
    bce_loss_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sum, global_numel, loss);
    LAUNCH_OK("bce_loss_kernel");
    return 0;
}
extern "C" int b200_bce_backward_t(const float* probs, const void* target, int target_dtype, const float* grad_out,
                                   float bg_weight, double global_numel, float* grad_probs, long long numel, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int bx = dice_blocks() * 2;
    const float inv = (float)(1.0 / global_numel);
    if (target_dtype == B200_F32)
        bce_bwd_kernel<float><<<bx, kEwThreads, 0, st>>>(probs, (const float*)target, grad_out, bg_weight, inv, grad_probs, numel);
    else if (target_dtype == B200_U8)
        bce_bwd_kernel<uint8_t><<<bx, kEwThreads, 0, st>>>(probs, (const uint8_t*)target, grad_out, bg_weight, inv, grad_probs, numel);
    else
        return fail("bce: target dtype %d unsupported (fp32 or uint8)", target_dtype);
    LAUNCH_OK("bce_bwd_kernel");
    return 0;
}
extern "C" int b200_bce_backward(const float* probs, const float* target, const float* grad_out, float bg_weight,
                                 double global_numel, float* grad_probs, long long numel, void* stream) {
    return b200_bce_backward_t(probs, target, B200_F32, grad_out, bg_weight, global_numel, grad_probs, numel, stream);
}

// ---------------------------------------------------------------------------------------
// fused Adam (optimizer.cuh): main.py:133-138, train.py:220
// ---------------------------------------------------------------------------------------
extern "C" int b200_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                              const long long* seg_begin, const long long* seg_len, int n_segs, const float* lr,
                              float* step, unsigned* ticket, double beta1, double beta2, float eps, float weight_decay,
                              int lr_step_size, float lr_gamma, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !lr || !step || !ticket) return fail("adam: null argument");
    if (n_segs < 1 || n_segs > kAdamMaxSegs) return fail("adam: 1..%d segments supported, got %d", kAdamMaxSegs, n_segs);
    if (check_ptr16(params, "adam params") || check_ptr16(grads, "adam grads") || check_ptr16(exp_avg, "adam exp_avg") ||
        check_ptr16(exp_avg_sq, "adam exp_avg_sq") || check_ptr16(max_exp_avg_sq, "adam max_exp_avg_sq"))
        return 1;
    AdamParams q;
    memset(&q, 0, sizeof(q));
    q.p = params; q.g = grads; q.m = exp_avg; q.v = exp_avg_sq; q.vmax = max_exp_avg_sq;
    q.lr = lr; q.step = step; q.ticket = ticket;
    q.beta1 = (float)beta1; q.beta2 = (float)beta2; q.eps = eps; q.weight_decay = weight_decay;
    q.beta1d = beta1; q.beta2d = beta2;
    q.one_minus_beta1 = (float)(1.0 - beta1); q.one_minus_beta2 = (float)(1.0 - beta2);
    q.lr_step_size = lr_step_size; q.lr_gamma = lr_gamma;
    q.segs.n = n_segs;
    long long cum = 0;
    for (int i = 0; i < n_segs; ++i) {
        if (seg_begin[i] % 4 || seg_len[i] % 4 || seg_len[i] <= 0) return fail("adam: segments must be non-empty multiples of 4 floats");
        q.segs.begin[i] = seg_begin[i] / 4;
        q.segs.cum[i] = cum;
        cum += seg_len[i] / 4;
    }
    q.segs.cum[n_segs] = cum;
    // ~2 float4 per thread per pass, at most 4 CTAs per SM: the kernel is a pure stream over 36 bytes per element
    const long long want = (cum + 2LL * kAdamThreads - 1) / (2LL * kAdamThreads);
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, 4LL * num_sms()));
    adam_step_kernel<<<grid, kAdamThreads, 0, (cudaStream_t)stream>>>(q);
    LAUNCH_OK("adam_step_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------
// plan introspection (tests / DESIGN.md): lets a CPU test replay the exact addressing the
// kernels use.  out[] layout is documented in tests/test_plan_cpu.py.
// ---------------------------------------------------------------------------------------
extern "C" int b200_conv_plan_debug(const b200_conv_desc* d, int* out, int n_out) {
    ConvKParams p;
    if (plan_conv(d, p)) return 1;
    const int fold_flag = conv_fold(d) ? 1 : 0, RBv = fold_flag ? 126 : 128;
    const int vals[] = {p.BD, p.MB, p.TR, p.Q0, p.QN, p.tiles_q, p.tiles_d, p.num_tiles, p.whole, p.n_jobs,
                        p.KG, p.KGa, p.KC, p.NTG, p.TG, p.x_stages, p.w_stages, (int)p.x_stage_bytes,
                        (int)p.w_stage_bytes, (int)p.x_plane_bytes, p.SRp, p.nslices, p.halo_rows,
                        (int)p.tmem_cols, (int)(p.smem_bar_off + kConvTailBytes), conv_grid_ctas(p), p.Wp, p.SS, fold_flag, RBv};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv + kMaxTaps) return fail("plan_debug: need %d ints", nv + kMaxTaps);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    for (int i = 0; i < kMaxTaps; ++i) out[nv + i] = p.tap_off[i];
    return 0;
}

extern "C" int b200_band_plan_debug(const b200_conv_desc* d, int* out, int n_out) {
    if (check_conv_desc(d)) return 1;
    BandParams p;
    if (plan_band(d, p)) return fail("band-marching conv does not apply to this descriptor");
    const int vals[] = {kBandBH, p.n_bands, p.n_cols, (int)p.units, p.nslots, (int)p.plane_bytes, (int)p.slot_bytes,
                        (int)p.w_bytes, (int)p.wimg_bytes, (int)(p.smem_bar_off + kBandTailBytes), band_ctas(p), p.Wp, p.SS};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv) return fail("band_plan_debug: need %d ints", nv);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    return 0;
}

#ifdef B200_PROBES
extern "C" int b200_wgl_prof_read(unsigned long long* host_out, int n) {
    if (n > 160 * 8) n = 160 * 8;
    CUDA_OK(cudaMemcpyFromSymbol(host_out, g_wgl_prof, (size_t)n * sizeof(unsigned long long)));
    return 0;
}
#endif
extern "C" int b200_march_prof_read(unsigned long long* host_out, int n) {
    if (n > 160 * 16) n = 160 * 16;
    CUDA_OK(cudaMemcpyFromSymbol(host_out, g_march_prof, (size_t)n * sizeof(unsigned long long)));
    return 0;
}

extern "C" int b200_march_plan_debug(const b200_conv_desc* d, int* out, int n_out) {
    if (check_conv_desc(d)) return 1;
    MarchParams p;
    if (plan_march(d, p)) return fail("marching conv does not apply to this descriptor");
    const int vals[] = {p.MB, p.TR, p.Q0, p.QN, p.n_strips, (int)p.units, p.KS, p.SRp, p.nslots, (int)p.plane_bytes,
                        (int)p.slot_bytes, (int)p.w_bytes, (int)p.wtile_bytes, (int)p.tmem_cols,
                        (int)(p.smem_bar_off + kMarchTailBytes), march_ctas(p), p.Wp, p.SS};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv) return fail("march_plan_debug: need %d ints", nv);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    return 0;
}

extern "C" int b200_wgrad_march_plan_debug(const b200_wgrad_desc* d, int* out, int n_out) {
    WgradMarchPlan P;
    const bool ok = plan_wgrad_march(d, P, true);       // the plan itself does not depend on the opt-in switch
    if (!ok) return fail("wgrad_march: does not apply (needs mode 0, 16 x 16 channels, W %% 16 == 0)");
    const WgradMarchParams& k = P.k;
    const int vals[] = {k.BH, k.n_bands, (int)k.units, k.ksteps, k.BHs, k.nsub, k.Wp, (int)k.y_plane_bytes,
                        (int)k.y_sub_bytes, (int)k.y_slot_bytes, (int)k.x_blk_bytes, (int)k.x_stage_bytes,
                        (int)k.smem_y_off, (int)k.smem_stg_off, (int)k.smem_x_off, (int)P.smem, P.grid};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv) return fail("wgrad_march_plan_debug: need %d ints", nv);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    return 0;
}

static int wgrad_line_plan_debug(const b200_wgrad_desc* d, int real_out, int real_in, int* out, int n_out);
extern "C" int b200_wgrad_line_plan_debug(const b200_wgrad_desc* d, int* out, int n_out) {
    return wgrad_line_plan_debug(d, 0, 0, out, n_out);
}
// the plan b200_wgrad_run takes for a gradient with `real_out` x `real_in` PyTorch channels (<= 8 on one side of a
// 16-channel GEMM: that side's upper chunk is skipped)
extern "C" int b200_wgrad_line_plan_debug2(const b200_wgrad_desc* d, int real_out, int real_in, int* out, int n_out) {
    return wgrad_line_plan_debug(d, real_out, real_in, out, n_out);
}
static int wgrad_line_plan_debug(const b200_wgrad_desc* d, int real_out, int real_in, int* out, int n_out) {
    WgradLinePlan P;
    if (!plan_wgrad_line(d, P, true, real_out, real_in)) return fail("wgrad_line: does not apply (needs mode 0, 16 x 16 or 32 x 32 channels, W %% 16 == 0)");
    const WgradLineParams& k = P.k;
    const int vals[] = {k.LH, k.n_bands, (int)k.units, k.ksteps, k.R, k.Ny, k.Wp, (int)k.Lp, (int)k.smem_x_off,
                        (int)k.smem_y_off, (int)k.smem_bar_off, (int)P.smem, P.grid, k.mirror, kWglNB, kWglND, k.NR,
                        (int)k.smem_raw_off, P.nchy, P.nchx, P.pair};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv) return fail("wgrad_line_plan_debug: need %d ints", nv);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    return 0;
}

extern "C" int b200_wgrad_plan_debug(const b200_wgrad_desc* d, int* out, int n_out) {
    WgradPlan P;
    if (plan_wgrad(d, P)) return 1;
    const WgradKParams& k = P.k;
    const int vals[] = {k.KT, k.XR, k.nband_loaded, k.CoC, k.CiC, k.nfold, k.nacc, k.M, k.Nmma,
                        k.n_jobs, k.splits, k.stages_per_split, k.y_planes, k.x_planes, (int)k.y_plane_bytes,
                        (int)k.x_plane_bytes, (int)k.stage_bytes, (int)k.stage_tx_bytes, k.stages, (int)k.tmem_cols,
                        (int)P.smem, P.grid, P.banded, P.folded, P.accs, k.Wp, k.SS, k.nh, P.kw_acc};
    const int nv = (int)(sizeof(vals) / sizeof(int));
    if (n_out < nv + 4 * kMaxJobs) return fail("wgrad_plan_debug: need %d ints", nv + 4 * kMaxJobs);
    for (int i = 0; i < nv; ++i) out[i] = vals[i];
    for (int j = 0; j < kMaxJobs; ++j) {
        out[nv + 4 * j + 0] = k.job_kd[j];
        out[nv + 4 * j + 1] = k.job_kh[j];
        out[nv + 4 * j + 2] = k.job_kw[j];
        out[nv + 4 * j + 3] = k.job_xch[j];
    }
    return 0;
}

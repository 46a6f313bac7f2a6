/* C ABI of libbrats_b200.so — the B200-native (sm_100a) kernels behind the reference's
 * `model.UNet` / `loss.Dice_loss_joint` hot path (lachinov/brats2019).
 *
 * The reference has no FFI of its own: its hot path is `torch.nn` calls made from Python
 * (model.py, loss.py).  Each entry point below therefore replaces one ATen operator family the
 * reference reaches, cited as reference file:line.  The Python mirror of the reference's
 * `nn.Module` interface (brats2019_b200/model.py, loss.py) binds these with ctypes; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller; the library owns nothing persistent.
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous and never synchronise.
 *  - Return value 0 = ok; non-zero = error, text available from b200_last_error().
 *  - No CPU fallback, no cuDNN fallback: unsupported shapes or a non-sm_100 device are errors.
 *  - Activation tensors ("act") are bf16, chunk-planar with zero halos and zero guard rows:
 *        act[C/8][guard + rows + guard][8],   rows = N*(D+2)*(H+2)*(W+2),
 *        row(n,dp,hp,wp) = ((n*(D+2)+dp)*(H+2)+hp)*(W+2)+wp, interior at dp,hp,wp >= 1,
 *        guard = b200_act_guard_rows(D,H,W); the pointer passed is the start of the allocation.
 *    Every halo and guard element must be 0 when a buffer is first handed to the library
 *    (allocate zero-filled); kernels never write them.  Channel counts are multiples of 16.
 */
#ifndef BRATS_B200_H
#define BRATS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of the model's EXTERNAL tensors (everything inside the library is bf16 activations / fp32 sums):
 * fp32 is what the reference's loaders and trainer pass (train.py:194-199); bf16 images and uint8 binary targets are
 * compact staging formats a host pipeline may use instead - converted in registers by the first kernel that reads them */
enum { B200_F32 = 0, B200_BF16 = 1, B200_U8 = 2 };

const char* b200_last_error(void);
/* 0 if device `dev` is compute capability 10.x with tcgen05/TMA available. */
int b200_device_check(int dev);
int b200_num_sms(void);

/* ---- activation geometry ----------------------------------------------------------------- */
long long b200_act_guard_rows(int D, int H, int W);
long long b200_act_plane_rows(int N, int D, int H, int W);   /* rows per chunk plane incl. both guards */

/* ---- layout (model.py:410-412 `input = x[0]`) -------------------------------------------- */
/* x: fp32 (N,Creal,D,H,W) contiguous  ->  act with Cpad channels (channels >= Creal are 0). */
int b200_pack_input(const float* x, void* act_out, int N, int D, int H, int W, int Creal, int Cpad, void* stream);
/* same with x_dtype = B200_F32 or B200_BF16 */
int b200_pack_input_t(const void* x, int x_dtype, void* act_out, int N, int D, int H, int W, int Creal, int Cpad,
                      void* stream);

/* ---- convolution, forward and data-gradient (aten::convolution, model.py:72-73, 336, 348,
 *      362, 393, 401; data half of convolution_backward, train.py:210) ------------------- */
typedef struct b200_conv_desc {
    int mode;          /* 0: 3x3x3 stride 1 pad 1;  1: 1x1x1 */
    int epi;           /* 0: store bf16 act;  1: + bias, sigmoid, store fp32 NCDHW (conv_output);
                        * 2 (mode 1, Cout = 8 * Cf): depth-to-space store - the data gradient of the k2 s2 conv
                        *   (model.py:360-363 backward) written straight to the FINE tensor: `out` and `residual` of
                        *   b200_conv_run are (N, 2D, 2H, 2W) activations of Cf channels, column block
                        *   ((kd*2+kh)*2+kw) * Cf of coarse voxel (d,h,w) lands on fine voxel (2d+kd, 2h+kh, 2w+kw), the
                        *   residual (the skip-connection gradient) is added in fp32 before the one bf16 rounding;
                        *   replaces b200_conv_run(epi 0) + b200_depth_to_space and their coarse intermediate */
    int N, D, H, W;    /* output volume (== input volume) */
    int Cin_a, Cin_b;  /* channels of the two K sources (Cin_b = 0 if unused): cat([a,b]) fused */
    int Cout;          /* GEMM N: stored channels per output row (multiple of 16) */
} b200_conv_desc;

/* weight "kind" for b200_conv_pack_weight (how the GEMM reads the PyTorch tensor) */
enum { B200_W_FWD = 0, B200_W_DGRAD = 1, B200_W_FWD_S2D = 2, B200_W_DGRAD_S2D = 3 };

size_t b200_conv_packed_weight_bytes(const b200_conv_desc* d);
/* CTAs the conv will launch per column job == leading extent of stats_partial. */
int b200_conv_ctas(const b200_conv_desc* d);
/* w: fp32 PyTorch layout (Cout_w, Cin_w, taps_w) -> bf16 UMMA operand images in `packed`.
 * K_real/N_real: logical GEMM extents; rows/cols beyond them are zero padding.
 * ci_off: first PyTorch input channel addressed (halves of the cat conv). */
int b200_conv_pack_weight(const b200_conv_desc* d, int kind, const float* w, int Cout_w, int Cin_w, int taps_w,
                          int ci_off, int K_real, int N_real, void* packed, void* stream);
/* All packed weights of a model in ONE launch.  The host builds a table once (pointers and layouts are
 * stable across steps), copies it to the device, and re-runs it whenever the weights changed:
 *   b200_pack_table_build: fills `table_host` (n_jobs * b200_pack_table_entry_bytes()) and *total_blocks;
 *   b200_pack_table_run:   launches the packing kernel over a DEVICE copy of that table. */
typedef struct b200_pack_job {
    b200_conv_desc desc;
    int kind, Cout_w, Cin_w, taps_w, ci_off, K_real, N_real;
    const float* w;      /* device, fp32 PyTorch layout */
    void* packed;        /* device, b200_conv_packed_weight_bytes(&desc) bytes, 16-byte aligned */
} b200_pack_job;
size_t b200_pack_table_entry_bytes(void);
int b200_pack_table_build(const b200_pack_job* jobs, int n_jobs, void* table_host, size_t table_bytes, int* total_blocks);
int b200_pack_table_run(const void* table_device, int n_jobs, int total_blocks, void* stream);

/* out = conv(cat[src_a, src_b]) (+ residual) (lrelu optional).
 * stats_partial (optional, epi 0, mode 0): fp32 [b200_conv_ctas][N][16] per-CTA GroupNorm
 *   partial sums (8 group sums, 8 group sums of squares) of the fp32 accumulators.
 * epi 1: bias[n_out_real], probs/logits fp32 (N,n_out_real,D,H,W); logits may be NULL. */
int b200_conv_run(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed, void* out,
                  const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                  float* logits, int n_out_real, void* stream);

/* The same convolution with the NEXT GroupNorm's backward sums folded into the epilogue (replaces the reduction pass of
 * native_group_norm_backward, model.py:105-112 backward): when this conv PRODUCES the gradient dy of y = lrelu(GN(c)),
 * pass gnb_x = the saved conv output c (act, same volume and channels as `out`) and gnb_coef = the [N][3][C] table
 * b200_gn_finalize_coef wrote in the forward pass; stats_partial ([b200_conv_ctas][N][16]) then receives, per group,
 * sum gamma*dz and sum gamma*dz*c (dz = dy * lrelu'), the input of b200_gn_backward_folded.  Only the band / marching
 * 3x3x3 kernels have this epilogue: b200_conv_supports_gnbwd(desc) != 0. */
int b200_conv_supports_gnbwd(const b200_conv_desc* d);
int b200_conv_run_gnbwd(const b200_conv_desc* d, const void* src_a, const void* src_b, const void* packed, void* out,
                        const void* residual, int lrelu_out, float* stats_partial, const float* bias, float* probs,
                        float* logits, int n_out_real, const void* gnb_x, const float* gnb_coef, void* stream);

/* 3x3x3 conv (epi 0, one source) + the statistics of the GroupNorm that follows it (aten::native_group_norm's mean /
 * rstd, model.py:95-96, 105-106, 338) in ONE launch: the epilogue leaves per-CTA partial sums in stats_partial
 * ([b200_conv_ctas][N][16]) and the last CTA to finish reduces them to mean / rstd ([N][8] each) - bit for bit what
 * b200_conv_run(stats_partial) followed by b200_gn_finalize writes, without the second launch.
 * ticket: one 32-bit word of device memory that is ZERO before the first launch; every launch leaves it zero again
 * (one word serves all launches on a stream and all replays of a CUDA graph). */
int b200_conv_run_gn(const b200_conv_desc* d, const void* src_a, const void* packed, void* out, float* stats_partial,
                     float* mean, float* rstd, unsigned int* ticket, float eps, void* stream);

/* ---- convolution weight gradient (weight half of convolution_backward, train.py:210) ---- */
typedef struct b200_wgrad_desc {
    int mode;          /* 0: 3x3x3;  1: 1x1x1 */
    int N, D, H, W;
    int Cout;          /* channels of dy (padded, multiple of 16) */
    int Cin;           /* channels of x  (padded, multiple of 16) */
} b200_wgrad_desc;
enum { B200_G_K3 = 0, B200_G_K1 = 1, B200_G_S2D = 2 };
size_t b200_wgrad_workspace_bytes(const b200_wgrad_desc* d);
/* grad: fp32 PyTorch layout (Cout_w, Cin_w, taps_w).  kind: B200_G_*.  accumulate != 0 adds. */
int b200_wgrad_run(const b200_wgrad_desc* d, const void* dy, const void* x, void* workspace, float* grad, int kind,
                   int Cout_w, int Cin_w, int taps_w, int ci_off, int accumulate, void* stream);

/* ---- GroupNorm(8) + LeakyReLU(0.01) + residual add (aten::native_group_norm, leaky_relu_,
 *      add; model.py:95-96, 105-115, 338, 413) ------------------------------------------- */
int b200_gn_finalize(const float* stats_partial, int ctas, int N, int C, int D, int H, int W, float eps, float* mean,
                     float* rstd, void* stream);
/* b200_gn_finalize + the coefficient table [N][3][C] = (rstd*gamma | beta - mean*rstd*gamma | gamma) that
 * b200_conv_run_gnbwd reads in the backward pass (do_lrelu == 0: the table encodes "no activation"). */
int b200_gn_finalize_coef(const float* stats_partial, int ctas, int N, int C, int D, int H, int W, float eps,
                          const float* gamma, const float* beta, int do_lrelu, float* mean, float* rstd, float* coef,
                          void* stream);
int b200_gn_apply(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  const void* residual, void* out, int N, int D, int H, int W, int C, int do_lrelu, void* stream);
size_t b200_gn_backward_workspace_floats(int N, int C);
/* which kernels b200_gn_backward launches for this tensor: 0 = reduction, finalize, apply (three launches, x and dy read
 * twice); 1 / 2 = ONE thread-block-cluster launch that keeps the tensors on chip between the two passes (1: in shared
 * memory, opt-in; 2: in registers - the 16^3 and 32^3 levels of the U-Net) */
int b200_gn_backward_form(int N, int D, int H, int W, int C);
/* dx, dgamma, dbeta of y = lrelu?(GN(x)); dy is the gradient w.r.t. y.  Two launches: reduction (whose last CTAs also
 * turn the partial sums into the per-group coefficients and dgamma / dbeta) and apply.
 * workspace: 8-byte aligned, b200_gn_backward_workspace_floats floats, ZERO-FILLED when allocated (its first words are
 * the tickets of the last-CTA pass; every call leaves them zero, so the buffer is reusable call after call). */
int b200_gn_backward(const void* x, const void* dy, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, void* dx, float* dgamma, float* dbeta, float* workspace, int N, int D, int H,
                     int W, int C, int do_lrelu, void* stream);

/* GroupNorm backward from the group sums a producing conv left in `gpart` ([ctas][N][16], b200_conv_run_gnbwd): no pass
 * over x and dy for the reduction; dgamma / dbeta come from per-CTA sums the apply kernel writes on its way. */
size_t b200_gn_backward_folded_workspace_floats(int N, int D, int H, int W, int C);
int b200_gn_backward_folded(const void* x, const void* dy, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const float* gpart, int ctas, void* dx, float* dgamma, float* dbeta,
                            float* workspace, int N, int D, int H, int W, int C, int do_lrelu, void* stream);

/* ---- trilinear x2 (aten::upsample_trilinear3d, model.py:7-14) fused with LeakyReLU
 *      (model.py:422); (N,D,H,W) is the COARSE volume -------------------------------------- */
int b200_upsample2x(const void* coarse, void* fine, int N, int D, int H, int W, int C, int do_lrelu, void* stream);
/* backward = adjoint of the interpolation applied to dfine * lrelu'(fine_out); separable in two passes
 * through `workspace` (b200_upsample2x_backward_workspace_bytes, 16-byte aligned, no initialisation needed) */
size_t b200_upsample2x_backward_workspace_bytes(int N, int D, int H, int W, int C);
int b200_upsample2x_backward(const void* dfine, const void* fine_out, void* dcoarse, void* workspace, int N, int D,
                             int H, int W, int C, int do_lrelu, void* stream);

/* ---- space-to-depth helpers for the 2x2x2 stride-2 conv (model.py:360-363) ---------------
 * (N,D,H,W) is the COARSE volume; coarse has 8*C channels, fine has C. */
int b200_space_to_depth(const void* fine, void* coarse, int N, int D, int H, int W, int C, void* stream);
int b200_depth_to_space(const void* coarse, const void* residual, void* fine, int N, int D, int H, int W, int C,
                        void* stream);
int b200_add(const void* a, const void* b, void* out, int N, int D, int H, int W, int C, void* stream);

/* ---- sigmoid backward (model.py:431) -> act gradient + conv_output.bias gradient -------- */
size_t b200_sigmoid_backward_workspace_floats(int N, int D, int H);
int b200_sigmoid_backward(const float* grad_probs, const float* probs, void* dlogit_act, float* dbias,
                          float* workspace, int N, int D, int H, int W, int Creal, int Cpad, void* stream);

/* ---- Dice_loss_joint (loss.py:98-122) ---------------------------------------------------- */
size_t b200_dice_workspace_floats(int B, int C);
/* sums[8]: I_c = sum p*g at [c], U_c = sum (p^2+g) at [4+c] (no epsilons; all-reducible). */
int b200_dice_sums(const float* probs, const float* target, float* sums, float* workspace, int B, int C, long long S,
                   void* stream);
int b200_dice_sums_t(const float* probs, const void* target, int target_dtype, float* sums, float* workspace, int B,
                     int C, long long S, void* stream);                 /* target_dtype = B200_F32 or B200_U8 */
int b200_dice_loss(const float* sums, int C, float priority, float* loss, void* stream);
int b200_dice_backward(const float* probs, const float* target, const float* sums, const float* grad_out,
                       float priority, float* grad_probs, int B, int C, long long S, void* stream);

int b200_dice_backward_t(const float* probs, const void* target, int target_dtype, const float* sums,
                         const float* grad_out, float priority, float* grad_probs, int B, int C, long long S, void* stream);

/* ---- BCE_Loss (loss.py:64-79) ------------------------------------------------
 * sum[1] = sum over elements of g log(p+1e-6) + bg_weight (1-g) log(1+1e-6-p)  (all-reducible);
 * loss = -sum / global_numel. */
size_t b200_bce_workspace_floats(void);
int b200_bce_sum(const float* probs, const float* target, float bg_weight, float* sum, float* workspace,
                 long long numel, void* stream);
int b200_bce_sum_t(const float* probs, const void* target, int target_dtype, float bg_weight, float* sum,
                   float* workspace, long long numel, void* stream);
int b200_bce_loss(const float* sum, double global_numel, float* loss, void* stream);
int b200_bce_backward(const float* probs, const float* target, const float* grad_out, float bg_weight,
                      double global_numel, float* grad_probs, long long numel, void* stream);

/* ---- fused Adam step (torch.optim.Adam(amsgrad, weight_decay) of main.py:133-138, stepped at
 *      train.py:220) ------------------------------------------------------------
 * One launch over flat fp32 buffers that share offsets: params, grads, exp_avg, exp_avg_sq and (amsgrad, may be
 * NULL) max_exp_avg_sq.  Only the element ranges [seg_begin[i], seg_begin[i] + seg_len[i]) are touched (host
 * arrays, multiples of 4 floats, <= 64 ranges): parameters without a gradient are skipped like torch does.
 * `lr`, `step` (fp32 count of steps taken, advanced by the kernel) and `ticket` (unsigned, zero) are DEVICE
 * scalars so the launch is CUDA-graph replayable.  lr_step_size > 0 applies StepLR(lr_step_size, lr_gamma)
 * stepped once per call (main.py:139-142, train.py:222-223) inside the kernel. */
int b200_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                   const long long* seg_begin, const long long* seg_len, int n_segs, const float* lr, float* step,
                   unsigned* ticket, double beta1, double beta2, float eps, float weight_decay, int lr_step_size,
                   float lr_gamma, void* stream);

int b200_bce_backward_t(const float* probs, const void* target, int target_dtype, const float* grad_out, float bg_weight,
                        double global_numel, float* grad_probs, long long numel, void* stream);

/* ---- plan introspection (tests): integer dump of the launch plan ---------------- */
int b200_conv_plan_debug(const b200_conv_desc* d, int* out, int n_out);
int b200_wgrad_plan_debug(const b200_wgrad_desc* d, int* out, int n_out);
/* plan of the line-marching weight-gradient kernel (3x3x3, 16 x 16 channels, W % 16 == 0; the default form for
 * those convs); error if it does not apply */
int b200_wgrad_line_plan_debug(const b200_wgrad_desc* d, int* out, int n_out);
int b200_wgrad_line_plan_debug2(const b200_wgrad_desc* d, int real_out, int real_in, int* out, int n_out);
/* plan of the marching weight-gradient kernel (3x3x3, 16 x 16 channels, W % 16 == 0); error if it does not apply */
int b200_wgrad_march_plan_debug(const b200_wgrad_desc* d, int* out, int n_out);
/* plan of the marching 3x3x3 kernel (Cin, Cout <= 32); error if it does not apply to `d` */
int b200_march_plan_debug(const b200_conv_desc* d, int* out, int n_out);
/* plan of the band-marching 3x3x3 kernel (16 -> 16 channels, wide lines); error if it does not apply */
int b200_band_plan_debug(const b200_conv_desc* d, int* out, int n_out);
/* perf probes: per-CTA cycle counters of the last marching-conv launch run with B200_CONV_DEBUG bit 256
 * (host buffer of n <= 2560 counters; synchronises the device) */
int b200_march_prof_read(unsigned long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* BRATS_B200_H */

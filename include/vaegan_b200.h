/*
 * vaegan_b200.h -- C ABI of libvaegan_sm100.so: hand-written sm_100a kernels for the VAE-GAN
 * training step of Don-Yin/VAE-GAN (reference: /root/reference/README.md, cited per entry).
 *
 * The reference has no FFI of its own: its "operator API" is the set of torch calls its
 * nn.Modules make (nn.Conv2d, nn.BatchNorm2d, ...).  Each entry point below replaces one of
 * those call sites (forward and the autograd backward torch would run for it); the
 * reference-side binding is a torch.autograd.Function that passes raw device pointers through
 * ctypes (see INTEGRATION.md and vae_gan_b200/_lib.py).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - activations are NHWC (channels innermost), dtype VG_F32 or VG_BF16;
 *     parameters, statistics and gradients of parameters are fp32 in torch's own layouts;
 *   - the caller owns every buffer (incl. workspaces); no compute entry point allocates device
 *     memory or keeps a pointer after it returns (the only allocator calls are the setup-time
 *     vg_peer_alloc / vg_peer_free helpers of the NVLink exchange buffers, which must be plain
 *     cudaMalloc memory so that their IPC handles can be opened by the other ranks; the only pointers kept
 *     are the caller-owned scratch / turn counters registered with vg_set_deterministic, until it is switched off);
 *   - kernels are launched with programmatic stream serialization and wait (griddepcontrol.wait) before their first
 *     global-memory access: stream-order semantics are unchanged, the launch latency overlaps the previous kernel's
 *     tail (VG_PDL=0 in the environment restores plain launches);
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation, CUDA-graph
 *     capturable;
 *   - return value: VG_OK or a negative VgStatus; vg_last_error() gives a thread-local
 *     message.  There is NO CPU fallback and no cuDNN/cuBLAS fallback: unsupported
 *     configurations are VG_EUNSUPPORTED.
 */
#ifndef VAEGAN_B200_H_
#define VAEGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vg_stream_t; /* cudaStream_t */

typedef enum {
  VG_OK = 0,
  VG_EINVAL = -1,       /* bad argument (null pointer, misaligned, inconsistent shape) */
  VG_ECUDA = -2,        /* CUDA runtime / launch failure */
  VG_EARCH = -3,        /* device is not sm_100 */
  VG_EUNSUPPORTED = -4  /* valid but not implemented by these kernels */
} VgStatus;

typedef enum { VG_F32 = 0, VG_BF16 = 1 } VgDtype;

/* ---- library ------------------------------------------------------------------------- */
int vg_version(void);
int vg_init(int device);                 /* checks compute capability 10.x, sets smem attrs */
const char* vg_last_error(void);
unsigned long long vg_launch_count(void);  /* kernels launched by this library so far */
/* force every convolution through the SIMT kernels (debug/bisect); returns previous value */
int vg_set_force_simt(int on);
/* Deterministic reductions (hard part 1; torch.use_deterministic_algorithms is what a user of the
 * reference would reach for).  While on, no floating-point sum that crosses thread blocks uses atomics:
 *   - per-channel / per-tap / scalar sums (BatchNorm statistics and backward sums, the second-order BatchNorm sums of
 *     the gradient penalty, single-channel weight gradients, bias gradients, spectral-norm dots, generator-loss
 *     scalars) are written as per-block partial sums into `scratch` and added in block order by a second small kernel;
 *   - split-K tensor-core convolutions / Linear layers, the tensor-core and CUDA-core weight gradients and the small
 *     GEMM keep their reductions but the splits of one output tile take turns in split order (`locks`: one int32 turn
 *     counter per output tile, zeroed by the caller, left zeroed by every launch);
 *   - BatchNorm statistics fused into a convolution epilogue and the bulk-copy single-channel weight gradient are
 *     switched off; configurations without an ordered variant (BatchNorm with C % 8 != 0 and C != 1) return
 *     VG_EUNSUPPORTED.
 * Both buffers are caller-owned device memory and must stay valid (and `scratch` unused by anything else on the
 * streams the library is called on) while the mode is on: scratch >= 1 MiB, 16-byte aligned (64 MiB covers every
 * BASELINE configuration), n_locks >= 1024.  Process-wide, one device.  on = 0 switches back (buffers may be NULL). */
int vg_set_deterministic(int on, void* scratch, size_t scratch_bytes, void* locks, int n_locks);
int vg_get_deterministic(void);

/* ---- convolutions: nn.Conv2d (README.md:148,151,163,165,170,379,383,387,441,556-571) and
 *      nn.ConvTranspose2d (README.md:156,158) -------------------------------------------- */
typedef struct {
  int n, h_in, w_in, c_in;      /* input activation, NHWC */
  int h_out, w_out, c_out;      /* output activation */
  int kh, kw, stride, pad;      /* square stride/pad as the reference uses */
  int transposed;               /* 0: Conv2d (weight OIHW), 1: ConvTranspose2d (weight IOHW) */
  int act_dtype;                /* VgDtype of x, dy, dx and of the weight packs */
  int out_dtype;                /* VgDtype of y (VG_F32 allowed with bf16 inputs) */
} VgConvDesc;

/* elements in one weight pack = kh*kw*c_in*c_out */
/* Re-layout (+ optional 1/sigma scaling for spectral norm, README.md:378) of a torch-layout
 * fp32 weight into the two GEMM packs the kernels consume, in act_dtype:
 *   pack_kn  [tap][c_out][c_in]   (reduction over c_in contiguous)
 *   pack_nk  [tap][c_in][c_out]   (reduction over c_out contiguous)
 * sigma: device scalar or NULL. */
int vg_conv_pack_weights(const VgConvDesc* d, const float* w, const float* sigma,
                         void* pack_kn, void* pack_nk, vg_stream_t stream);

/* The same re-layout for EVERY weight of a network in one launch (no sigma: with cached packs the spectral
 * norm is applied in the convolution epilogue, vg_conv_forward_scaled).  `items` is a HOST array; n0, n1 are the
 * first two dims of the torch weight ([c_out][c_in] for Conv2d / Linear, [c_in][c_out] for ConvTranspose2d),
 * taps = kh*kw.  A pack stays valid until the optimizer changes the weight. */
#define VG_PACK_MAX 32
typedef struct {
  const float* w;
  void* pack_kn;
  void* pack_nk;
  int n0, n1, taps, transposed;
} VgPackItem;
int vg_conv_pack_weights_batched(const VgPackItem* items, int n_items, int dtype, vg_stream_t stream);

/* y = conv(x) [+ bias] [* colscale[n][c_out]]  (colscale = Dropout2d keep/(1-p), README.md:413).
 * stats (nullable): double[2*c_out], accumulates per-channel sum(y) and sum(y^2) of the
 * values written, for the BatchNorm that follows (README.md:192).  Must be zeroed by caller. */
int vg_conv_forward(const VgConvDesc* d, const void* x, const void* pack_kn, const void* pack_nk,
                    const float* bias, const float* colscale, void* y, double* stats,
                    vg_stream_t stream);
/* dx = conv^T(dy): gradient w.r.t. the input. */
int vg_conv_dgrad(const VgConvDesc* d, const void* dy, const void* pack_kn, const void* pack_nk,
                  void* dx, vg_stream_t stream);
/* The same two with the spectral norm (README.md:378,383,387: weight = weight_orig / sigma) applied in the
 * epilogue instead of in the pack: y = conv(x, W) / sigma[g] (+ bias) (* colscale), dx = conv^T(dy, W) / sigma[g],
 * g = sample / sigma_group_n (sigma_group_n = 0: one sigma for the whole batch).  sigma NULL = plain call. */
int vg_conv_forward_scaled(const VgConvDesc* d, const void* x, const void* pack_kn, const void* pack_nk,
                           const float* bias, const float* colscale, const float* sigma, int sigma_group_n,
                           void* y, double* stats, vg_stream_t stream);
int vg_conv_dgrad_scaled(const VgConvDesc* d, const void* dy, const void* pack_kn, const void* pack_nk,
                         const float* sigma, int sigma_group_n, void* dx, vg_stream_t stream);
/* General forward epilogue - what the eval-mode / sampling path (README.md:661-664 decode, :1215-1226
 * visualize_reconstructions: generator.eval()) needs once the BatchNorms are folded into the weights:
 *   t  = conv(x, W) / sigma + bias ; t *= colscale ; [t = bf16(t) + residual] ; y = leaky_relu(t, act_slope)
 *   y2 = leaky_relu(post_scale[c] * y + post_shift[c], post_slope)          (second output, nullable)
 * so a pre-activation block becomes three convolution launches with no elementwise pass: conv1 carries the folded
 * bn2 as bias + LeakyReLU, the shortcut carries its folded BatchNorm as bias, conv2 adds the shortcut (`residual`)
 * and also emits the NEXT block's leaky_relu(bn1(.)) as y2.  act_slope / residual / y2 need a bf16 output on the
 * tensor-core path (channels % 64 == 0); otherwise VG_EUNSUPPORTED. */
typedef struct {
  const float* bias;
  const float* colscale;
  const float* sigma;
  int sigma_group_n;
  float act_slope;          /* 1 = no activation */
  const void* residual;     /* nullable; y's shape and dtype */
  void* y2;                 /* nullable; y's shape and dtype */
  const float* post_scale;  /* float[c_out], required with y2 */
  const float* post_shift;
  float post_slope;
} VgConvEpilogue;
int vg_conv_forward_fused(const VgConvDesc* d, const void* x, const void* pack_kn, const void* pack_nk,
                          const VgConvEpilogue* ep, void* y, double* stats, vg_stream_t stream);
/* Fold an eval-mode BatchNorm that FOLLOWS a convolution into it (sampling path): w_out = w * s[c_out],
 * bias_out = beta - running_mean * s (+ s * bias_in), s = gamma / sqrt(running_var + eps).  w in torch layout
 * ([c_out][c_in][k][k], or [c_in][c_out][k][k] when transposed); inner = elements per (c_out, c_in) pair = k*k. */
int vg_fold_bn_into_conv(const float* w, const float* bias_in, const float* gamma, const float* beta,
                         const float* running_mean, const float* running_var, float eps, int c_out, int c_in,
                         int inner, int transposed, float* w_out, float* bias_out, vg_stream_t stream);
/* (scale, shift) of an eval-mode BatchNorm as two float[c] vectors: scale = gamma / sqrt(running_var + eps),
 * shift = beta - running_mean * scale  (the post_scale / post_shift of VgConvEpilogue) */
int vg_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int c, float* scale, float* shift, vg_stream_t stream);
/* dw += x (*) dy in torch's weight layout, fp32 (caller zeroes dw for a fresh gradient);
 * dbias (nullable) += per-channel sum of dy.
 * workspace (nullable): float[kh*kw*c_in*c_out] scratch.  When given, the tensor-core kernel
 * accumulates its split-K partial sums into it with coalesced 16-byte vector reductions and a
 * second kernel adds the result into dw; without it the partial sums go straight into dw with
 * scalar atomics (slower).  Its contents are undefined afterwards. */
int vg_conv_wgrad(const VgConvDesc* d, const void* x, const void* dy, float* dw, float* dbias,
                  float* workspace, vg_stream_t stream);

/* Tile table ("shapes outside the tuned set ... autotuned tile table").  For one layer shape
 * (every field of `d` that describes geometry, including the batch n) and direction (dgrad = 0: vg_conv_forward*,
 * 1: vg_conv_dgrad*) pin the tensor-core kernel's N tile `bn` (64 / 128 / 256, must divide the output channels; 0 = keep
 * the heuristic) and its form (0 heuristic, 1 one-shot CTAs, 2 persistent CTAs, 3 persistent CTA pairs / cta_group::2; a
 * form the shape cannot take falls back to the next one).  bn = 0 and form = 0 removes the entry.  Shapes without an
 * entry use the built-in heuristics.  vg_conv_tune_record(1) starts collecting the distinct
 * (shape, direction) keys the tensor-core path is called with, vg_conv_tune_seen copies up to max_keys of them as
 * int[10] = {dgrad, n, h_in, w_in, c_in, c_out, k, stride, pad, transposed} and returns the total in *count -
 * vae_gan_b200/tune.py uses the pair to autotune whatever model is run.  Host-side state only; thread-safe. */
int vg_conv_tune_set(const VgConvDesc* d, int dgrad, int bn, int form);
int vg_conv_tune_clear(void);
int vg_conv_tune_record(int on);
int vg_conv_tune_seen(int* keys, int max_keys, int* count);

/* ---- BatchNorm2d (+LeakyReLU +Dropout) fused family: README.md:143,152,159,166,169,172,
 *      144/180/190 (nn.Dropout), 376,382,388,394,442 --------------------------------------- */
typedef struct {
  long long rows;               /* N*H*W of this rank's tensor */
  int c;                        /* channels */
  int hw;                       /* H*W (row / hw = local sample index) */
  int dtype;                    /* VgDtype of x, y, dy, dx */
  float slope;                  /* LeakyReLU negative slope; 1.0f = no activation */
  float drop_p;                 /* elementwise dropout probability; 0 = none */
  unsigned long long seed;      /* Philox key */
  unsigned long long offset;    /* Philox stream id (one per dropout site and step) */
  long long sample_offset;      /* global index of local sample 0 (partition invariance) */
  int training;                 /* 1: batch statistics; 0: running statistics (eval) */
  /* Optional DEVICE step counter: effective Philox stream = offset + 65536 * (*step_ptr).
   * Lets a captured CUDA graph draw fresh masks on every replay. */
  const unsigned long long* step_ptr;
} VgBnDesc;

/* sums[0..c) += sum_x, sums[c..2c) += sum_x^2 over rows (double, caller zeroes). */
int vg_bn_stats(const void* x, const VgBnDesc* d, double* sums, vg_stream_t stream);
/* mean_rstd[0..c) = mean, [c..2c) = 1/sqrt(var_biased+eps); running stats updated with
 * momentum and the UNBIASED variance when running_mean != NULL (torch semantics). */
int vg_bn_finalize(const double* sums, double count, int c, float eps, float momentum,
                   float* running_mean, float* running_var, float* mean_rstd, vg_stream_t stream);
/* eval mode: mean_rstd from running statistics. */
int vg_bn_eval_stats(const float* running_mean, const float* running_var, int c, float eps,
                     float* mean_rstd, vg_stream_t stream);
/* y = dropout(leaky_relu(gamma * (x - mean) * rstd + beta)) */
int vg_bn_act_forward(const void* x, const float* mean_rstd, const float* gamma, const float* beta,
                      const VgBnDesc* d, void* y, vg_stream_t stream);
/* g = dy * dropout' * lrelu'; sums[0..c) += sum g, sums[c..2c) += sum g*xhat (double). */
int vg_bn_act_backward_reduce(const void* dy, const void* x, const float* mean_rstd,
                              const float* gamma, const float* beta, const VgBnDesc* d,
                              double* sums, vg_stream_t stream);
/* dx = out_colscale[n][c] * gamma*rstd*(g - sum_g/count - xhat*sum_gx/count) + addend
 * (training=0: dx = gamma*rstd*g).  out_colscale, addend nullable. `count` is the GLOBAL
 * element count per channel (all ranks). */
int vg_bn_act_backward_apply(const void* dy, const void* x, const float* mean_rstd,
                             const float* gamma, const float* beta, const double* sums,
                             double count, const VgBnDesc* d, const float* out_colscale,
                             const void* addend, void* dx, vg_stream_t stream);
/* dgamma += sum g*xhat, dbeta += sum g  (fp32 params grads from the double sums) */
int vg_bn_param_grads(const double* sums, int c, float* dgamma, float* dbeta, vg_stream_t stream);

/* out = leaky_relu( bnA(a) + bnB(b) ), either BN optional (mean_rstd_* NULL = identity):
 * the `out += shortcut(x)` of README.md:183,195,405,417.  stats (nullable, double[2c]) +=
 * per-channel sum / sum of squares of `out` for the next block's bn1. */
int vg_bn_add_forward(const void* a, const float* mean_rstd_a, const float* gamma_a,
                      const float* beta_a, const void* b, const float* mean_rstd_b,
                      const float* gamma_b, const float* beta_b, const VgBnDesc* d, void* out,
                      double* stats, vg_stream_t stream);
/* ---- fused-statistics forms (round 2): the same BatchNorm call sites (README.md:143,152,159,166,169,
 *      376,382,388,442 and the residual adds :183,195,405,417) with the small per-layer kernels folded in.
 * A VgBnChannel says where one BatchNorm's (mean, rstd) come from:
 *   mean_rstd_in != NULL           : use them as given (float[2c]);
 *   else training && sums != NULL  : finalize from the FINAL (already all-reduced under SyncBN) double[2c]
 *                                    sum / sum of squares over `count` elements per channel; running_mean /
 *                                    running_var (nullable) are updated with `momentum` and the unbiased
 *                                    variance exactly like vg_bn_finalize;
 *   else (eval)                    : from running_mean / running_var like vg_bn_eval_stats.
 * In the last two cases mean_rstd_out (nullable, float[2c]) receives (mean, rstd) for the backward. */
typedef struct {
  const float* gamma;
  const float* beta;
  const float* mean_rstd_in;
  const double* sums;
  double count;
  float* running_mean;
  float* running_var;
  float* mean_rstd_out;
  float eps, momentum;
} VgBnChannel;
/* y = dropout(leaky_relu(batch_norm(x))): vg_bn_finalize | vg_bn_eval_stats + vg_bn_act_forward in ONE launch */
int vg_bn_act_forward_fused(const void* x, const VgBnChannel* bn, const VgBnDesc* d, void* y, vg_stream_t stream);
/* out = leaky_relu(bnA(a) + bnB(b)) (+ stats of out), bn_a / bn_b nullable = identity: both finalizes folded in */
int vg_bn_add_forward_fused(const void* a, const VgBnChannel* bn_a, const void* b, const VgBnChannel* bn_b,
                            const VgBnDesc* d, void* out, double* stats, vg_stream_t stream);
/* vg_bn_act_backward_apply + vg_bn_param_grads in ONE launch: additionally dgamma += param_scale * sum g*xhat,
 * dbeta += param_scale * sum g (both nullable; param_scale = 1/world under data parallelism, where the sums are
 * already global and the flat gradient buffer is summed over the ranks afterwards). */
int vg_bn_act_backward_apply_fused(const void* dy, const void* x, const float* mean_rstd,
                                   const float* gamma, const float* beta, const double* sums,
                                   double count, const VgBnDesc* d, const float* out_colscale,
                                   const void* addend, void* dx, float* dgamma, float* dbeta,
                                   float param_scale, vg_stream_t stream);
/* dgamma += scale * sum g*xhat, dbeta += scale * sum g (vg_bn_param_grads with the 1/world factor) */
int vg_bn_param_grads_scaled(const double* sums, int c, float scale, float* dgamma, float* dbeta, vg_stream_t stream);
/* Eval / sampling path: out = leaky_relu(a + b, d->slope) and, from out AS STORED, out2 = leaky_relu(post_scale[c] *
 * out + post_shift[c], post_slope) - the residual add of one block (README.md:195) and the next block's eval-mode
 * bn1 + LeakyReLU (README.md:188-189) in one pass (4 streams instead of 3 + 2). */
int vg_add_dual_forward(const void* a, const void* b, const float* post_scale, const float* post_shift, float post_slope,
                        const VgBnDesc* d, void* out, void* out2, vg_stream_t stream);
/* y = leaky_relu(x) on a flat fp32 tensor (discriminator head, README.md:475-481) */
int vg_lrelu_forward(const void* x, long long n, int dtype, float slope, void* y, vg_stream_t stream);
/* dx = dy * (y_ref > 0 ? 1 : slope) */
int vg_lrelu_backward(const void* dy, const void* y_ref, long long n, int dtype, float slope,
                      void* dx, vg_stream_t stream);
/* out = a + b (elementwise, same dtype) -- gradient accumulation at residual forks */
int vg_add(const void* a, const void* b, long long n, int dtype, void* out, vg_stream_t stream);

/* Second-order backward of BatchNorm(+LeakyReLU) for the gradient penalty (README.md:717-739,
   torch's batch_norm double backward): G = dL/d(dx) of vg_bn_act_backward_apply's dx.  reduce fills
   sums5 = double[5*C] (zeroed by the caller; all-reduced by the caller under SyncBN); apply writes
   g_dy = dL/d(dy) and g_x = dL/dx.  `colscale` = the out_colscale given to the first backward.
   dL/dgamma[c] = rstd[c] * (sums5[4][c] - sums5[0][c]*sums5[2][c]/M - sums5[1][c]*sums5[3][c]/M). */
int vg_bn_act_double_backward_reduce(const void* dy, const void* x, const void* G, const float* mean_rstd,
                                     const float* gamma, const float* beta, const VgBnDesc* d,
                                     const float* colscale, double* sums5, vg_stream_t stream);
int vg_bn_act_double_backward_apply(const void* dy, const void* x, const void* G, const float* mean_rstd,
                                    const float* gamma, const float* beta, const double* sums5, double count,
                                    const VgBnDesc* d, const float* colscale, void* g_dy, void* g_x,
                                    vg_stream_t stream);

/* ---- Philox4x32-10 randomness (replaces torch's bernoulli_/randn_like: README.md:144,381,581) */
/* keep-mask bytes for the elementwise dropout that vg_bn_act_forward applies */
int vg_dropout_mask(const VgBnDesc* d, uint8_t* mask, vg_stream_t stream);
/* Dropout2d scale per (n, c): 0 or 1/(1-p); index = (sample_offset + n)*c + ch */
int vg_dropout2d_scale(float* scale, int n, int c, float p, unsigned long long seed,
                       unsigned long long offset, const unsigned long long* step_ptr,
                       long long sample_offset, vg_stream_t stream);
/* standard normal noise, element i uses Philox index start+i */
int vg_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long offset,
                     const unsigned long long* step_ptr, long long start, vg_stream_t stream);
/* uniform [0,1) noise, element i = (Philox word of index start+i >> 8) * 2^-24: the gradient-penalty
   interpolation weights (replaces np.random.random, README.md:719) */
int vg_philox_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* step_ptr, long long start, vg_stream_t stream);
/* *counter += inc (device-side step counter used by the step_ptr arguments) */
int vg_counter_add(unsigned long long* counter, unsigned long long inc, vg_stream_t stream);

/* ---- avg_pool2d(k) + flatten in NCHW order (README.md:471-473) -------------------------- */
/* x / dx are NHWC in `dtype`; the pooled, flattened features (out / dout) are fp32 */
int vg_avgpool_flatten_forward(const void* x, int n, int h, int w, int c, int k, int dtype,
                               void* out, vg_stream_t stream);
int vg_avgpool_flatten_backward(const void* dout, int n, int h, int w, int c, int k, int dtype,
                                void* dx, vg_stream_t stream);

/* ---- nn.Linear (+LeakyReLU 0.2) README.md:458-461,474-483 ------------------------------- */
/* y[m][n] = lrelu(x[m][k] . w[n][k]^T + bias[n]).  The head's activations (x, y, dy, dx) are
 * fp32 (they are tiny); only the weight w is stored in `dtype`.  slope 1 = no activation. */
int vg_linear_forward(const void* x, const void* w, const float* bias, int m, int n, int k,
                      int dtype, float slope, void* y, vg_stream_t stream);
int vg_linear_dgrad(const void* dy, const void* w, int m, int n, int k, int dtype, void* dx,
                    vg_stream_t stream);
/* dw[n][k] += dy^T x (fp32), dbias[n] += column sums of dy */
int vg_linear_wgrad(const void* x, const void* dy, int m, int n, int k, int dtype, float* dw,
                    float* dbias, vg_stream_t stream);

/* ---- spectral norm (legacy torch hook, 1 power iteration; README.md:378,383,387) -------- */
/* w_orig [rows][cols] fp32.  training=1: v <- normalize(W^T u), u <- normalize(W v) in place.
 * sigma (device scalar) = u^T W v.  workspace: float[rows + cols + 4]. */
int vg_spectral_norm_sigma(const float* w_orig, int rows, int cols, float* u, float* v,
                           int training, float eps, float* sigma, float* workspace,
                           vg_stream_t stream);
/* The power iteration of EVERY spectral-normed weight of a network in three launches (+ one memset).  `items`
 * is a HOST array.  u, v: the module buffers (updated in place in training, like the hook does); u_out, v_out
 * (nullable): the copies of the post-iteration vectors that this forward's backward needs (later forwards of the
 * same iteration move the buffers on); sigma: device scalar per item.  workspace: sum(rows + cols) floats. */
#define VG_SN_MAX 16
typedef struct {
  const float* w;
  float* u;
  float* v;
  float* u_out;
  float* v_out;
  float* sigma;
  int rows, cols;
} VgSnItem;
int vg_spectral_norm_sigma_batched(const VgSnItem* items, int n_items, int training, float eps,
                                   float* workspace, size_t workspace_floats, vg_stream_t stream);
/* dw_orig += (dw_hat - <dw_hat, w_orig/sigma> u v^T) / sigma ; workspace: float[1] (callers may pass more) */
int vg_spectral_norm_backward(const float* dw_hat, const float* w_orig, const float* u,
                              const float* v, const float* sigma, int rows, int cols,
                              float* dw_orig, float* workspace, vg_stream_t stream);
/* Weight gradient of a spectral-normed convolution in ONE call (README.md:378-387: `utils.spectral_norm(nn.Conv2d(...))`
 * under autograd): dw_orig += (dW - <dW, w_orig / sigma> u v^T) / sigma with dW = vg_conv_wgrad(x, dy), for tensor-core
 * layers only (vg_conv_wgrad_sn_supported(d) == 1: bf16 activations, channels % 64 == 0; otherwise VG_EUNSUPPORTED - use
 * vg_conv_wgrad + vg_spectral_norm_backward).  The gradient stays in the tensor-core kernel's own layout
 * [tap][c_s][c_u] and the spectral-norm correction is applied while it is transposed into torch's layout, so the
 * unpack pass, the temporary dW and two of the three memsets of the two-call sequence disappear (7 -> 4 graph nodes per
 * layer and backward pass).  u, v, sigma: as saved by the forward.  workspace: float[kh*kw*c_in*c_out + 4], contents
 * undefined afterwards.  dbias (nullable) += column sums of dy. */
int vg_conv_wgrad_sn_supported(const VgConvDesc* d);
int vg_conv_wgrad_sn(const VgConvDesc* d, const void* x, const void* dy, const float* w_orig, const float* u, const float* v,
                     const float* sigma, float* dw_orig, float* dbias, float* workspace, vg_stream_t stream);


/* ---- reparameterisation (README.md:575-582) --------------------------------------------- */
/* lv = clamp(lv_raw, -50, 50); z = mu + exp(0.5 lv) * eps (training) or mu (eval). z in z_dtype */
int vg_reparam_forward(const float* mu, const float* lv_raw, const float* eps, long long n,
                       int training, int z_dtype, void* z, float* lv_clamped, vg_stream_t stream);
/* d_mu = dz ; d_lv_raw = (dz * 0.5*exp(0.5 lv)*eps + dlv_in) inside the clamp, 0 outside.
 * dlv_in (nullable): gradient arriving on the clamped log_var output (e.g. the KL term). */
int vg_reparam_backward(const void* dz, const float* lv_raw, const float* eps, const float* dlv_in,
                        long long n, int training, int z_dtype, float* d_mu, float* d_lv_raw,
                        vg_stream_t stream);

/* ---- fused generator loss + gradients (README.md:816-831) ------------------------------- */
typedef struct {
  long long n_pix;        /* elements of xhat / x on this rank */
  long long n_pix_global; /* mean denominators use the GLOBAL batch (data parallel) */
  long long n_lat;        /* elements of mu / log_var on this rank */
  int n_logits;           /* D(xhat) logits on this rank */
  int n_logits_global;
  int adv_mode;           /* 0: BCE-with-logits vs target 1 (north_star); 1: -mean(D) (ref) */
  float w_adv, w_recon, w_kl;
  int xhat_dtype;         /* VgDtype of xhat and d_xhat */
} VgLossDesc;
/* losses (device, double[4], caller zeroes): [0] total, [1] L1+MSE, [2] KL (sum), [3] adv.
 * Writes d_xhat (recon part only; the adversarial part arrives through D's dgrad),
 * d_mu, d_lv (KL part), d_logits -- all already multiplied by the loss weights. */
int vg_generator_loss(const void* xhat, const float* x, const float* mu, const float* lv,
                      const float* logits, const VgLossDesc* d, void* d_xhat, float* d_mu,
                      float* d_lv, float* d_logits, double* losses, vg_stream_t stream);
/* discriminator loss (README.md:792-793 or BCE): losses double[3] = total, real, fake */
int vg_discriminator_loss(const float* d_real, const float* d_fake, int n, int n_global,
                          int adv_mode, float* g_real, float* g_fake, double* losses,
                          vg_stream_t stream);

/* ---- fused optimizers over flat fp32 buffers (README.md:802-806,834,918-919) ------------ */
typedef struct {
  int kind;               /* 0 Adam (north_star), 1 RMSprop (reference) */
  float lr, beta1, beta2, alpha, eps, weight_decay;
  float bias_corr1, bias_corr2;   /* Adam: 1-beta^t */
  float clamp;            /* > 0: clamp params to +-clamp after the update (README.md:805) */
  float grad_scale;       /* multiplies the gradient first (1/world_size after allreduce-sum) */
} VgOptDesc;
/* step_ptr (nullable, device): when given, Adam's bias corrections are computed on the device
 * as 1 - beta^(*step_ptr) instead of taken from the descriptor (CUDA-graph replay). */
int vg_optimizer_step(float* p, const float* g, float* m, float* v, long long n,
                      const VgOptDesc* d, const unsigned long long* step_ptr, vg_stream_t stream);

/* ---- SyncBN statistics exchange over NVLink peer memory ------------------
 * One-shot all-reduce(sum) of a small fp64 vector (the 2C BatchNorm sums) without NCCL: every rank
 * stores its vector into slot [slot][rank] of EVERY peer's exchange buffer (plain stores through
 * peer-mapped pointers over NVLink/NVSwitch), publishes a flag, waits for all peers' flags and sums
 * the `world` sub-slots in rank order (bit-identical on all ranks).  Latency ~ one NVLink round trip
 * instead of a collective launch; 96 of these sit on the critical path of one training iteration.
 *   peer_data[r]  : rank r's data buffer,  double[n_slots][world][VG_PEER_MAX_N]   (peer-mapped)
 *   peer_flags[r] : rank r's flag buffer,  unsigned long long[n_slots][world]      (zero-initialised)
 *   epoch_ptr     : device counter, identical on all ranks, strictly increasing between two uses of
 *                   the same slot (the exchange's own per-iteration counter, >= 1; a flag passes the wait only
 *                   when it EQUALS the epoch, so stale and future values are both rejected)
 * The vector is reduced IN PLACE.  A rank must not reuse a slot before every peer has consumed it; using
 * each slot once per step with at least two exchanges per step guarantees that. */
#define VG_PEER_MAX_WORLD 8
#define VG_PEER_MAX_N 2048
typedef struct {
  void* peer_data[VG_PEER_MAX_WORLD];
  void* peer_flags[VG_PEER_MAX_WORLD];
  int rank, world, n_slots;
} VgPeerDesc;
/* enable P2P access from the current device to `peer_device` (idempotent) */
int vg_enable_peer_access(int peer_device);
/* Setup-time helpers (NOT on the hot path; the only calls that touch the allocator): the exchange
 * buffers are plain cudaMalloc allocations whose 64-byte CUDA IPC handles the other ranks open. */
int vg_peer_alloc(size_t bytes, void** ptr);                   /* cudaMalloc + zero fill */
int vg_peer_free(void* ptr);
int vg_peer_get_handle(void* ptr, void* handle64);             /* HOST buffer of 64 bytes */
int vg_peer_open_handle(const void* handle64, void** ptr);     /* maps a peer's buffer into this process */
int vg_peer_close_handle(void* ptr);
int vg_peer_allreduce_f64(double* vec, int n, const VgPeerDesc* pd, int slot,
                          const unsigned long long* epoch_ptr, vg_stream_t stream);

/* ---- layout / dtype helpers at the module boundary --------------------------------------- */
int vg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, vg_stream_t stream);
/* NCHW fp32 <-> NHWC dtype */
int vg_nchw_to_nhwc(const float* src, int n, int c, int h, int w, int dst_dtype, void* dst, vg_stream_t stream);
int vg_nhwc_to_nchw(const void* src, int src_dtype, int n, int c, int h, int w, float* dst, vg_stream_t stream);
int vg_fill_zero(void* p, size_t bytes, vg_stream_t stream);
/* ---- input pipeline (README.md:79-90 NiftyDataset.__getitem__, :785 imgs.type(Tensor)) -------------------
 * Per-image min-max normalisation to [0, 1] in float64 arithmetic, `(img - img.min()) / (img.max() - img.min())`
 * (README.md:87), fused with the cast to the fp32 tensor the step consumes (README.md:785) and, optionally, the
 * bf16 copy the convolutions read.  `raw`: n images of `pixels_per_image` voxel values in their STORAGE dtype
 * (so the host->device copy moves 1-2 B per pixel instead of the 8 B of the dataset's float64 arrays).
 * numpy semantics are kept: a constant image yields NaN (0/0), a NaN pixel poisons its image. */
typedef enum { VG_RAW_U8 = 0, VG_RAW_U16 = 1, VG_RAW_I16 = 2, VG_RAW_F32 = 3, VG_RAW_F64 = 4 } VgRawDtype;
int vg_normalize_images(const void* raw, int raw_dtype, int n, long long pixels_per_image, float* out_f32,
                        void* out_bf16, vg_stream_t stream);
/* dst = src * (*scale): `scale` is a DEVICE scalar (no host sync) - rescales the loss kernels' stored gradients
 * when the loss is back-propagated with a gradient other than 1 */
int vg_scale(const void* src, const float* scale, long long n, int dtype, void* dst, vg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAN_B200_H_ */

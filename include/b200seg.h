/*
 * b200seg.h -- C ABI of the B200-native U-Net + Dice hot path.
 *
 * One shared library (libb200seg.so), built for sm_100a only.  Plain pointers,
 * sizes and POD descriptors; no C++ or torch types cross this boundary.
 *
 * The reference (MrinalJain17/CT-image-segmentation) is pure Python and has no
 * native interface of its own: its hot path bottoms out in torch/MONAI calls.
 * Each entry point therefore cites the reference CALL SITE whose arithmetic it
 * replaces (file:line relative to the reference repository root); the Python
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - Activations are channels-last ("NDHWC"): element (n, d, h, w, c) of a
 *    tensor lives at  base[((((n*D + d)*H + h)*W + w) * ld) + c]  where `ld`
 *    (>= C) is the voxel stride in ELEMENTS.  ld > C addresses a channel slice
 *    of a wider buffer (zero-copy skip concatenation).  2-D data is D = 1.
 *  - dtype selects the storage type of activations and packed weights:
 *    B200SEG_BF16 (production; fp32 accumulation) or B200SEG_F32 (check mode).
 *  - All device memory, including workspaces, is owned by the caller.  The
 *    library never allocates or frees device memory and never synchronises;
 *    every launch is asynchronous on the `stream` argument (a cudaStream_t
 *    passed as void*).
 *  - Return value: 0 on success, negative on error; b200seg_last_error()
 *    returns a thread-local message.  No exceptions, no abort, no CPU fallback.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_VERSION 100 /* 0.1.0 */

enum b200seg_status {
  B200SEG_OK = 0,
  B200SEG_ERR_ARG = -1,         /* bad argument / shape / alignment */
  B200SEG_ERR_UNSUPPORTED = -2, /* configuration not implemented */
  B200SEG_ERR_WORKSPACE = -3,   /* workspace too small */
  B200SEG_ERR_CUDA = -4,        /* CUDA launch / runtime error */
  B200SEG_ERR_DEVICE = -5       /* not an sm_100 device */
};

enum b200seg_dtype { B200SEG_F32 = 0, B200SEG_BF16 = 1 };
enum b200seg_label_dtype { B200SEG_LABEL_U8 = 0, B200SEG_LABEL_I64 = 1 };

/* which of the four data-path GEMMs a packed weight feeds */
enum b200seg_weight_kind {
  B200SEG_W_CONV_FPROP = 0,
  B200SEG_W_CONV_DGRAD = 1,
  B200SEG_W_CONVTR_FPROP = 2,
  B200SEG_W_CONVTR_DGRAD = 3
};

enum b200seg_conv_flags {
  B200SEG_CONV_ACCUMULATE = 1,   /* destination += result (gradient fan-in) */
  B200SEG_CONV_FORCE_GENERIC = 2, /* use the CUDA-core kernel even where a tcgen05 kernel exists */
  /* x-, y- and residual-shaped tensors whose channel count C is not a multiple of 16 carry
   * ZERO padding in channels [C, round_up(C,16)) (so ld >= round_up(C,16)); kernels may read the
   * padding and rewrite it with zeros.  Lets the tcgen05 kernels take the 10-class layers. */
  B200SEG_CONV_PADDED_CHANNELS = 4,
  B200SEG_CONV_NO_SLIDE = 8, /* tcgen05 streaming kernel even where the sliding-window kernel applies (tests) */
  /* Layers with few output tiles (<= 74 CTAs) run on the split-K cluster kernel -- 2 or 4 CTAs of a thread-block
   * cluster share one output tile and reduce their partial accumulators through distributed shared memory.
   * Ignored where the kernel does not apply.  Default behaviour since 0.2 (environment B200SEG_CONV_SPLITK=0
   * switches the default off; the flag then opts a call in). */
  B200SEG_CONV_SPLIT_K = 16,
  B200SEG_CONV_NO_SPLIT_K = 32 /* streaming kernel even where the split-K kernel applies (tests, A/B timing) */
};

/*
 * Geometry of one Conv{2,3}d / ConvTranspose{2,3}d layer, in the layer's own
 * terms: x is the layer INPUT (cin channels, in_* extent), y its OUTPUT (cout
 * channels, out_* extent).  Conv:  out = floor((in + 2p - k)/s) + 1.
 * ConvTranspose: out = (in - 1)s - 2p + k + output_padding (= s*in here).
 */
typedef struct b200seg_conv_desc {
  int32_t n, cin, cout;
  int32_t in_d, in_h, in_w;
  int32_t out_d, out_h, out_w;
  int32_t kd, kh, kw; /* 1 or 3 */
  int32_t sd, sh, sw; /* 1 or 2 */
  int32_t pd, ph, pw;
  int32_t x_ld, y_ld, r_ld; /* voxel strides (elements) of x-shaped, y-shaped, residual tensors */
  int32_t dtype;            /* enum b200seg_dtype */
  int32_t flags;            /* enum b200seg_conv_flags */
} b200seg_conv_desc;

typedef struct b200seg_norm_desc {
  int32_t n, c;
  int64_t spatial;          /* D*H*W */
  int32_t x_ld, y_ld, r_ld; /* voxel strides (elements) */
  int32_t dtype;
  float eps;
} b200seg_norm_desc;

typedef struct b200seg_dice_desc {
  int32_t n, c;
  int64_t spatial;
  int32_t ld;               /* voxel stride of logits / dlogits (elements) */
  int32_t dtype;            /* logits dtype */
  int32_t label_dtype;      /* enum b200seg_label_dtype */
  int32_t include_background;
} b200seg_dice_desc;

/* ---- library ------------------------------------------------------------ */
int b200seg_version(void);
const char* b200seg_last_error(void);
/* number of CUDA kernels this library has launched in the process so far (all threads) */
long long b200seg_launch_count(void);
/* ... of which tcgen05 (tensor-core) kernels */
long long b200seg_tc_launch_count(void);
/* Name of the kernel family of the most recent launch made through this library (diagnostics and
 * tests: proves which implementation a layer was dispatched to).  Static storage, never NULL. */
const char* b200seg_last_launch(void);
/* 0 if `device` is an sm_100 part this library can run on */
int b200seg_check_device(int device);

/* ---- weights -------------------------------------------------------------
 * Repack an fp32 PyTorch-layout parameter (Conv: (cout,cin,kd,kh,kw);
 * ConvTranspose: (cin,cout,kd,kh,kw); reference state_dict)
 * into the kernel layout for one of the four data-path uses. */
size_t b200seg_packed_weight_bytes(const b200seg_conv_desc* d, int kind);
int b200seg_pack_weight(const b200seg_conv_desc* d, int kind, const float* w_torch,
                        void* w_packed, void* stream);

/* All bf16 weights of a network in ONE launch.  `table` is an array of n_entries b200seg_pack_entry
 * in DEVICE memory (built once; parameter and packed-buffer addresses are stable across steps):
 * entry i repacks fp32 parameter `w` into `packed` exactly as b200seg_pack_weight(kind) would
 * (packed must hold b200seg_packed_weight_bytes; tc_offset = offset of the tcgen05 layout in it,
 * as returned by b200seg_packed_weight_tc_offset). */
#define B200SEG_PACK_TC_ONLY 0x100 /* OR-ed into `kind`: skip the generic layout (layer always runs on tcgen05) */
typedef struct b200seg_pack_entry {
  uint64_t w;          /* const float*  (device) */
  uint64_t packed;     /* void*         (device) */
  uint64_t tc_offset;  /* bytes */
  int32_t taps, cin, cout, kind;
} b200seg_pack_entry;
size_t b200seg_packed_weight_tc_offset(const b200seg_conv_desc* d, int kind);
int b200seg_pack_weights_batched(const b200seg_pack_entry* table, int32_t n_entries, void* stream);

/* ---- convolutions ---------------------------------------------------------
 * Replace torch.nn.Conv{2,3}d / ConvTranspose{2,3}d forward + autograd as called
 * by monai.networks.nets.UNet, which the reference instantiates at
 * capstone/volumetric/base_trainer.py:65-72 and capstone/training/base_trainer.py:72-79
 * and calls at capstone/volumetric/base_trainer.py:74-78 (forward) / PL backward.
 *
 * fprop:  y = conv(x, w) [+ bias] [+ residual]          (residual is y-shaped, voxel stride r_ld)
 * dgrad:  dx = conv^T(dy, w) [+ residual] [+ dx if ACCUMULATE]   (residual is x-shaped)
 * wgrad:  gw (fp32, PyTorch layout) = d loss / d w ; gbias (fp32, may be NULL) = sum dy
 */
int b200seg_conv_fprop(const b200seg_conv_desc* d, const void* x, const void* w_packed,
                       const float* bias, const void* residual, void* y, void* stream);
int b200seg_conv_dgrad(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                       const void* residual, void* dx, void* stream);
size_t b200seg_conv_wgrad_workspace_bytes(const b200seg_conv_desc* d);
int b200seg_conv_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw,
                       float* gbias, void* workspace, size_t workspace_bytes, void* stream);

/* im2col of a small-Cin conv layer (taps*cin <= 32, the network's first convolutions, Cin = 1 or 3):
 * col[n, o, ci*taps + tap] = x[n, o*stride - pad + tap, ci] for every output voxel o, zero outside the
 * volume; `col` is (N, out_d, out_h, out_w, col_ld) channels-last with col_ld >= taps*cin (the rest is
 * written as zero padding).  With the PyTorch weight (cout, cin, k..) read as a (cout, cin*taps) matrix
 * the layer is a 1x1x1 convolution on `col`, which b200seg_conv_fprop / _wgrad run on the tcgen05
 * kernels (replaces the F.conv3d call of the first monai Convolution, capstone/volumetric/base_trainer.py:65-78). */
int b200seg_im2col(const b200seg_conv_desc* d, const void* x, void* col, int32_t col_ld, void* stream);

/* fprop fused with the statistics pass of the InstanceNorm that follows (no residual): the conv
 * epilogue emits per-CTA partial sums of y and y^2 (from the fp32 accumulators) into the workspace
 * and a small kernel reduces them (fixed order, double) to mean[n*stat_ld], rstd[n*stat_ld]
 * (stat_ld >= cout entries per sample; entries >= cout describe zero padding channels).
 * Returns B200SEG_OK when the statistics were produced, B200SEG_STATS_NOT_FUSED (1) when the
 * convolution ran on a kernel without the fusion (y is valid; call b200seg_instnorm_stats). */
#define B200SEG_STATS_NOT_FUSED 1
size_t b200seg_conv_fprop_stats_workspace_bytes(const b200seg_conv_desc* d);
size_t b200seg_convtr_fprop_stats_workspace_bytes(const b200seg_conv_desc* d);
int b200seg_conv_fprop_stats(const b200seg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                             void* y, float* mean, float* rstd, int32_t stat_ld, float eps, void* workspace,
                             size_t workspace_bytes, void* stream);
int b200seg_convtr_fprop_stats(const b200seg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                               void* y, float* mean, float* rstd, int32_t stat_ld, float eps, void* workspace,
                               size_t workspace_bytes, void* stream);

/* The same fusion with the finalisation moved into the CONSUMER: conv_fprop_partials runs the convolution
 * (transposed_layer != 0: ConvTranspose) and leaves the per-CTA partial statistics in `partials`
 * (b200seg_conv[tr]_fprop_stats_workspace_bytes bytes), reporting their layout; returns
 * B200SEG_STATS_NOT_FUSED (1) when the layer ran on a kernel without the fusion (y is complete, use
 * b200seg_instnorm_stats).  instnorm_prelu_fwd_partials = InstanceNorm + PReLU (+ residual) whose blocks
 * reduce those partials themselves (cstat = the convolution's cout; channels [cstat, d->c) are zero padding)
 * and write mean / rstd (n * d->c floats each) for the backward pass. */
int b200seg_conv_fprop_partials(const b200seg_conv_desc* d, int32_t transposed_layer, const void* x,
                                const void* w_packed, const float* bias, void* y, float* partials,
                                size_t partials_bytes, int32_t* ncls, int64_t* tiles, void* stream);
int b200seg_instnorm_prelu_fwd_partials(const b200seg_norm_desc* d, const void* x, const float* partials,
                                        int32_t ncls, int64_t tiles, int32_t cstat, float* mean, float* rstd,
                                        const float* alpha, const void* residual, void* y, void* stream);

int b200seg_convtr_fprop(const b200seg_conv_desc* d, const void* x, const void* w_packed,
                         const float* bias, const void* residual, void* y, void* stream);
int b200seg_convtr_dgrad(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                         const void* residual, void* dx, void* stream);
size_t b200seg_convtr_wgrad_workspace_bytes(const b200seg_conv_desc* d);
int b200seg_convtr_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw,
                         float* gbias, void* workspace, size_t workspace_bytes, void* stream);

/* ---- InstanceNorm(affine=False, eps, biased var) + PReLU(one shared alpha) ---
 * Replace the `norm` and `act` children of monai Convolution and the
 * residual sum of monai ResidualUnit (A.3), reached from the same call sites as above.
 *
 * stats:  mean[n*c], rstd[n*c]   (fp32)
 * fwd:    y = prelu((x - mean) * rstd, alpha) [+ residual]
 * bwd:    dx = d loss / d x  given dy = d loss / d y ;  dalpha[0] = d loss / d alpha
 * alpha / dalpha are DEVICE pointers to one float.
 * Ordering contract of the backward: for instances of up to 32^3 voxels dalpha is finished inside the kernel that
 * produces the per-(sample, channel) sums (the CTA that draws the last ticket of a per-device counter adds them up in
 * a fixed order), so two b200seg_instnorm_prelu_bwd calls on ONE device must be stream-ordered with respect to each
 * other (same stream, or event dependencies) -- as they are in any backward pass.
 */
size_t b200seg_instnorm_workspace_bytes(const b200seg_norm_desc* d);
int b200seg_instnorm_stats(const b200seg_norm_desc* d, const void* x, float* mean, float* rstd,
                           void* workspace, size_t workspace_bytes, void* stream);
int b200seg_instnorm_prelu_fwd(const b200seg_norm_desc* d, const void* x, const float* mean,
                               const float* rstd, const float* alpha, const void* residual,
                               void* y, void* stream);
int b200seg_instnorm_prelu_bwd(const b200seg_norm_desc* d, const void* x, const float* mean,
                               const float* rstd, const float* alpha, const void* dy, void* dx,
                               float* dalpha, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ---- softmax + Dice ---------------------------------------------------------
 * Replace monai.losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
 * as built at capstone/models/losses.py:78-85 and capstone/volumetric/losses.py:70-77
 * (formula identical to the in-tree capstone/models/temp.py:137-157 with uniform weights).
 *
 * fwd:  sums[n][c][3] (fp32) = { I = sum_v p*t, G = sum_v t, P = sum_v p }, p = softmax(logits),
 *       t = one_hot(labels).  (With include_background == 0 the c = 0 row is still produced.)
 * bwd:  dlogits given gI[n][c] = d loss / d I and gP[n][c] = d loss / d P (fp32; rows of
 *       excluded classes must be 0):  g_c = gI*t_c + gP ;  dz_c = p_c (g_c - sum_k g_k p_k).
 * The per-(n,c) epilogue  f = 1 - (2I + s)/(G + P + s), its reduction and the
 * missing-annotation weighting (capstone/models/losses.py:206-221) operate on n*c numbers and
 * stay in host-side PyTorch.
 */
size_t b200seg_softmax_dice_workspace_bytes(const b200seg_dice_desc* d);
int b200seg_softmax_dice_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                             float* sums, void* workspace, size_t workspace_bytes, void* stream);
/* Loss sums AND the Dice-metric counts of the same step in one pass over the logits: the reference runs its metric on
 * every training step (`_log_dice_scores`, capstone/volumetric/base_trainer.py:116-132 -> `_squash_predictions`
 * capstone/training/utils.py:19-20 -> `compute_meandice` capstone/models/temp.py:173-214).  sums as
 * b200seg_softmax_dice_fwd; counts[n][c][3] = {tp, |pred|, |target|} (int64) as b200seg_argmax_dice_counts. */
size_t b200seg_softmax_dice_metric_workspace_bytes(const b200seg_dice_desc* d);
int b200seg_softmax_dice_metric_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float* sums,
                                    int64_t* counts, void* workspace, size_t workspace_bytes, void* stream);
int b200seg_softmax_dice_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                             const float* gI, const float* gP, void* dlogits, void* stream);

/* dgrad of a 3x3x3 stride-1 layer FUSED with the reduction pass of the InstanceNorm + PReLU backward that consumes
 * its result (the mirror of b200seg_conv_fprop_partials): monai Convolution = conv -> InstanceNorm -> PReLU, so the
 * input gradient dx of layer L+1 is the output gradient of layer L's PReLU/InstanceNorm, whose backward starts with
 * three per-(sample, channel) sums over all voxels (sum g~, sum g~ xhat, sum dy xhat [xhat <= 0], g~ = dy prelu'(xhat)).
 * The dgrad epilogue has dx in registers: it loads layer L's pre-norm tensor `norm_x` (same voxels), forms the sums
 * from the value AS STORED and leaves per-CTA partials; b200seg_instnorm_prelu_bwd_from_partials reduces them (fixed
 * order, double) and runs the apply pass.  One full read of dy and norm_x (a bandwidth pass over the layer) is gone.
 * Returns B200SEG_STATS_NOT_FUSED (1) WITHOUT launching anything where the fused kernel does not apply (it exists for
 * the sliding-window tcgen05 kernel with 16 padded channels on both sides: the head layer and the 16->16 layers):
 * the caller then runs b200seg_conv_dgrad + b200seg_instnorm_prelu_bwd.  mean / rstd hold `stat_ld` (>= 16) entries
 * per sample; partials: b200seg_conv_dgrad_instnorm_partials_bytes(d) (0 = not fusable). */
size_t b200seg_conv_dgrad_instnorm_partials_bytes(const b200seg_conv_desc* d);
int b200seg_conv_dgrad_instnorm_partials(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                                         const void* residual, void* dx, const void* norm_x, int32_t norm_x_ld,
                                         const float* mean, const float* rstd, int32_t stat_ld, const float* alpha,
                                         float* partials, size_t partials_bytes, int64_t* rows_per_sample,
                                         void* stream);
int b200seg_instnorm_prelu_bwd_from_partials(const b200seg_norm_desc* d, const void* x, const float* mean,
                                             const float* rstd, const float* alpha, const void* dy,
                                             const float* partials, int64_t rows_per_sample, void* dx, float* dalpha,
                                             void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam step (the reference's configure_optimizers, capstone/volumetric/base_trainer.py:178-182:
 * Adam(lr), default betas / eps, no weight decay, no amsgrad) on ONE flat fp32 parameter buffer whose
 * gradient is the flat all-reduce bucket; `step` is the 1-based update count (bias correction). */
int b200seg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, int64_t step, void* stream);

/* Every voxel-wise loss the reference offers through ONE softmax pass (capstone/models/losses.py:45-124,
 * LOSSES table :160-167; 3-D twins capstone/volumetric/losses.py:37-126): sums5[n][c][5] = {I, G, P, F, N},
 *   I = sum p t, G = sum t, P = sum p             -> monai DiceLoss / in-tree GeneralizedDiceLoss
 *   F = sum t (1 - p)^gamma (-log p)              -> monai FocalLoss(gamma) on the one-hot target (AsDiscrete)
 *   N = sum t (-log p)                            -> F.cross_entropy, plain or with the class WEIGHT table
 * with p = softmax(logits), t = one_hot(labels), log p as log_softmax computes it.  bwd: dlogits from
 * d loss / d{I, P, F, N} per (n, c) (n*c floats each).  At most 16 classes. */
size_t b200seg_softmax_loss_workspace_bytes(const b200seg_dice_desc* d);
int b200seg_softmax_loss_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float gamma,
                             float* sums5, void* workspace, size_t workspace_bytes, void* stream);
int b200seg_softmax_loss_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float gamma,
                             const float* gI, const float* gP, const float* gF, const float* gN, void* dlogits,
                             void* stream);

/* The same pass with the Boundary loss added (capstone/models/losses.py:127-157: BoundaryLossWrapper =
 * mean(softmax(x)[:, 1:] * dist_maps); dispatched at :186-191 with the batch's pre-computed signed distance maps,
 * capstone/data/utils.py:10-26): sums6[n][c][6] = {I, G, P, F, N, B},  B = sum_v p_c(v) * dist_maps[n][c-1][v]
 * for c >= 1 (B[.,0] = 0).  dist_maps is PLANAR (n, c-1, spatial) fp32 as the dataset delivers it.  bwd takes the
 * extra coefficient gB = d loss / d B (n*c floats; entry 0 of every row ignored). */
size_t b200seg_softmax_boundary_loss_workspace_bytes(const b200seg_dice_desc* d);
int b200seg_softmax_boundary_loss_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                                      const float* dist_maps, float gamma, float* sums6, void* workspace,
                                      size_t workspace_bytes, void* stream);
int b200seg_softmax_boundary_loss_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                                      const float* dist_maps, float gamma, const float* gI, const float* gP,
                                      const float* gF, const float* gN, const float* gB, void* dlogits,
                                      void* stream);

/* Dice loss value and gradient coefficients from the (n, c, 3) sums of b200seg_softmax_dice_fwd, in one
 * launch: the arithmetic of monai.losses.DiceLoss.forward after the spatial sums
 * (capstone/models/losses.py:80-85 configures it; formula in SURVEY.md A.6):
 *   f = 1 - (2 I + smooth) / (G + P + smooth); loss = mean (mean != 0) or sum of f over (n, foreground c);
 *   gI = d loss / d I, gP = d loss / d P (n*c floats each, zero for an excluded background) -- the
 *   inputs of b200seg_softmax_dice_bwd. */
int b200seg_dice_loss_epilogue(const float* sums, int32_t n, int32_t c, int32_t include_background,
                               float smooth, int32_t mean, float* loss, float* gI, float* gP, void* stream);

/* ---- label maps and the Dice metric -------------------------------------------
 * argmax:  pred[v] = argmax_c softmax(logits)[c], first maximum wins
 *          (capstone/training/utils.py:19-20 `_squash_predictions`).
 * counts:  counts[n][c][3] (int64) = { |pred==c & target==c|, |pred==c|, |target==c| }, the
 *          integer form of compute_meandice's sums (capstone/models/temp.py:173-214), used by
 *          DiceMetricWrapper (capstone/models/metrics.py:15-21).  `pred_out` (uint8) may be NULL;
 *          `target` may be NULL (then only the label map is produced and counts is untouched).
 * squash:  labels[v] = max_c masks[c][v] * (c+1)  (capstone/volumetric/utils.py:4-7,
 *          capstone/training/utils.py:13-16); masks are (n, n_struct, spatial) uint8, NCDHW.
 */
int b200seg_argmax_dice_counts(const b200seg_dice_desc* d, const void* logits, const void* target,
                               uint8_t* pred_out, int64_t* counts, void* stream);
int b200seg_label_dice_counts(int32_t n, int64_t spatial, int32_t c, const uint8_t* pred,
                              const void* target, int32_t target_dtype, int64_t* counts,
                              void* stream);
int b200seg_squash_masks(int32_t n, int32_t n_struct, int64_t spatial, const uint8_t* masks,
                         uint8_t* labels, void* stream);

/* ---- HU windowing + normalisation ------------------------------------------------
 * out[v][k] = ((clip(hu[v], lo_k, hi_k) - lo_k) / (hi_k - lo_k + 1e-8) - mean_k) / std_k
 * (capstone/transforms/transforms_2d.py:97-107 `apply_window`, then albumentations Normalize
 * with the constants of capstone/transforms/predefined.py:5-29).  hu: int16; out: channels-last
 * `dtype`, voxel stride out_ld.  lo/hi/mean/std are HOST arrays of n_windows (<= 4) floats.
 */
int b200seg_hu_window_norm(int64_t n_vox, int32_t n_windows, const int16_t* hu, const float* lo,
                           const float* hi, const float* mean, const float* std_, void* out,
                           int32_t out_ld, int32_t dtype, void* stream);

/* ---- sliding-window inference (MONAI `sliding_window_inference` semantics,
 * Appendix A.7 -- absent from the reference, which resizes whole volumes instead,
 * capstone/volumetric/transforms.py:9-23) --------------------------------------------------------
 * window_accumulate: acc[(d0+d, h0+h, w0+w)][c] += window_logits[(d,h,w)][c]; cnt[...] += 1
 *                    (fp32 channels-last accumulators of the whole D x H x W volume, constant
 *                    importance map; the part of a window overhanging the volume is dropped)
 * accum_argmax:      labels[v] = argmax_c softmax(acc[v]/cnt[v]) (first maximum);
 *                    mean_logits (fp32, may be NULL) receives acc/cnt.
 */
int b200seg_window_accumulate(int32_t dtype, const void* window_logits, int32_t src_ld, float* acc, float* cnt,
                              int32_t c, int32_t wd, int32_t wh, int32_t ww, int32_t D, int32_t H, int32_t W,
                              int32_t d0, int32_t h0, int32_t w0, void* stream);
/* ... with an importance map: acc += importance[(d,h,w)] * logits, cnt += importance[(d,h,w)] (fp32, one value per
 * window voxel, MONAI mode="gaussian"; NULL = constant 1).  `window_logits` / `importance` may point INTO a window
 * (a run of its d-slices: wd = the run's depth), which is how a window is split between the d-slabs owned by
 * different ranks (inference.py). */
int b200seg_window_accumulate_weighted(int32_t dtype, const void* window_logits, int32_t src_ld,
                                       const float* importance, float* acc, float* cnt, int32_t c, int32_t wd,
                                       int32_t wh, int32_t ww, int32_t D, int32_t H, int32_t W, int32_t d0, int32_t h0,
                                       int32_t w0, void* stream);
int b200seg_accum_argmax(const float* acc, const float* cnt, uint8_t* labels, float* mean_logits, int64_t n_vox,
                         int32_t c, void* stream);

/* ---- patch sampler (replaces the whole-volume Resize3D of
 * capstone/volumetric/transforms.py:9-23 / datasets.py:24-48) ---------------------------------------
 * For each of n_patches origins (int32 triples d,h,w on the DEVICE; may be negative / overhang):
 * img_out[b] = window+normalise(hu[origin_b + .]) as `dtype`, lab_out[b] = labels[origin_b + .]
 * (out-of-volume voxels read as pad_hu / label 0).  labels / lab_out may be NULL.
 */
int b200seg_crop_window_norm(int32_t dtype, const int16_t* hu, const uint8_t* labels, const int32_t* origins,
                             int32_t n_patches, void* img_out, uint8_t* lab_out, int32_t D, int32_t H, int32_t W,
                             int32_t pd, int32_t ph, int32_t pw, float lo, float hi, float mean, float std_,
                             int32_t pad_hu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */

// Internal launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200seg.h"

namespace b200seg {

struct GatherParams {
  int n;
  int sD, sH, sW;  // extent of the gathered (source) tensor
  int dD, dH, dW;  // extent of the destination tensor
  int src_c, dst_c;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int src_ld, dst_ld, res_ld;
  int transposed;  // 0: src = dst*s - p + k ; 1: src = (dst + p - k)/s
  int accumulate;
  // filled by launch_gather
  int csd, csh, csw;  // parity classes per dim (stride for the transposed gather, else 1)
  int cD, cH, cW;     // per-class grid
  int64_t per_class;
  int tiles_per_class;
  int src_vec;
};

struct WgradParams {
  int n;
  int sD, sH, sW;  // extent of S (gathered operand)
  int tD, tH, tW;  // extent of T (indexed by output position)
  int a_c, b_c;    // channels of S and T
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int s_ld, t_ld;
  // filled by wgrad_plan
  int taps, tiles_a, tiles_b, splits;
  int64_t vox_total, vox_per_split;
};

int launch_gather(GatherParams p, int dtype, const void* src, const void* w, const float* bias,
                  const void* res, void* dst, cudaStream_t st);
void wgrad_plan(WgradParams& p);
size_t wgrad_partial_bytes(const WgradParams& p);
int launch_wgrad(const WgradParams& p, int dtype, const void* S, const void* T, float* gw,
                 float* partial, cudaStream_t st);
size_t colsum_workspace_bytes(int64_t nvox, int c);
int launch_colsum(int dtype, const void* x, int64_t nvox, int c, int ld, float* out, float* partial,
                  cudaStream_t st);
int launch_pack_weight(int dtype, int kind, const float* w, void* packed, int taps, int cin,
                       int cout, cudaStream_t st);

// conv_small_cin.cu  (Cin <= 4 first-layer convolutions)
bool small_cin_supported(const b200seg_conv_desc* d);
size_t small_cin_wgrad_workspace(const b200seg_conv_desc* d);
int launch_small_cin_fprop(const b200seg_conv_desc* d, const void* x, const void* w, const float* bias,
                           const void* res, void* y, cudaStream_t st);
int launch_im2col(const b200seg_conv_desc* d, const void* x, void* col, int col_ld, cudaStream_t st);
int launch_small_cin_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw, float* partial,
                           cudaStream_t st);

// norm.cu
int norm_blocks(const b200seg_norm_desc& d);
size_t norm_workspace_bytes(const b200seg_norm_desc& d);
int launch_instnorm_stats(const b200seg_norm_desc& d, const void* x, float* mean, float* rstd,
                          void* ws, cudaStream_t st);
int launch_instnorm_stats_from_partials(const float* partial, int n, int c, int c_out, int ncls, int64_t tiles,
                                        int64_t spatial, float eps, float* mean, float* rstd, cudaStream_t st);
int launch_instnorm_prelu_fwd(const b200seg_norm_desc& d, const void* x, const float* mean,
                              const float* rstd, const float* alpha, const void* res, void* y,
                              cudaStream_t st);
int launch_instnorm_prelu_fwd_partials(const b200seg_norm_desc& d, const void* x, const float* partial, int ncls,
                                       int64_t tiles, int cstat, const float* alpha, const void* res, void* y,
                                       float* mean, float* rstd, cudaStream_t st);
int launch_instnorm_prelu_bwd(const b200seg_norm_desc& d, const void* x, const float* mean,
                              const float* rstd, const float* alpha, const void* dy, void* dx,
                              float* dalpha, void* ws, cudaStream_t st);
int launch_instnorm_prelu_bwd_from_partials(const b200seg_norm_desc& d, const void* x, const float* mean,
                                            const float* rstd, const float* alpha, const void* dy,
                                            const float* partial, int64_t rows, void* dx, float* dalpha, void* ws,
                                            cudaStream_t st);

int launch_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t step, cudaStream_t st);

// dice.cu
size_t dice_workspace_bytes(const b200seg_dice_desc& d);
int launch_softmax_dice_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                            float* sums, void* ws, cudaStream_t st);
size_t dice_metric_workspace_bytes(const b200seg_dice_desc& d);
int launch_softmax_dice_metric_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float* sums,
                                   int64_t* counts, void* ws, cudaStream_t st);
int launch_softmax_dice_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                            const float* gI, const float* gP, void* dlogits, cudaStream_t st);
size_t loss_workspace_bytes(const b200seg_dice_desc& d);
int launch_softmax_loss_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float gamma,
                            float* sums5, void* ws, cudaStream_t st);
int launch_softmax_loss_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float gamma,
                            const float* gI, const float* gP, const float* gF, const float* gN, void* dlogits,
                            cudaStream_t st);
size_t boundary_workspace_bytes(const b200seg_dice_desc& d);
int launch_softmax_boundary_loss_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                                     const float* dist, float gamma, float* sums6, void* ws, cudaStream_t st);
int launch_softmax_boundary_loss_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                                     const float* dist, float gamma, const float* gI, const float* gP,
                                     const float* gF, const float* gN, const float* gB, void* dlogits,
                                     cudaStream_t st);
int launch_dice_loss_epilogue(const float* sums, int n, int c, int c0, float smooth, float inv_count, float* loss,
                              float* gI, float* gP, cudaStream_t st);
int launch_argmax_dice_counts(const b200seg_dice_desc& d, const void* logits, const void* target,
                              uint8_t* pred_out, int64_t* counts, cudaStream_t st);
int launch_label_dice_counts(int n, int64_t spatial, int c, const uint8_t* pred, const void* target,
                             int target_dtype, int64_t* counts, cudaStream_t st);
int launch_squash_masks(int n, int n_struct, int64_t spatial, const uint8_t* masks, uint8_t* labels,
                        cudaStream_t st);
int launch_window_accumulate(int dtype, const void* src, int src_ld, const float* imp, float* acc, float* cnt, int C,
                             int wd, int wh, int ww, int D, int H, int W, int d0, int h0, int w0, cudaStream_t st);
int launch_accum_argmax(const float* acc, const float* cnt, uint8_t* labels, float* mean_out, int64_t nvox, int C,
                        cudaStream_t st);
int launch_crop_window_norm(int dtype, const int16_t* hu, const uint8_t* lab, const int* origins, int nb_patches,
                            void* img_out, uint8_t* lab_out, int D, int H, int W, int pd, int ph, int pw, float lo,
                            float hi, float mean, float stdv, int pad_hu, cudaStream_t st);
int launch_hu_window_norm(int64_t n_vox, int n_windows, const int16_t* hu, const float* lo,
                          const float* hi, const float* mean, const float* std_, void* out,
                          int out_ld, int dtype, cudaStream_t st);

}  // namespace b200seg

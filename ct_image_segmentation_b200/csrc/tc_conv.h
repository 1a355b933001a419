// tcgen05 / TMEM / TMA implicit-GEMM convolution family (bf16, fp32 accumulation).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/b200seg.h"

namespace b200seg {

enum TcConvOp { TC_CONV_FPROP = 0, TC_CONV_DGRAD = 1, TC_CONVTR_FPROP = 2, TC_CONVTR_DGRAD = 3 };

// true when the tcgen05 kernel takes this layer/op (bf16, channel counts multiple of 16 or
// zero-padded to it, 16-byte aligned pointers, ...)
bool tc_conv_supported(const b200seg_conv_desc* d, int op, const void* src, const void* dst, const void* res);
// stats (optional, fprop without residual only): [blockIdx.x][cout][2] fp32 per-CTA sum / sum of squares;
// blockIdx.x = (class * n + sample) * tiles + tile (tc_conv_grid) or sample-major (tc_slide_conv_grid)
void tc_conv_grid(const b200seg_conv_desc* d, int op, int* ncls_out, int64_t* tiles_out);
int64_t tc_slide_conv_grid(const b200seg_conv_desc* d, int op, bool bst = false);  // rows of statistic partials
int tc_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                const void* residual, void* dst, float* stats, cudaStream_t st);
// sliding-window variant for small-channel, high-resolution 3x3x3 stride-1 layers (tc_slide.cu)
bool tc_slide_conv_supported(const b200seg_conv_desc* d, int op);
// optional fusion of a dgrad with the reduction pass of the InstanceNorm + PReLU backward its result feeds:
// nx = that layer's pre-norm tensor (voxel stride nx_ld), mean / rstd with nstat_ld entries per sample, alpha its
// slope; partials [CTA][16][3] (CTAs of one sample contiguous: tc_slide_conv_grid(d, op) / n rows per sample)
struct TcBwdStats {
  const void* nx;
  int nx_ld, nstat_ld;
  const float* mean;
  const float* rstd;
  const float* alpha;
  float* partials;
};
bool tc_slide_conv_bwdstats_supported(const b200seg_conv_desc* d, int op);
int tc_slide_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                      const void* residual, void* dst, float* stats, cudaStream_t st,
                      const TcBwdStats* bst = nullptr);
// line-tiled variant of the sliding kernel (tc_line.cu): 16 -> 16 (padded) channels, row length 32 / 64 / 128, source
// voxel stride exactly 16 elements; tc_slide_conv_supported / _grid / _run route to it where it applies
bool tc_line_conv_supported(const b200seg_conv_desc* d, int op);
int64_t tc_line_conv_rows(const b200seg_conv_desc* d, int op);
int tc_line_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                     const void* residual, void* dst, float* stats, cudaStream_t st, const TcBwdStats* bst);
// sliding-window kernels for the high-resolution stride-2 layers, ConvTranspose and Conv (tc_convtr.cu)
bool tc_convtr_slide_supported(const b200seg_conv_desc* d, int op, const void* residual);
int64_t tc_convtr_slide_grid(const b200seg_conv_desc* d, int op);
int tc_convtr_slide_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                        void* dst, float* stats, cudaStream_t st);
bool tc_convtr_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer);
size_t tc_convtr_wgrad_workspace(const b200seg_conv_desc* d, bool transposed_layer);
int tc_convtr_wgrad_run(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy, float* gw,
                        float* G32, cudaStream_t st);
bool tc_slide_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer);
size_t tc_slide_wgrad_workspace(const b200seg_conv_desc* d);
int tc_slide_wgrad_run(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw, float* G32,
                       cudaStream_t st);
size_t tc_packed_weight_bytes(const b200seg_conv_desc* d);
// packs the tcgen05 layout into `out` and (if gen != NULL) the generic bf16 layout into `gen`, one launch
int tc_pack_weight(const b200seg_conv_desc* d, int kind, const float* w, void* out, void* gen, cudaStream_t st);
int tc_pack_weights_batched(const b200seg_pack_entry* table_dev, int n_entries, cudaStream_t st);
size_t tc_wgrad_extra_workspace(const b200seg_conv_desc* d);
// weight gradient on tcgen05: x / dy in the layer's own terms, gw in PyTorch layout, G32 = fp32 scratch
bool tc_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy);
int tc_wgrad_run(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy, float* gw,
                 float* G32, cudaStream_t st);

}  // namespace b200seg

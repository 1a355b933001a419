// extern "C" surface of libb200seg.so (see include/b200seg.h): argument validation, mapping of
// layer descriptors onto the kernel families, error reporting.  No device allocation, no sync.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_conv.h"

namespace b200seg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static std::atomic<long long> g_tc_launches{0};
static std::atomic<const char*> g_last_launch{""};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void note_launch(const char* what) { g_last_launch.store(what, std::memory_order_relaxed); }
void count_tc_launch() { g_tc_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

int check_conv_desc(const b200seg_conv_desc* d, bool transposed_layer) {
  B200SEG_CHECK_ARG(d != nullptr, "conv desc is NULL");
  B200SEG_CHECK_ARG(d->n > 0 && d->cin > 0 && d->cout > 0, "conv desc: n/cin/cout must be positive");
  B200SEG_CHECK_ARG(d->dtype == B200SEG_F32 || d->dtype == B200SEG_BF16, "conv desc: bad dtype %d", d->dtype);
  const int k[3] = {d->kd, d->kh, d->kw}, s[3] = {d->sd, d->sh, d->sw}, p[3] = {d->pd, d->ph, d->pw};
  const int in[3] = {d->in_d, d->in_h, d->in_w}, out[3] = {d->out_d, d->out_h, d->out_w};
  for (int i = 0; i < 3; ++i) {
    B200SEG_CHECK_ARG(k[i] == 1 || k[i] == 3, "conv desc: kernel extent must be 1 or 3, got %d", k[i]);
    B200SEG_CHECK_ARG(s[i] == 1 || s[i] == 2, "conv desc: stride must be 1 or 2, got %d", s[i]);
    B200SEG_CHECK_ARG(p[i] == (k[i] - 1) / 2, "conv desc: padding must be (k-1)/2");
    B200SEG_CHECK_ARG(in[i] > 0 && out[i] > 0, "conv desc: non-positive extent");
    if (!transposed_layer) {
      B200SEG_CHECK_ARG(out[i] == (in[i] + 2 * p[i] - k[i]) / s[i] + 1,
                        "conv desc: out extent %d inconsistent with in %d (k=%d s=%d p=%d)", out[i],
                        in[i], k[i], s[i], p[i]);
    } else {
      // output_padding = s - 1  (MONAI Convolution)
      B200SEG_CHECK_ARG(out[i] == (in[i] - 1) * s[i] - 2 * p[i] + k[i] + (s[i] - 1),
                        "convtr desc: out extent %d inconsistent with in %d (k=%d s=%d p=%d)", out[i],
                        in[i], k[i], s[i], p[i]);
    }
  }
  B200SEG_CHECK_ARG(d->x_ld >= d->cin && d->y_ld >= d->cout, "conv desc: ld smaller than channel count");
  return B200SEG_OK;
}

void fill_geom(GatherParams& g, const b200seg_conv_desc* d) {
  g.n = d->n;
  g.kd = d->kd; g.kh = d->kh; g.kw = d->kw;
  g.sd = d->sd; g.sh = d->sh; g.sw = d->sw;
  g.pd = d->pd; g.ph = d->ph; g.pw = d->pw;
  g.accumulate = (d->flags & B200SEG_CONV_ACCUMULATE) ? 1 : 0;
}

}  // namespace
}  // namespace b200seg

using namespace b200seg;

extern "C" {

int b200seg_version(void) { return B200SEG_VERSION; }
const char* b200seg_last_error(void) { return g_err; }
long long b200seg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
long long b200seg_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }
const char* b200seg_last_launch(void) { return g_last_launch.load(std::memory_order_relaxed); }

int b200seg_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    return B200SEG_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libb200seg is built for sm_100a only", device, prop.major, prop.minor);
    return B200SEG_ERR_DEVICE;
  }
  return B200SEG_OK;
}

// ---- weights ---------------------------------------------------------------------------------
static size_t generic_weight_bytes(const b200seg_conv_desc* d) {
  size_t esz = d->dtype == B200SEG_BF16 ? 2 : 4;
  return align_up((size_t)d->kd * d->kh * d->kw * d->cin * d->cout * esz, 256);
}
static const void* tc_weights(const b200seg_conv_desc* d, const void* w_packed) {
  return (const char*)w_packed + generic_weight_bytes(d);
}

// packed buffer = [ generic layout [tap][src][dst] | tcgen05 layout (bf16 only) ]
size_t b200seg_packed_weight_bytes(const b200seg_conv_desc* d, int kind) {
  if (!d) return 0;
  (void)kind;
  return generic_weight_bytes(d) + tc_packed_weight_bytes(d);
}

int b200seg_pack_weight(const b200seg_conv_desc* d, int kind, const float* w_torch, void* w_packed,
                        void* stream) {
  B200SEG_CHECK_ARG(d && w_torch && w_packed, "pack_weight: NULL argument");
  B200SEG_CHECK_ARG(kind >= 0 && kind <= 3, "pack_weight: bad kind %d", kind);
  if (d->dtype == B200SEG_BF16)  // both layouts in one launch
    return tc_pack_weight(d, kind, w_torch, (char*)w_packed + generic_weight_bytes(d), w_packed, as_stream(stream));
  return launch_pack_weight(d->dtype, kind, w_torch, w_packed, d->kd * d->kh * d->kw, d->cin, d->cout,
                            as_stream(stream));
}

extern "C" size_t b200seg_packed_weight_tc_offset(const b200seg_conv_desc* d, int kind) {
  (void)kind;
  return d ? generic_weight_bytes(d) : 0;
}

extern "C" int b200seg_pack_weights_batched(const b200seg_pack_entry* table, int32_t n_entries, void* stream) {
  B200SEG_CHECK_ARG(table && n_entries > 0 && n_entries < 65536, "pack_weights_batched: bad argument");
  return tc_pack_weights_batched(table, n_entries, as_stream(stream));
}

// tcgen05 dispatch: sliding-window kernel where it applies, else the streaming kernel
static int tc_dispatch(const b200seg_conv_desc* d, int op, const void* src, const void* w_packed,
                       const float* bias, const void* residual, void* dst, void* stream, float* stats = nullptr) {
  if (!(d->flags & B200SEG_CONV_NO_SLIDE) && tc_slide_conv_supported(d, op))
    return tc_slide_conv_run(d, op, src, tc_weights(d, w_packed), bias, residual, dst, stats, as_stream(stream));
  if (tc_convtr_slide_supported(d, op, residual))
    return tc_convtr_slide_run(d, op, src, tc_weights(d, w_packed), bias, dst, stats, as_stream(stream));
  return tc_conv_run(d, op, src, tc_weights(d, w_packed), bias, residual, dst, stats, as_stream(stream));
}

// ---- conv ---------------------------------------------------------------------------------------
int b200seg_conv_fprop(const b200seg_conv_desc* d, const void* x, const void* w_packed,
                       const float* bias, const void* residual, void* y, void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && w_packed && y, "conv_fprop: NULL pointer");
  if (tc_conv_supported(d, TC_CONV_FPROP, x, y, residual)) return tc_dispatch(d, TC_CONV_FPROP, x, w_packed, bias, residual, y, stream);
  if (small_cin_supported(d)) return launch_small_cin_fprop(d, x, w_packed, bias, residual, y, as_stream(stream));
  GatherParams g{};
  fill_geom(g, d);
  g.sD = d->in_d; g.sH = d->in_h; g.sW = d->in_w;
  g.dD = d->out_d; g.dH = d->out_h; g.dW = d->out_w;
  g.src_c = d->cin; g.dst_c = d->cout;
  g.src_ld = d->x_ld; g.dst_ld = d->y_ld; g.res_ld = d->r_ld;
  g.transposed = 0;
  return launch_gather(g, d->dtype, x, w_packed, bias, residual, y, as_stream(stream));
}

// fprop fused with the statistics of the InstanceNorm that follows: the conv epilogue writes per-CTA
// partial sums into the workspace, a small kernel turns them into mean / rstd.  Returns
// B200SEG_STATS_NOT_FUSED (1) when the convolution ran on a kernel without the fusion.
static void stats_layout(const b200seg_conv_desc* d, int op, int* ncls, int64_t* tiles) {
  if (!(d->flags & B200SEG_CONV_NO_SLIDE) && tc_slide_conv_supported(d, op)) {
    *ncls = 1;
    *tiles = tc_slide_conv_grid(d, op) / d->n;
  } else if (tc_convtr_slide_supported(d, op, nullptr)) {
    *ncls = 1;
    *tiles = tc_convtr_slide_grid(d, op) / d->n;
  } else {
    tc_conv_grid(d, op, ncls, tiles);
  }
}

static size_t fprop_stats_ws(const b200seg_conv_desc* d, bool transposed_layer) {
  int ncls;
  int64_t tiles;
  stats_layout(d, transposed_layer ? TC_CONVTR_FPROP : TC_CONV_FPROP, &ncls, &tiles);
  return (size_t)ncls * d->n * tiles * d->cout * 2 * sizeof(float) + 256;
}

// deferred != NULL: leave the partials in `ws` and report their layout (ncls, tiles) instead of finalising them
struct StatsLayout { int ncls; int64_t tiles; };
static int fprop_stats_common(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* w_packed,
                              const float* bias, void* y, float* mean, float* rstd, int stat_ld, float eps,
                              void* ws, size_t ws_bytes, void* stream, StatsLayout* deferred = nullptr) {
  int op = transposed_layer ? TC_CONVTR_FPROP : TC_CONV_FPROP;
  // Only the sliding-window kernel fuses the statistics: it accumulates them in registers across
  // the slabs of a column and emits one partial per CTA (a few hundred per sample).  In the
  // streaming kernel (one tile per CTA, up to 32 k CTAs) the per-tile warp reductions and the
  // reduction over tens of thousands of partials cost more than the separate statistics pass
  // (measured: ConvTranspose 32->10 fprop 258 -> 440 us).
  // The streaming kernel (one tile per CTA) fuses them only while the partials stay few (deep layers).
  bool fuse = false;
  if ((deferred || (mean && rstd)) && ws && tc_conv_supported(d, op, x, y, nullptr)) {
    fuse = (!(d->flags & B200SEG_CONV_NO_SLIDE) && tc_slide_conv_supported(d, op)) ||
           tc_convtr_slide_supported(d, op, nullptr);
    if (!fuse) {
      int ncls;
      int64_t tiles;
      tc_conv_grid(d, op, &ncls, &tiles);
      fuse = (int64_t)ncls * d->n * tiles <= 512;
    }
  }
  if (fuse) {
    if (ws_bytes < fprop_stats_ws(d, transposed_layer)) {
      set_error("conv_fprop_stats: workspace %zu < required %zu", ws_bytes, fprop_stats_ws(d, transposed_layer));
      return B200SEG_ERR_WORKSPACE;
    }
    int rc = tc_dispatch(d, op, x, w_packed, bias, nullptr, y, stream, (float*)ws);
    if (rc) return rc;
    int ncls;
    int64_t tiles;
    stats_layout(d, op, &ncls, &tiles);
    if (deferred) {
      deferred->ncls = ncls;
      deferred->tiles = tiles;
      return B200SEG_OK;
    }
    const int64_t spatial = (int64_t)d->out_d * d->out_h * d->out_w;
    return launch_instnorm_stats_from_partials((const float*)ws, d->n, d->cout, stat_ld > d->cout ? stat_ld : d->cout,
                                               ncls, tiles, spatial, eps, mean, rstd, as_stream(stream));
  }
  int rc = transposed_layer ? b200seg_convtr_fprop(d, x, w_packed, bias, nullptr, y, stream)
                            : b200seg_conv_fprop(d, x, w_packed, bias, nullptr, y, stream);
  return rc ? rc : B200SEG_STATS_NOT_FUSED;
}

size_t b200seg_conv_fprop_stats_workspace_bytes(const b200seg_conv_desc* d) { return d ? fprop_stats_ws(d, false) : 0; }
size_t b200seg_convtr_fprop_stats_workspace_bytes(const b200seg_conv_desc* d) { return d ? fprop_stats_ws(d, true) : 0; }

int b200seg_conv_fprop_stats(const b200seg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                             void* y, float* mean, float* rstd, int32_t stat_ld, float eps, void* workspace,
                             size_t workspace_bytes, void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && w_packed && y, "conv_fprop_stats: NULL pointer");
  return fprop_stats_common(d, false, x, w_packed, bias, y, mean, rstd, stat_ld, eps, workspace, workspace_bytes, stream);
}

int b200seg_convtr_fprop_stats(const b200seg_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                               void* y, float* mean, float* rstd, int32_t stat_ld, float eps, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = check_conv_desc(d, true);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && w_packed && y, "convtr_fprop_stats: NULL pointer");
  return fprop_stats_common(d, true, x, w_packed, bias, y, mean, rstd, stat_ld, eps, workspace, workspace_bytes, stream);
}

static int check_norm_desc(const b200seg_norm_desc* d);

int b200seg_conv_fprop_partials(const b200seg_conv_desc* d, int32_t transposed_layer, const void* x,
                                const void* w_packed, const float* bias, void* y, float* partials,
                                size_t partials_bytes, int32_t* ncls, int64_t* tiles, void* stream) {
  int rc = check_conv_desc(d, transposed_layer != 0);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && w_packed && y && partials && ncls && tiles, "conv_fprop_partials: NULL pointer");
  StatsLayout lay{0, 0};
  rc = fprop_stats_common(d, transposed_layer != 0, x, w_packed, bias, y, nullptr, nullptr, 0, 0.f, partials,
                          partials_bytes, stream, &lay);
  *ncls = lay.ncls;
  *tiles = lay.tiles;
  return rc;
}

int b200seg_instnorm_prelu_fwd_partials(const b200seg_norm_desc* d, const void* x, const float* partials,
                                        int32_t ncls, int64_t tiles, int32_t cstat, float* mean, float* rstd,
                                        const float* alpha, const void* residual, void* y, void* stream) {
  int rc = check_norm_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && partials && mean && rstd && alpha && y, "instnorm_prelu_fwd_partials: NULL pointer");
  B200SEG_CHECK_ARG(ncls > 0 && tiles > 0 && cstat > 0 && cstat <= d->c && d->c <= 256,
                    "instnorm_prelu_fwd_partials: bad partial layout (ncls %d, cstat %d, c %d)", ncls, cstat, d->c);
  B200SEG_CHECK_ARG(d->y_ld >= d->c && (!residual || d->r_ld >= d->c), "instnorm_prelu_fwd_partials: bad ld");
  return launch_instnorm_prelu_fwd_partials(*d, x, partials, ncls, tiles, cstat, alpha, residual, y, mean, rstd,
                                            as_stream(stream));
}

int b200seg_conv_dgrad(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                       const void* residual, void* dx, void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  B200SEG_CHECK_ARG(dy && w_packed && dx, "conv_dgrad: NULL pointer");
  if (tc_conv_supported(d, TC_CONV_DGRAD, dy, dx, residual)) return tc_dispatch(d, TC_CONV_DGRAD, dy, w_packed, nullptr, residual, dx, stream);
  GatherParams g{};
  fill_geom(g, d);
  g.sD = d->out_d; g.sH = d->out_h; g.sW = d->out_w;
  g.dD = d->in_d; g.dH = d->in_h; g.dW = d->in_w;
  g.src_c = d->cout; g.dst_c = d->cin;
  g.src_ld = d->y_ld; g.dst_ld = d->x_ld; g.res_ld = d->r_ld;
  g.transposed = 1;
  return launch_gather(g, d->dtype, dy, w_packed, nullptr, residual, dx, as_stream(stream));
}

int b200seg_im2col(const b200seg_conv_desc* d, const void* x, void* col, int32_t col_ld, void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && col, "im2col: NULL pointer");
  const int J = d->kd * d->kh * d->kw * d->cin;
  const int ve = d->dtype == B200SEG_BF16 ? 8 : 4;
  B200SEG_CHECK_ARG(d->cin <= 4 && J <= 32, "im2col: only small-Cin layers (taps*cin <= 32), got %d", J);
  B200SEG_CHECK_ARG(col_ld >= J && col_ld % ve == 0 && ((uintptr_t)col % 16) == 0,
                    "im2col: col_ld must cover taps*cin and keep 16-byte rows");
  return launch_im2col(d, x, col, col_ld, as_stream(stream));
}

static void conv_wgrad_params(const b200seg_conv_desc* d, WgradParams& w) {
  w.n = d->n;
  w.sD = d->in_d; w.sH = d->in_h; w.sW = d->in_w;
  w.tD = d->out_d; w.tH = d->out_h; w.tW = d->out_w;
  w.a_c = d->cin; w.b_c = d->cout;
  w.kd = d->kd; w.kh = d->kh; w.kw = d->kw;
  w.sd = d->sd; w.sh = d->sh; w.sw = d->sw;
  w.pd = d->pd; w.ph = d->ph; w.pw = d->pw;
  w.s_ld = d->x_ld; w.t_ld = d->y_ld;
  wgrad_plan(w);
}

static void convtr_wgrad_params(const b200seg_conv_desc* d, WgradParams& w) {
  // G[k][co][ci] = sum_i dy[i*s - p + k, co] * x[i, ci]  : S = dy (gathered), T = x
  w.n = d->n;
  w.sD = d->out_d; w.sH = d->out_h; w.sW = d->out_w;
  w.tD = d->in_d; w.tH = d->in_h; w.tW = d->in_w;
  w.a_c = d->cout; w.b_c = d->cin;
  w.kd = d->kd; w.kh = d->kh; w.kw = d->kw;
  w.sd = d->sd; w.sh = d->sh; w.sw = d->sw;
  w.pd = d->pd; w.ph = d->ph; w.pw = d->pw;
  w.s_ld = d->y_ld; w.t_ld = d->x_ld;
  wgrad_plan(w);
}

static size_t wgrad_main_bytes(const b200seg_conv_desc* d, const WgradParams& w) {
  size_t a = wgrad_partial_bytes(w), b = small_cin_wgrad_workspace(d);
  return align_up(a > b ? a : b, 256);
}
static size_t wgrad_colsum_bytes(const b200seg_conv_desc* d) {
  int64_t nvox_y = (int64_t)d->n * d->out_d * d->out_h * d->out_w;
  return align_up(colsum_workspace_bytes(nvox_y, d->cout), 256);
}
static size_t wgrad_ws_bytes(const b200seg_conv_desc* d, const WgradParams& w) {
  return wgrad_main_bytes(d, w) + wgrad_colsum_bytes(d) + tc_wgrad_extra_workspace(d);
}

size_t b200seg_conv_wgrad_workspace_bytes(const b200seg_conv_desc* d) {
  if (!d) return 0;
  WgradParams w{};
  conv_wgrad_params(d, w);
  return wgrad_ws_bytes(d, w);
}

size_t b200seg_convtr_wgrad_workspace_bytes(const b200seg_conv_desc* d) {
  if (!d) return 0;
  WgradParams w{};
  convtr_wgrad_params(d, w);
  return wgrad_ws_bytes(d, w);
}

static int wgrad_common(const b200seg_conv_desc* d, bool transposed_layer, const WgradParams& w, const void* S,
                        const void* T, const void* x, const void* dy, float* gw, float* gbias, void* ws,
                        size_t ws_bytes, void* stream) {
  B200SEG_CHECK_ARG(S && T && gw && ws, "wgrad: NULL pointer");
  size_t need = wgrad_ws_bytes(d, w);
  if (ws_bytes < need) {
    set_error("wgrad: workspace %zu < required %zu bytes", ws_bytes, need);
    return B200SEG_ERR_WORKSPACE;
  }
  float* partial = (float*)ws;
  int rc;
  if (tc_wgrad_supported(d, transposed_layer, x, dy)) {
    float* g32 = (float*)((char*)ws + wgrad_main_bytes(d, w) + wgrad_colsum_bytes(d));
    if (tc_slide_wgrad_supported(d, transposed_layer))
      rc = tc_slide_wgrad_run(d, x, dy, gw, g32, as_stream(stream));
    else if (tc_convtr_wgrad_supported(d, transposed_layer))
      rc = tc_convtr_wgrad_run(d, transposed_layer, x, dy, gw, g32, as_stream(stream));
    else
      rc = tc_wgrad_run(d, transposed_layer, x, dy, gw, g32, as_stream(stream));
  } else if (!transposed_layer && small_cin_supported(d)) {
    rc = launch_small_cin_wgrad(d, x, dy, gw, partial, as_stream(stream));
  } else {
    rc = launch_wgrad(w, d->dtype, S, T, gw, partial, as_stream(stream));
  }
  if (rc) return rc;
  if (gbias) {
    float* cs = (float*)((char*)ws + wgrad_main_bytes(d, w));
    int64_t nvox_y = (int64_t)d->n * d->out_d * d->out_h * d->out_w;
    rc = launch_colsum(d->dtype, dy, nvox_y, d->cout, d->y_ld, gbias, cs, as_stream(stream));
  }
  return rc;
}

int b200seg_conv_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw,
                       float* gbias, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  WgradParams w{};
  conv_wgrad_params(d, w);
  return wgrad_common(d, false, w, x, dy, x, dy, gw, gbias, workspace, workspace_bytes, stream);
}

int b200seg_convtr_fprop(const b200seg_conv_desc* d, const void* x, const void* w_packed,
                         const float* bias, const void* residual, void* y, void* stream) {
  int rc = check_conv_desc(d, true);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && w_packed && y, "convtr_fprop: NULL pointer");
  if (tc_conv_supported(d, TC_CONVTR_FPROP, x, y, residual)) return tc_dispatch(d, TC_CONVTR_FPROP, x, w_packed, bias, residual, y, stream);
  GatherParams g{};
  fill_geom(g, d);
  g.sD = d->in_d; g.sH = d->in_h; g.sW = d->in_w;
  g.dD = d->out_d; g.dH = d->out_h; g.dW = d->out_w;
  g.src_c = d->cin; g.dst_c = d->cout;
  g.src_ld = d->x_ld; g.dst_ld = d->y_ld; g.res_ld = d->r_ld;
  g.transposed = 1;
  return launch_gather(g, d->dtype, x, w_packed, bias, residual, y, as_stream(stream));
}

int b200seg_convtr_dgrad(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                         const void* residual, void* dx, void* stream) {
  int rc = check_conv_desc(d, true);
  if (rc) return rc;
  B200SEG_CHECK_ARG(dy && w_packed && dx, "convtr_dgrad: NULL pointer");
  if (tc_conv_supported(d, TC_CONVTR_DGRAD, dy, dx, residual)) return tc_dispatch(d, TC_CONVTR_DGRAD, dy, w_packed, nullptr, residual, dx, stream);
  GatherParams g{};
  fill_geom(g, d);
  g.sD = d->out_d; g.sH = d->out_h; g.sW = d->out_w;
  g.dD = d->in_d; g.dH = d->in_h; g.dW = d->in_w;
  g.src_c = d->cout; g.dst_c = d->cin;
  g.src_ld = d->y_ld; g.dst_ld = d->x_ld; g.res_ld = d->r_ld;
  g.transposed = 0;
  return launch_gather(g, d->dtype, dy, w_packed, nullptr, residual, dx, as_stream(stream));
}

int b200seg_convtr_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw,
                         float* gbias, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_conv_desc(d, true);
  if (rc) return rc;
  WgradParams w{};
  convtr_wgrad_params(d, w);
  return wgrad_common(d, true, w, dy, x, x, dy, gw, gbias, workspace, workspace_bytes, stream);
}

// ---- InstanceNorm + PReLU ---------------------------------------------------------------------------
static int check_norm_desc(const b200seg_norm_desc* d) {
  B200SEG_CHECK_ARG(d != nullptr, "norm desc is NULL");
  B200SEG_CHECK_ARG(d->n > 0 && d->c > 0 && d->spatial > 0, "norm desc: n/c/spatial must be positive");
  B200SEG_CHECK_ARG(d->dtype == B200SEG_F32 || d->dtype == B200SEG_BF16, "norm desc: bad dtype");
  B200SEG_CHECK_ARG(d->x_ld >= d->c, "norm desc: x_ld smaller than channel count");
  return B200SEG_OK;
}

size_t b200seg_instnorm_workspace_bytes(const b200seg_norm_desc* d) {
  return d ? norm_workspace_bytes(*d) : 0;
}

int b200seg_instnorm_stats(const b200seg_norm_desc* d, const void* x, float* mean, float* rstd,
                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_norm_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && mean && rstd && workspace, "instnorm_stats: NULL pointer");
  if (workspace_bytes < norm_workspace_bytes(*d)) {
    set_error("instnorm_stats: workspace %zu < required %zu", workspace_bytes, norm_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_instnorm_stats(*d, x, mean, rstd, workspace, as_stream(stream));
}

int b200seg_instnorm_prelu_fwd(const b200seg_norm_desc* d, const void* x, const float* mean,
                               const float* rstd, const float* alpha, const void* residual, void* y,
                               void* stream) {
  int rc = check_norm_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && mean && rstd && alpha && y, "instnorm_prelu_fwd: NULL pointer");
  B200SEG_CHECK_ARG(d->y_ld >= d->c && (!residual || d->r_ld >= d->c), "instnorm_prelu_fwd: bad ld");
  return launch_instnorm_prelu_fwd(*d, x, mean, rstd, alpha, residual, y, as_stream(stream));
}

int b200seg_instnorm_prelu_bwd(const b200seg_norm_desc* d, const void* x, const float* mean,
                               const float* rstd, const float* alpha, const void* dy, void* dx,
                               float* dalpha, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_norm_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && mean && rstd && alpha && dy && dx && dalpha && workspace,
                    "instnorm_prelu_bwd: NULL pointer");
  B200SEG_CHECK_ARG(d->y_ld >= d->c && d->r_ld >= d->c, "instnorm_prelu_bwd: y_ld (dy) / r_ld (dx) too small");
  if (workspace_bytes < norm_workspace_bytes(*d)) {
    set_error("instnorm_prelu_bwd: workspace %zu < required %zu", workspace_bytes, norm_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_instnorm_prelu_bwd(*d, x, mean, rstd, alpha, dy, dx, dalpha, workspace, as_stream(stream));
}

// ---- dgrad fused with the reduction pass of the InstanceNorm + PReLU backward it feeds -------------------------
size_t b200seg_conv_dgrad_instnorm_partials_bytes(const b200seg_conv_desc* d) {
  if (!d || !tc_slide_conv_bwdstats_supported(d, TC_CONV_DGRAD)) return 0;
  return (size_t)tc_slide_conv_grid(d, TC_CONV_DGRAD, true) * 16 * 3 * sizeof(float);
}

int b200seg_conv_dgrad_instnorm_partials(const b200seg_conv_desc* d, const void* dy, const void* w_packed,
                                         const void* residual, void* dx, const void* norm_x, int32_t norm_x_ld,
                                         const float* mean, const float* rstd, int32_t stat_ld, const float* alpha,
                                         float* partials, size_t partials_bytes, int64_t* rows_per_sample,
                                         void* stream) {
  int rc = check_conv_desc(d, false);
  if (rc) return rc;
  B200SEG_CHECK_ARG(dy && w_packed && dx && norm_x && mean && rstd && alpha && partials && rows_per_sample,
                    "conv_dgrad_instnorm_partials: NULL pointer");
  if (!tc_conv_supported(d, TC_CONV_DGRAD, dy, dx, residual) || !tc_slide_conv_bwdstats_supported(d, TC_CONV_DGRAD) ||
      norm_x_ld < 16 || (norm_x_ld % 8) || ((uintptr_t)norm_x % 16) || stat_ld < 16)
    return B200SEG_STATS_NOT_FUSED;  // nothing was launched: the caller runs the two plain entry points
  const size_t need = b200seg_conv_dgrad_instnorm_partials_bytes(d);
  if (partials_bytes < need) {
    set_error("conv_dgrad_instnorm_partials: partials %zu < required %zu", partials_bytes, need);
    return B200SEG_ERR_WORKSPACE;
  }
  TcBwdStats bst{norm_x, norm_x_ld, stat_ld, mean, rstd, alpha, partials};
  *rows_per_sample = tc_slide_conv_grid(d, TC_CONV_DGRAD, true) / d->n;
  return tc_slide_conv_run(d, TC_CONV_DGRAD, dy, tc_weights(d, w_packed), nullptr, residual, dx, nullptr,
                           as_stream(stream), &bst);
}

int b200seg_instnorm_prelu_bwd_from_partials(const b200seg_norm_desc* d, const void* x, const float* mean,
                                             const float* rstd, const float* alpha, const void* dy,
                                             const float* partials, int64_t rows_per_sample, void* dx, float* dalpha,
                                             void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_norm_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(x && mean && rstd && alpha && dy && partials && dx && dalpha && workspace,
                    "instnorm_prelu_bwd_from_partials: NULL pointer");
  B200SEG_CHECK_ARG(d->c == 16 && rows_per_sample > 0, "instnorm_prelu_bwd_from_partials: partial rows hold 16 channels");
  B200SEG_CHECK_ARG(d->y_ld >= d->c && d->r_ld >= d->c, "instnorm_prelu_bwd_from_partials: y_ld (dy) / r_ld (dx) too small");
  if (workspace_bytes < norm_workspace_bytes(*d)) {
    set_error("instnorm_prelu_bwd_from_partials: workspace %zu < required %zu", workspace_bytes, norm_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_instnorm_prelu_bwd_from_partials(*d, x, mean, rstd, alpha, dy, partials, rows_per_sample, dx, dalpha,
                                                 workspace, as_stream(stream));
}

// ---- optimiser --------------------------------------------------------------------------------------
int b200seg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, int64_t step, void* stream) {
  B200SEG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad argument");
  B200SEG_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0,
                    "adam_step: buffers must be 16-byte aligned");
  return launch_adam_flat(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, as_stream(stream));
}

// ---- softmax + Dice -------------------------------------------------------------------------------------
static int check_dice_desc(const b200seg_dice_desc* d) {
  B200SEG_CHECK_ARG(d != nullptr, "dice desc is NULL");
  B200SEG_CHECK_ARG(d->n > 0 && d->c > 0 && d->spatial > 0, "dice desc: n/c/spatial must be positive");
  B200SEG_CHECK_ARG(d->ld >= d->c, "dice desc: ld smaller than class count");
  B200SEG_CHECK_ARG(d->dtype == B200SEG_F32 || d->dtype == B200SEG_BF16, "dice desc: bad dtype");
  B200SEG_CHECK_ARG(d->label_dtype == B200SEG_LABEL_U8 || d->label_dtype == B200SEG_LABEL_I64,
                    "dice desc: bad label dtype");
  return B200SEG_OK;
}

size_t b200seg_softmax_dice_workspace_bytes(const b200seg_dice_desc* d) {
  return d ? dice_workspace_bytes(*d) : 0;
}

int b200seg_softmax_dice_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                             float* sums, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && sums && workspace, "softmax_dice_fwd: NULL pointer");
  if (workspace_bytes < dice_workspace_bytes(*d)) {
    set_error("softmax_dice_fwd: workspace %zu < required %zu", workspace_bytes, dice_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_softmax_dice_fwd(*d, logits, labels, sums, workspace, as_stream(stream));
}

size_t b200seg_softmax_dice_metric_workspace_bytes(const b200seg_dice_desc* d) {
  return d ? dice_metric_workspace_bytes(*d) : 0;
}

int b200seg_softmax_dice_metric_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float* sums,
                                    int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && sums && counts && workspace, "softmax_dice_metric_fwd: NULL pointer");
  if (workspace_bytes < dice_metric_workspace_bytes(*d)) {
    set_error("softmax_dice_metric_fwd: workspace %zu < required %zu", workspace_bytes,
              dice_metric_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_softmax_dice_metric_fwd(*d, logits, labels, sums, counts, workspace, as_stream(stream));
}

int b200seg_softmax_dice_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                             const float* gI, const float* gP, void* dlogits, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && gI && gP && dlogits, "softmax_dice_bwd: NULL pointer");
  return launch_softmax_dice_bwd(*d, logits, labels, gI, gP, dlogits, as_stream(stream));
}

size_t b200seg_softmax_loss_workspace_bytes(const b200seg_dice_desc* d) { return d ? loss_workspace_bytes(*d) : 0; }

int b200seg_softmax_loss_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float gamma,
                             float* sums5, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && sums5 && workspace && gamma >= 0.f, "softmax_loss_fwd: bad argument");
  if (workspace_bytes < loss_workspace_bytes(*d)) {
    set_error("softmax_loss_fwd: workspace %zu < required %zu", workspace_bytes, loss_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_softmax_loss_fwd(*d, logits, labels, gamma, sums5, workspace, as_stream(stream));
}

int b200seg_softmax_loss_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels, float gamma,
                             const float* gI, const float* gP, const float* gF, const float* gN, void* dlogits,
                             void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && gI && gP && gF && gN && dlogits, "softmax_loss_bwd: NULL pointer");
  return launch_softmax_loss_bwd(*d, logits, labels, gamma, gI, gP, gF, gN, dlogits, as_stream(stream));
}

size_t b200seg_softmax_boundary_loss_workspace_bytes(const b200seg_dice_desc* d) {
  return d ? boundary_workspace_bytes(*d) : 0;
}

int b200seg_softmax_boundary_loss_fwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                                      const float* dist_maps, float gamma, float* sums6, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && dist_maps && sums6 && workspace && gamma >= 0.f,
                    "softmax_boundary_loss_fwd: bad argument");
  if (workspace_bytes < boundary_workspace_bytes(*d)) {
    set_error("softmax_boundary_loss_fwd: workspace %zu < required %zu", workspace_bytes,
              boundary_workspace_bytes(*d));
    return B200SEG_ERR_WORKSPACE;
  }
  return launch_softmax_boundary_loss_fwd(*d, logits, labels, dist_maps, gamma, sums6, workspace, as_stream(stream));
}

int b200seg_softmax_boundary_loss_bwd(const b200seg_dice_desc* d, const void* logits, const void* labels,
                                      const float* dist_maps, float gamma, const float* gI, const float* gP,
                                      const float* gF, const float* gN, const float* gB, void* dlogits,
                                      void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits && labels && dist_maps && gI && gP && gF && gN && gB && dlogits,
                    "softmax_boundary_loss_bwd: NULL pointer");
  return launch_softmax_boundary_loss_bwd(*d, logits, labels, dist_maps, gamma, gI, gP, gF, gN, gB, dlogits,
                                          as_stream(stream));
}

int b200seg_dice_loss_epilogue(const float* sums, int32_t n, int32_t c, int32_t include_background, float smooth,
                               int32_t mean, float* loss, float* gI, float* gP, void* stream) {
  B200SEG_CHECK_ARG(sums && loss && gI && gP && n > 0 && c > 0, "dice_loss_epilogue: bad argument");
  const int c0 = include_background ? 0 : 1;
  B200SEG_CHECK_ARG(c > c0, "dice_loss_epilogue: no foreground class");
  const float inv_count = mean ? 1.f / (float)(n * (c - c0)) : 1.f;
  return launch_dice_loss_epilogue(sums, n, c, c0, smooth, inv_count, loss, gI, gP, as_stream(stream));
}

int b200seg_argmax_dice_counts(const b200seg_dice_desc* d, const void* logits, const void* target,
                               uint8_t* pred_out, int64_t* counts, void* stream) {
  int rc = check_dice_desc(d);
  if (rc) return rc;
  B200SEG_CHECK_ARG(logits, "argmax_dice_counts: NULL logits");
  B200SEG_CHECK_ARG(!target || counts, "argmax_dice_counts: target given but counts is NULL");
  B200SEG_CHECK_ARG(target || pred_out, "argmax_dice_counts: nothing to produce");
  return launch_argmax_dice_counts(*d, logits, target, pred_out, counts, as_stream(stream));
}

int b200seg_label_dice_counts(int32_t n, int64_t spatial, int32_t c, const uint8_t* pred,
                              const void* target, int32_t target_dtype, int64_t* counts, void* stream) {
  B200SEG_CHECK_ARG(n > 0 && spatial > 0 && c > 0 && pred && target && counts, "label_dice_counts: bad argument");
  return launch_label_dice_counts(n, spatial, c, pred, target, target_dtype, counts, as_stream(stream));
}

int b200seg_squash_masks(int32_t n, int32_t n_struct, int64_t spatial, const uint8_t* masks,
                         uint8_t* labels, void* stream) {
  B200SEG_CHECK_ARG(n > 0 && n_struct > 0 && n_struct < 255 && spatial > 0 && masks && labels,
                    "squash_masks: bad argument");
  return launch_squash_masks(n, n_struct, spatial, masks, labels, as_stream(stream));
}

int b200seg_hu_window_norm(int64_t n_vox, int32_t n_windows, const int16_t* hu, const float* lo,
                           const float* hi, const float* mean, const float* std_, void* out,
                           int32_t out_ld, int32_t dtype, void* stream) {
  B200SEG_CHECK_ARG(n_vox > 0 && n_windows >= 1 && n_windows <= 4 && hu && lo && hi && mean && std_ && out,
                    "hu_window_norm: bad argument");
  B200SEG_CHECK_ARG(out_ld >= n_windows, "hu_window_norm: out_ld < n_windows");
  return launch_hu_window_norm(n_vox, n_windows, hu, lo, hi, mean, std_, out, out_ld, dtype, as_stream(stream));
}

int b200seg_window_accumulate(int32_t dtype, const void* window_logits, int32_t src_ld, float* acc, float* cnt,
                              int32_t c, int32_t wd, int32_t wh, int32_t ww, int32_t D, int32_t H, int32_t W,
                              int32_t d0, int32_t h0, int32_t w0, void* stream) {
  B200SEG_CHECK_ARG(window_logits && acc && cnt && c > 0 && src_ld >= c && wd > 0 && wh > 0 && ww > 0,
                    "window_accumulate: bad argument");
  B200SEG_CHECK_ARG(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 < D && h0 < H && w0 < W, "window_accumulate: origin outside");
  return launch_window_accumulate(dtype, window_logits, src_ld, nullptr, acc, cnt, c, wd, wh, ww, D, H, W, d0, h0, w0,
                                  as_stream(stream));
}

int b200seg_window_accumulate_weighted(int32_t dtype, const void* window_logits, int32_t src_ld,
                                       const float* importance, float* acc, float* cnt, int32_t c, int32_t wd,
                                       int32_t wh, int32_t ww, int32_t D, int32_t H, int32_t W, int32_t d0, int32_t h0,
                                       int32_t w0, void* stream) {
  B200SEG_CHECK_ARG(window_logits && acc && cnt && c > 0 && src_ld >= c && wd > 0 && wh > 0 && ww > 0,
                    "window_accumulate_weighted: bad argument");
  B200SEG_CHECK_ARG(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 < D && h0 < H && w0 < W,
                    "window_accumulate_weighted: origin outside");
  return launch_window_accumulate(dtype, window_logits, src_ld, importance, acc, cnt, c, wd, wh, ww, D, H, W, d0, h0,
                                  w0, as_stream(stream));
}

int b200seg_accum_argmax(const float* acc, const float* cnt, uint8_t* labels, float* mean_logits, int64_t n_vox,
                         int32_t c, void* stream) {
  B200SEG_CHECK_ARG(acc && cnt && labels && n_vox > 0 && c > 0 && c <= 32, "accum_argmax: bad argument");
  return launch_accum_argmax(acc, cnt, labels, mean_logits, n_vox, c, as_stream(stream));
}

int b200seg_crop_window_norm(int32_t dtype, const int16_t* hu, const uint8_t* labels, const int32_t* origins,
                             int32_t n_patches, void* img_out, uint8_t* lab_out, int32_t D, int32_t H, int32_t W,
                             int32_t pd, int32_t ph, int32_t pw, float lo, float hi, float mean, float std_,
                             int32_t pad_hu, void* stream) {
  B200SEG_CHECK_ARG(hu && origins && img_out && n_patches > 0 && pd > 0 && ph > 0 && pw > 0, "crop_window_norm: bad argument");
  return launch_crop_window_norm(dtype, hu, labels, origins, n_patches, img_out, lab_out, D, H, W, pd, ph, pw, lo, hi,
                                 mean, std_, pad_hu, as_stream(stream));
}

}  // extern "C"

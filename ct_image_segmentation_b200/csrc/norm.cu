// InstanceNorm (affine=False, biased variance, eps inside the sqrt) fused with PReLU (one shared
// alpha) and the residual sum of the enclosing ResidualUnit.  Channels-last, bandwidth-bound:
// vectorised coalesced access (up to 16 B per thread), fp32 statistics, deterministic two-stage
// reductions (per-block partials, then a fixed-order finalisation in double).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace b200seg {

namespace {

struct NormGeom {
  int V;      // elements per thread access
  int L;      // threads per voxel  (= C / V)
  int VB;     // voxels per block iteration (= 256 / L)
  int nblk;   // blocks per sample for the reduction kernels
  int nblk_apply;
};

inline bool aligned(const void* p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; }

// largest V in {8,4,2,1} with V*esz <= 16 that divides C and every ld and keeps 'ptrs' aligned
int pick_vec(const b200seg_norm_desc& d, std::initializer_list<const void*> ptrs,
             std::initializer_list<int> lds) {
  size_t esz = d.dtype == B200SEG_BF16 ? 2 : 4;
  for (int V = (int)(16 / esz); V > 1; V >>= 1) {
    bool ok = (d.c % V) == 0;
    for (int ld : lds) ok = ok && (ld % V) == 0;
    for (const void* p : ptrs) ok = ok && aligned(p, V * esz);
    if (ok) return V;
  }
  return 1;
}

}  // namespace

int norm_blocks(const b200seg_norm_desc& d) {
  // >= 16384 elements per block (256 threads x 16 B x a few iterations): deep layers with few voxels
  // but many channels still spread over tens of blocks instead of looping in one
  int64_t nb = cdiv64(d.spatial * (int64_t)d.c, 16384);
  int64_t cap = 1184 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

size_t norm_workspace_bytes(const b200seg_norm_desc& d) {
  size_t nb = norm_blocks(d);
  return ((size_t)d.n * nb * d.c * 3 + (size_t)d.n * d.c * 3) * sizeof(float) + 256;
}

// ------------------------------------------------------------------------------------------
template <typename T, int V, int NACC, typename F>
__device__ __forceinline__ void block_reduce_store(float (&acc)[NACC][V], int L, int VB, int vi,
                                                   int l, float* red, float* out, int c, F) {
  // red: [256][NACC*V] floats in shared memory
  const int t = threadIdx.x;
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < V; ++i) red[t * (NACC * V) + a * V + i] = acc[a][i];
  __syncthreads();
  int s = 1;
  while (s < VB) s <<= 1;
  for (s >>= 1; s >= 1; s >>= 1) {
    if (vi < s && vi + s < VB) {
#pragma unroll
      for (int j = 0; j < NACC * V; ++j) red[t * (NACC * V) + j] += red[(t + s * L) * (NACC * V) + j];
    }
    __syncthreads();
  }
  if (vi == 0) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int i = 0; i < V; ++i) out[(l * V + i) * NACC + a] = red[t * (NACC * V) + a * V + i];
  }
}

// partial[n][blk][c][2] = { sum x, sum x^2 }
template <typename T, int V>
__global__ void __launch_bounds__(256)
instnorm_stats_partial_kernel(const T* __restrict__ x, int64_t spatial, int c, int ld, int L, int VB,
                              int64_t vox_per_block, float* __restrict__ partial) {
  extern __shared__ float red[];
  const int t = threadIdx.x, vi = t / L, l = t % L;
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float acc[2][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (vi < VB) {
    const T* base = x + ((int64_t)n * spatial) * ld + l * V;
    for (int64_t v = v_begin + vi; v < v_end; v += VB) {
      Vec<T, V> xv;
      xv.load(base + v * ld);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        acc[0][i] += xv.v[i];
        acc[1][i] = fmaf(xv.v[i], xv.v[i], acc[1][i]);
      }
    }
  }
  float* out = partial + ((int64_t)n * gridDim.x + blockIdx.x) * c * 2;
  block_reduce_store<T, V, 2>(acc, L, VB, vi, l, red, out, c, 0);
}

// one warp per (n, c): mean, rstd from the block partials (double accumulation, fixed order)
__global__ void instnorm_stats_final_kernel(const float* __restrict__ partial, int nblk, int c,
                                            int nc_total, int64_t spatial, float eps,
                                            float* __restrict__ mean, float* __restrict__ rstd) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= nc_total) return;
  int n = warp / c, ch = warp % c;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    const float* p = partial + (((int64_t)n * nblk + b) * c + ch) * 2;
    s1 += (double)p[0];
    s2 += (double)p[1];
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane == 0) {
    double m = s1 / (double)spatial;
    double var = s2 / (double)spatial - m * m;
    if (var < 0.0) var = 0.0;
    mean[warp] = (float)m;
    rstd[warp] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
instnorm_prelu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ alpha,
                          const T* __restrict__ res, T* __restrict__ y, int64_t spatial, int c,
                          int x_ld, int y_ld, int r_ld, int L, int VB, int64_t vox_per_block) {
  const int t = threadIdx.x, vi = t / L, l = t % L;
  if (vi >= VB) return;
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float m[V], r[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m[i] = mean[n * c + l * V + i];
    r[i] = rstd[n * c + l * V + i];
  }
  const float a = alpha[0];
  const int64_t vox0 = (int64_t)n * spatial;
  auto body = [&](const Vec<T, V>& xv, const Vec<T, V>& rv, int64_t v) {
    Vec<T, V> ov;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float h = (xv.v[i] - m[i]) * r[i];
      ov.v[i] = h > 0.f ? h : a * h;
    }
    if (res) {
#pragma unroll
      for (int i = 0; i < V; ++i) ov.v[i] += rv.v[i];
    }
    ov.store(y + (vox0 + v) * y_ld + l * V);
  };
  int64_t v = v_begin + vi;
  for (; v + VB < v_end; v += 2 * VB) {  // two voxels per iteration, loads first (bytes in flight per thread x 2)
    Vec<T, V> xv0, xv1, rv0, rv1;
    xv0.load(x + (vox0 + v) * x_ld + l * V);
    xv1.load(x + (vox0 + v + VB) * x_ld + l * V);
    if (res) {
      rv0.load(res + (vox0 + v) * r_ld + l * V);
      rv1.load(res + (vox0 + v + VB) * r_ld + l * V);
    }
    body(xv0, rv0, v);
    body(xv1, rv1, v + VB);
  }
  if (v < v_end) {
    Vec<T, V> xv, rv;
    xv.load(x + (vox0 + v) * x_ld + l * V);
    if (res) rv.load(res + (vox0 + v) * r_ld + l * V);
    body(xv, rv, v);
  }
}

// InstanceNorm + PReLU forward straight from the PARTIAL statistics a convolution epilogue produced
// (row (o * N + n) * tiles + t holds cstat x {sum, sum of squares}): every block first reduces the
// ncls * tiles rows of its sample (256 threads = (row group, entry), double, fixed order) into
// mean / rstd in shared memory, then applies them.  Saves the separate few-microsecond finalisation
// launch that sat on the forward chain after every convolution; block 0 of each sample also writes
// mean / rstd out for the backward pass.  Channels >= cstat are zero padding (mean 0, rstd 1/sqrt(eps)).
template <typename T, int V>
__global__ void __launch_bounds__(256)
instnorm_prelu_fwd_partials_kernel(const T* __restrict__ x, const float* __restrict__ partial, int ncls, int n_total,
                                   int64_t tiles, int cstat, float eps, const float* __restrict__ alpha,
                                   const T* __restrict__ res, T* __restrict__ y, int64_t spatial, int c, int x_ld,
                                   int y_ld, int r_ld, int L, int VB, int64_t vox_per_block,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ double dsum[512];
  __shared__ float smean[256], srstd[256];
  const int t = threadIdx.x, n = blockIdx.y;
  const int E = 2 * cstat;
  const int G = E <= 256 ? 256 / E : 1;
  const int ntile = (int)tiles;
  for (int sl = t; sl < G * E; sl += 256) {
    const int g = sl / E, e = sl % E;
    double acc = 0.0;
    for (int o = 0; o < ncls; ++o) {
      const float* base = partial + ((int64_t)(o * n_total + n) * ntile) * E + e;
      // row group g takes tiles g, g + G, ...: the same rows in the same order whatever the unrolling
#pragma unroll 8
      for (int tt = g; tt < ntile; tt += G) acc += (double)__ldg(base + (int64_t)tt * E);
    }
    dsum[sl] = acc;
  }
  __syncthreads();
  for (int ch = t; ch < c; ch += 256) {
    float m = 0.f, rs = rsqrtf(eps);
    if (ch < cstat) {
      double s1 = 0.0, s2 = 0.0;
      for (int g = 0; g < G; ++g) {
        s1 += dsum[g * E + 2 * ch];
        s2 += dsum[g * E + 2 * ch + 1];
      }
      const double mm = s1 / (double)spatial;
      double var = s2 / (double)spatial - mm * mm;
      if (var < 0.0) var = 0.0;
      m = (float)mm;
      rs = (float)(1.0 / sqrt(var + (double)eps));
    } else {
      rs = (float)(1.0 / sqrt((double)eps));
    }
    smean[ch] = m;
    srstd[ch] = rs;
    if (blockIdx.x == 0) {
      mean_out[n * c + ch] = m;
      rstd_out[n * c + ch] = rs;
    }
  }
  __syncthreads();
  const int vi = t / L, l = t % L;
  if (vi >= VB) return;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float m[V], r[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m[i] = smean[l * V + i];
    r[i] = srstd[l * V + i];
  }
  const float a = alpha[0];
  const int64_t vox0 = (int64_t)n * spatial;
  for (int64_t v = v_begin + vi; v < v_end; v += VB) {
    Vec<T, V> xv, ov;
    xv.load(x + (vox0 + v) * x_ld + l * V);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float h = (xv.v[i] - m[i]) * r[i];
      ov.v[i] = h > 0.f ? h : a * h;
    }
    if (res) {
      Vec<T, V> rv;
      rv.load(res + (vox0 + v) * r_ld + l * V);
#pragma unroll
      for (int i = 0; i < V; ++i) ov.v[i] += rv.v[i];
    }
    ov.store(y + (vox0 + v) * y_ld + l * V);
  }
}

// partial[n][blk][c][3] = { sum g~, sum g~*xhat, sum dy*xhat*[xhat<=0] },  g~ = dy * prelu'(xhat)
template <typename T, int V>
__global__ void __launch_bounds__(256)
instnorm_prelu_bwd_partial_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ alpha, int64_t spatial, int c, int x_ld,
                                  int dy_ld, int L, int VB, int64_t vox_per_block,
                                  float* __restrict__ partial) {
  extern __shared__ float red[];
  const int t = threadIdx.x, vi = t / L, l = t % L;
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float acc[3][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = acc[2][i] = 0.f;
  if (vi < VB) {
    float m[V], r[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      m[i] = mean[n * c + l * V + i];
      r[i] = rstd[n * c + l * V + i];
    }
    const float a = alpha[0];
    const int64_t vox0 = (int64_t)n * spatial;
    auto body = [&](const Vec<T, V>& xv, const Vec<T, V>& gv) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float h = (xv.v[i] - m[i]) * r[i];
        bool pos = h > 0.f;
        float g = pos ? gv.v[i] : a * gv.v[i];
        acc[0][i] += g;
        acc[1][i] = fmaf(g, h, acc[1][i]);
        acc[2][i] += pos ? 0.f : gv.v[i] * h;
      }
    };
    // two voxels per iteration, all four 16-byte loads issued before the arithmetic: twice the bytes in flight per
    // thread (the pass is latency bound at one row pair per thread: 4.2 TB/s of reads, ncu r2)
    int64_t v = v_begin + vi;
    for (; v + VB < v_end; v += 2 * VB) {
      Vec<T, V> xv0, gv0, xv1, gv1;
      xv0.load(x + (vox0 + v) * x_ld + l * V);
      gv0.load(dy + (vox0 + v) * dy_ld + l * V);
      xv1.load(x + (vox0 + v + VB) * x_ld + l * V);
      gv1.load(dy + (vox0 + v + VB) * dy_ld + l * V);
      body(xv0, gv0);
      body(xv1, gv1);
    }
    if (v < v_end) {
      Vec<T, V> xv, gv;
      xv.load(x + (vox0 + v) * x_ld + l * V);
      gv.load(dy + (vox0 + v) * dy_ld + l * V);
      body(xv, gv);
    }
  }
  float* out = partial + ((int64_t)n * gridDim.x + blockIdx.x) * c * 3;
  block_reduce_store<T, V, 3>(acc, L, VB, vi, l, red, out, c, 0);
}

// sums[nc][3] = { s1/S, s2/S, dalpha contribution }
__global__ void instnorm_bwd_final_kernel(const float* __restrict__ partial, int nblk, int c,
                                          int nc_total, int64_t spatial, float* __restrict__ sums) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= nc_total) return;
  int n = warp / c, ch = warp % c;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    const float* p = partial + (((int64_t)n * nblk + b) * c + ch) * 3;
    s1 += (double)p[0];
    s2 += (double)p[1];
    s3 += (double)p[2];
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  s3 = warp_sum_d(s3);
  if (lane == 0) {
    sums[warp * 3 + 0] = (float)(s1 / (double)spatial);
    sums[warp * 3 + 1] = (float)(s2 / (double)spatial);
    sums[warp * 3 + 2] = (float)s3;
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
instnorm_prelu_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ alpha, const float* __restrict__ sums,
                                T* __restrict__ dx, int64_t spatial, int c, int x_ld, int dy_ld,
                                int dx_ld, int L, int VB, int64_t vox_per_block, int nc_total,
                                float* __restrict__ dalpha) {
  const int t = threadIdx.x, vi = t / L, l = t % L;
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    // d loss / d alpha = sum over (n, c) of the per-instance terms: one block, fixed order
    __shared__ double red[256];
    double s = 0.0;
    for (int i = t; i < nc_total; i += 256) s += (double)sums[i * 3 + 2];
    red[t] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) dalpha[0] = (float)red[0];
  }
  if (vi >= VB) return;
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float m[V], r[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    int idx = n * c + l * V + i;
    m[i] = mean[idx];
    r[i] = rstd[idx];
    s1[i] = sums[idx * 3 + 0];
    s2[i] = sums[idx * 3 + 1];
  }
  const float a = alpha[0];
  const int64_t vox0 = (int64_t)n * spatial;
  auto body = [&](const Vec<T, V>& xv, const Vec<T, V>& gv, int64_t v) {
    Vec<T, V> ov;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float h = (xv.v[i] - m[i]) * r[i];
      float g = h > 0.f ? gv.v[i] : a * gv.v[i];
      ov.v[i] = r[i] * (g - s1[i] - h * s2[i]);
    }
    ov.store(dx + (vox0 + v) * dx_ld + l * V);
  };
  int64_t v = v_begin + vi;
  for (; v + VB < v_end; v += 2 * VB) {  // (loads of two voxels first: see the partial kernel)
    Vec<T, V> xv0, gv0, xv1, gv1;
    xv0.load(x + (vox0 + v) * x_ld + l * V);
    gv0.load(dy + (vox0 + v) * dy_ld + l * V);
    xv1.load(x + (vox0 + v + VB) * x_ld + l * V);
    gv1.load(dy + (vox0 + v + VB) * dy_ld + l * V);
    body(xv0, gv0, v);
    body(xv1, gv1, v + VB);
  }
  if (v < v_end) {
    Vec<T, V> xv, gv;
    xv.load(x + (vox0 + v) * x_ld + l * V);
    gv.load(dy + (vox0 + v) * dy_ld + l * V);
    body(xv, gv, v);
  }
}

// d loss / d alpha = sum over (n, c) of sums[.][2], finished INSIDE the kernel that produced the sums: every producer
// CTA takes a ticket after its sums are visible (fence + atomicInc, which wraps back to 0 for the next launch), the
// CTA that draws the last ticket adds all nc terms in a fixed order (strided per thread, then a shared-memory tree in
// double), so the result does not depend on which CTA comes last.  Saves the one-block dalpha_final launch that sat
// on the serial backward chain after each of these kernels (12 per step of the 16-256 net) although nothing on
// that chain reads dalpha.  Call with the whole block; `writer` = this CTA wrote sums and takes a ticket.
__device__ unsigned int g_dalpha_ticket = 0;

template <int NT>
__device__ __forceinline__ void dalpha_last_block(bool writer, unsigned int total, const float* sums, int nc_total,
                                                  float* dalpha) {
  __shared__ double dred[NT];
  __shared__ unsigned int is_last;
  if (!writer) return;  // (block-uniform)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicInc(&g_dalpha_ticket, total - 1) == total - 1 ? 1u : 0u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  for (int i = threadIdx.x; i < nc_total; i += NT) s += (double)__ldcg(sums + i * 3 + 2);
  dred[threadIdx.x] = s;
  __syncthreads();
  for (int o = NT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dalpha[0] = (float)dred[0];
}

// Small-instance backward in ONE launch (deep layers: <= 4096 voxels per sample, tensors of a MB that
// live in L2): a CTA owns V channels of one sample for ALL voxels, so the three spatial sums never
// leave the block -- pass 1 accumulates them, a shared-memory tree makes them block-wide, pass 2
// re-reads x / dy (L2 hits) and writes dx.  Replaces partial + final + apply (three launches of
// ~5-10 us each at these sizes).  sums[nc][3] is still written for the dalpha reduction.
template <typename T, int V>
__global__ void __launch_bounds__(512)
instnorm_prelu_bwd_small_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ alpha, T* __restrict__ dx,
                                int64_t spatial, int c, int x_ld, int dy_ld, int dx_ld, float* __restrict__ sums,
                                float* __restrict__ dalpha) {
  __shared__ float red[16][3 * V];
  __shared__ float tot[3 * V];
  const int t = threadIdx.x, n = blockIdx.y, c0 = blockIdx.x * V;
  float m[V], r[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m[i] = mean[n * c + c0 + i];
    r[i] = rstd[n * c + c0 + i];
  }
  const float a = alpha[0];
  const int64_t vox0 = (int64_t)n * spatial;
  float acc[3][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = acc[2][i] = 0.f;
  for (int64_t v = t; v < spatial; v += 512) {
    Vec<T, V> xv, gv;
    xv.load(x + (vox0 + v) * x_ld + c0);
    gv.load(dy + (vox0 + v) * dy_ld + c0);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float h = (xv.v[i] - m[i]) * r[i];
      const bool pos = h > 0.f;
      const float g = pos ? gv.v[i] : a * gv.v[i];
      acc[0][i] += g;
      acc[1][i] = fmaf(g, h, acc[1][i]);
      acc[2][i] += pos ? 0.f : gv.v[i] * h;
    }
  }
  const int warp = t >> 5, lane = t & 31;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float w = warp_sum(acc[k][i]);
      if (lane == 0) red[warp][k * V + i] = w;
    }
  __syncthreads();
  if (t < 3 * V) {
    float s_ = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) s_ += red[w][t];  // fixed order
    const int k = t / V, i = t % V;
    const float o = k < 2 ? s_ / (float)spatial : s_;
    tot[t] = o;
    sums[(n * c + c0 + i) * 3 + k] = o;
  }
  __syncthreads();
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s1[i] = tot[i];
    s2[i] = tot[V + i];
  }
  for (int64_t v = t; v < spatial; v += 512) {
    Vec<T, V> xv, gv, ov;
    xv.load(x + (vox0 + v) * x_ld + c0);
    gv.load(dy + (vox0 + v) * dy_ld + c0);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float h = (xv.v[i] - m[i]) * r[i];
      const float g = h > 0.f ? gv.v[i] : a * gv.v[i];
      ov.v[i] = r[i] * (g - s1[i] - h * s2[i]);
    }
    ov.store(dx + (vox0 + v) * dx_ld + c0);
  }
  dalpha_last_block<512>(true, gridDim.x * gridDim.y, sums, (int)gridDim.y * c, dalpha);
}

// Mid-size instances (up to 32^3 voxels per sample) in ONE launch on a thread-block CLUSTER: the CL CTAs of a cluster
// own the same V channels of one sample and 1/CL of its voxels each.  Pass 1 leaves a CTA's three partial sums in its
// shared memory, after the cluster barrier every CTA adds up the CL partials through distributed shared memory in rank
// order (so all of them hold bit-identical totals), pass 2 re-reads the CTA's own voxels (L2 hits) and writes dx.
// Replaces, for 16^3 instances, the single-CTA-per-channel-group kernel above (16 CTAs walking 4096 voxels each:
// 20 us in the r2 launch list for a 1 MB tensor, pure load latency) and, for 32^3 instances, the three launches
// partial + final + apply (22 us).  Loads of two iterations are issued before the arithmetic.
template <typename T, int V>
__global__ void __launch_bounds__(512)
instnorm_prelu_bwd_cluster_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const float* __restrict__ alpha, T* __restrict__ dx,
                                  int64_t spatial, int c, int x_ld, int dy_ld, int dx_ld, int CL,
                                  float* __restrict__ sums, float* __restrict__ dalpha) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[16][3 * V];
  __shared__ float part[3 * V];  // this CTA's partial sums: read by the whole cluster
  __shared__ float tot[3 * V];
  const int t = threadIdx.x, n = blockIdx.y;
  const int rank = (int)cluster.block_rank();
  const int c0 = (blockIdx.x / CL) * V;
  const int64_t per = (spatial + CL - 1) / CL;
  const int64_t v_begin = (int64_t)rank * per, v_end = min(spatial, v_begin + per);
  float m[V], r[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m[i] = mean[n * c + c0 + i];
    r[i] = rstd[n * c + c0 + i];
  }
  const float a = alpha[0];
  const int64_t vox0 = (int64_t)n * spatial;
  float acc[3][V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = acc[2][i] = 0.f;
  auto reduce = [&](const Vec<T, V>& xv, const Vec<T, V>& gv) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float h = (xv.v[i] - m[i]) * r[i];
      const bool pos = h > 0.f;
      const float g = pos ? gv.v[i] : a * gv.v[i];
      acc[0][i] += g;
      acc[1][i] = fmaf(g, h, acc[1][i]);
      acc[2][i] += pos ? 0.f : gv.v[i] * h;
    }
  };
  {
    int64_t v = v_begin + t;
    for (; v + 512 < v_end; v += 1024) {
      Vec<T, V> x0, g0, x1, g1;
      x0.load(x + (vox0 + v) * x_ld + c0);
      g0.load(dy + (vox0 + v) * dy_ld + c0);
      x1.load(x + (vox0 + v + 512) * x_ld + c0);
      g1.load(dy + (vox0 + v + 512) * dy_ld + c0);
      reduce(x0, g0);
      reduce(x1, g1);
    }
    if (v < v_end) {
      Vec<T, V> x0, g0;
      x0.load(x + (vox0 + v) * x_ld + c0);
      g0.load(dy + (vox0 + v) * dy_ld + c0);
      reduce(x0, g0);
    }
  }
  const int warp = t >> 5, lane = t & 31;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float w = warp_sum(acc[k][i]);
      if (lane == 0) red[warp][k * V + i] = w;
    }
  __syncthreads();
  if (t < 3 * V) {
    float s_ = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) s_ += red[w][t];  // fixed order
    part[t] = s_;
  }
  cluster.sync();
  if (t < 3 * V) {
    float s_ = 0.f;
    for (int rk = 0; rk < CL; ++rk) s_ += *cluster.map_shared_rank(&part[t], rk);  // rank order: same bits in every CTA
    const int k = t / V, i = t % V;
    const float o = k < 2 ? s_ / (float)spatial : s_;
    tot[t] = o;
    if (rank == 0) sums[(n * c + c0 + i) * 3 + k] = o;
  }
  cluster.sync();  // totals visible to the block; no CTA leaves while a peer may still read its partials
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s1[i] = tot[i];
    s2[i] = tot[V + i];
  }
  auto apply = [&](const Vec<T, V>& xv, const Vec<T, V>& gv, int64_t v) {
    Vec<T, V> ov;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float h = (xv.v[i] - m[i]) * r[i];
      const float g = h > 0.f ? gv.v[i] : a * gv.v[i];
      ov.v[i] = r[i] * (g - s1[i] - h * s2[i]);
    }
    ov.store(dx + (vox0 + v) * dx_ld + c0);
  };
  int64_t v = v_begin + t;
  for (; v + 512 < v_end; v += 1024) {
    Vec<T, V> x0, g0, x1, g1;
    x0.load(x + (vox0 + v) * x_ld + c0);
    g0.load(dy + (vox0 + v) * dy_ld + c0);
    x1.load(x + (vox0 + v + 512) * x_ld + c0);
    g1.load(dy + (vox0 + v + 512) * dy_ld + c0);
    apply(x0, g0, v);
    apply(x1, g1, v + 512);
  }
  if (v < v_end) {
    Vec<T, V> x0, g0;
    x0.load(x + (vox0 + v) * x_ld + c0);
    g0.load(dy + (vox0 + v) * dy_ld + c0);
    apply(x0, g0, v);
  }
  dalpha_last_block<512>(rank == 0, (gridDim.x / CL) * gridDim.y, sums, (int)gridDim.y * c, dalpha);
}

// ------------------------------------------------------------------------------------------
namespace {

int make_geom(const b200seg_norm_desc& d, int V, NormGeom& g) {
  g.V = V;
  g.L = d.c / V;
  if (g.L > 256 || g.L < 1) {
    set_error("instnorm: %d channels with vector width %d needs more than 256 threads per voxel",
              d.c, V);
    return B200SEG_ERR_UNSUPPORTED;
  }
  g.VB = 256 / g.L;
  g.nblk = norm_blocks(d);
  int64_t nb = cdiv64(d.spatial, (int64_t)g.VB * 8);
  int64_t cap = 2368 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  g.nblk_apply = (int)nb;
  return B200SEG_OK;
}

#define DISPATCH_TV(dtype, V, ...)                                                       \
  do {                                                                                   \
    if (dtype == B200SEG_BF16) {                                                         \
      using T = __nv_bfloat16;                                                           \
      switch (V) {                                                                       \
        case 8: { constexpr int VV = 8; __VA_ARGS__; } break;                            \
        case 4: { constexpr int VV = 4; __VA_ARGS__; } break;                            \
        case 2: { constexpr int VV = 2; __VA_ARGS__; } break;                            \
        default: { constexpr int VV = 1; __VA_ARGS__; } break;                           \
      }                                                                                  \
    } else {                                                                             \
      using T = float;                                                                   \
      switch (V) {                                                                       \
        case 4: { constexpr int VV = 4; __VA_ARGS__; } break;                            \
        case 2: { constexpr int VV = 2; __VA_ARGS__; } break;                            \
        default: { constexpr int VV = 1; __VA_ARGS__; } break;                           \
      }                                                                                  \
    }                                                                                    \
  } while (0)

}  // namespace

int launch_instnorm_stats(const b200seg_norm_desc& d, const void* x, float* mean, float* rstd,
                          void* ws, cudaStream_t st) {
  NormGeom g;
  int V = pick_vec(d, {x}, {d.x_ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  float* partial = (float*)ws;
  int64_t per = cdiv64(d.spatial, g.nblk);
  dim3 grid(g.nblk, d.n);
  size_t smem = 256 * 2 * V * sizeof(float);
  DISPATCH_TV(d.dtype, V,
              (instnorm_stats_partial_kernel<T, VV><<<grid, 256, smem, st>>>(
                  (const T*)x, d.spatial, d.c, d.x_ld, g.L, g.VB, per, partial)));
  B200SEG_CHECK_LAUNCH("instnorm_stats_partial");
  int nc = d.n * d.c;
  instnorm_stats_final_kernel<<<(nc * 32 + 255) / 256, 256, 0, st>>>(partial, g.nblk, d.c, nc,
                                                                     d.spatial, d.eps, mean, rstd);
  B200SEG_CHECK_LAUNCH("instnorm_stats_final");
  return B200SEG_OK;
}

int launch_instnorm_prelu_fwd(const b200seg_norm_desc& d, const void* x, const float* mean,
                              const float* rstd, const float* alpha, const void* res, void* y,
                              cudaStream_t st) {
  NormGeom g;
  int V = res ? pick_vec(d, {x, y, res}, {d.x_ld, d.y_ld, d.r_ld})
              : pick_vec(d, {x, y}, {d.x_ld, d.y_ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  int64_t per = cdiv64(d.spatial, g.nblk_apply);
  dim3 grid(g.nblk_apply, d.n);
  DISPATCH_TV(d.dtype, V,
              (instnorm_prelu_fwd_kernel<T, VV><<<grid, 256, 0, st>>>(
                  (const T*)x, mean, rstd, alpha, (const T*)res, (T*)y, d.spatial, d.c, d.x_ld,
                  d.y_ld, d.r_ld, g.L, g.VB, per)));
  B200SEG_CHECK_LAUNCH("instnorm_prelu_fwd");
  return B200SEG_OK;
}

int launch_instnorm_prelu_fwd_partials(const b200seg_norm_desc& d, const void* x, const float* partial, int ncls,
                                       int64_t tiles, int cstat, const float* alpha, const void* res, void* y,
                                       float* mean, float* rstd, cudaStream_t st) {
  NormGeom g;
  int V = res ? pick_vec(d, {x, y, res}, {d.x_ld, d.y_ld, d.r_ld}) : pick_vec(d, {x, y}, {d.x_ld, d.y_ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  int64_t per = cdiv64(d.spatial, g.nblk_apply);
  dim3 grid(g.nblk_apply, d.n);
  DISPATCH_TV(d.dtype, V,
              (instnorm_prelu_fwd_partials_kernel<T, VV><<<grid, 256, 0, st>>>(
                  (const T*)x, partial, ncls, d.n, tiles, cstat, d.eps, alpha, (const T*)res, (T*)y, d.spatial, d.c,
                  d.x_ld, d.y_ld, d.r_ld, g.L, g.VB, per, mean, rstd)));
  B200SEG_CHECK_LAUNCH("instnorm_prelu_fwd_partials");
  return B200SEG_OK;
}

int launch_instnorm_prelu_bwd(const b200seg_norm_desc& d, const void* x, const float* mean,
                              const float* rstd, const float* alpha, const void* dy, void* dx,
                              float* dalpha, void* ws, cudaStream_t st) {
  // descriptor roles in the backward: x_ld -> x, y_ld -> dy, r_ld -> dx
  NormGeom g;
  int V = pick_vec(d, {x, dy, dx}, {d.x_ld, d.y_ld, d.r_ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  float* partial = (float*)ws;
  float* sums = partial + (size_t)d.n * g.nblk * d.c * 3;
  // B200SEG_NORM_CLUSTER=0: the r1 / r2 paths (A/B runs)
  static const bool use_cluster = [] { const char* e = getenv("B200SEG_NORM_CLUSTER"); return !(e && e[0] == '0'); }();
  if (use_cluster && d.spatial > 1024 && d.spatial <= 32768 && V >= 4) {  // 16^3 .. 32^3: one launch on clusters
    int CL = 2;
    while (CL < 8 && (int64_t)CL * 1024 < d.spatial) CL *= 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(d.c / V * CL), (unsigned)d.n);
    cfg.blockDim = dim3(512);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    DISPATCH_TV(d.dtype, V,
                (e = cudaLaunchKernelEx(&cfg, instnorm_prelu_bwd_cluster_kernel<T, VV>, (const T*)x, (const T*)dy, mean,
                                        rstd, alpha, (T*)dx, d.spatial, d.c, d.x_ld, d.y_ld, d.r_ld, CL, sums, dalpha)));
    if (e != cudaSuccess) {
      set_error("instnorm_prelu_bwd_cluster launch failed: %s", cudaGetErrorString(e));
      return B200SEG_ERR_CUDA;
    }
    B200SEG_CHECK_LAUNCH("instnorm_prelu_bwd_cluster");
    return B200SEG_OK;
  }
  if (d.spatial <= 4096 && V >= 4) {  // deep layers: one launch, sums stay in the block
    dim3 gs(d.c / V, d.n);
    DISPATCH_TV(d.dtype, V,
                (instnorm_prelu_bwd_small_kernel<T, VV><<<gs, 512, 0, st>>>(
                    (const T*)x, (const T*)dy, mean, rstd, alpha, (T*)dx, d.spatial, d.c, d.x_ld, d.y_ld,
                    d.r_ld, sums, dalpha)));
    B200SEG_CHECK_LAUNCH("instnorm_prelu_bwd_small");
    return B200SEG_OK;
  }
  int64_t per = cdiv64(d.spatial, g.nblk);
  dim3 grid(g.nblk, d.n);
  size_t smem = 256 * 3 * V * sizeof(float);
  DISPATCH_TV(d.dtype, V,
              (instnorm_prelu_bwd_partial_kernel<T, VV><<<grid, 256, smem, st>>>(
                  (const T*)x, (const T*)dy, mean, rstd, alpha, d.spatial, d.c, d.x_ld, d.y_ld,
                  g.L, g.VB, per, partial)));
  B200SEG_CHECK_LAUNCH("instnorm_prelu_bwd_partial");
  int nc = d.n * d.c;
  instnorm_bwd_final_kernel<<<(nc * 32 + 255) / 256, 256, 0, st>>>(partial, g.nblk, d.c, nc,
                                                                   d.spatial, sums);
  B200SEG_CHECK_LAUNCH("instnorm_bwd_final");
  int64_t per2 = cdiv64(d.spatial, g.nblk_apply);
  dim3 grid2(g.nblk_apply, d.n);
  DISPATCH_TV(d.dtype, V,
              (instnorm_prelu_bwd_apply_kernel<T, VV><<<grid2, 256, 0, st>>>(
                  (const T*)x, (const T*)dy, mean, rstd, alpha, sums, (T*)dx, d.spatial, d.c,
                  d.x_ld, d.y_ld, d.r_ld, g.L, g.VB, per2, nc, dalpha)));
  B200SEG_CHECK_LAUNCH("instnorm_prelu_bwd_apply");
  return B200SEG_OK;
}

// InstanceNorm + PReLU backward whose REDUCTION pass already happened in the epilogue of the dgrad that produced
// dy (tc_slide_conv_kernel<.., BST>): partial[n][rows][c][3] per-CTA sums -> sums (fixed order, double) -> apply.
int launch_instnorm_prelu_bwd_from_partials(const b200seg_norm_desc& d, const void* x, const float* mean,
                                            const float* rstd, const float* alpha, const void* dy,
                                            const float* partial, int64_t rows, void* dx, float* dalpha, void* ws,
                                            cudaStream_t st) {
  NormGeom g;
  int V = pick_vec(d, {x, dy, dx}, {d.x_ld, d.y_ld, d.r_ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  float* sums = (float*)ws;
  int nc = d.n * d.c;
  instnorm_bwd_final_kernel<<<(nc * 32 + 255) / 256, 256, 0, st>>>(partial, (int)rows, d.c, nc, d.spatial, sums);
  B200SEG_CHECK_LAUNCH("instnorm_bwd_final");
  int64_t per2 = cdiv64(d.spatial, g.nblk_apply);
  dim3 grid2(g.nblk_apply, d.n);
  DISPATCH_TV(d.dtype, V,
              (instnorm_prelu_bwd_apply_kernel<T, VV><<<grid2, 256, 0, st>>>(
                  (const T*)x, (const T*)dy, mean, rstd, alpha, sums, (T*)dx, d.spatial, d.c,
                  d.x_ld, d.y_ld, d.r_ld, g.L, g.VB, per2, nc, dalpha)));
  B200SEG_CHECK_LAUNCH("instnorm_prelu_bwd_apply");
  return B200SEG_OK;
}

// mean / rstd from per-(n, c) sum and sum of squares accumulated by a convolution epilogue
// mean / rstd from the per-CTA partial sums a convolution epilogue produced.  partial index of
// (class o, sample n, tile t) = (o * N + n) * T + t;  one warp per (n, channel), fixed order, double.
// mean / rstd are written with `c_out` (>= c) entries per sample: entries >= c describe zero
// padding channels (mean 0, rstd 1/sqrt(eps)), as the statistics kernel would produce for them.
__global__ void instnorm_stats_from_partials_kernel(const float* __restrict__ partial, int n, int c, int c_out,
                                                    int ncls, int64_t tiles, int64_t spatial, float eps,
                                                    float* __restrict__ mean, float* __restrict__ rstd) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= n * c_out) return;
  const int nn = warp / c_out, ch = warp % c_out;
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
    for (int o = 0; o < ncls; ++o) {
      const float* base = partial + (((int64_t)o * n + nn) * tiles) * c * 2;
      for (int64_t t = lane; t < tiles; t += 32) {
        s1 += (double)base[(t * c + ch) * 2];
        s2 += (double)base[(t * c + ch) * 2 + 1];
      }
    }
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane == 0) {
    double m = s1 / (double)spatial;
    double var = s2 / (double)spatial - m * m;
    if (var < 0.0) var = 0.0;
    mean[warp] = (float)m;
    rstd[warp] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int launch_instnorm_stats_from_partials(const float* partial, int n, int c, int c_out, int ncls, int64_t tiles,
                                        int64_t spatial, float eps, float* mean, float* rstd, cudaStream_t st) {
  const int total = n * c_out;
  instnorm_stats_from_partials_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(partial, n, c, c_out, ncls, tiles,
                                                                                spatial, eps, mean, rstd);
  B200SEG_CHECK_LAUNCH("instnorm_stats_from_partials");
  return B200SEG_OK;
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradient):  out[c] = sum_v x[v*ld + c]   -- the statistics kernel with the
// batch folded into the voxel axis; per-block partials, fixed-order finalisation in double
// ------------------------------------------------------------------------------------------
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblk, int c,
                                    float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= c) return;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += (double)partial[((int64_t)b * c + warp) * 2];
  s = warp_sum_d(s);
  if (lane == 0) out[warp] = (float)s;
}

static b200seg_norm_desc colsum_desc(int dtype, int64_t nvox, int c, int ld) {
  b200seg_norm_desc d;
  d.n = 1; d.c = c; d.spatial = nvox; d.x_ld = ld; d.y_ld = ld; d.r_ld = ld; d.dtype = dtype; d.eps = 0.f;
  return d;
}

size_t colsum_workspace_bytes(int64_t nvox, int c) {
  b200seg_norm_desc d = colsum_desc(B200SEG_F32, nvox, c, c);
  return (size_t)norm_blocks(d) * c * 2 * sizeof(float) + 256;
}

int launch_colsum(int dtype, const void* x, int64_t nvox, int c, int ld, float* out, float* partial,
                  cudaStream_t st) {
  b200seg_norm_desc d = colsum_desc(dtype, nvox, c, ld);
  NormGeom g;
  int V = pick_vec(d, {x}, {ld});
  int rc = make_geom(d, V, g);
  if (rc) return rc;
  int64_t per = cdiv64(nvox, g.nblk);
  dim3 grid(g.nblk, 1);
  size_t smem = 256 * 2 * V * sizeof(float);
  DISPATCH_TV(dtype, V,
              (instnorm_stats_partial_kernel<T, VV><<<grid, 256, smem, st>>>((const T*)x, nvox, c, ld, g.L,
                                                                            g.VB, per, partial)));
  B200SEG_CHECK_LAUNCH("colsum_partial");
  colsum_final_kernel<<<(c * 32 + 255) / 256, 256, 0, st>>>(partial, g.nblk, c, out);
  B200SEG_CHECK_LAUNCH("colsum_final");
  return B200SEG_OK;
}

// ------------------------------------------------------------------------------------------
// Adam on ONE flat fp32 parameter / gradient / moment buffer (torch.optim.Adam semantics: no weight
// decay, no amsgrad):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//                      p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// 128-bit accesses, 7 x 4 bytes of traffic per parameter.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float beta1, float beta2, float step_size, float inv_bc2_sqrt, float eps) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
             vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ma[k] = beta1 * ma[k] + (1.f - beta1) * ga[k];
        va[k] = beta2 * va[k] + (1.f - beta2) * ga[k] * ga[k];
        pa[k] -= step_size * (ma[k] / (sqrtf(va[k]) * inv_bc2_sqrt + eps));
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n; ++j) {
        const float gj = g[j];
        m[j] = beta1 * m[j] + (1.f - beta1) * gj;
        v[j] = beta2 * v[j] + (1.f - beta2) * gj * gj;
        p[j] -= step_size * (m[j] / (sqrtf(v[j]) * inv_bc2_sqrt + eps));
      }
    }
  }
}

int launch_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t step, cudaStream_t st) {
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int64_t nb = cdiv64(n, 256 * 4 * 4);
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  adam_flat_kernel<<<(unsigned)nb, 256, 0, st>>>(p, g, m, v, n, beta1, beta2, step_size, inv_bc2_sqrt, eps);
  B200SEG_CHECK_LAUNCH("adam_flat");
  return B200SEG_OK;
}

}  // namespace b200seg

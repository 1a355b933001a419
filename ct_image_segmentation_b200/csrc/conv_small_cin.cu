// Direct (CUDA-core) kernels for the network's first convolutions, Cin <= 4 (CT volumes have one
// channel; the 2-D windowed input has three).  27*Cin multiply-adds per output value: these layers
// are bandwidth-bound (AI ~ 18 FLOP/B), a tensor-core tile would be > 90 %
// zero padding.  fprop: one thread per output voxel and 16 output channels.  wgrad: the gathered
// inputs of a voxel chunk (an im2col row of 27*Cin values per voxel) and the output gradients are
// staged in shared memory, every thread owns a few (tap, ci, co) sums; per-block partials are
// reduced in a fixed order (deterministic).  No dgrad: the network input needs no gradient.
#include "common.cuh"
#include "kernels.h"

namespace b200seg {

struct SmallCinParams {
  int n, cin, cout;
  int iD, iH, iW, oD, oH, oW;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int x_ld, y_ld, r_ld;
  int64_t nvox;  // n*oD*oH*oW
};

// y[v][co] = bias[co] + sum_{tap,ci} x[in(v,tap)][ci] * w[tap][ci][co]   (+ residual)
template <typename T>
__global__ void __launch_bounds__(128)
conv_small_cin_fprop_kernel(SmallCinParams p, const T* __restrict__ x, const T* __restrict__ w,
                            const float* __restrict__ bias, const T* __restrict__ res, T* __restrict__ y) {
  extern __shared__ float ws[];  // [taps*cin][16] weights of this channel group
  const int taps = p.kd * p.kh * p.kw;
  const int J = taps * p.cin;
  const int co0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < J * 16; i += blockDim.x) {
    int j = i / 16, c = i % 16;
    ws[i] = (co0 + c < p.cout) ? to_f<T>(w[(int64_t)j * p.cout + co0 + c]) : 0.f;
  }
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.nvox) return;
  int64_t r = v;
  const int ow = (int)(r % p.oW); r /= p.oW;
  const int oh = (int)(r % p.oH); r /= p.oH;
  const int od = (int)(r % p.oD); r /= p.oD;
  const int n = (int)r;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = (bias && co0 + c < p.cout) ? bias[co0 + c] : 0.f;
  int j = 0;
  for (int kd = 0; kd < p.kd; ++kd) {
    const int id = od * p.sd - p.pd + kd;
    for (int kh = 0; kh < p.kh; ++kh) {
      const int ih = oh * p.sh - p.ph + kh;
      for (int kw = 0; kw < p.kw; ++kw) {
        const int iw = ow * p.sw - p.pw + kw;
        const bool in = id >= 0 && id < p.iD && ih >= 0 && ih < p.iH && iw >= 0 && iw < p.iW;
        const T* xp = x + ((((int64_t)n * p.iD + id) * p.iH + ih) * p.iW + iw) * (int64_t)p.x_ld;
        for (int ci = 0; ci < p.cin; ++ci, ++j) {
          const float xv = in ? to_f<T>(xp[ci]) : 0.f;
          const float* wr = ws + j * 16;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = fmaf(xv, wr[c], acc[c]);
        }
      }
    }
  }
  T* yp = y + v * (int64_t)p.y_ld + co0;
  const T* rp = res ? res + v * (int64_t)p.r_ld + co0 : nullptr;
  const bool vec = (co0 + 16 <= p.cout) && (((uintptr_t)yp % (8 * sizeof(T))) == 0) &&
                   (!rp || ((uintptr_t)rp % (8 * sizeof(T))) == 0);
  if (vec) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      Vec<T, 8> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = acc[h * 8 + i];
      if (rp) {
        Vec<T, 8> rv;
        rv.load(rp + h * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] += rv.v[i];
      }
      o.store(yp + h * 8);
    }
  } else {
    for (int c = 0; c < 16 && co0 + c < p.cout; ++c) {
      float o = acc[c];
      if (rp) o += to_f<T>(rp[c]);
      yp[c] = from_f<T>(o);
    }
  }
}

// partial[blk][j][co] = sum over the block's voxels of x[in(v,tap)][ci] * dy[v][co],  j = tap*cin+ci
template <typename T, int R>
__global__ void __launch_bounds__(256)
conv_small_cin_wgrad_kernel(SmallCinParams p, const T* __restrict__ x, const T* __restrict__ dy,
                            float* __restrict__ partial, int64_t vox_per_block) {
  constexpr int VC = 64;             // voxels per chunk
  extern __shared__ float sm[];      // X[J][VC] | DY[VC][cout] | per-voxel base offsets | tap table
  const int taps = p.kd * p.kh * p.kw;
  const int J = taps * p.cin;
  const int CO = p.cout;
  float* Xs = sm;                    // [J][VC]
  float* Ds = sm + (VC + 1) * J;     // [VC][CO]
  int* vbase = reinterpret_cast<int*>(Ds + VC * CO);   // [VC][4]: n, id0, ih0, iw0 (n < 0: no voxel)
  int* jtab = vbase + VC * 4;        // [J][4]: kd, kh, kw, ci
  for (int j = threadIdx.x; j < J; j += 256) {
    const int tap = j / p.cin;
    jtab[j * 4 + 0] = tap / (p.kw * p.kh);
    jtab[j * 4 + 1] = (tap / p.kw) % p.kh;
    jtab[j * 4 + 2] = tap % p.kw;
    jtab[j * 4 + 3] = j % p.cin;
  }
  const int nout = J * CO;
  // R = outputs per thread (J*CO <= 256*R)
  float acc[R];
  int oj[R], oc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    acc[r] = 0.f;
    int o = threadIdx.x + r * 256;
    oj[r] = o < nout ? (o / CO) * (VC + 1) : 0;
    oc[r] = o < nout ? o % CO : 0;
  }
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, p.nvox);
  for (int64_t v0 = v_begin; v0 < v_end; v0 += VC) {
    __syncthreads();
    if (threadIdx.x < VC) {
      const int64_t v = v0 + threadIdx.x;
      int nn = -1, id0 = 0, ih0 = 0, iw0 = 0;
      if (v < v_end) {
        int64_t r = v;
        const int ow = (int)(r % p.oW); r /= p.oW;
        const int oh = (int)(r % p.oH); r /= p.oH;
        const int od = (int)(r % p.oD); r /= p.oD;
        nn = (int)r;
        id0 = od * p.sd - p.pd; ih0 = oh * p.sh - p.ph; iw0 = ow * p.sw - p.pw;
      }
      vbase[threadIdx.x * 4 + 0] = nn; vbase[threadIdx.x * 4 + 1] = id0;
      vbase[threadIdx.x * 4 + 2] = ih0; vbase[threadIdx.x * 4 + 3] = iw0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VC * J; i += 256) {
      const int lv = i % VC, j = i / VC;  // consecutive threads -> consecutive voxels
      const int nn = vbase[lv * 4];
      float val = 0.f;
      if (nn >= 0) {
        const int id = vbase[lv * 4 + 1] + jtab[j * 4], ih = vbase[lv * 4 + 2] + jtab[j * 4 + 1],
                  iw = vbase[lv * 4 + 3] + jtab[j * 4 + 2];
        if (id >= 0 && id < p.iD && ih >= 0 && ih < p.iH && iw >= 0 && iw < p.iW)
          val = to_f<T>(x[((((int64_t)nn * p.iD + id) * p.iH + ih) * p.iW + iw) * (int64_t)p.x_ld + jtab[j * 4 + 3]]);
      }
      Xs[j * (VC + 1) + lv] = val;  // Xs[j][lv], padded rows: conflict-free column reads
    }
    for (int i = threadIdx.x; i < VC * CO; i += 256) {
      const int lv = i / CO, c = i % CO;
      const int64_t v = v0 + lv;
      Ds[i] = v < v_end ? to_f<T>(dy[v * (int64_t)p.y_ld + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int lv = 0; lv < VC; ++lv) {
      const float* dr = Ds + lv * CO;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(Xs[oj[r] + lv], dr[oc[r]], acc[r]);
    }
  }
  float* out = partial + (int64_t)blockIdx.x * nout;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int o = threadIdx.x + r * 256;
    if (o < nout) out[o] = acc[r];
  }
}

// Register-tiled variant for J = taps*cin <= 32 and cout in {16, 32, 64}: a chunk of 128 voxels is
// staged in shared memory (X[j][voxel] gathered, DY[voxel][co]); a "team" of cout threads owns the
// whole (j, co) output as 8 x 4 register tiles and walks its share of the chunk's voxels (8 scalar
// + one 128-bit shared load per 32 FMA); the 256/cout teams of a block are summed in a fixed order.
template <typename T, int CO>
__global__ void __launch_bounds__(256)
conv_small_cin_wgrad_tile_kernel(SmallCinParams p, const T* __restrict__ x, const T* __restrict__ dy,
                                 float* __restrict__ partial, int64_t vox_per_block, int dy_vec) {
  constexpr int VC = 128, XP = VC + 1, JP = 32;
  constexpr int NTEAMS = 256 / CO, CQ = CO / 4;
  extern __shared__ float sm[];
  float* Xs = sm;                                        // [JP][XP]
  float* Ds = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sm + JP * XP) + 15) & ~uintptr_t(15));  // [VC][CO]
  int4* vbase = reinterpret_cast<int4*>(Ds + VC * CO);   // [VC]: n (< 0: no voxel), id0, ih0, iw0
  const int J = p.kd * p.kh * p.kw * p.cin;
  const int rows = p.kd * p.kh, kwc = p.kw * p.cin;
  const int t = threadIdx.x;
  const int team = t / CO, jq = (t % CO) / CQ, cq = t % CQ;
  for (int i = t; i < JP * XP; i += 256) Xs[i] = 0.f;    // rows >= J stay zero
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, p.nvox);
  for (int64_t v0 = v_begin; v0 < v_end; v0 += VC) {
    __syncthreads();
    if (t < VC) {
      const int64_t v = v0 + t;
      int4 e = make_int4(-1, 0, 0, 0);
      if (v < v_end) {
        int64_t r = v;
        const int ow = (int)(r % p.oW); r /= p.oW;
        const int oh = (int)(r % p.oH); r /= p.oH;
        const int od = (int)(r % p.oD); r /= p.oD;
        e = make_int4((int)r, od * p.sd - p.pd, oh * p.sh - p.ph, ow * p.sw - p.pw);
      }
      vbase[t] = e;
    }
    __syncthreads();
    for (int i = t; i < VC * rows; i += 256) {
      const int lv = i % VC, row = i / VC;   // consecutive threads -> consecutive voxels
      const int kd = row / p.kh, kh = row - kd * p.kh;
      const int4 e = vbase[lv];
      const int id = e.y + kd, ih = e.z + kh;
      const bool ok = e.x >= 0 && id >= 0 && id < p.iD && ih >= 0 && ih < p.iH;
      const T* xr = x + ((((int64_t)e.x * p.iD + id) * p.iH + ih) * p.iW) * (int64_t)p.x_ld;
      float* xs = Xs + (row * kwc) * XP + lv;
      for (int kw = 0; kw < p.kw; ++kw) {
        const int iw = e.w + kw;
        const bool okw = ok && iw >= 0 && iw < p.iW;
        for (int ci = 0; ci < p.cin; ++ci)
          xs[(kw * p.cin + ci) * XP] = okw ? to_f<T>(xr[(int64_t)iw * p.x_ld + ci]) : 0.f;
      }
    }
    for (int i = t; i < VC * CQ; i += 256) {
      const int lv = i / CQ, c4 = i % CQ;
      const int64_t v = v0 + lv;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < v_end) {
        const T* dp = dy + v * (int64_t)p.y_ld + 4 * c4;
        if (dy_vec) {
          Vec<T, 4> dv;
          dv.load(dp);
          o = make_float4(dv.v[0], dv.v[1], dv.v[2], dv.v[3]);
        } else {
          o = make_float4(to_f<T>(dp[0]), to_f<T>(dp[1]), to_f<T>(dp[2]), to_f<T>(dp[3]));
        }
      }
      reinterpret_cast<float4*>(Ds)[i] = o;
    }
    __syncthreads();
    const float* xt = Xs + (8 * jq) * XP;
#pragma unroll 2
    for (int lv = team; lv < VC; lv += NTEAMS) {
      const float4 dv = reinterpret_cast<const float4*>(Ds)[lv * CQ + cq];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xv = xt[i * XP + lv];
        acc[i][0] = fmaf(xv, dv.x, acc[i][0]);
        acc[i][1] = fmaf(xv, dv.y, acc[i][1]);
        acc[i][2] = fmaf(xv, dv.z, acc[i][2]);
        acc[i][3] = fmaf(xv, dv.w, acc[i][3]);
      }
    }
  }
  __syncthreads();
  float* red = sm;  // [NTEAMS][JP*CO]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) red[team * (JP * CO) + (8 * jq + i) * CO + 4 * cq + c] = acc[i][c];
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (J * CO);
  for (int o = t; o < J * CO; o += 256) {
    float s_ = 0.f;
#pragma unroll
    for (int k = 0; k < NTEAMS; ++k) s_ += red[k * (JP * CO) + o];
    out[o] = s_;
  }
}

// gw[co][ci][tap] = sum_blk partial[blk][tap*cin+ci][co]; one warp per output, fixed order
__global__ void conv_small_cin_wgrad_final_kernel(const float* __restrict__ partial, int nblk, int taps,
                                                  int cin, int cout, float* __restrict__ gw) {
  const int nout = taps * cin * cout;
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (o >= nout) return;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += (double)partial[(int64_t)b * nout + o];
  s = warp_sum_d(s);
  if (lane == 0) {
    const int co = o % cout, j = o / cout;
    const int tap = j / cin, ci = j % cin;
    gw[((int64_t)co * cin + ci) * taps + tap] = (float)s;
  }
}

// im2col of a small-Cin layer: col[v][ci*taps + tap] = x[in(v, tap)][ci] (zero outside the volume) for
// every OUTPUT voxel v.  J = taps*cin <= 32 values per voxel, written as 16-byte vectors; channels
// [J, col_ld) are the buffer's zero padding and are rewritten as zeros.  With the PyTorch weight
// (cout, cin, taps) read as a (cout, J) matrix the layer becomes a 1x1x1 convolution with J input
// channels on `col`: fprop and wgrad then run on the tcgen05 kernels, and the two first-layer convs of
// a residual unit (same input, same geometry) share one im2col.
template <typename T>
__global__ void __launch_bounds__(128)
im2col_small_cin_kernel(SmallCinParams p, const T* __restrict__ x, T* __restrict__ col, int col_ld) {
  extern __shared__ int4 jtab[];  // (kd - pd, kh - ph, kw - pw, ci) of column j
  const int taps = p.kd * p.kh * p.kw;
  const int J = taps * p.cin;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const int ci = j / taps, tap = j % taps;
    jtab[j] = make_int4(tap / (p.kw * p.kh) - p.pd, (tap / p.kw) % p.kh - p.ph, tap % p.kw - p.pw, ci);
  }
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.nvox) return;
  int64_t r = v;
  const int ow = (int)(r % p.oW); r /= p.oW;
  const int oh = (int)(r % p.oH); r /= p.oH;
  const int od = (int)(r % p.oD); r /= p.oD;
  const int id0 = od * p.sd, ih0 = oh * p.sh, iw0 = ow * p.sw;
  const T* xn = x + r * (int64_t)p.iD * p.iH * p.iW * p.x_ld;
  T* out = col + v * (int64_t)col_ld;
  constexpr int VE = 16 / sizeof(T);  // elements per 16-byte store
  for (int j0 = 0; j0 < col_ld; j0 += VE) {
    Vec<T, VE> o;
#pragma unroll
    for (int i = 0; i < VE; ++i) {
      const int j = j0 + i;
      float val = 0.f;
      if (j < J) {
        const int4 e = jtab[j];
        const int id = id0 + e.x, ih = ih0 + e.y, iw = iw0 + e.z;
        if (id >= 0 && id < p.iD && ih >= 0 && ih < p.iH && iw >= 0 && iw < p.iW)
          val = to_f<T>(xn[(((int64_t)id * p.iH + ih) * p.iW + iw) * p.x_ld + e.w]);
      }
      o.v[i] = val;
    }
    o.store(out + j0);
  }
}

namespace {
void fill(SmallCinParams& p, const b200seg_conv_desc* d) {
  p.n = d->n; p.cin = d->cin; p.cout = d->cout;
  p.iD = d->in_d; p.iH = d->in_h; p.iW = d->in_w; p.oD = d->out_d; p.oH = d->out_h; p.oW = d->out_w;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.sd = d->sd; p.sh = d->sh; p.sw = d->sw;
  p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.x_ld = d->x_ld; p.y_ld = d->y_ld; p.r_ld = d->r_ld;
  p.nvox = (int64_t)d->n * d->out_d * d->out_h * d->out_w;
}
}  // namespace

bool small_cin_supported(const b200seg_conv_desc* d) {
  const int J = d->kd * d->kh * d->kw * d->cin;
  return d->cin <= 4 && J * d->cout <= 2048 && !(d->flags & B200SEG_CONV_FORCE_GENERIC);
}

int small_cin_wgrad_blocks(const b200seg_conv_desc* d) {
  int64_t nvox = (int64_t)d->n * d->out_d * d->out_h * d->out_w;
  int64_t nb = cdiv64(nvox, 256);
  if (nb > 592) nb = 592;
  if (nb < 1) nb = 1;
  return (int)nb;
}

size_t small_cin_wgrad_workspace(const b200seg_conv_desc* d) {
  if (!small_cin_supported(d)) return 0;
  return (size_t)small_cin_wgrad_blocks(d) * d->kd * d->kh * d->kw * d->cin * d->cout * sizeof(float);
}

int launch_small_cin_fprop(const b200seg_conv_desc* d, const void* x, const void* w, const float* bias,
                           const void* res, void* y, cudaStream_t st) {
  SmallCinParams p;
  fill(p, d);
  const int J = d->kd * d->kh * d->kw * d->cin;
  dim3 grid((unsigned)cdiv64(p.nvox, 128), (unsigned)((d->cout + 15) / 16));
  size_t smem = (size_t)J * 16 * sizeof(float);
  if (d->dtype == B200SEG_BF16)
    conv_small_cin_fprop_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(
        p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (const __nv_bfloat16*)res, (__nv_bfloat16*)y);
  else
    conv_small_cin_fprop_kernel<float><<<grid, 128, smem, st>>>(p, (const float*)x, (const float*)w, bias,
                                                                (const float*)res, (float*)y);
  B200SEG_CHECK_LAUNCH("conv_small_cin_fprop");
  return B200SEG_OK;
}

int launch_small_cin_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw, float* partial,
                           cudaStream_t st) {
  SmallCinParams p;
  fill(p, d);
  const int taps = d->kd * d->kh * d->kw, J = taps * d->cin;
  const int nb = small_cin_wgrad_blocks(d);
  int64_t per = cdiv64(cdiv64(p.nvox, nb), 64) * 64;
  size_t smem = (size_t)(65 * J + 64 * d->cout) * sizeof(float) + (size_t)(64 * 4 + J * 4) * sizeof(int);
  const int nout_ = J * d->cout;
  if (J <= 32 && (d->cout == 16 || d->cout == 32 || d->cout == 64)) {
    const size_t esz = d->dtype == B200SEG_BF16 ? 2 : 4;
    const int dy_vec = (d->y_ld % 4 == 0) && ((uintptr_t)dy % (4 * esz) == 0);
    const int64_t per128 = cdiv64(cdiv64(p.nvox, nb), 128) * 128;
    // staging: X[32][129] + DY[128][cout] + voxel table; reused for the cross-team reduction [256/cout][32*cout]
    size_t stage = (size_t)(32 * 129 + 4 + 128 * d->cout) * sizeof(float) + 128 * sizeof(int4);
    size_t redb = (size_t)256 * 32 * sizeof(float);
    size_t smem_t = stage > redb ? stage : redb;
#define SMALL_CIN_WGRAD_TILE(TT, CC)                                                                                \
  do {                                                                                                              \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      cudaFuncSetAttribute(conv_small_cin_wgrad_tile_kernel<TT, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           64 * 1024);                                                                              \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    conv_small_cin_wgrad_tile_kernel<TT, CC><<<nb, 256, smem_t, st>>>(p, (const TT*)x, (const TT*)dy, partial,      \
                                                                      per128, dy_vec);                              \
  } while (0)
    if (d->dtype == B200SEG_BF16) {
      if (d->cout == 16) SMALL_CIN_WGRAD_TILE(__nv_bfloat16, 16);
      else if (d->cout == 32) SMALL_CIN_WGRAD_TILE(__nv_bfloat16, 32);
      else SMALL_CIN_WGRAD_TILE(__nv_bfloat16, 64);
    } else {
      if (d->cout == 16) SMALL_CIN_WGRAD_TILE(float, 16);
      else if (d->cout == 32) SMALL_CIN_WGRAD_TILE(float, 32);
      else SMALL_CIN_WGRAD_TILE(float, 64);
    }
#undef SMALL_CIN_WGRAD_TILE
    B200SEG_CHECK_LAUNCH("conv_small_cin_wgrad_tile");
    conv_small_cin_wgrad_final_kernel<<<(nout_ * 32 + 255) / 256, 256, 0, st>>>(partial, nb, taps, d->cin, d->cout, gw);
    B200SEG_CHECK_LAUNCH("conv_small_cin_wgrad_final");
    return B200SEG_OK;
  }
#define SMALL_CIN_WGRAD(TT, RR) \
  conv_small_cin_wgrad_kernel<TT, RR><<<nb, 256, smem, st>>>(p, (const TT*)x, (const TT*)dy, partial, per)
  if (d->dtype == B200SEG_BF16) {
    if (nout_ <= 512) SMALL_CIN_WGRAD(__nv_bfloat16, 2);
    else if (nout_ <= 1024) SMALL_CIN_WGRAD(__nv_bfloat16, 4);
    else SMALL_CIN_WGRAD(__nv_bfloat16, 8);
  } else {
    if (nout_ <= 512) SMALL_CIN_WGRAD(float, 2);
    else if (nout_ <= 1024) SMALL_CIN_WGRAD(float, 4);
    else SMALL_CIN_WGRAD(float, 8);
  }
#undef SMALL_CIN_WGRAD
  B200SEG_CHECK_LAUNCH("conv_small_cin_wgrad");
  const int nout = J * d->cout;
  conv_small_cin_wgrad_final_kernel<<<(nout * 32 + 255) / 256, 256, 0, st>>>(partial, nb, taps, d->cin, d->cout, gw);
  B200SEG_CHECK_LAUNCH("conv_small_cin_wgrad_final");
  return B200SEG_OK;
}

int launch_im2col(const b200seg_conv_desc* d, const void* x, void* col, int col_ld, cudaStream_t st) {
  SmallCinParams p;
  fill(p, d);
  const int J = d->kd * d->kh * d->kw * d->cin;
  const unsigned grid = (unsigned)cdiv64(p.nvox, 128);
  const size_t smem = (size_t)J * sizeof(int4);
  if (d->dtype == B200SEG_BF16)
    im2col_small_cin_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(p, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, col_ld);
  else
    im2col_small_cin_kernel<float><<<grid, 128, smem, st>>>(p, (const float*)x, (float*)col, col_ld);
  B200SEG_CHECK_LAUNCH("im2col_small_cin");
  return B200SEG_OK;
}

}  // namespace b200seg

// CUDA-core implicit-GEMM convolution family (any channel count, fp32 or bf16 storage, fp32
// accumulation).  This is the fp32 "check mode" path and the path for layer shapes the tcgen05
// kernels do not take (Cin = 1, odd channel counts).
//
// Everything is expressed as ONE gather-GEMM:
//     dst[n, o, t] = sum_{tap} sum_{s} src[n, g(o, tap), s] * Wp[tap][s][t]
// with  g(o,k) = o*stride - pad + k            ("forward" gather: conv fprop, convtr dgrad)
//  or   g(o,k) = (o + pad - k) / stride        ("transposed" gather: conv dgrad, convtr fprop;
//                                               only where divisible and in range).
// For the transposed gather with stride 2 the destination voxels are enumerated parity-class
// major, so that every voxel of a tile shares the same set of contributing taps and the
// non-contributing taps are skipped block-wide (no multiplies with inserted zeros).
//
// and one weight-gradient GEMM (always in forward-gather orientation):
//     G[tap][a][b] = sum_{n,o} S[n, o*stride - pad + tap, a] * T[n, o, b]
#include "common.cuh"
#include "kernels.h"

namespace b200seg {

// ------------------------------------------------------------------------------------------
// gather-GEMM
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* out);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* out) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x;
    out[2 * i + 1] = f.y;
  }
}

template <typename T, int BN>
__global__ void __launch_bounds__(128)
conv_gather_kernel(GatherParams p, const T* __restrict__ src, const T* __restrict__ w,
                   const float* __restrict__ bias, const T* __restrict__ res, T* dst) {
  constexpr int BM = 64, BK = 16, TN = BN / 8;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int s_n[BM], s_d[BM], s_h[BM], s_w[BM];

  const int tid = threadIdx.x;
  const int cls = blockIdx.x / p.tiles_per_class;
  const int tile = blockIdx.x % p.tiles_per_class;
  const int cls_w = cls % p.csw, cls_h = (cls / p.csw) % p.csh, cls_d = cls / (p.csw * p.csh);

  if (tid < BM) {
    int64_t j = (int64_t)tile * BM + tid;
    if (j < p.per_class) {
      int jw = (int)(j % p.cW); j /= p.cW;
      int jh = (int)(j % p.cH); j /= p.cH;
      int jd = (int)(j % p.cD); j /= p.cD;
      s_n[tid] = (int)j;
      s_d[tid] = jd * p.csd + cls_d;
      s_h[tid] = jh * p.csh + cls_h;
      s_w[tid] = jw * p.csw + cls_w;
    } else {
      s_n[tid] = -1; s_d[tid] = 0; s_h[tid] = 0; s_w[tid] = 0;
    }
  }
  __syncthreads();

  const int n0 = blockIdx.y * BN;
  const int tx = tid % 8, ty = tid / 8;
  const int lv = tid >> 1, lhalf = (tid & 1) * 8;  // loader: voxel and channel half
  const int ln = s_n[lv], ld_ = s_d[lv], lh = s_h[lv], lw = s_w[lv];

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int kd = 0; kd < p.kd; ++kd) {
    if (p.transposed && ((cls_d + p.pd - kd) % p.sd) != 0) continue;  // block-uniform
    for (int kh = 0; kh < p.kh; ++kh) {
      if (p.transposed && ((cls_h + p.ph - kh) % p.sh) != 0) continue;
      for (int kw = 0; kw < p.kw; ++kw) {
        if (p.transposed && ((cls_w + p.pw - kw) % p.sw) != 0) continue;
        const int tap = (kd * p.kh + kh) * p.kw + kw;
        // source voxel of this thread's loader voxel
        bool valid = ln >= 0;
        int id, ih, iw;
        if (!p.transposed) {
          id = ld_ * p.sd - p.pd + kd;
          ih = lh * p.sh - p.ph + kh;
          iw = lw * p.sw - p.pw + kw;
        } else {
          int td = ld_ + p.pd - kd, th = lh + p.ph - kh, tw = lw + p.pw - kw;
          valid = valid && td >= 0 && th >= 0 && tw >= 0;
          id = td / p.sd; ih = th / p.sh; iw = tw / p.sw;  // divisible by construction
        }
        valid = valid && id >= 0 && id < p.sD && ih >= 0 && ih < p.sH && iw >= 0 && iw < p.sW;
        const T* sp = nullptr;
        if (valid)
          sp = src + ((((int64_t)ln * p.sD + id) * p.sH + ih) * p.sW + iw) * (int64_t)p.src_ld;
        const T* wt = w + (int64_t)tap * p.src_c * p.dst_c;

        for (int c0 = 0; c0 < p.src_c; c0 += BK) {
          // ---- A tile: As[k][voxel]
          float av[8];
          if (valid && p.src_vec && (c0 + lhalf + 8 <= p.src_c)) {
            load8<T>(sp + c0 + lhalf, av);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              int c = c0 + lhalf + i;
              av[i] = (valid && c < p.src_c) ? to_f<T>(sp[c]) : 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) As[lhalf + i][lv] = av[i];
          // ---- B tile: Bs[k][n]
#pragma unroll
          for (int r = 0; r < TN; ++r) {
            int idx = tid + r * 128;
            int k = idx / BN, nn = idx % BN;
            int c = c0 + k, t = n0 + nn;
            Bs[k][nn] = (c < p.src_c && t < p.dst_c) ? to_f<T>(wt[(int64_t)c * p.dst_c + t]) : 0.f;
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          }
          __syncthreads();
        }
      }
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = ty * 4 + i;
    int n = s_n[m];
    if (n < 0) continue;
    int64_t lin = (((int64_t)n * p.dD + s_d[m]) * p.dH + s_h[m]) * p.dW + s_w[m];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int t = n0 + tx * TN + j;
      if (t >= p.dst_c) continue;
      float v = acc[i][j];
      if (bias) v += bias[t];
      if (res) v += to_f<T>(res[lin * p.res_ld + t]);
      T* o = dst + lin * p.dst_ld + t;
      if (p.accumulate) v += to_f<T>(*o);
      *o = from_f<T>(v);
    }
  }
}

template <typename T>
static int launch_gather_t(const GatherParams& p, const void* src, const void* w, const float* bias,
                           const void* res, void* dst, cudaStream_t st) {
  int nclass = p.csd * p.csh * p.csw;
  int64_t gx = (int64_t)nclass * p.tiles_per_class;
  if (gx > 0x7fffffffLL) {
    set_error("conv_gather: grid too large");
    return B200SEG_ERR_ARG;
  }
  if (p.dst_c <= 16) {
    dim3 grid((unsigned)gx, (p.dst_c + 15) / 16);
    conv_gather_kernel<T, 16><<<grid, 128, 0, st>>>(p, (const T*)src, (const T*)w, bias, (const T*)res, (T*)dst);
  } else if (p.dst_c <= 32) {
    dim3 grid((unsigned)gx, (p.dst_c + 31) / 32);
    conv_gather_kernel<T, 32><<<grid, 128, 0, st>>>(p, (const T*)src, (const T*)w, bias, (const T*)res, (T*)dst);
  } else {
    dim3 grid((unsigned)gx, (p.dst_c + 63) / 64);
    conv_gather_kernel<T, 64><<<grid, 128, 0, st>>>(p, (const T*)src, (const T*)w, bias, (const T*)res, (T*)dst);
  }
  B200SEG_CHECK_LAUNCH("conv_gather");
  return B200SEG_OK;
}

int launch_gather(GatherParams p, int dtype, const void* src, const void* w, const float* bias,
                  const void* res, void* dst, cudaStream_t st) {
  // class decomposition of the destination grid
  if (p.transposed) {
    if ((p.dD % p.sd) || (p.dH % p.sh) || (p.dW % p.sw)) {
      set_error("transposed gather needs destination extents divisible by the stride");
      return B200SEG_ERR_UNSUPPORTED;
    }
    p.csd = p.sd; p.csh = p.sh; p.csw = p.sw;
  } else {
    p.csd = p.csh = p.csw = 1;
  }
  p.cD = p.dD / p.csd; p.cH = p.dH / p.csh; p.cW = p.dW / p.csw;
  p.per_class = (int64_t)p.n * p.cD * p.cH * p.cW;
  p.tiles_per_class = (int)cdiv64(p.per_class, 64);
  size_t esz = dtype == B200SEG_BF16 ? 2 : 4;
  p.src_vec = (p.src_ld % 8 == 0) && (p.src_c % 8 == 0) && (((uintptr_t)src) % 16 == 0) &&
              ((p.src_ld * esz) % 16 == 0);
  if (p.per_class == 0) return B200SEG_OK;
  if (dtype == B200SEG_BF16) return launch_gather_t<__nv_bfloat16>(p, src, w, bias, res, dst, st);
  return launch_gather_t<float>(p, src, w, bias, res, dst, st);
}

// ------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(WgradParams p, const T* __restrict__ S, const T* __restrict__ Tt,
                  float* __restrict__ partial) {
  constexpr int BA = 32, BB = 32, BV = 32;
  __shared__ float Ss[BV][BA + 1];
  __shared__ float Ts[BV][BB + 1];

  const int tid = threadIdx.x;
  int bx = blockIdx.x;
  const int tb_i = bx % p.tiles_b; bx /= p.tiles_b;
  const int ta_i = bx % p.tiles_a; bx /= p.tiles_a;
  const int tap = bx;
  const int kw = tap % p.kw, kh = (tap / p.kw) % p.kh, kd = tap / (p.kw * p.kh);
  const int a0 = ta_i * BA, b0 = tb_i * BB;
  const int split = blockIdx.y;
  const int64_t v_begin = (int64_t)split * p.vox_per_split;
  const int64_t v_end = min(v_begin + p.vox_per_split, p.vox_total);

  const int lv = tid >> 3, lc = (tid & 7) * 4;  // loader: voxel in chunk, first of 4 channels
  const int ca = tid >> 4, cb = tid & 15;       // compute: rows ca*2.., cols cb*2..
  // a weight gradient sums up to millions of voxel products: a plain fp32 running sum is off by a few 1e-4 at
  // 2 x 128^3 (r2: 1.6e-4 against a float64 reference).  Products of one 32-voxel chunk are summed in fp32, the
  // chunks in double (4 DADD per thread and chunk: nothing next to the 128 FMAs).
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};

  for (int64_t v0 = v_begin; v0 < v_end; v0 += BV) {
    int64_t o = v0 + lv;
    bool in = o < v_end;
    float tv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
    if (in) {
      int64_t r = o;
      int ow = (int)(r % p.tW); r /= p.tW;
      int oh = (int)(r % p.tH); r /= p.tH;
      int od = (int)(r % p.tD); r /= p.tD;
      int n = (int)r;
      const T* tp = Tt + o * (int64_t)p.t_ld;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int c = b0 + lc + i;
        if (c < p.b_c) tv[i] = to_f<T>(tp[c]);
      }
      int id = od * p.sd - p.pd + kd, ih = oh * p.sh - p.ph + kh, iw = ow * p.sw - p.pw + kw;
      if (id >= 0 && id < p.sD && ih >= 0 && ih < p.sH && iw >= 0 && iw < p.sW) {
        const T* sp = S + ((((int64_t)n * p.sD + id) * p.sH + ih) * p.sW + iw) * (int64_t)p.s_ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int c = a0 + lc + i;
          if (c < p.a_c) sv[i] = to_f<T>(sp[c]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Ts[lv][lc + i] = tv[i];
      Ss[lv][lc + i] = sv[i];
    }
    __syncthreads();
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll
    for (int v = 0; v < BV; ++v) {
      float s0 = Ss[v][ca * 2], s1 = Ss[v][ca * 2 + 1];
      float t0 = Ts[v][cb * 2], t1 = Ts[v][cb * 2 + 1];
      c00 = fmaf(s0, t0, c00);
      c01 = fmaf(s0, t1, c01);
      c10 = fmaf(s1, t0, c10);
      c11 = fmaf(s1, t1, c11);
    }
    acc[0][0] += (double)c00;
    acc[0][1] += (double)c01;
    acc[1][0] += (double)c10;
    acc[1][1] += (double)c11;
    __syncthreads();
  }
  float* out = partial + ((int64_t)split * p.taps + tap) * p.a_c * p.b_c;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int a = a0 + ca * 2 + i, b = b0 + cb * 2 + j;
      if (a < p.a_c && b < p.b_c) out[(int64_t)a * p.b_c + b] = (float)acc[i][j];
    }
}

// gw[b][a][tap] = sum_s partial[s][tap][a][b]   (fixed order => deterministic)
__global__ void wgrad_reduce_unpack_kernel(const float* __restrict__ partial, float* __restrict__ gw,
                                           int splits, int taps, int a_c, int b_c) {
  int64_t total = (int64_t)taps * a_c * b_c;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int b = (int)(idx % b_c);
  int64_t r = idx / b_c;
  int a = (int)(r % a_c);
  int tap = (int)(r / a_c);
  double s = 0.0;
  for (int i = 0; i < splits; ++i) s += (double)partial[(int64_t)i * total + idx];
  gw[((int64_t)b * a_c + a) * taps + tap] = (float)s;
}

void wgrad_plan(WgradParams& p) {
  p.taps = p.kd * p.kh * p.kw;
  p.tiles_a = (p.a_c + 31) / 32;
  p.tiles_b = (p.b_c + 31) / 32;
  p.vox_total = (int64_t)p.n * p.tD * p.tH * p.tW;
  int64_t base = (int64_t)p.taps * p.tiles_a * p.tiles_b;
  int64_t want = cdiv64(148 * 8, base);
  int64_t max_splits = cdiv64(p.vox_total, 512);
  int64_t splits = want < 1 ? 1 : want;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 4096) splits = 4096;
  p.vox_per_split = cdiv64(cdiv64(p.vox_total, splits), 32) * 32;
  if (p.vox_per_split < 32) p.vox_per_split = 32;
  p.splits = (int)cdiv64(p.vox_total, p.vox_per_split);
  if (p.splits < 1) p.splits = 1;
}

size_t wgrad_partial_bytes(const WgradParams& p) {
  return (size_t)p.splits * p.taps * p.a_c * p.b_c * sizeof(float);
}

int launch_wgrad(const WgradParams& p, int dtype, const void* S, const void* T, float* gw,
                 float* partial, cudaStream_t st) {
  dim3 grid((unsigned)(p.taps * p.tiles_a * p.tiles_b), (unsigned)p.splits);
  if (dtype == B200SEG_BF16)
    conv_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)S,
                                                           (const __nv_bfloat16*)T, partial);
  else
    conv_wgrad_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)S, (const float*)T, partial);
  B200SEG_CHECK_LAUNCH("conv_wgrad");
  int64_t total = (int64_t)p.taps * p.a_c * p.b_c;
  wgrad_reduce_unpack_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(partial, gw, p.splits,
                                                                         p.taps, p.a_c, p.b_c);
  B200SEG_CHECK_LAUNCH("wgrad_reduce_unpack");
  return B200SEG_OK;
}

// ------------------------------------------------------------------------------------------
// weight packing:  packed[tap][s][t]  from the fp32 PyTorch parameter
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ packed, int taps,
                                   int src_c, int dst_c, int cin, int cout, int kind) {
  int64_t total = (int64_t)taps * src_c * dst_c;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int t = (int)(idx % dst_c);
  int64_t r = idx / dst_c;
  int s = (int)(r % src_c);
  int tap = (int)(r / src_c);
  int64_t wi;
  switch (kind) {
    case B200SEG_W_CONV_FPROP:   wi = ((int64_t)t * cin + s) * taps + tap; break;   // w[co=t][ci=s]
    case B200SEG_W_CONV_DGRAD:   wi = ((int64_t)s * cin + t) * taps + tap; break;   // w[co=s][ci=t]
    case B200SEG_W_CONVTR_FPROP: wi = ((int64_t)s * cout + t) * taps + tap; break;  // w[ci=s][co=t]
    default:                     wi = ((int64_t)t * cout + s) * taps + tap; break;  // w[ci=t][co=s]
  }
  packed[idx] = from_f<T>(w[wi]);
}

int launch_pack_weight(int dtype, int kind, const float* w, void* packed, int taps, int cin,
                       int cout, cudaStream_t st) {
  bool src_is_cin = (kind == B200SEG_W_CONV_FPROP || kind == B200SEG_W_CONVTR_FPROP);
  int src_c = src_is_cin ? cin : cout, dst_c = src_is_cin ? cout : cin;
  int64_t total = (int64_t)taps * cin * cout;
  unsigned blocks = (unsigned)cdiv64(total, 256);
  if (dtype == B200SEG_BF16)
    pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)packed, taps, src_c, dst_c, cin, cout, kind);
  else
    pack_weight_kernel<float><<<blocks, 256, 0, st>>>(w, (float*)packed, taps, src_c, dst_c, cin, cout, kind);
  B200SEG_CHECK_LAUNCH("pack_weight");
  return B200SEG_OK;
}

}  // namespace b200seg

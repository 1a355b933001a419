// Sliding-window tcgen05 kernels for the high-resolution ConvTranspose (k3, s2, p1, op1) layers of
// the decoder (32 -> 10 at full resolution, 64 -> 16 at 1/2): the up-sampling layers whose output
// is 8x the input and whose arithmetic intensity is far too low for one-tile-per-CTA streaming.
//
// Output voxel o = 2 i - 1 + k per dimension: an even output (o = 2 j) takes tap k = 1 from input j;
// an odd output (o = 2 j + 1) takes tap k = 2 from input j and tap k = 0 from input j + 1.  So the 8
// output parity classes of an input tile need the tile itself and its +1 shifted copies in d, h, w.
//
// fprop: a CTA owns a column of the INPUT volume, 16 lines x 8 voxels in (h, w), and sweeps along d.
//   Per input slab two w-shifted copies of the 17-line halo tile are TMA-loaded into a ring (box rows
//   of exactly 8 voxels = one swizzle atom: the h shift is a descriptor offset, the d shift the next
//   ring slot, the w shift the copy).  The 8 classes are 8 x 16 TMEM columns; an input tile with
//   shift (sd, sh, sw) feeds every class whose parity is odd where the shift is 1, so its MMAs are
//   folded along N over runs of adjacent classes (N = 128, 64, 32, 16): 14 tcgen05.mma per K slice
//   instead of 27, i.e. half the shared-memory operand reads of the A tiles.  Two accumulator
//   buffers overlap the epilogue (bias, InstanceNorm partial statistics, bf16, 8 x 32-byte stores
//   per input voxel) with the MMAs of the next slab.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace b200seg {

using bf16 = __nv_bfloat16;

int tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, int row_bytes);

namespace {
constexpr int TH = 16;   // input lines per tile
constexpr int TWV = 8;   // input voxels per line (one swizzle atom of rows)
inline int round16(int c) { return (c + 15) / 16 * 16; }

// runs of adjacent output classes (class = pd*4 + ph*2 + pw) fed by the input tile of one shift
struct Run { int sd, sh, sw, cls0, ncls, slot0; };
__host__ __device__ constexpr Run run_at(int r) {
  constexpr Run runs[14] = {
      {0, 0, 0, 0, 8, 0},  {1, 0, 0, 4, 4, 8},  {0, 1, 0, 2, 2, 12}, {0, 1, 0, 6, 2, 14}, {0, 0, 1, 1, 1, 16},
      {0, 0, 1, 3, 1, 17}, {0, 0, 1, 5, 1, 18}, {0, 0, 1, 7, 1, 19}, {1, 1, 0, 6, 2, 20}, {1, 0, 1, 5, 1, 22},
      {1, 0, 1, 7, 1, 23}, {0, 1, 1, 3, 1, 24}, {0, 1, 1, 7, 1, 25}, {1, 1, 1, 7, 1, 26}};
  return runs[r];
}
// kernel tap (0..2) along one dimension for an input shift s and an output parity p
__host__ __device__ constexpr int tap_of(int s, int p) { return s ? 0 : (p ? 2 : 1); }
}  // namespace

struct alignas(64) TcConvTrFpropParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int n, D, H, W;  // input extent
  int tilesH, tilesW, dseg, nseg;
  int cout, dst_ld;
  int accumulate;  // dst += result (gradient fan-in of a strided conv's dgrad)
  const float* bias;
  bf16* dst;
  float* stats;  // optional [CTA][cout][2]
};

template <int KC>
__global__ void __launch_bounds__(192)
tc_convtr_fprop_kernel(const __grid_constant__ TcConvTrFpropParams p) {
  constexpr int BN = 16;
  constexpr int RING = KC == 32 ? 4 : 3;
  constexpr int PITCH = KC * 2;
  constexpr int LINE = TWV * PITCH;
  constexpr int COPY_BYTES = (TH + 1) * LINE;
  constexpr int SLAB_BYTES = 2 * COPY_BYTES;
  constexpr int WT_BYTES = BN * PITCH;
  constexpr int W_BYTES = (27 * WT_BYTES + 1023) / 1024 * 1024;
  constexpr uint32_t TMEM_COLS = 256;  // 2 buffers x 8 classes x 16 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RING * SLAB_BYTES);
  uint64_t* empty = full + RING;
  uint64_t* acc_full = empty + RING;   // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint64_t* wbar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  int bx = blockIdx.x;
  const int seg = bx % p.nseg; bx /= p.nseg;
  const int tw_i = bx % p.tilesW; bx /= p.tilesW;
  const int th_i = bx % p.tilesH; bx /= p.tilesH;
  const int n = bx;
  const int h0 = th_i * TH, w0 = tw_i * TWV;
  const int d_begin = seg * p.dseg;
  const int nd = min(p.dseg, p.D - d_begin);  // input slabs (= steps) of this CTA

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_init(wbar, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmA);
    tc::prefetch_tmap(&p.tmB);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // weights: 27 (class, tap) tiles in run order
      tc::mbar_expect_tx(wbar, 27 * WT_BYTES);
#pragma unroll 1
      for (int r = 0; r < 14; ++r) {
        const Run ru = run_at(r);
        for (int i = 0; i < ru.ncls; ++i) {
          const int c = ru.cls0 + i;
          const int tap = (tap_of(ru.sd, (c >> 2) & 1) * 3 + tap_of(ru.sh, (c >> 1) & 1)) * 3 + tap_of(ru.sw, c & 1);
          tc::tma_load_2d(wsm + (ru.slot0 + i) * WT_BYTES, &p.tmB, wbar, 0, tap * BN);
        }
      }
      for (int s = 0; s <= nd; ++s) {
        const int slot = s % RING;
        tc::mbar_wait(&empty[slot], (((uint32_t)(s / RING)) & 1u) ^ 1u);
        uint8_t* dst = ring + slot * SLAB_BYTES;
        tc::mbar_expect_tx(&full[slot], SLAB_BYTES);
        tc::tma_load_5d(dst, &p.tmA, &full[slot], 0, w0, h0, d_begin + s, n);
        tc::tma_load_5d(dst + COPY_BYTES, &p.tmA, &full[slot], 0, w0 + 1, h0, d_begin + s, n);
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform issue loop, one elected lane issues
      const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform register: see tc::warp_uniform
      constexpr uint64_t layout = tc::layout_for_row_bytes(PITCH);
      const uint32_t w_addr = tc::smem_u32(wsm), r_addr = tc::smem_u32(ring);
      const uint64_t tmpl = tc::make_smem_desc(0, 16, 8 * PITCH, layout);
      const uint64_t w_desc = tmpl + (w_addr >> 4);
      tc::mbar_wait(wbar, 0);
      int waited = 0;
      for (int j = 0; j < nd; ++j) {
        const int buf = j & 1;
        tc::mbar_wait(&acc_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);
        while (waited <= j + 1) {
          tc::mbar_wait(&full[waited % RING], ((uint32_t)(waited / RING)) & 1u);
          ++waited;
        }
        tc::tc_fence_after();
        const uint64_t slab0 = tmpl + ((r_addr + (j % RING) * SLAB_BYTES) >> 4);
        const uint64_t slab1 = tmpl + ((r_addr + ((j + 1) % RING) * SLAB_BYTES) >> 4);
        const uint32_t acc = tmem_acc + buf * 128;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
          for (int r = 0; r < 14; ++r) {
            constexpr uint32_t idesc8 = tc::make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t idesc4 = tc::make_idesc_bf16(128, 64, false, false);
            constexpr uint32_t idesc2 = tc::make_idesc_bf16(128, 32, false, false);
            constexpr uint32_t idesc1 = tc::make_idesc_bf16(128, 16, false, false);
            const Run ru = run_at(r);
            const uint64_t a = (ru.sd ? slab1 : slab0) + ((ru.sw * COPY_BYTES + ru.sh * LINE) >> 4) + 2 * k;
            const uint64_t b = w_desc + ((ru.slot0 * WT_BYTES) >> 4) + 2 * k;
            const uint32_t idesc = ru.ncls == 8 ? idesc8 : (ru.ncls == 4 ? idesc4 : (ru.ncls == 2 ? idesc2 : idesc1));
            tc::umma_bf16_warp(acc + ru.cls0 * BN, a, b, idesc, (r == 0 && k == 0) ? 0u : 1u);
          }
        }
        tc::umma_commit_warp(&acc_full[buf]);
        tc::umma_commit_warp(&empty[j % RING]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lh = row / TWV, lw = row % TWV;
    const int ih = h0 + lh, iw = w0 + lw;
    const bool valid = ih < p.H && iw < p.W;
    const int oD = 2 * p.D, oH = 2 * p.H, oW = 2 * p.W;
    float bias[BN];
#pragma unroll
    for (int c = 0; c < BN; ++c) bias[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    float ssum[BN], ssq[BN];
#pragma unroll
    for (int c = 0; c < BN; ++c) ssum[c] = ssq[c] = 0.f;
    for (int j = 0; j < nd; ++j) {
      const int buf = j & 1;
      tc::mbar_wait(&acc_full[buf], ((uint32_t)j >> 1) & 1u);
      tc::tc_fence_after();
      const int id = d_begin + j;
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {  // (pd, ph); the two pw classes are adjacent output voxels
        uint32_t v0[16], v1[16];
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + buf * 128 + cp * 32;
        tc::tmem_ld16(taddr, v0);
        tc::tmem_ld16(taddr + 16, v1);
        tc::tmem_ld_wait();
        if (valid) {
          const int od = 2 * id + (cp >> 1), oh = 2 * ih + (cp & 1);
          bf16* op = p.dst + ((((int64_t)n * oD + od) * oH + oh) * oW + 2 * iw) * p.dst_ld;
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(pw ? v1[i] : v0[i]) + bias[i];
            uint4* o4 = reinterpret_cast<uint4*>(op + (int64_t)pw * p.dst_ld);
            if (p.accumulate) {
              const uint4 r0 = o4[0], r1 = o4[1];
              const __nv_bfloat162* g0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
              const __nv_bfloat162* g1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 a = __bfloat1622float2(g0[i]), b = __bfloat1622float2(g1[i]);
                f[2 * i] += a.x; f[2 * i + 1] += a.y;
                f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
              }
            }
            if (p.stats) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                ssum[i] += f[i];
                ssq[i] = fmaf(f[i], f[i], ssq[i]);
              }
            }
            uint4 o0, o1;
            __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
            __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              q0[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              q1[i] = __floats2bfloat162_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
            }
            o4[0] = o0;
            o4[1] = o1;
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
    if (p.stats) {
      float* sred = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][BN][2]
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        const float a = warp_sum(ssum[c]), b = warp_sum(ssq[c]);
        if (lane == 0) {
          sred[(q * BN + c) * 2] = a;
          sred[(q * BN + c) * 2 + 1] = b;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (p.stats && threadIdx.x < 2 * BN) {
    const float* sred = reinterpret_cast<const float*>(tmem_slot + 4);
    const int c = threadIdx.x >> 1, m = threadIdx.x & 1;
    if (c < p.cout)
      p.stats[((int64_t)blockIdx.x * p.cout + c) * 2 + m] =
          sred[(0 * BN + c) * 2 + m] + sred[(1 * BN + c) * 2 + m] + sred[(2 * BN + c) * 2 + m] +
          sred[(3 * BN + c) * 2 + m];
  }
  if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
// dgrad / wgrad: dy (the 2x-resolution tensor) is read as 8 parity-decimated sub-volumes (tensor maps
// with doubled strides, one per (pd, ph, pw)).  Per dimension the input voxel j pairs with
//   tap 1 -> even-class element j;  tap 0 -> odd-class element j - 1;  tap 2 -> odd-class element j.
// A "unit" in the shared-memory ring holds, for one step of the sweep along d, the odd-d class slab
// (taps kd = 0 of the next step and kd = 2 of this one) and the even-d class slab (kd = 1).  Each
// d-class slab = section ph=0 (16 lines) and section ph=1 (17 lines, from line h0-1), each with three
// copies in w: even class (kw = 1), odd class from w0-1 (kw = 0), odd class from w0 (kw = 2).
// So every tap's 128 x 16 tile is a (unit, section, line offset, copy) descriptor offset.
namespace {
constexpr int DY_PITCH = 32;                         // 16 channels bf16
constexpr int DY_LINE = TWV * DY_PITCH;              // 256 B = one 32-byte-swizzle atom
constexpr int DY_SEC1 = 3 * TH * DY_LINE;            // offset of the ph=1 section inside a d-class slab
constexpr int DY_DSLAB = (3 * TH + 3 * (TH + 1)) * DY_LINE;   // 99 lines
constexpr int DY_UNIT = 2 * DY_DSLAB;                // [odd-d slab | even-d slab]

// byte offset of tap (kh, kw)'s tile inside a d-class slab
__host__ __device__ constexpr int dy_tap_offset(int kh, int kw) {
  const int copy = kw == 1 ? 0 : (kw == 0 ? 1 : 2);
  return kh == 1 ? copy * TH * DY_LINE : DY_SEC1 + copy * (TH + 1) * DY_LINE + (kh == 2 ? DY_LINE : 0);
}

// TMA loads of one d-class slab (parity pd, class element e along d) into `dst`; 6 box loads
__device__ __forceinline__ void load_dy_dslab(uint8_t* dst, const CUtensorMap* maps, int pd, uint64_t* bar, int w0,
                                              int h0, int e, int n) {
  const CUtensorMap* m = maps + pd * 4;
  tc::tma_load_5d(dst, m + 0, bar, 0, w0, h0, e, n);
  tc::tma_load_5d(dst + TH * DY_LINE, m + 1, bar, 0, w0 - 1, h0, e, n);
  tc::tma_load_5d(dst + 2 * TH * DY_LINE, m + 1, bar, 0, w0, h0, e, n);
  tc::tma_load_5d(dst + DY_SEC1, m + 2, bar, 0, w0, h0 - 1, e, n);
  tc::tma_load_5d(dst + DY_SEC1 + (TH + 1) * DY_LINE, m + 3, bar, 0, w0 - 1, h0 - 1, e, n);
  tc::tma_load_5d(dst + DY_SEC1 + 2 * (TH + 1) * DY_LINE, m + 3, bar, 0, w0, h0 - 1, e, n);
}
}  // namespace

struct alignas(64) TcConvTrDgradParams {
  CUtensorMap tmA[8];  // dy classes, index pd*4 + ph*2 + pw
  CUtensorMap tmB;     // weights [27][CI][16]
  int n, D, H, W;      // input (dx) extent
  int tilesH, tilesW, dseg, nseg;
  int dst_ld;
  int cout;            // real destination channels (bias / statistics)
  const float* bias;   // optional (strided-conv fprop)
  bf16* dst;
  float* stats;        // optional [CTA][cout][2]
};

// dx[j, ci] = sum_taps dy[2j - 1 + k, co] * W[ci, co, k]:  M = 128 input voxels, K = 16 (co), N = CI
template <int CI>
__global__ void __launch_bounds__(192)
tc_convtr_dgrad_kernel(const __grid_constant__ TcConvTrDgradParams p) {
  constexpr int RING = 3;
  constexpr int WT_BYTES = CI * DY_PITCH;
  constexpr int W_BYTES = (27 * WT_BYTES + 1023) / 1024 * 1024;
  constexpr uint32_t TMEM_COLS = 2 * CI < 32 ? 32 : 2 * CI;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RING * DY_UNIT);
  uint64_t* empty = full + RING;
  uint64_t* acc_full = empty + RING;   // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint64_t* wbar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  int bx = blockIdx.x;
  const int seg = bx % p.nseg; bx /= p.nseg;
  const int tw_i = bx % p.tilesW; bx /= p.tilesW;
  const int th_i = bx % p.tilesH; bx /= p.tilesH;
  const int n = bx;
  const int h0 = th_i * TH, w0 = tw_i * TWV;
  const int d_begin = seg * p.dseg;
  const int nd = min(p.dseg, p.D - d_begin);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_init(wbar, 1);
    tc::fence_barrier_init();
    for (int i = 0; i < 8; ++i) tc::prefetch_tmap(&p.tmA[i]);
    tc::prefetch_tmap(&p.tmB);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(wbar, 27 * WT_BYTES);
      for (int t = 0; t < 27; ++t) tc::tma_load_2d(wsm + t * WT_BYTES, &p.tmB, wbar, 0, t * CI);
      // unit s: odd-d slab of class element d_begin-1+s, even-d slab of element d_begin+s-1 (s >= 1)
      for (int s = 0; s <= nd; ++s) {
        const int slot = s % RING;
        tc::mbar_wait(&empty[slot], (((uint32_t)(s / RING)) & 1u) ^ 1u);
        uint8_t* dst = ring + slot * DY_UNIT;
        tc::mbar_expect_tx(&full[slot], s ? DY_UNIT : DY_DSLAB);
        load_dy_dslab(dst, p.tmA, 1, &full[slot], w0, h0, d_begin - 1 + s, n);
        if (s) load_dy_dslab(dst + DY_DSLAB, p.tmA, 0, &full[slot], w0, h0, d_begin + s - 1, n);
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform issue loop, one elected lane issues
      const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform register: see tc::warp_uniform
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, CI, false, false);
      const uint32_t w_addr = tc::smem_u32(wsm), r_addr = tc::smem_u32(ring);
      const uint64_t tmpl = tc::make_smem_desc(0, 16, 8 * DY_PITCH, tc::LAYOUT_SW32);
      const uint64_t w_desc = tmpl + (w_addr >> 4);
      tc::mbar_wait(wbar, 0);
      int waited = 0;
      for (int j = 0; j < nd; ++j) {
        const int buf = j & 1;
        tc::mbar_wait(&acc_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);
        while (waited <= j + 1) {
          tc::mbar_wait(&full[waited % RING], ((uint32_t)(waited / RING)) & 1u);
          ++waited;
        }
        tc::tc_fence_after();
        const uint64_t u0 = tmpl + ((r_addr + (j % RING) * DY_UNIT) >> 4);        // odd-d slab: kd = 0
        const uint64_t u1 = tmpl + ((r_addr + ((j + 1) % RING) * DY_UNIT) >> 4);  // odd: kd = 2, even (+DSLAB): kd = 1
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const uint64_t part = kd == 0 ? u0 : (kd == 1 ? u1 + (DY_DSLAB >> 4) : u1);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int tap = (kd * 3 + kh) * 3 + kw;
              tc::umma_bf16_warp(tmem_acc + buf * CI, part + (dy_tap_offset(kh, kw) >> 4), w_desc + ((tap * WT_BYTES) >> 4),
                            idesc, tap ? 1u : 0u);
            }
        }
        tc::umma_commit_warp(&acc_full[buf]);
        tc::umma_commit_warp(&empty[j % RING]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ih = h0 + row / TWV, iw = w0 + row % TWV;
    const bool valid = ih < p.H && iw < p.W;
    float bias[CI], ssum[CI], ssq[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) {
      bias[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
      ssum[c] = ssq[c] = 0.f;
    }
    for (int j = 0; j < nd; ++j) {
      const int buf = j & 1;
      tc::mbar_wait(&acc_full[buf], ((uint32_t)j >> 1) & 1u);
      tc::tc_fence_after();
      const int64_t lin = (((int64_t)n * p.D + d_begin + j) * p.H + ih) * p.W + iw;
#pragma unroll
      for (int ch = 0; ch < CI / 16; ++ch) {
        uint32_t v[16];
        tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + buf * CI + ch * 16, v);
        tc::tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + bias[ch * 16 + i];
          if (p.stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              ssum[ch * 16 + i] += f[i];
              ssq[ch * 16 + i] = fmaf(f[i], f[i], ssq[ch * 16 + i]);
            }
          }
          uint4 o0, o1;
          __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
          __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            q0[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            q1[i] = __floats2bfloat162_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
          }
          uint4* op = reinterpret_cast<uint4*>(p.dst + lin * p.dst_ld + ch * 16);
          op[0] = o0;
          op[1] = o1;
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
    if (p.stats) {
      float* sred = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][CI][2]
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const float a = warp_sum(ssum[c]), b = warp_sum(ssq[c]);
        if (lane == 0) {
          sred[(q * CI + c) * 2] = a;
          sred[(q * CI + c) * 2 + 1] = b;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (p.stats && threadIdx.x < 2 * CI) {
    const float* sred = reinterpret_cast<const float*>(tmem_slot + 4);
    const int c = threadIdx.x >> 1, m = threadIdx.x & 1;
    if (c < p.cout)
      p.stats[((int64_t)blockIdx.x * p.cout + c) * 2 + m] =
          sred[(0 * CI + c) * 2 + m] + sred[(1 * CI + c) * 2 + m] + sred[(2 * CI + c) * 2 + m] +
          sred[(3 * CI + c) * 2 + m];
  }
  if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
// wgrad: G[tap][co][ci] = sum_j dy[2j - 1 + k, co] * x[j, ci].  Voxels are the K dimension: both
// operands MN-major as TMA delivers them.  A = dy tap tiles, M = 64 = 4 slots of 16 channels at the
// copy stride of a section (slots 0..2 = the three w copies = taps kw 1, 0, 2; slot 3 is never read
// back), B = the x tile (N = CI).  One accumulator per (kd, kh), resident in TMEM for the whole
// sweep; three issuing threads (one per kd).  Per-CTA partial tiles, summed by the unpack kernel.
struct alignas(64) TcConvTrWgradParams {
  CUtensorMap tmA[8];  // dy classes
  CUtensorMap tmX;     // x, box (CI, 8, 16, 1, 1)
  int n, D, H, W;
  int tilesH, tilesW, dseg, nseg;
  float* out;          // [CTA][27][16][CI]
};

template <int CI>
__global__ void __launch_bounds__(192)
tc_convtr_wgrad_kernel(const __grid_constant__ TcConvTrWgradParams p) {
  constexpr int RING = 3, XR = 3;
  constexpr int PX = CI * 2;
  constexpr int X_BYTES = TH * TWV * PX;
  constexpr uint32_t TMEM_COLS = 9 * CI <= 256 ? 256 : 512;
  static_assert(9 * CI <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* xring = smem + RING * DY_UNIT;  // also the read slack behind the last dy copy (slot 3 of the A tiles)
  uint64_t* full = reinterpret_cast<uint64_t*>(xring + XR * X_BYTES);
  uint64_t* empty = full + RING;
  uint64_t* fullX = empty + RING;
  uint64_t* emptyX = fullX + XR;
  uint64_t* acc_full = emptyX + XR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  int bx = blockIdx.x;
  const int seg = bx % p.nseg; bx /= p.nseg;
  const int tw_i = bx % p.tilesW; bx /= p.tilesW;
  const int th_i = bx % p.tilesH; bx /= p.tilesH;
  const int n = bx;
  const int h0 = th_i * TH, w0 = tw_i * TWV;
  const int d_begin = seg * p.dseg;
  const int nd = min(p.dseg, p.D - d_begin);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 3); }
    for (int i = 0; i < XR; ++i) { tc::mbar_init(&fullX[i], 1); tc::mbar_init(&emptyX[i], 3); }
    tc::mbar_init(acc_full, 3);
    tc::fence_barrier_init();
    for (int i = 0; i < 8; ++i) tc::prefetch_tmap(&p.tmA[i]);
    tc::prefetch_tmap(&p.tmX);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s <= nd; ++s) {
        const int slot = s % RING;
        tc::mbar_wait(&empty[slot], (((uint32_t)(s / RING)) & 1u) ^ 1u);
        uint8_t* dst = ring + slot * DY_UNIT;
        tc::mbar_expect_tx(&full[slot], s ? DY_UNIT : DY_DSLAB);
        load_dy_dslab(dst, p.tmA, 1, &full[slot], w0, h0, d_begin - 1 + s, n);
        if (s) {
          load_dy_dslab(dst + DY_DSLAB, p.tmA, 0, &full[slot], w0, h0, d_begin + s - 1, n);
          const int j = s - 1, xs = j % XR;
          tc::mbar_wait(&emptyX[xs], (((uint32_t)(j / XR)) & 1u) ^ 1u);
          tc::mbar_expect_tx(&fullX[xs], X_BYTES);
          tc::tma_load_5d(xring + xs * X_BYTES, &p.tmX, &fullX[xs], 0, w0, h0, d_begin + j, n);
        }
      }
    }
  } else if (warp <= 3) {  // warp-uniform issue loop, one elected lane issues (`else`: see tc::warp_index)
    constexpr uint32_t idesc = tc::make_idesc_bf16(64, CI, true, true);
    constexpr uint64_t layB = tc::layout_for_row_bytes(PX);
    const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform registers: see tc::warp_uniform
    const int kd = (int)tc::warp_uniform((uint32_t)warp) - 1;
    const uint32_t r_addr = tc::smem_u32(ring), x_addr = tc::smem_u32(xring);
    // A (MN-major): LBO = stride between 16-channel slots = the copy stride of the section, SBO = one line (8 voxels)
    const uint64_t a_tmpl0 = tc::make_smem_desc(0, TH * DY_LINE, DY_LINE, tc::LAYOUT_SW32);        // section ph=0
    const uint64_t a_tmpl1 = tc::make_smem_desc(0, (TH + 1) * DY_LINE, DY_LINE, tc::LAYOUT_SW32);  // section ph=1
    const uint64_t b_tmpl = tc::make_smem_desc(0, 16, TWV * PX, layB);
    int waited = 0;
    for (int j = 0; j < nd; ++j) {
      while (waited <= j + 1) {
        tc::mbar_wait(&full[waited % RING], ((uint32_t)(waited / RING)) & 1u);
        ++waited;
      }
      tc::mbar_wait(&fullX[j % XR], ((uint32_t)(j / XR)) & 1u);
      tc::tc_fence_after();
      const uint32_t part = r_addr + (kd == 0 ? (j % RING) * DY_UNIT
                                              : ((j + 1) % RING) * DY_UNIT + (kd == 1 ? DY_DSLAB : 0));
      const uint64_t xb = b_tmpl + ((x_addr + (j % XR) * X_BYTES) >> 4);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint64_t a0 = (kh == 1 ? a_tmpl0 + (part >> 4)
                                     : a_tmpl1 + ((part + DY_SEC1 + (kh == 2 ? DY_LINE : 0)) >> 4));
        const uint32_t acc = tmem_acc + (kd * 3 + kh) * CI;
#pragma unroll
        for (int t = 0; t < TH / 2; ++t)
          tc::umma_bf16_warp(acc, a0 + (((2 * t) * DY_LINE) >> 4), xb + (((2 * t) * (TWV * PX)) >> 4), idesc,
                        (j > 0 || t > 0) ? 1u : 0u);
      }
      tc::umma_commit_warp(&empty[j % RING]);
      tc::umma_commit_warp(&emptyX[j % XR]);
    }
    tc::umma_commit_warp(acc_full);
  }
  if (warp >= 2) {
    const int q = warp & 3;  // TMEM quarter = slot = w copy: kw = 1, 0, 2
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    if (q < 3) {
      const int kw = q == 0 ? 1 : (q == 1 ? 0 : 2);
      const int co = lane & 15;
      const bool valid = lane < 16;
      for (int a = 0; a < 9; ++a) {
        const int tap = a * 3 + kw;  // a = kd*3 + kh
        float* orow = p.out + (((int64_t)blockIdx.x * 27 + tap) * 16 + co) * CI;
#pragma unroll
        for (int ch = 0; ch < CI / 16; ++ch) {
          uint32_t v[16];
          tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + a * CI + ch * 16, v);
          tc::tmem_ld_wait();
          if (valid) {
            float4* o4 = reinterpret_cast<float4*>(orow + ch * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                  __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_acc);
}

// defined in tc_slide_wgrad.cu: gw[b][a][tap] = sum_cta G[cta][tap][a][b]
int tc_slide_wgrad_unpack(const float* G, float* gw, int taps, int a_c, int b_c, int a_pad, int b_pad, int nparts,
                          cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// host side.  The three kernels are written in terms of a LOW-resolution tensor (32 / 64 channels)
// and a HIGH-resolution tensor (2x in every dimension, <= 16 channels, padded rows) related by
// hi = 2 * lo - 1 + tap.  Both stride-2 layer types of the network map onto them:
//   ConvTranspose k3 s2 (lo = x, hi = y):  fprop = lo->hi (F),  dgrad = hi->lo (D),  wgrad = W
//   Conv          k3 s2 (hi = x, lo = y):  fprop = hi->lo (D),  dgrad = lo->hi (F),  wgrad = W
// with the weight tiles of the layer's own packed layout (tap index identical in all six cases).
namespace {

struct HiLo {
  int n, D, H, W;      // low-resolution extent
  int lo_c, hi_c;      // channels (hi_c <= 16)
  int lo_ld, hi_ld;
  bool transposed_layer;
};

bool hilo(const b200seg_conv_desc* d, bool transposed_layer, HiLo& g) {
  if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 2 || d->sh != 2 || d->sw != 2) return false;
  g.transposed_layer = transposed_layer;
  g.n = d->n;
  if (transposed_layer) {
    if (d->out_d != 2 * d->in_d || d->out_h != 2 * d->in_h || d->out_w != 2 * d->in_w) return false;
    g.D = d->in_d; g.H = d->in_h; g.W = d->in_w;
    g.lo_c = d->cin; g.hi_c = d->cout; g.lo_ld = d->x_ld; g.hi_ld = d->y_ld;
  } else {
    if (d->in_d != 2 * d->out_d || d->in_h != 2 * d->out_h || d->in_w != 2 * d->out_w) return false;
    g.D = d->out_d; g.H = d->out_h; g.W = d->out_w;
    g.lo_c = d->cout; g.hi_c = d->cin; g.lo_ld = d->y_ld; g.hi_ld = d->x_ld;
  }
  if (g.D < 4 || (int64_t)g.H * g.W < 512) return false;
  if ((g.lo_c != 32 && g.lo_c != 64) || round16(g.hi_c) != 16 || g.hi_c < 8) return false;
  return true;
}

// column / d-segment decomposition: as many CTAs as fit in ONE wave (`per_sm` resident CTAs per SM)
void hilo_grid(const HiLo& g, int per_sm, int& tilesH, int& tilesW, int& dseg, int& nseg, int64_t& grid) {
  tilesH = (g.H + TH - 1) / TH;
  tilesW = (g.W + TWV - 1) / TWV;
  const int64_t cols = (int64_t)g.n * tilesH * tilesW;
  int ns = (int)((148 * per_sm) / cols);
  if (ns < 1) ns = 1;
  dseg = (g.D + ns - 1) / ns;
  if (dseg < 4) dseg = 4;
  if (dseg > g.D) dseg = g.D;
  nseg = (g.D + dseg - 1) / dseg;
  grid = cols * nseg;
}

enum { KERNEL_F = 0, KERNEL_D = 1 };
// which kernel an op of a layer runs on (-1: none)
int kernel_of(bool transposed_layer, int op) {
  if (transposed_layer) return op == TC_CONVTR_FPROP ? KERNEL_F : (op == TC_CONVTR_DGRAD ? KERNEL_D : -1);
  return op == TC_CONV_FPROP ? KERNEL_D : (op == TC_CONV_DGRAD ? KERNEL_F : -1);
}

// the 8 parity-class tensor maps of the high-resolution tensor (16 padded channels), boxes of 16 / 17 lines
int make_hi_maps(CUtensorMap* maps, const HiLo& g, const void* hi) {
  const int oD = 2 * g.D, oH = 2 * g.H, oW = 2 * g.W, ld = g.hi_ld;
  for (int m = 0; m < 8; ++m) {
    const int pd = (m >> 2) & 1, ph = (m >> 1) & 1, pw = m & 1;
    const bf16* base = (const bf16*)hi + (((int64_t)pd * oH + ph) * oW + pw) * ld;
    uint64_t dims[5] = {16, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.D, (uint64_t)g.n};
    uint64_t strides[4] = {(uint64_t)ld * 2 * 2, (uint64_t)oW * ld * 2 * 2, (uint64_t)oH * oW * ld * 2 * 2,
                           (uint64_t)oD * oH * oW * ld * 2};
    uint32_t box[5] = {16, (uint32_t)TWV, (uint32_t)(ph ? TH + 1 : TH), 1, 1};
    int rc = tc_make_map(&maps[m], base, 5, dims, strides, box, 32);
    if (rc) return rc;
  }
  return B200SEG_OK;
}

int make_lo_map(CUtensorMap* map, const HiLo& g, const void* lo, int lines) {
  uint64_t dims[5] = {(uint64_t)g.lo_c, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.D, (uint64_t)g.n};
  uint64_t strides[4] = {(uint64_t)g.lo_ld * 2, (uint64_t)g.W * g.lo_ld * 2, (uint64_t)g.H * g.W * g.lo_ld * 2,
                         (uint64_t)g.D * g.H * g.W * g.lo_ld * 2};
  uint32_t box[5] = {(uint32_t)g.lo_c, (uint32_t)TWV, (uint32_t)lines, 1, 1};
  return tc_make_map(map, lo, 5, dims, strides, box, g.lo_c * 2);
}

template <int KC>
int launch_f(const TcConvTrFpropParams& p, unsigned grid, cudaStream_t st) {
  constexpr int RING = KC == 32 ? 4 : 3;
  constexpr int SLAB = 2 * (TH + 1) * TWV * KC * 2;
  constexpr int WB = (27 * 16 * KC * 2 + 1023) / 1024 * 1024;
  const size_t smem = 1024 + WB + RING * SLAB + 16 * 8 + 64 + 4 * 16 * 2 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_convtr_fprop_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  tc_convtr_fprop_kernel<KC><<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_convtr_fprop");
  count_tc_launch();
  return B200SEG_OK;
}

template <int CI>
int launch_d(const TcConvTrDgradParams& p, unsigned grid, cudaStream_t st) {
  constexpr int WB = (27 * CI * DY_PITCH + 1023) / 1024 * 1024;
  const size_t smem = 1024 + WB + 3 * DY_UNIT + 16 * 8 + 64 + 4 * CI * 2 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_convtr_dgrad_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  tc_convtr_dgrad_kernel<CI><<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_convtr_dgrad");
  count_tc_launch();
  return B200SEG_OK;
}

// lo -> hi
int run_f(const HiLo& g, const void* lo, const void* w_tc, const float* bias, void* hi, float* stats, int accumulate,
          cudaStream_t st) {
  TcConvTrFpropParams p;
  memset(&p, 0, sizeof(p));
  p.n = g.n; p.D = g.D; p.H = g.H; p.W = g.W;
  int64_t grid;
  hilo_grid(g, g.lo_c == 32 ? 2 : 1, p.tilesH, p.tilesW, p.dseg, p.nseg, grid);
  p.cout = g.hi_c; p.dst_ld = g.hi_ld; p.accumulate = accumulate;
  p.bias = bias; p.dst = (bf16*)hi; p.stats = stats;
  int rc = make_lo_map(&p.tmA, g, lo, TH + 1);
  if (rc) return rc;
  {
    uint64_t dims[2] = {(uint64_t)g.lo_c, (uint64_t)27 * 16};
    uint64_t strides[1] = {(uint64_t)g.lo_c * 2};
    uint32_t box[2] = {(uint32_t)g.lo_c, 16};
    rc = tc_make_map(&p.tmB, w_tc, 2, dims, strides, box, g.lo_c * 2);
    if (rc) return rc;
  }
  if (grid > 0x7fffffffLL) { set_error("tc_convtr_slide: grid too large"); return B200SEG_ERR_ARG; }
  if (g.lo_c == 32) return launch_f<32>(p, (unsigned)grid, st);
  return launch_f<64>(p, (unsigned)grid, st);
}

// hi -> lo
int run_d(const HiLo& g, const void* hi, const void* w_tc, const float* bias, void* lo, float* stats, cudaStream_t st) {
  TcConvTrDgradParams p;
  memset(&p, 0, sizeof(p));
  p.n = g.n; p.D = g.D; p.H = g.H; p.W = g.W;
  int64_t grid;
  hilo_grid(g, 1, p.tilesH, p.tilesW, p.dseg, p.nseg, grid);
  p.dst_ld = g.lo_ld; p.dst = (bf16*)lo; p.cout = g.lo_c; p.bias = bias; p.stats = stats;
  int rc = make_hi_maps(p.tmA, g, hi);
  if (rc) return rc;
  {
    uint64_t dims[2] = {16, (uint64_t)27 * g.lo_c};
    uint64_t strides[1] = {32};
    uint32_t box[2] = {16, (uint32_t)g.lo_c};
    rc = tc_make_map(&p.tmB, w_tc, 2, dims, strides, box, 32);
    if (rc) return rc;
  }
  if (grid > 0x7fffffffLL) { set_error("tc_convtr_dgrad: grid too large"); return B200SEG_ERR_ARG; }
  if (g.lo_c == 32) return launch_d<32>(p, (unsigned)grid, st);
  return launch_d<64>(p, (unsigned)grid, st);
}

}  // namespace

// Stride-2 layers the sliding kernels take (on top of tc_conv_supported): 3-D k3 s2 with exact 2x
// extents, low-resolution side 32 / 64 channels, high-resolution side <= 16 channels (padded rows).
// No fused residual; in-place accumulation only for the lo->hi kernel (strided-conv dgrad fan-in).
bool tc_convtr_slide_supported(const b200seg_conv_desc* d, int op, const void* residual) {
  if (d->flags & B200SEG_CONV_NO_SLIDE) return false;
  const bool transposed_layer = (op == TC_CONVTR_FPROP || op == TC_CONVTR_DGRAD);
  HiLo g;
  if (!hilo(d, transposed_layer, g)) return false;
  const int k = kernel_of(transposed_layer, op);
  if (k < 0 || residual) return false;
  if ((d->flags & B200SEG_CONV_ACCUMULATE) && k != KERNEL_F) return false;
  return true;
}

// number of CTAs (= per-CTA statistic partials, fprop only); CTAs of one sample are contiguous
int64_t tc_convtr_slide_grid(const b200seg_conv_desc* d, int op) {
  const bool transposed_layer = (op == TC_CONVTR_FPROP || op == TC_CONVTR_DGRAD);
  HiLo g;
  hilo(d, transposed_layer, g);
  int th, tw, dseg, nseg;
  int64_t grid;
  hilo_grid(g, (kernel_of(transposed_layer, op) == KERNEL_F && g.lo_c == 32) ? 2 : 1, th, tw, dseg, nseg, grid);
  return grid;
}

int tc_convtr_slide_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                        void* dst, float* stats, cudaStream_t st) {
  const bool transposed_layer = (op == TC_CONVTR_FPROP || op == TC_CONVTR_DGRAD);
  HiLo g;
  if (!hilo(d, transposed_layer, g)) { set_error("tc_convtr_slide: unsupported layer"); return B200SEG_ERR_UNSUPPORTED; }
  if (kernel_of(transposed_layer, op) == KERNEL_F)
    return run_f(g, src, w_tc, bias, dst, stats, (d->flags & B200SEG_CONV_ACCUMULATE) ? 1 : 0, st);
  return run_d(g, src, w_tc, bias, dst, stats, st);
}

// wgrad of the same layers: low-resolution side exactly 32 channels (9 accumulators x 32 TMEM columns)
bool tc_convtr_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer) {
  if (d->flags & B200SEG_CONV_NO_SLIDE) return false;
  HiLo g;
  return hilo(d, transposed_layer, g) && g.lo_c == 32;
}

size_t tc_convtr_wgrad_workspace(const b200seg_conv_desc* d, bool transposed_layer) {
  HiLo g;
  if (!hilo(d, transposed_layer, g)) return 0;
  int th, tw, dseg, nseg;
  int64_t grid;
  hilo_grid(g, 1, th, tw, dseg, nseg, grid);
  return (size_t)grid * 27 * 16 * g.lo_c * sizeof(float);
}

int tc_convtr_wgrad_run(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy, float* gw,
                        float* G32, cudaStream_t st) {
  HiLo g;
  if (!hilo(d, transposed_layer, g)) { set_error("tc_convtr_wgrad: unsupported layer"); return B200SEG_ERR_UNSUPPORTED; }
  const void* hi = transposed_layer ? dy : x;
  const void* lo = transposed_layer ? x : dy;
  TcConvTrWgradParams p;
  memset(&p, 0, sizeof(p));
  constexpr int CI = 32;
  p.n = g.n; p.D = g.D; p.H = g.H; p.W = g.W;
  int64_t grid;
  hilo_grid(g, 1, p.tilesH, p.tilesW, p.dseg, p.nseg, grid);
  p.out = G32;
  int rc = make_hi_maps(p.tmA, g, hi);
  if (rc) return rc;
  rc = make_lo_map(&p.tmX, g, lo, TH);
  if (rc) return rc;
  const size_t smem = 1024 + 3 * DY_UNIT + 3 * (TH * TWV * CI * 2) + 16 * 8 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_convtr_wgrad_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  tc_convtr_wgrad_kernel<CI><<<(unsigned)grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_convtr_wgrad");
  count_tc_launch();
  // G[cta][tap][hi channel][lo channel]; both PyTorch layouts are [lo channel][hi channel][tap]:
  // ConvTranspose (cin = lo, cout = hi, taps), Conv (cout = lo, cin = hi, taps)
  return tc_slide_wgrad_unpack(G32, gw, 27, g.hi_c, g.lo_c, 16, CI, (int)grid, st);
}

}  // namespace b200seg

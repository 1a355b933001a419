// Bandwidth-bound voxel-wise kernels at the end of the network: fused softmax + Dice sums
// (forward and backward), fused softmax-argmax + Dice-metric counts, mask squashing and HU
// windowing.  One thread per voxel, channels-last logits (C <= 32), fp32 math.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace b200seg {

namespace {

constexpr int kDiceThreads = 256;

inline int dice_blocks(int64_t spatial, int n) {
  int64_t nb = cdiv64(spatial, kDiceThreads * 4);
  int64_t cap = 1184 / (n > 0 ? n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

template <int LT>
__device__ __forceinline__ int load_label(const void* labels, int64_t idx) {
  if constexpr (LT == B200SEG_LABEL_U8) return (int)reinterpret_cast<const uint8_t*>(labels)[idx];
  else return (int)reinterpret_cast<const long long*>(labels)[idx];
}

// softmax of one voxel: p[0..C) (entries >= C are 0).  Plain expf / division in fp32, i.e. the
// arithmetic `torch.softmax` performs (max-subtracted exponentials over their sum).
// vec16: the voxel row is 16 bf16 (32 B, 16-byte aligned): two 128-bit loads instead of C scalar ones
// CE > 0: the class count is the compile-time constant CE (C is ignored): all loops run over exactly CE classes
template <typename T, int CMAX, int CE = 0>
__device__ __forceinline__ void voxel_softmax(const T* z, int C, float (&p)[CMAX], bool vec16 = false,
                                              float* mx_out = nullptr, float* logs_out = nullptr) {
  if constexpr (CE > 0) C = CE;
  float mx = -INFINITY;
  if constexpr (sizeof(T) == 2 && CMAX == 16) {
    if (vec16) {
      const uint4* zp = reinterpret_cast<const uint4*>(z);
      uint4 r0 = zp[0], r1 = zp[1];
      const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
      const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 a = __bfloat1622float2(h0[i]), b = __bfloat1622float2(h1[i]);
        p[2 * i] = a.x; p[2 * i + 1] = a.y; p[8 + 2 * i] = b.x; p[8 + 2 * i + 1] = b.y;
      }
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        p[c] = c < C ? p[c] : -INFINITY;
        mx = fmaxf(mx, p[c]);
      }
    }
  }
  if (mx == -INFINITY) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      p[c] = c < C ? to_f<T>(z[c]) : -INFINITY;
      mx = fmaxf(mx, p[c]);
    }
  }
  float s = 0.f;
  if constexpr (sizeof(T) == 2) {
    // bf16 logits carry 8 significant bits: ex2.approx and one reciprocal are far below that
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      p[c] = c < C ? exp2f((p[c] - mx) * 1.4426950408889634f) : 0.f;
      s += p[c];
    }
    const float inv = __frcp_rn(s);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) p[c] *= inv;
    if (mx_out) { *mx_out = mx; *logs_out = __logf(s); }
  } else {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      p[c] = c < C ? expf(p[c] - mx) : 0.f;
      s += p[c];
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) p[c] = p[c] / s;
    if (mx_out) { *mx_out = mx; *logs_out = logf(s); }
  }
}

}  // namespace

// bf16 logits stored as 16-channel rows (10 classes zero-padded, B200SEG padded buffers): the row
// can be moved with two 128-bit accesses; a destination row may be fully rewritten (padding = 0)
static bool vec16_ok(const b200seg_dice_desc& d, const void* a, const void* b) {
  return d.dtype == B200SEG_BF16 && d.ld == 16 && d.c <= 16 && ((uintptr_t)a % 16) == 0 &&
         ((uintptr_t)b % 16) == 0;
}

size_t dice_workspace_bytes(const b200seg_dice_desc& d) {
  return (size_t)d.n * dice_blocks(d.spatial, d.n) * d.c * 3 * sizeof(float) + 256;
}

// partial[n][blk][c][3] = { I, G, P }
template <typename T, int CMAX, int LT, int CE = 0>
__global__ void __launch_bounds__(kDiceThreads)
softmax_dice_fwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                        int64_t spatial, int C, int ld, int64_t vox_per_block,
                        float* __restrict__ partial, bool vec16) {
  if constexpr (CE > 0) C = CE;
  constexpr int CL = CE > 0 ? CE : CMAX;  // classes that exist at compile time
  __shared__ float red[kDiceThreads / 32][CMAX * 3];
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float aI[CMAX], aG[CMAX], aP[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aG[c] = aP[c] = 0.f;
  for (int64_t v = v_begin + threadIdx.x; v < v_end; v += kDiceThreads) {
    int64_t vox = (int64_t)n * spatial + v;
    float p[CMAX];
    voxel_softmax<T, CMAX, CE>(logits + vox * ld, C, p, vec16);
    int lab = load_label<LT>(labels, vox);
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      bool hit = (lab == c);
      aI[c] += hit ? p[c] : 0.f;
      aG[c] += hit ? 1.f : 0.f;
      aP[c] += p[c];
    }
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    float i_ = warp_sum(aI[c]), g_ = warp_sum(aG[c]), p_ = warp_sum(aP[c]);
    if (lane == 0) {
      red[warp][c * 3 + 0] = i_;
      red[warp][c * 3 + 1] = g_;
      red[warp][c * 3 + 2] = p_;
    }
  }
  __syncthreads();
  if (threadIdx.x < C * 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kDiceThreads / 32; ++w) s += red[w][threadIdx.x];
    partial[((int64_t)n * gridDim.x + blockIdx.x) * C * 3 + threadIdx.x] = s;
  }
}

__global__ void dice_sums_final_kernel(const float* __restrict__ partial, int nblk, int per_n,
                                       int total, float* __restrict__ sums) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= total) return;
  int n = warp / per_n, j = warp % per_n;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += (double)partial[((int64_t)n * nblk + b) * per_n + j];
  s = warp_sum_d(s);
  if (lane == 0) sums[warp] = (float)s;
}

template <typename T, int CMAX, int LT, int CE = 0>
__global__ void __launch_bounds__(kDiceThreads)
softmax_dice_bwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                        const float* __restrict__ gI, const float* __restrict__ gP,
                        T* __restrict__ dlogits, int64_t spatial, int C, int ld, bool vec16) {
  if constexpr (CE > 0) C = CE;
  constexpr int CL = CE > 0 ? CE : CMAX;
  const int n = blockIdx.y;
  float cI[CMAX], cP[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    cI[c] = c < C ? gI[n * C + c] : 0.f;
    cP[c] = c < C ? gP[n * C + c] : 0.f;
  }
  for (int64_t v = (int64_t)blockIdx.x * kDiceThreads + threadIdx.x; v < spatial;
       v += (int64_t)gridDim.x * kDiceThreads) {
    int64_t vox = (int64_t)n * spatial + v;
    float p[CMAX];
    voxel_softmax<T, CMAX, CE>(logits + vox * ld, C, p, vec16);
    int lab = load_label<LT>(labels, vox);
    float g[CMAX];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < CL) {
        g[c] = (lab == c ? cI[c] : 0.f) + cP[c];
        dot = fmaf(g[c], p[c], dot);
      } else {
        g[c] = 0.f;
      }
    }
    T* o = dlogits + vox * ld;
    bool done = false;
    if constexpr (sizeof(T) == 2 && CMAX == 16) {
      if (vec16) {  // 16-channel padded row: write all 16 (padding = 0) with two 128-bit stores
        uint4 o0, o1;
        __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
        float dz[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) dz[c] = c < C ? p[c] * (g[c] - dot) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          q0[i] = __floats2bfloat162_rn(dz[2 * i], dz[2 * i + 1]);
          q1[i] = __floats2bfloat162_rn(dz[8 + 2 * i], dz[8 + 2 * i + 1]);
        }
        reinterpret_cast<uint4*>(o)[0] = o0;
        reinterpret_cast<uint4*>(o)[1] = o1;
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) o[c] = from_f<T>(p[c] * (g[c] - dot));
    }
  }
}

// ---- production layout (bf16 logits in 16-channel rows, 10 classes): bulk-copy staged kernels -------------------
// The kernels above keep one 32-byte row in flight per thread; at 64 registers that is 32 KB per SM, i.e. latency
// bound at ~2.9 TB/s (r2 measurement: fwd 49 us, bwd 94 us at 2 x 128^3).  Here a CTA streams its voxel range through
// a shared-memory ring filled by 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map): RING_STAGES x
// 8 KB in flight per CTA whatever the register count, completion on mbarriers; every thread then reads its own row
// from shared memory.  The forward kernel optionally adds the Dice METRIC counts (reference `_log_dice_scores`,
// capstone/volumetric/base_trainer.py:116-132: argmax of the softmax, first maximum, then |pred|, |target|, tp per
// class) to the same pass -- they need nothing but the probabilities and the label the loss sums already use.
namespace {
constexpr int RING_VOX = 256;     // voxels per stage = threads per CTA
constexpr int RING_STAGES = 4;
constexpr int RING_ROW = 32;      // bytes per voxel row (16 x bf16)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// softmax of a 16-wide bf16 row held in two 128-bit registers; CE classes, the rest of the row is padding
template <int CE>
__device__ __forceinline__ void row_softmax(const uint4& r0, const uint4& r1, float (&p)[16]) {
  const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
  const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 a = __bfloat1622float2(h0[i]), b = __bfloat1622float2(h1[i]);
    p[2 * i] = a.x; p[2 * i + 1] = a.y; p[8 + 2 * i] = b.x; p[8 + 2 * i + 1] = b.y;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CE; ++c) mx = fmaxf(mx, p[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    p[c] = c < CE ? exp2f((p[c] - mx) * 1.4426950408889634f) : 0.f;
    s += p[c];
  }
  const float inv = __frcp_rn(s);
#pragma unroll
  for (int c = 0; c < CE; ++c) p[c] *= inv;
}
}  // namespace

// partial[n][blk][CE][NS]: NS = 3 {I, G, P} or 5 {I, G, P, NP, TP} (NP = #voxels predicted c, TP = predicted and labelled c)
template <int CE, int LT, bool COUNTS>
__global__ void __launch_bounds__(RING_VOX)
softmax_dice_fwd_ring_kernel(const __nv_bfloat16* __restrict__ logits, const void* __restrict__ labels, int64_t spatial,
                             int64_t vox_per_block, float* __restrict__ partial) {
  constexpr int NS = COUNTS ? 5 : 3;
  __shared__ alignas(128) uint8_t ring[RING_STAGES][RING_VOX * RING_ROW];
  __shared__ uint64_t full[RING_STAGES];
  __shared__ float red[RING_VOX / 32][CE * NS];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  const int nchunks = (int)((v_end - v_begin + RING_VOX - 1) / RING_VOX);
  const __nv_bfloat16* base = logits + ((int64_t)n * spatial + v_begin) * 16;
  if (tid == 0) {
    for (int i = 0; i < RING_STAGES; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int k) {
    const int64_t v0 = (int64_t)k * RING_VOX;
    const uint32_t bytes = (uint32_t)min((int64_t)RING_VOX, v_end - v_begin - v0) * RING_ROW;
    tc::mbar_expect_tx(&full[k % RING_STAGES], bytes);
    bulk_load(ring[k % RING_STAGES], base + v0 * 16, bytes, &full[k % RING_STAGES]);
  };
  if (tid == 0)
    for (int k = 0; k < RING_STAGES && k < nchunks; ++k) issue(k);
  static_assert(CE * 6 <= 64, "packed per-class counters: 6 bits per class in one 64-bit word");
  float aI[CE], aP[CE];
  unsigned int cG[CE], cN[COUNTS ? CE : 1], cT[COUNTS ? CE : 1];
#pragma unroll
  for (int c = 0; c < CE; ++c) { aI[c] = aP[c] = 0.f; cG[c] = 0u; }
#pragma unroll
  for (int c = 0; c < (COUNTS ? CE : 1); ++c) cN[c] = cT[c] = 0u;
  // The three COUNTS (|target|, |pred|, tp per class) are kept as 6-bit fields of one 64-bit word each -- one shift
  // and one add per voxel instead of a compare + predicated add per class -- and spilled into the 32-bit per-class
  // counters every 32 voxels (a field holds up to 63).  The kernel is instruction bound (ncu r2: 71 % SM
  // throughput at 29 % DRAM), every instruction per voxel counts.
  unsigned long long pkG = 0ull, pkN = 0ull, pkT = 0ull;
  auto spill = [&]() {
#pragma unroll
    for (int c = 0; c < CE; ++c) {
      cG[c] += (unsigned int)(pkG >> (6 * c)) & 63u;
      if constexpr (COUNTS) {
        cN[c] += (unsigned int)(pkN >> (6 * c)) & 63u;
        cT[c] += (unsigned int)(pkT >> (6 * c)) & 63u;
      }
    }
    pkG = pkN = pkT = 0ull;
  };
  for (int k = 0; k < nchunks; ++k) {
    const int st = k % RING_STAGES;
    const int64_t v = v_begin + (int64_t)k * RING_VOX + tid;
    const bool live = v < v_end;
    const int lab = live ? load_label<LT>(labels, (int64_t)n * spatial + v) : -1;
    tc::mbar_wait(&full[st], (uint32_t)(k / RING_STAGES) & 1u);
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (live) {
      const uint4* rp = reinterpret_cast<const uint4*>(ring[st] + tid * RING_ROW);
      r0 = rp[0];
      r1 = rp[1];
    }
    if (live) {
      float p[16];
      row_softmax<CE>(r0, r1, p);
      int best = 0;
      if constexpr (COUNTS) {
        float bv = p[0];
#pragma unroll
        for (int c = 1; c < CE; ++c)
          if (p[c] > bv) { bv = p[c]; best = c; }
      }
#pragma unroll
      for (int c = 0; c < CE; ++c) {
        aI[c] += (lab == c) ? p[c] : 0.f;
        aP[c] += p[c];
      }
      if ((unsigned)lab < (unsigned)CE) {
        const unsigned long long one = 1ull << (6 * lab);
        pkG += one;
        if constexpr (COUNTS) pkT += (best == lab) ? one : 0ull;
      }
      if constexpr (COUNTS) pkN += 1ull << (6 * best);
    }
    if ((k & 31) == 31) spill();
    // Hand the stage back to the bulk-copy engine only AFTER the rows have been consumed: the shared-memory loads
    // above are then complete (their values were used), every thread orders its generic-proxy reads before the
    // async-proxy refill (fence.proxy.async), and the CTA barrier publishes that to the issuing thread.  (With the
    // refill issued right after the loads were merely ISSUED, the copy could overtake them: r2 found the backward
    // kernel non-deterministic at 4 x 160^3, 7 CTAs per SM.)
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && k + RING_STAGES < nchunks) issue(k + RING_STAGES);
  }
  spill();
  const int warp = tid / 32, lane = tid % 32;
#pragma unroll
  for (int c = 0; c < CE; ++c) {
    const float i_ = warp_sum(aI[c]), p_ = warp_sum(aP[c]);
    const unsigned g_ = __reduce_add_sync(0xffffffffu, cG[c]);
    if (lane == 0) {
      red[warp][c * NS + 0] = i_;
      red[warp][c * NS + 1] = (float)g_;   // exact: a block covers far fewer than 2^24 voxels
      red[warp][c * NS + 2] = p_;
    }
    if constexpr (COUNTS) {
      const unsigned a = __reduce_add_sync(0xffffffffu, cN[c]), b = __reduce_add_sync(0xffffffffu, cT[c]);
      if (lane == 0) {
        red[warp][c * NS + 3] = (float)a;   // exact: a block covers far fewer than 2^24 voxels
        red[warp][c * NS + 4] = (float)b;
      }
    }
  }
  __syncthreads();
  if (tid < CE * NS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RING_VOX / 32; ++w) s += red[w][tid];
    partial[((int64_t)n * gridDim.x + blockIdx.x) * CE * NS + tid] = s;
  }
}

// sums[n][c][3] = {I, G, P} (float) and counts[n][c][3] = {tp, |pred|, |target|} (int64, exact) from the 5-column partials
__global__ void dice_metric_final_kernel(const float* __restrict__ partial, int nblk, int C, int total,
                                         float* __restrict__ sums, long long* __restrict__ counts) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= total) return;  // total = n * C * 5
  const int n = warp / (C * 5), j = warp % (C * 5), c = j / 5, k = j % 5;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += (double)partial[((int64_t)n * nblk + b) * C * 5 + j];
  s = warp_sum_d(s);
  if (lane == 0) {
    if (k < 3) sums[(n * C + c) * 3 + k] = (float)s;
    const long long iv = (long long)(s + 0.5);
    if (k == 1) counts[(n * C + c) * 3 + 2] = iv;   // G = |target|
    if (k == 3) counts[(n * C + c) * 3 + 1] = iv;   // NP = |pred|
    if (k == 4) counts[(n * C + c) * 3 + 0] = iv;   // TP
  }
}

template <int CE, int LT>
__global__ void __launch_bounds__(RING_VOX)
softmax_dice_bwd_ring_kernel(const __nv_bfloat16* __restrict__ logits, const void* __restrict__ labels,
                             const float* __restrict__ gI, const float* __restrict__ gP,
                             __nv_bfloat16* __restrict__ dlogits, int64_t spatial, int64_t vox_per_block) {
  __shared__ alignas(128) uint8_t ring[RING_STAGES][RING_VOX * RING_ROW];
  __shared__ uint64_t full[RING_STAGES];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  const int nchunks = (int)((v_end - v_begin + RING_VOX - 1) / RING_VOX);
  const __nv_bfloat16* base = logits + ((int64_t)n * spatial + v_begin) * 16;
  if (tid == 0) {
    for (int i = 0; i < RING_STAGES; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int k) {
    const int64_t v0 = (int64_t)k * RING_VOX;
    const uint32_t bytes = (uint32_t)min((int64_t)RING_VOX, v_end - v_begin - v0) * RING_ROW;
    tc::mbar_expect_tx(&full[k % RING_STAGES], bytes);
    bulk_load(ring[k % RING_STAGES], base + v0 * 16, bytes, &full[k % RING_STAGES]);
  };
  if (tid == 0)
    for (int k = 0; k < RING_STAGES && k < nchunks; ++k) issue(k);
  float cI[CE], cP[CE];
#pragma unroll
  for (int c = 0; c < CE; ++c) {
    cI[c] = gI[n * CE + c];
    cP[c] = gP[n * CE + c];
  }
  for (int k = 0; k < nchunks; ++k) {
    const int st = k % RING_STAGES;
    const int64_t v = v_begin + (int64_t)k * RING_VOX + tid;
    const bool live = v < v_end;
    const int lab = live ? load_label<LT>(labels, (int64_t)n * spatial + v) : -1;
    tc::mbar_wait(&full[st], (uint32_t)(k / RING_STAGES) & 1u);
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (live) {
      const uint4* rp = reinterpret_cast<const uint4*>(ring[st] + tid * RING_ROW);
      r0 = rp[0];
      r1 = rp[1];
    }
    if (live) {
      float p[16];
      row_softmax<CE>(r0, r1, p);
      float g[CE], dot = 0.f;
#pragma unroll
      for (int c = 0; c < CE; ++c) {
        g[c] = (lab == c ? cI[c] : 0.f) + cP[c];
        dot = fmaf(g[c], p[c], dot);
      }
      float dz[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) dz[c] = c < CE ? p[c] * (g[c < CE ? c : 0] - dot) : 0.f;
      uint4 o0, o1;
      __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
      __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q0[i] = __floats2bfloat162_rn(dz[2 * i], dz[2 * i + 1]);
        q1[i] = __floats2bfloat162_rn(dz[8 + 2 * i], dz[8 + 2 * i + 1]);
      }
      uint4* op = reinterpret_cast<uint4*>(dlogits + ((int64_t)n * spatial + v) * 16);
      op[0] = o0;
      op[1] = o1;
    }
    tc::fence_proxy_async();  // see the forward kernel: rows consumed, reads ordered before the async refill
    __syncthreads();
    if (tid == 0 && k + RING_STAGES < nchunks) issue(k + RING_STAGES);
  }
}

// ---- all voxel-wise losses of the reference through ONE softmax pass --------------------------------
// sums5[n][c][5] = { I, G, P, F, N } per sample and class over the voxels of class c:
//   I = sum p_c t_c, G = sum t_c, P = sum p_c                          (Dice / GeneralizedDice)
//   F = sum t_c (1 - p_c)^gamma (-log p_c)                              (monai FocalLoss, one-hot target)
//   N = sum t_c (-log p_c)                                              (F.cross_entropy, plain or class-weighted)
// log p = (z - max) - log sum exp, as log_softmax computes it.
// HAS_B: a sixth sum for the Boundary loss (capstone/models/losses.py:127-157),
//   B = sum p_c * dist_{c-1}  (c >= 1; dist = pre-computed signed distance maps, planar (N, C-1, spatial) fp32)
// and the row stride of `partial` / sums becomes 6.
template <typename T, int CMAX, int LT, int CE = 0, bool HAS_B = false>
__global__ void __launch_bounds__(kDiceThreads)
softmax_loss_fwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels, int64_t spatial, int C,
                        int ld, int64_t vox_per_block, float gamma, float* __restrict__ partial, bool vec16,
                        const float* __restrict__ dist = nullptr) {
  if constexpr (CE > 0) C = CE;
  constexpr int CL = CE > 0 ? CE : CMAX;
  constexpr int NS = HAS_B ? 6 : 5;
  __shared__ float red[kDiceThreads / 32][CMAX * NS];
  const int n = blockIdx.y;
  const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v_end = min(v_begin + vox_per_block, spatial);
  float aI[CMAX], aG[CMAX], aP[CMAX], aF[CMAX], aN[CMAX], aB[HAS_B ? CMAX : 1];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aG[c] = aP[c] = aF[c] = aN[c] = 0.f;
#pragma unroll
  for (int c = 0; c < (HAS_B ? CMAX : 1); ++c) aB[c] = 0.f;
  for (int64_t v = v_begin + threadIdx.x; v < v_end; v += kDiceThreads) {
    const int64_t vox = (int64_t)n * spatial + v;
    float p[CMAX], mx, ls;
    const T* zrow = logits + vox * ld;
    voxel_softmax<T, CMAX, CE>(zrow, C, p, vec16, &mx, &ls);
    const int lab = load_label<LT>(labels, vox);
    const bool in = lab >= 0 && lab < C;
    const float nll = in ? ls - (to_f<T>(zrow[in ? lab : 0]) - mx) : 0.f;
    const float om = 1.f - __expf(-nll);
    const float fw = gamma == 2.f ? om * om : __powf(fmaxf(om, 0.f), gamma);
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      const bool hit = (lab == c);
      aI[c] += hit ? p[c] : 0.f;
      aG[c] += hit ? 1.f : 0.f;
      aP[c] += p[c];
      aF[c] += hit ? fw * nll : 0.f;
      aN[c] += hit ? nll : 0.f;
    }
    if constexpr (HAS_B) {
      // consecutive threads read consecutive voxels of one distance-map plane: coalesced
      const float* dv = dist + (int64_t)n * (C - 1) * spatial + v;
#pragma unroll
      for (int c = 1; c < CL; ++c)
        if (c < C) aB[c] = fmaf(p[c], dv[(int64_t)(c - 1) * spatial], aB[c]);
    }
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    const float i_ = warp_sum(aI[c]), g_ = warp_sum(aG[c]), p_ = warp_sum(aP[c]), f_ = warp_sum(aF[c]),
                n_ = warp_sum(aN[c]);
    if (lane == 0) {
      red[warp][c * NS + 0] = i_;
      red[warp][c * NS + 1] = g_;
      red[warp][c * NS + 2] = p_;
      red[warp][c * NS + 3] = f_;
      red[warp][c * NS + 4] = n_;
    }
    if constexpr (HAS_B) {
      const float b_ = warp_sum(aB[c]);
      if (lane == 0) red[warp][c * NS + 5] = b_;
    }
  }
  __syncthreads();
  if (threadIdx.x < C * NS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kDiceThreads / 32; ++w) s += red[w][threadIdx.x];
    partial[((int64_t)n * gridDim.x + blockIdx.x) * C * NS + threadIdx.x] = s;
  }
}

// dlogits for d(loss)/d(sums5) = (gI, -, gP, gF, gN) per (n, c):
//   dz_j = p_j (g_j - sum_k g_k p_k) + (delta_{lab,j} - p_j) * A,   g_c = gI_c t_c + gP_c,
//   A = gF_lab * u f'(u) - gN_lab,  u = p_lab,  u f'(u) = -gamma (1-u)^(gamma-1) u nll - (1-u)^gamma
// HAS_B: g_c additionally carries gB_c * dist_{c-1}(v) (Boundary loss, see the forward kernel)
template <typename T, int CMAX, int LT, int CE = 0, bool HAS_B = false>
__global__ void __launch_bounds__(kDiceThreads)
softmax_loss_bwd_kernel(const T* __restrict__ logits, const void* __restrict__ labels, const float* __restrict__ gI,
                        const float* __restrict__ gP, const float* __restrict__ gF, const float* __restrict__ gN,
                        T* __restrict__ dlogits, int64_t spatial, int C, int ld, float gamma, bool vec16,
                        const float* __restrict__ gB = nullptr, const float* __restrict__ dist = nullptr) {
  if constexpr (CE > 0) C = CE;
  constexpr int CL = CE > 0 ? CE : CMAX;
  __shared__ float sF[CMAX], sN[CMAX], sB[CMAX];
  const int n = blockIdx.y;
  if (threadIdx.x < CMAX) {
    sF[threadIdx.x] = threadIdx.x < C ? gF[n * C + threadIdx.x] : 0.f;
    sN[threadIdx.x] = threadIdx.x < C ? gN[n * C + threadIdx.x] : 0.f;
    if constexpr (HAS_B) sB[threadIdx.x] = (threadIdx.x < C && threadIdx.x > 0) ? gB[n * C + threadIdx.x] : 0.f;
  }
  float cI[CMAX], cP[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    cI[c] = c < C ? gI[n * C + c] : 0.f;
    cP[c] = c < C ? gP[n * C + c] : 0.f;
  }
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * kDiceThreads + threadIdx.x; v < spatial;
       v += (int64_t)gridDim.x * kDiceThreads) {
    const int64_t vox = (int64_t)n * spatial + v;
    float p[CMAX], mx, ls;
    const T* zrow = logits + vox * ld;
    voxel_softmax<T, CMAX, CE>(zrow, C, p, vec16, &mx, &ls);
    const int lab = load_label<LT>(labels, vox);
    const bool in = lab >= 0 && lab < C;
    const float nll = in ? ls - (to_f<T>(zrow[in ? lab : 0]) - mx) : 0.f;
    const float u = __expf(-nll), om = 1.f - u;
    float ufp;  // u * f'(u)
    if (gamma == 2.f) ufp = -2.f * om * u * nll - om * om;
    else {
      const float omc = fmaxf(om, 0.f);
      ufp = -gamma * __powf(omc, gamma - 1.f) * u * nll - __powf(omc, gamma);
    }
    const float A = in ? sF[lab] * ufp - sN[lab] : 0.f;
    float g[CMAX];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < CL) {
        g[c] = (lab == c ? cI[c] : 0.f) + cP[c];
        if constexpr (HAS_B) {
          if (c >= 1 && c < C)
            g[c] = fmaf(sB[c], dist[((int64_t)n * (C - 1) + (c - 1)) * spatial + v], g[c]);
        }
        dot = fmaf(g[c], p[c], dot);
      } else {
        g[c] = 0.f;
      }
    }
    T* o = dlogits + vox * ld;
    float dz[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      dz[c] = c < C ? p[c] * (g[c] - dot) + ((lab == c ? 1.f : 0.f) - p[c]) * A : 0.f;
    bool done = false;
    if constexpr (sizeof(T) == 2 && CMAX == 16) {
      if (vec16) {
        uint4 o0, o1;
        __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          q0[i] = __floats2bfloat162_rn(dz[2 * i], dz[2 * i + 1]);
          q1[i] = __floats2bfloat162_rn(dz[8 + 2 * i], dz[8 + 2 * i + 1]);
        }
        reinterpret_cast<uint4*>(o)[0] = o0;
        reinterpret_cast<uint4*>(o)[1] = o1;
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) o[c] = from_f<T>(dz[c]);
    }
  }
}

// Dice loss from the (N, C, 3) sums {I, G, P} and its gradient coefficients in ONE tiny launch
// (replaces ~25 framework kernels on a few dozen numbers):
//   f[n,c] = 1 - (2 I + s) / (G + P + s),  loss = sum_{n, c >= c0} f / count      (count = 1 for "sum")
//   gI[n,c] = d loss / d I = -2 / (G + P + s) / count,  gP[n,c] = d loss / d P = (2 I + s) / (G + P + s)^2 / count
// Classes below c0 (the excluded background) get zero coefficients.  One block, fixed order, double sum.
__global__ void dice_loss_epilogue_kernel(const float* __restrict__ sums, int n, int c, int c0, float smooth,
                                          float inv_count, float* __restrict__ loss, float* __restrict__ gI,
                                          float* __restrict__ gP) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n * c; i += 256) {
    const int ch = i % c;
    float a = 0.f, b = 0.f;
    if (ch >= c0) {
      const float I = sums[i * 3], G = sums[i * 3 + 1], P = sums[i * 3 + 2];
      const float num = 2.f * I + smooth, den = G + P + smooth;
      acc += (double)(1.f - num / den);
      a = -2.f / den * inv_count;
      b = num / (den * den) * inv_count;
    }
    gI[i] = a;
    gP[i] = b;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] * (double)inv_count);
}

int launch_dice_loss_epilogue(const float* sums, int n, int c, int c0, float smooth, float inv_count, float* loss,
                              float* gI, float* gP, cudaStream_t st) {
  dice_loss_epilogue_kernel<<<1, 256, 0, st>>>(sums, n, c, c0, smooth, inv_count, loss, gI, gP);
  B200SEG_CHECK_LAUNCH("dice_loss_epilogue");
  return B200SEG_OK;
}

// pred = argmax softmax (first maximum), optional counts[n][c][3] = {tp, |pred|, |target|}
template <typename T, int CMAX, int LT>
__global__ void __launch_bounds__(kDiceThreads)
argmax_counts_kernel(const T* __restrict__ logits, const void* __restrict__ target,
                     uint8_t* __restrict__ pred_out, unsigned long long* __restrict__ counts,
                     int64_t spatial, int C, int ld, bool vec16) {
  __shared__ unsigned int red[CMAX * 3];
  const int n = blockIdx.y;
  if (threadIdx.x < CMAX * 3) red[threadIdx.x] = 0u;
  __syncthreads();
  unsigned int tp[CMAX], np_[CMAX], nt[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) tp[c] = np_[c] = nt[c] = 0u;
  for (int64_t v = (int64_t)blockIdx.x * kDiceThreads + threadIdx.x; v < spatial;
       v += (int64_t)gridDim.x * kDiceThreads) {
    int64_t vox = (int64_t)n * spatial + v;
    float p[CMAX];
    voxel_softmax<T, CMAX>(logits + vox * ld, C, p, vec16);
    int best = 0;
    float bv = p[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C && p[c] > bv) { bv = p[c]; best = c; }
    if (pred_out) pred_out[vox] = (uint8_t)best;
    if (target) {
      int lab = load_label<LT>(target, vox);
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        np_[c] += (best == c);
        nt[c] += (lab == c);
        tp[c] += (best == c && lab == c);
      }
    }
  }
  if (target) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      unsigned a = __reduce_add_sync(0xffffffffu, tp[c]);
      unsigned b = __reduce_add_sync(0xffffffffu, np_[c]);
      unsigned d = __reduce_add_sync(0xffffffffu, nt[c]);
      if ((threadIdx.x % 32) == 0 && c < C) {
        if (a) atomicAdd(&red[c * 3 + 0], a);
        if (b) atomicAdd(&red[c * 3 + 1], b);
        if (d) atomicAdd(&red[c * 3 + 2], d);
      }
    }
    __syncthreads();
    if (threadIdx.x < C * 3 && red[threadIdx.x])
      atomicAdd(&counts[(int64_t)n * C * 3 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
  }
}

template <int LT>
__global__ void __launch_bounds__(kDiceThreads)
label_counts_kernel(const uint8_t* __restrict__ pred, const void* __restrict__ target,
                    unsigned long long* __restrict__ counts, int64_t spatial, int C) {
  constexpr int CMAX = 32;
  __shared__ unsigned int red[CMAX * 3];
  const int n = blockIdx.y;
  if (threadIdx.x < CMAX * 3) red[threadIdx.x] = 0u;
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * kDiceThreads + threadIdx.x; v < spatial;
       v += (int64_t)gridDim.x * kDiceThreads) {
    int64_t vox = (int64_t)n * spatial + v;
    int pr = pred[vox], lab = load_label<LT>(target, vox);
    if (pr >= 0 && pr < C) atomicAdd(&red[pr * 3 + 1], 1u);
    if (lab >= 0 && lab < C) atomicAdd(&red[lab * 3 + 2], 1u);
    if (pr == lab && pr >= 0 && pr < C) atomicAdd(&red[pr * 3 + 0], 1u);
  }
  __syncthreads();
  if (threadIdx.x < C * 3 && red[threadIdx.x])
    atomicAdd(&counts[(int64_t)n * C * 3 + threadIdx.x], (unsigned long long)red[threadIdx.x]);
}

// labels[n][v] = max_c masks[n][c][v] * (c+1)
__global__ void squash_masks_kernel(const uint8_t* __restrict__ masks, uint8_t* __restrict__ labels,
                                    int n_struct, int64_t spatial, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / spatial, v = i % spatial;
    const uint8_t* m = masks + (n * n_struct) * spatial + v;
    int best = 0;
    for (int c = 0; c < n_struct; ++c) {
      int val = (int)m[(int64_t)c * spatial] * (c + 1);
      best = val > best ? val : best;
    }
    labels[i] = (uint8_t)best;
  }
}

struct WindowCfg {
  float lo[4], hi[4], mean[4], std_[4];
  int n;
};

template <typename T>
__global__ void hu_window_norm_kernel(const int16_t* __restrict__ hu, T* __restrict__ out,
                                      int64_t n_vox, int out_ld, WindowCfg w) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vox;
       i += (int64_t)gridDim.x * blockDim.x) {
    float x = (float)hu[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < w.n) {
        // double, as numpy promotes int16 / float64 in apply_window; then float32 normalise
        double c = fmin(fmax((double)x, (double)w.lo[k]), (double)w.hi[k]);
        double s = (c - (double)w.lo[k]) / ((double)w.hi[k] - (double)w.lo[k] + 1e-8);
        float f = (float)s;
        f = (f - w.mean[k]) * (1.0f / w.std_[k]);
        out[i * out_ld + k] = from_f<T>(f);
      }
    }
  }
}

// ---- sliding-window inference: accumulate a window of logits, then average + arg-max --------
// acc[(d0+d, h0+h, w0+w)][c] += imp[(d,h,w)] * src[(d,h,w)][c];  cnt[(d0+d, h0+h, w0+w)] += imp[(d,h,w)]
// (fp32 accumulators; imp == nullptr: constant importance 1, the += 1 of MONAI's mode="constant")
template <typename T>
__global__ void window_accumulate_kernel(const T* __restrict__ src, int src_ld, const float* __restrict__ imp,
                                         float* __restrict__ acc, float* __restrict__ cnt, int C, int wd, int wh,
                                         int ww, int H, int W, int d0, int h0, int w0, int cd, int ch, int cw) {
  // (cd, ch, cw): extent of the window that lies inside the volume (windows may overhang a padded edge)
  const int64_t total = (int64_t)cd * ch * cw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % cw), h = (int)((i / cw) % ch), d = (int)(i / ((int64_t)cw * ch));
    const int64_t sv = ((int64_t)d * wh + h) * ww + w;
    const T* sp = src + sv * src_ld;
    const int64_t o = ((int64_t)(d0 + d) * H + (h0 + h)) * W + (w0 + w);
    float* ap = acc + o * C;
    if constexpr (sizeof(T) == 2) {
      // the reference's 10 classes, bf16 rows starting on 4-byte boundaries (padded 16-channel rows of the network
      // output as well as the compact 10-channel rows of the exchange buffers): five 32-bit loads of the prediction,
      // five 64-bit read-modify-writes of the fp32 accumulator row (40 bytes, 8-byte aligned)
      if (C == 10 && (src_ld % 2) == 0 && ((uintptr_t)src % 4) == 0 && ((uintptr_t)acc % 8) == 0) {
        const float m = imp ? imp[sv] : 1.f;
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(sp);
        uint32_t r[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) r[i] = rp[i];
        float2* a2 = reinterpret_cast<float2*>(ap);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r[i]));
          float2 a = a2[i];
          if (imp) { a.x = fmaf(m, f.x, a.x); a.y = fmaf(m, f.y, a.y); }
          else { a.x += f.x; a.y += f.y; }
          a2[i] = a;
        }
        cnt[o] += m;
        continue;
      }
    }
    if (imp) {
      const float m = imp[sv];
      for (int c = 0; c < C; ++c) ap[c] = fmaf(m, to_f<T>(sp[c]), ap[c]);
      cnt[o] += m;
    } else {
      for (int c = 0; c < C; ++c) ap[c] += to_f<T>(sp[c]);
      cnt[o] += 1.f;
    }
  }
}

// labels[v] = argmax_c softmax(acc[v] / cnt[v]) (first maximum); optional averaged logits out
__global__ void accum_argmax_kernel(const float* __restrict__ acc, const float* __restrict__ cnt,
                                    uint8_t* __restrict__ labels, float* __restrict__ mean_out,
                                    int64_t nvox, int C) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox;
       v += (int64_t)gridDim.x * blockDim.x) {
    const float inv_n = cnt[v];
    float z[32];
    float mx = -INFINITY;
    for (int c = 0; c < 32; ++c) {
      z[c] = c < C ? acc[v * C + c] / inv_n : -INFINITY;
      mx = fmaxf(mx, z[c]);
    }
    float s = 0.f, p[32];
    for (int c = 0; c < 32; ++c) { p[c] = c < C ? expf(z[c] - mx) : 0.f; s += p[c]; }
    int best = 0;
    float bv = p[0] / s;
    for (int c = 1; c < 32; ++c) {
      float pc = p[c] / s;
      if (c < C && pc > bv) { bv = pc; best = c; }
    }
    labels[v] = (uint8_t)best;
    if (mean_out)
      for (int c = 0; c < C; ++c) mean_out[v * C + c] = z[c];
  }
}

int launch_window_accumulate(int dtype, const void* src, int src_ld, const float* imp, float* acc, float* cnt, int C,
                             int wd, int wh, int ww, int D, int H, int W, int d0, int h0, int w0, cudaStream_t st) {
  int cd = min(wd, D - d0), ch = min(wh, H - h0), cw = min(ww, W - w0);
  if (cd <= 0 || ch <= 0 || cw <= 0) return B200SEG_OK;
  int64_t total = (int64_t)cd * ch * cw;
  int64_t nb = cdiv64(total, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  if (dtype == B200SEG_BF16)
    window_accumulate_kernel<__nv_bfloat16><<<(unsigned)nb, 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, imp, acc,
                                                                         cnt, C, wd, wh, ww, H, W, d0, h0, w0, cd, ch, cw);
  else
    window_accumulate_kernel<float><<<(unsigned)nb, 256, 0, st>>>((const float*)src, src_ld, imp, acc, cnt, C, wd, wh,
                                                                 ww, H, W, d0, h0, w0, cd, ch, cw);
  B200SEG_CHECK_LAUNCH("window_accumulate");
  return B200SEG_OK;
}

int launch_accum_argmax(const float* acc, const float* cnt, uint8_t* labels, float* mean_out, int64_t nvox, int C,
                        cudaStream_t st) {
  int64_t nb = cdiv64(nvox, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  accum_argmax_kernel<<<(unsigned)nb, 256, 0, st>>>(acc, cnt, labels, mean_out, nvox, C);
  B200SEG_CHECK_LAUNCH("accum_argmax");
  return B200SEG_OK;
}

// ---- patch sampler: crop + HU window + normalise (+ label crop), out-of-volume = padding -----
template <typename T>
__global__ void crop_window_norm_kernel(const int16_t* __restrict__ hu, const uint8_t* __restrict__ lab,
                                        const int* __restrict__ origins, T* __restrict__ img_out,
                                        uint8_t* __restrict__ lab_out, int D, int H, int W, int pd, int ph,
                                        int pw, float lo, float hi, float mean, float stdv, int16_t pad_hu) {
  const int b = blockIdx.y;
  const int od = origins[b * 3 + 0], oh = origins[b * 3 + 1], ow = origins[b * 3 + 2];
  const int64_t total = (int64_t)pd * ph * pw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % pw), h = (int)((i / pw) % ph), d = (int)(i / ((int64_t)pw * ph));
    const int sd = od + d, sh = oh + h, sw = ow + w;
    const bool in = sd >= 0 && sd < D && sh >= 0 && sh < H && sw >= 0 && sw < W;
    const int64_t si = ((int64_t)sd * H + sh) * W + sw;
    const float x = in ? (float)hu[si] : (float)pad_hu;
    double c = fmin(fmax((double)x, (double)lo), (double)hi);
    float f = (float)((c - (double)lo) / ((double)hi - (double)lo + 1e-8));
    f = (f - mean) * (1.0f / stdv);
    img_out[(int64_t)b * total + i] = from_f<T>(f);
    if (lab_out) lab_out[(int64_t)b * total + i] = (in && lab) ? lab[si] : (uint8_t)0;
  }
}

int launch_crop_window_norm(int dtype, const int16_t* hu, const uint8_t* lab, const int* origins, int nb_patches,
                            void* img_out, uint8_t* lab_out, int D, int H, int W, int pd, int ph, int pw, float lo,
                            float hi, float mean, float stdv, int pad_hu, cudaStream_t st) {
  int64_t total = (int64_t)pd * ph * pw;
  int64_t nb = cdiv64(total, 256 * 4);
  if (nb > 148 * 8) nb = 148 * 8;
  dim3 grid((unsigned)nb, (unsigned)nb_patches);
  if (dtype == B200SEG_BF16)
    crop_window_norm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(hu, lab, origins, (__nv_bfloat16*)img_out, lab_out, D, H,
                                                                 W, pd, ph, pw, lo, hi, mean, stdv, (int16_t)pad_hu);
  else
    crop_window_norm_kernel<float><<<grid, 256, 0, st>>>(hu, lab, origins, (float*)img_out, lab_out, D, H, W, pd, ph, pw,
                                                         lo, hi, mean, stdv, (int16_t)pad_hu);
  B200SEG_CHECK_LAUNCH("crop_window_norm");
  return B200SEG_OK;
}

// ------------------------------------------------------------------------------------------
#define DISPATCH_DICE(d, ...)                                                              \
  do {                                                                                     \
    if ((d).c > 32 || (d).c < 1) {                                                         \
      set_error("softmax/dice kernels support 1..32 classes, got %d", (d).c);              \
      return B200SEG_ERR_UNSUPPORTED;                                                      \
    }                                                                                      \
    const bool bf = (d).dtype == B200SEG_BF16, u8 = (d).label_dtype == B200SEG_LABEL_U8;   \
    const bool small = (d).c <= 16;                                                        \
    if (bf && u8 && small)        { using T = __nv_bfloat16; constexpr int CM = 16; constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else if (bf && u8)            { using T = __nv_bfloat16; constexpr int CM = 32; constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else if (bf && small)         { using T = __nv_bfloat16; constexpr int CM = 16; constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
    else if (bf)                  { using T = __nv_bfloat16; constexpr int CM = 32; constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
    else if (u8 && small)         { using T = float;         constexpr int CM = 16; constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else if (u8)                  { using T = float;         constexpr int CM = 32; constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else if (small)               { using T = float;         constexpr int CM = 16; constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
    else                          { using T = float;         constexpr int CM = 32; constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
  } while (0)

// the reference's 10 classes (9 structures + background) as a compile-time constant: the padded
// 16-wide loops shrink to exactly 10 (these kernels are instruction-bound, not memory-bound)
#define DISPATCH_DICE10(d, ...)                                                            \
  do {                                                                                     \
    const bool bf = (d).dtype == B200SEG_BF16, u8 = (d).label_dtype == B200SEG_LABEL_U8;   \
    if (bf && u8)      { using T = __nv_bfloat16; constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else if (bf)       { using T = __nv_bfloat16; constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
    else if (u8)       { using T = float;         constexpr int LT = B200SEG_LABEL_U8;  __VA_ARGS__; } \
    else               { using T = float;         constexpr int LT = B200SEG_LABEL_I64; __VA_ARGS__; } \
  } while (0)

// the bulk-copy staged kernels take the production layout: bf16, 16-channel rows, the reference's 10 classes
static bool ring_ok(const b200seg_dice_desc& d, const void* a, const void* b) {
  return d.c == 10 && d.ld == 16 && vec16_ok(d, a, b) && std::getenv("B200SEG_DICE_NO_RING") == nullptr;
}

static int64_t ring_vox_per_block(const b200seg_dice_desc& d, int nb) {
  return cdiv64(cdiv64(d.spatial, nb), RING_VOX) * RING_VOX;
}

int launch_softmax_dice_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                            float* sums, void* ws, cudaStream_t st) {
  int nb = dice_blocks(d.spatial, d.n);
  int64_t per = cdiv64(d.spatial, nb);
  dim3 grid(nb, d.n);
  float* partial = (float*)ws;
  if (ring_ok(d, logits, nullptr)) {
    per = ring_vox_per_block(d, nb);
    if (d.label_dtype == B200SEG_LABEL_U8)
      softmax_dice_fwd_ring_kernel<10, B200SEG_LABEL_U8, false><<<grid, RING_VOX, 0, st>>>(
          (const __nv_bfloat16*)logits, labels, d.spatial, per, partial);
    else
      softmax_dice_fwd_ring_kernel<10, B200SEG_LABEL_I64, false><<<grid, RING_VOX, 0, st>>>(
          (const __nv_bfloat16*)logits, labels, d.spatial, per, partial);
    B200SEG_CHECK_LAUNCH("softmax_dice_fwd_ring");
  } else {
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_dice_fwd_kernel<T, 16, LT, 10><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, d.spatial, d.c, d.ld, per, partial, vec16_ok(d, logits, nullptr))));
  } else
  DISPATCH_DICE(d, (softmax_dice_fwd_kernel<T, CM, LT><<<grid, kDiceThreads, 0, st>>>(
                       (const T*)logits, labels, d.spatial, d.c, d.ld, per, partial, vec16_ok(d, logits, nullptr))));
  B200SEG_CHECK_LAUNCH("softmax_dice_fwd");
  }
  int total = d.n * d.c * 3;
  dice_sums_final_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(partial, nb, d.c * 3, total, sums);
  B200SEG_CHECK_LAUNCH("dice_sums_final");
  return B200SEG_OK;
}

size_t dice_metric_workspace_bytes(const b200seg_dice_desc& d) {
  return (size_t)d.n * dice_blocks(d.spatial, d.n) * d.c * 5 * sizeof(float) + 256;
}

// Dice sums AND Dice-metric counts.  Production layout: ONE pass (ring kernel with COUNTS); any other layout /
// dtype (fp32 check mode, unpadded rows, other class counts): the loss kernel, then the metric kernel.
int launch_softmax_dice_metric_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float* sums,
                                   int64_t* counts, void* ws, cudaStream_t st) {
  if (!ring_ok(d, logits, nullptr)) {
    int rc = launch_softmax_dice_fwd(d, logits, labels, sums, ws, st);
    if (rc) return rc;
    return launch_argmax_dice_counts(d, logits, labels, nullptr, counts, st);
  }
  const int nb = dice_blocks(d.spatial, d.n);
  const int64_t per = ring_vox_per_block(d, nb);
  dim3 grid(nb, d.n);
  float* partial = (float*)ws;
  if (d.label_dtype == B200SEG_LABEL_U8)
    softmax_dice_fwd_ring_kernel<10, B200SEG_LABEL_U8, true><<<grid, RING_VOX, 0, st>>>(
        (const __nv_bfloat16*)logits, labels, d.spatial, per, partial);
  else
    softmax_dice_fwd_ring_kernel<10, B200SEG_LABEL_I64, true><<<grid, RING_VOX, 0, st>>>(
        (const __nv_bfloat16*)logits, labels, d.spatial, per, partial);
  B200SEG_CHECK_LAUNCH("softmax_dice_metric_fwd_ring");
  const int total = d.n * d.c * 5;
  dice_metric_final_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(partial, nb, d.c, total, sums,
                                                                      (long long*)counts);
  B200SEG_CHECK_LAUNCH("dice_metric_final");
  return B200SEG_OK;
}

int launch_softmax_dice_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                            const float* gI, const float* gP, void* dlogits, cudaStream_t st) {
  if (ring_ok(d, logits, dlogits)) {
    int64_t nb = cdiv64(d.spatial, RING_VOX * 8);
    int64_t cap = 2368 / (d.n > 0 ? d.n : 1);
    if (cap < 1) cap = 1;
    if (nb > cap) nb = cap;
    const int64_t per = ring_vox_per_block(d, (int)nb);
    dim3 grid((unsigned)cdiv64(d.spatial, per), d.n);
    if (d.label_dtype == B200SEG_LABEL_U8)
      softmax_dice_bwd_ring_kernel<10, B200SEG_LABEL_U8><<<grid, RING_VOX, 0, st>>>(
          (const __nv_bfloat16*)logits, labels, gI, gP, (__nv_bfloat16*)dlogits, d.spatial, per);
    else
      softmax_dice_bwd_ring_kernel<10, B200SEG_LABEL_I64><<<grid, RING_VOX, 0, st>>>(
          (const __nv_bfloat16*)logits, labels, gI, gP, (__nv_bfloat16*)dlogits, d.spatial, per);
    B200SEG_CHECK_LAUNCH("softmax_dice_bwd_ring");
    return B200SEG_OK;
  }
  int64_t nb = cdiv64(d.spatial, kDiceThreads * 2);
  int64_t cap = 4736 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  dim3 grid((unsigned)nb, d.n);
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_dice_bwd_kernel<T, 16, LT, 10><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, gI, gP, (T*)dlogits, d.spatial, d.c, d.ld,
                           vec16_ok(d, logits, dlogits))));
  } else
  DISPATCH_DICE(d, (softmax_dice_bwd_kernel<T, CM, LT><<<grid, kDiceThreads, 0, st>>>(
                       (const T*)logits, labels, gI, gP, (T*)dlogits, d.spatial, d.c, d.ld,
                       vec16_ok(d, logits, dlogits))));
  B200SEG_CHECK_LAUNCH("softmax_dice_bwd");
  return B200SEG_OK;
}

size_t loss_workspace_bytes(const b200seg_dice_desc& d) {
  return (size_t)d.n * dice_blocks(d.spatial, d.n) * d.c * 5 * sizeof(float) + 256;
}

int launch_softmax_loss_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float gamma,
                            float* sums5, void* ws, cudaStream_t st) {
  int nb = dice_blocks(d.spatial, d.n);
  int64_t per = cdiv64(d.spatial, nb);
  dim3 grid(nb, d.n);
  float* partial = (float*)ws;
  if (d.c > 16) { set_error("softmax_loss: at most 16 classes, got %d", d.c); return B200SEG_ERR_UNSUPPORTED; }
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_loss_fwd_kernel<T, 16, LT, 10><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, d.spatial, d.c, d.ld, per, gamma, partial,
                           vec16_ok(d, logits, nullptr))));
  } else {
    DISPATCH_DICE10(d, (softmax_loss_fwd_kernel<T, 16, LT, 0><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, d.spatial, d.c, d.ld, per, gamma, partial,
                           vec16_ok(d, logits, nullptr))));
  }
  B200SEG_CHECK_LAUNCH("softmax_loss_fwd");
  int total = d.n * d.c * 5;
  dice_sums_final_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(partial, nb, d.c * 5, total, sums5);
  B200SEG_CHECK_LAUNCH("dice_sums_final");
  return B200SEG_OK;
}

size_t boundary_workspace_bytes(const b200seg_dice_desc& d) {
  return (size_t)d.n * dice_blocks(d.spatial, d.n) * d.c * 6 * sizeof(float) + 256;
}

int launch_softmax_boundary_loss_fwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                                     const float* dist, float gamma, float* sums6, void* ws, cudaStream_t st) {
  int nb = dice_blocks(d.spatial, d.n);
  int64_t per = cdiv64(d.spatial, nb);
  dim3 grid(nb, d.n);
  float* partial = (float*)ws;
  if (d.c > 16 || d.c < 2) { set_error("softmax_boundary_loss: 2..16 classes, got %d", d.c); return B200SEG_ERR_UNSUPPORTED; }
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_loss_fwd_kernel<T, 16, LT, 10, true><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, d.spatial, d.c, d.ld, per, gamma, partial,
                           vec16_ok(d, logits, nullptr), dist)));
  } else {
    DISPATCH_DICE10(d, (softmax_loss_fwd_kernel<T, 16, LT, 0, true><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, d.spatial, d.c, d.ld, per, gamma, partial,
                           vec16_ok(d, logits, nullptr), dist)));
  }
  B200SEG_CHECK_LAUNCH("softmax_boundary_loss_fwd");
  int total = d.n * d.c * 6;
  dice_sums_final_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(partial, nb, d.c * 6, total, sums6);
  B200SEG_CHECK_LAUNCH("dice_sums_final");
  return B200SEG_OK;
}

int launch_softmax_boundary_loss_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels,
                                     const float* dist, float gamma, const float* gI, const float* gP,
                                     const float* gF, const float* gN, const float* gB, void* dlogits,
                                     cudaStream_t st) {
  int64_t nb = cdiv64(d.spatial, kDiceThreads * 2);
  int64_t cap = 4736 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  dim3 grid((unsigned)nb, d.n);
  if (d.c > 16 || d.c < 2) { set_error("softmax_boundary_loss: 2..16 classes, got %d", d.c); return B200SEG_ERR_UNSUPPORTED; }
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_loss_bwd_kernel<T, 16, LT, 10, true><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, gI, gP, gF, gN, (T*)dlogits, d.spatial, d.c, d.ld, gamma,
                           vec16_ok(d, logits, dlogits), gB, dist)));
  } else {
    DISPATCH_DICE10(d, (softmax_loss_bwd_kernel<T, 16, LT, 0, true><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, gI, gP, gF, gN, (T*)dlogits, d.spatial, d.c, d.ld, gamma,
                           vec16_ok(d, logits, dlogits), gB, dist)));
  }
  B200SEG_CHECK_LAUNCH("softmax_boundary_loss_bwd");
  return B200SEG_OK;
}

int launch_softmax_loss_bwd(const b200seg_dice_desc& d, const void* logits, const void* labels, float gamma,
                            const float* gI, const float* gP, const float* gF, const float* gN, void* dlogits,
                            cudaStream_t st) {
  int64_t nb = cdiv64(d.spatial, kDiceThreads * 2);
  int64_t cap = 4736 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  dim3 grid((unsigned)nb, d.n);
  if (d.c > 16) { set_error("softmax_loss: at most 16 classes, got %d", d.c); return B200SEG_ERR_UNSUPPORTED; }
  if (d.c == 10) {
    DISPATCH_DICE10(d, (softmax_loss_bwd_kernel<T, 16, LT, 10><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, gI, gP, gF, gN, (T*)dlogits, d.spatial, d.c, d.ld, gamma,
                           vec16_ok(d, logits, dlogits))));
  } else {
    DISPATCH_DICE10(d, (softmax_loss_bwd_kernel<T, 16, LT, 0><<<grid, kDiceThreads, 0, st>>>(
                           (const T*)logits, labels, gI, gP, gF, gN, (T*)dlogits, d.spatial, d.c, d.ld, gamma,
                           vec16_ok(d, logits, dlogits))));
  }
  B200SEG_CHECK_LAUNCH("softmax_loss_bwd");
  return B200SEG_OK;
}

int launch_argmax_dice_counts(const b200seg_dice_desc& d, const void* logits, const void* target,
                              uint8_t* pred_out, int64_t* counts, cudaStream_t st) {
  if (target) {
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)d.n * d.c * 3 * sizeof(int64_t), st);
    if (e != cudaSuccess) {
      set_error("argmax_dice_counts: memset failed: %s", cudaGetErrorString(e));
      return B200SEG_ERR_CUDA;
    }
  }
  int64_t nb = cdiv64(d.spatial, kDiceThreads * 4);
  int64_t cap = 2368 / (d.n > 0 ? d.n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  dim3 grid((unsigned)nb, d.n);
  DISPATCH_DICE(d, (argmax_counts_kernel<T, CM, LT><<<grid, kDiceThreads, 0, st>>>(
                       (const T*)logits, target, pred_out, (unsigned long long*)counts, d.spatial,
                       d.c, d.ld, vec16_ok(d, logits, nullptr))));
  B200SEG_CHECK_LAUNCH("argmax_counts");
  return B200SEG_OK;
}

int launch_label_dice_counts(int n, int64_t spatial, int c, const uint8_t* pred, const void* target,
                             int target_dtype, int64_t* counts, cudaStream_t st) {
  if (c > 32 || c < 1) {
    set_error("label_dice_counts supports 1..32 classes, got %d", c);
    return B200SEG_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)n * c * 3 * sizeof(int64_t), st);
  if (e != cudaSuccess) {
    set_error("label_dice_counts: memset failed: %s", cudaGetErrorString(e));
    return B200SEG_ERR_CUDA;
  }
  int64_t nb = cdiv64(spatial, kDiceThreads * 8);
  int64_t cap = 2368 / (n > 0 ? n : 1);
  if (cap < 1) cap = 1;
  if (nb > cap) nb = cap;
  dim3 grid((unsigned)nb, n);
  if (target_dtype == B200SEG_LABEL_U8)
    label_counts_kernel<B200SEG_LABEL_U8><<<grid, kDiceThreads, 0, st>>>(pred, target, (unsigned long long*)counts, spatial, c);
  else
    label_counts_kernel<B200SEG_LABEL_I64><<<grid, kDiceThreads, 0, st>>>(pred, target, (unsigned long long*)counts, spatial, c);
  B200SEG_CHECK_LAUNCH("label_counts");
  return B200SEG_OK;
}

int launch_squash_masks(int n, int n_struct, int64_t spatial, const uint8_t* masks, uint8_t* labels,
                        cudaStream_t st) {
  int64_t total = (int64_t)n * spatial;
  int64_t nb = cdiv64(total, 256 * 4);
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  squash_masks_kernel<<<(unsigned)nb, 256, 0, st>>>(masks, labels, n_struct, spatial, total);
  B200SEG_CHECK_LAUNCH("squash_masks");
  return B200SEG_OK;
}

int launch_hu_window_norm(int64_t n_vox, int n_windows, const int16_t* hu, const float* lo,
                          const float* hi, const float* mean, const float* std_, void* out,
                          int out_ld, int dtype, cudaStream_t st) {
  WindowCfg w;
  w.n = n_windows;
  for (int k = 0; k < 4; ++k) {
    w.lo[k] = k < n_windows ? lo[k] : 0.f;
    w.hi[k] = k < n_windows ? hi[k] : 1.f;
    w.mean[k] = k < n_windows ? mean[k] : 0.f;
    w.std_[k] = k < n_windows ? std_[k] : 1.f;
  }
  int64_t nb = cdiv64(n_vox, 256 * 4);
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  if (dtype == B200SEG_BF16)
    hu_window_norm_kernel<__nv_bfloat16><<<(unsigned)nb, 256, 0, st>>>(hu, (__nv_bfloat16*)out, n_vox, out_ld, w);
  else
    hu_window_norm_kernel<float><<<(unsigned)nb, 256, 0, st>>>(hu, (float*)out, n_vox, out_ld, w);
  B200SEG_CHECK_LAUNCH("hu_window_norm");
  return B200SEG_OK;
}

}  // namespace b200seg

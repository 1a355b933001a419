// tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 in, fp32 accumulate in TMEM, bf16 out).
//
// One kernel covers conv fprop (stride 1/2), conv dgrad (stride 1/2), ConvTranspose fprop and
// ConvTranspose dgrad: all are "for every destination voxel, sum over a list of (source offset,
// weight tile) steps".  Per CTA: an output tile of 128 destination voxels (a tD x tH x tW box of the
// destination class grid) x BN destination channels.
//   warp 0   : TMA producer  -- per step and K block, one 5-D box load of the source tile
//              (channels-last tensor map, out-of-bounds = zero = conv padding) and one 2-D box load
//              of the weight tile, into a multi-stage shared-memory ring (mbarrier full/empty).
//   warp 1   : allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16 per instruction) from the
//              swizzled K-major shared tiles; tcgen05.commit releases ring slots / signals epilogue.
//   warps 2-5: epilogue -- tcgen05.ld the fp32 accumulator rows, add bias / residual / previous
//              destination (gradient fan-in), convert to bf16, 32-byte vector stores.
// Stride-2 gathers never touch inserted zeros or skipped taps: forward gathers read one of 8
// parity-subsampled tensor maps (doubled strides), transposed gathers are decomposed into 8
// destination parity classes with 1..8 contributing taps each.
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace b200seg {

using bf16 = __nv_bfloat16;

struct TcStep {
  int8_t map, dd, dh, dw;
  int32_t wtile;
};

struct alignas(64) TcConvParams {
  CUtensorMap tmA[8];
  CUtensorMap tmB;
  TcStep steps[27];
  int cls_begin[9];
  int ncls_d, ncls_h, ncls_w;
  int n;
  int cD, cH, cW;
  int tD, tH, tW;
  int tilesD, tilesH, tilesW;
  int dD, dH, dW;
  int os_d, os_h, os_w;
  int kblocks;
  int cout, cout_pad;
  int dst_ld, res_ld;
  int accumulate;
  int stages;
  const float* bias;
  const bf16* res;
  bf16* dst;
  float* stats;  // optional [blockIdx.x][cout][2]: per-CTA sum / sum of squares of its outputs (InstanceNorm)
  int ksplit;    // split-K kernel only: CTAs per cluster sharing one output tile (appended: offsets above unchanged)
};

template <int BN, int KC>
struct TcCfg {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_STRIDE;
  static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int ROW_BYTES = KC * 2;
};

template <int BN, int KC>
__global__ void __launch_bounds__(192)
tc_conv_kernel(const __grid_constant__ TcConvParams p) {
  using Cfg = TcCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * Cfg::STAGE_BYTES);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;

  // ---- tile coordinates (block-uniform)
  int bx = blockIdx.x;
  const int tw_i = bx % p.tilesW; bx /= p.tilesW;
  const int th_i = bx % p.tilesH; bx /= p.tilesH;
  const int td_i = bx % p.tilesD; bx /= p.tilesD;
  const int n = bx % p.n;
  const int cls = bx / p.n;
  const int cls_w = cls % p.ncls_w, cls_h = (cls / p.ncls_w) % p.ncls_h, cls_d = cls / (p.ncls_w * p.ncls_h);
  const int d0 = td_i * p.tD, h0 = th_i * p.tH, w0 = tw_i * p.tW;
  const int n0 = blockIdx.y * BN;
  const int s_begin = p.cls_begin[cls], s_end = p.cls_begin[cls + 1];
  const int total_iters = (s_end - s_begin) * p.kblocks;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmB);
    tc::prefetch_tmap(&p.tmA[p.steps[s_begin].map]);
  }
  if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int s = s_begin; s < s_end; ++s) {
        const TcStep step = p.steps[s];
        const CUtensorMap* tm = &p.tmA[step.map];
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tc::mbar_wait(&empty[st], ph ^ 1u);
          uint8_t* a = smem + (size_t)st * Cfg::STAGE_BYTES;
          tc::mbar_expect_tx(&full[st], Cfg::A_BYTES + Cfg::B_BYTES);
          tc::tma_load_5d(a, tm, &full[st], kb * KC, w0 + step.dw, h0 + step.dh, d0 + step.dd, n);
          tc::tma_load_2d(a + Cfg::A_BYTES, &p.tmB, &full[st], 0,
                          (step.wtile * p.kblocks + kb) * p.cout_pad + n0);
          if (++st == stages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform issue loop, one elected lane issues (see umma_bf16_warp)
      const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform register: see tc::warp_uniform
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, BN, false, false);
      constexpr uint64_t layout = tc::layout_for_row_bytes(Cfg::ROW_BYTES);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < total_iters; ++it) {
        tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        const uint32_t a_addr = tc::smem_u32(smem + (size_t)st * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          const uint64_t ad = tc::make_smem_desc(a_addr + k * 32, 16, 8 * Cfg::ROW_BYTES, layout);
          const uint64_t bd = tc::make_smem_desc(b_addr + k * 32, 16, 8 * Cfg::ROW_BYTES, layout);
          tc::umma_bf16_warp(tmem_acc, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc::umma_commit_warp(&empty[st]);
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
      tc::umma_commit_warp(acc_full);
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lw = row % p.tW, lh = (row / p.tW) % p.tH, ldp = row / (p.tW * p.tH);
    const int jd = d0 + ldp, jh = h0 + lh, jw = w0 + lw;
    const bool valid = jd < p.cD && jh < p.cH && jw < p.cW;
    const int od = jd * p.os_d + cls_d, oh = jh * p.os_h + cls_h, ow = jw * p.os_w + cls_w;
    const int64_t lin = (((int64_t)n * p.dD + od) * p.dH + oh) * p.dW + ow;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < BN / 16; ++ch) {
      uint32_t v[16];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 16, v);
      tc::tmem_ld_wait();
      if (valid) {
        const int c0 = n0 + ch * 16;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.cout) f[i] += p.bias[c0 + i];
        }
        if (p.res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + lin * p.res_ld + c0);
          uint4 r0 = rp[0], r1 = rp[1];
          const __nv_bfloat162* h0_ = reinterpret_cast<const __nv_bfloat162*>(&r0);
          const __nv_bfloat162* h1_ = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 a = __bfloat1622float2(h0_[i]), b = __bfloat1622float2(h1_[i]);
            f[2 * i] += a.x; f[2 * i + 1] += a.y;
            f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
          }
        }
        uint4* op = reinterpret_cast<uint4*>(p.dst + lin * p.dst_ld + c0);
        if (p.accumulate) {
          uint4 r0 = op[0], r1 = op[1];
          const __nv_bfloat162* h0_ = reinterpret_cast<const __nv_bfloat162*>(&r0);
          const __nv_bfloat162* h1_ = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 a = __bfloat1622float2(h0_[i]), b = __bfloat1622float2(h1_[i]);
            f[2 * i] += a.x; f[2 * i + 1] += a.y;
            f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
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
        op[0] = o0;
        op[1] = o1;
      }
      if (p.stats) {  // warp-uniform: per-channel sum / sum of squares over the tile's valid rows
        float* sred = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][BN][2]
        const int c0 = n0 + ch * 16;
        float x[16], x2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          x[i] = valid ? __uint_as_float(v[i]) : 0.f;
          if (valid && p.bias && c0 + i < p.cout) x[i] += p.bias[c0 + i];
          x2[i] = x[i] * x[i];
        }
        // 32 shuffles per 16-channel chunk (recursive halving), not 160: with a handful of CTAs on the
        // deep layers this epilogue was as long as the whole main loop
        const float a = warp_sum16(x, lane), b = warp_sum16(x2, lane);
        if ((lane & 1) == 0) {
          const int c = (lane >> 1) & 15;
          sred[(q * BN + ch * 16 + c) * 2] = a;
          sred[(q * BN + ch * 16 + c) * 2 + 1] = b;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (p.stats) {
    const float* sred = reinterpret_cast<const float*>(tmem_slot + 4);
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
      const int c = i >> 1, m = i & 1;
      if (n0 + c < p.cout)
        p.stats[((int64_t)blockIdx.x * p.cout + n0 + c) * 2 + m] =
            sred[(0 * BN + c) * 2 + m] + sred[(1 * BN + c) * 2 + m] + sred[(2 * BN + c) * 2 + m] +
            sred[(3 * BN + c) * 2 + m];
    }
  }
  if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
// Split-K variant for the deep layers (default on; B200SEG_CONV_SPLITK=0 / B200SEG_CONV_NO_SPLIT_K switch it off).  The 8^3 layers have 8 output tiles x 4..8 channel tiles = 32..64 CTAs,
// each streaming 27 taps x Cin of A and B tiles through ONE SM's L2 port (2.2 MB per CTA for 256 -> 256:
// ~16 us at ~70 B/clk, whatever the tensor core does).  Here a thread-block cluster of `ksplit` CTAs
// (cluster dims (1, 1, ksplit), blockIdx.z = rank) shares one output tile: every CTA runs the same
// producer / MMA pipeline over 1/ksplit of the (tap, K-block) iterations into its own TMEM accumulator;
// ranks > 0 then push their fp32 partial tile into rank 0's shared memory (st.shared::cluster, staged
// column-major so that a warp writes 128 contiguous bytes), the cluster synchronises, and rank 0 adds the
// partials in rank order (deterministic) and runs the usual epilogue (bias / residual / accumulate /
// statistics).  Per-CTA traffic and MMA count drop by ksplit; the cost is one cluster barrier and
// (ksplit - 1) x 128 x BN x 4 bytes over the SM-to-SM network.
// ------------------------------------------------------------------------------------------------
namespace tcx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
}  // namespace tcx

template <int BN, int KC>
__global__ void __launch_bounds__(192)
tc_conv_splitk_kernel(const __grid_constant__ TcConvParams p) {
  using Cfg = TcCfg<BN, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * Cfg::STAGE_BYTES);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* sred = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][BN][2]
  float* staging = sred + 4 * BN * 2;                     // [ksplit - 1][BN][128]: partial tiles of ranks 1..

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  const int nsplit = p.ksplit;
  const int z = (int)tcx::cluster_ctarank();  // == blockIdx.z (cluster dims (1, 1, ksplit), gridDim.z == ksplit)

  int bx = blockIdx.x;
  const int tw_i = bx % p.tilesW; bx /= p.tilesW;
  const int th_i = bx % p.tilesH; bx /= p.tilesH;
  const int td_i = bx % p.tilesD; bx /= p.tilesD;
  const int n = bx % p.n;
  const int cls = bx / p.n;
  const int cls_w = cls % p.ncls_w, cls_h = (cls / p.ncls_w) % p.ncls_h, cls_d = cls / (p.ncls_w * p.ncls_h);
  const int d0 = td_i * p.tD, h0 = th_i * p.tH, w0 = tw_i * p.tW;
  const int n0 = blockIdx.y * BN;
  const int s_begin = p.cls_begin[cls], s_end = p.cls_begin[cls + 1];
  // this rank's share of the (step, K-block) iterations: [it0, it1)
  const int all_iters = (s_end - s_begin) * p.kblocks;
  const int per = (all_iters + nsplit - 1) / nsplit;
  const int it0 = min(z * per, all_iters), it1 = min(it0 + per, all_iters);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmB);
  }
  if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  // every CTA of the cluster is running before anyone touches a peer's shared memory
  tcx::cluster_arrive();
  tcx::cluster_wait();

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int it = it0; it < it1; ++it) {
        const int s = s_begin + it / p.kblocks, kb = it % p.kblocks;
        const TcStep step = p.steps[s];
        tc::mbar_wait(&empty[st], ph ^ 1u);
        uint8_t* a = smem + (size_t)st * Cfg::STAGE_BYTES;
        tc::mbar_expect_tx(&full[st], Cfg::A_BYTES + Cfg::B_BYTES);
        tc::tma_load_5d(a, &p.tmA[step.map], &full[st], kb * KC, w0 + step.dw, h0 + step.dh, d0 + step.dd, n);
        tc::tma_load_2d(a + Cfg::A_BYTES, &p.tmB, &full[st], 0, (step.wtile * p.kblocks + kb) * p.cout_pad + n0);
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform register: see tc::warp_uniform
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, BN, false, false);
    constexpr uint64_t layout = tc::layout_for_row_bytes(Cfg::ROW_BYTES);
    int st = 0;
    uint32_t ph = 0;
    for (int it = it0; it < it1; ++it) {
      tc::mbar_wait(&full[st], ph);
      tc::tc_fence_after();
      const uint32_t a_addr = tc::smem_u32(smem + (size_t)st * Cfg::STAGE_BYTES);
      const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
      for (int k = 0; k < KC / 16; ++k) {
        const uint64_t ad = tc::make_smem_desc(a_addr + k * 32, 16, 8 * Cfg::ROW_BYTES, layout);
        const uint64_t bd = tc::make_smem_desc(b_addr + k * 32, 16, 8 * Cfg::ROW_BYTES, layout);
        tc::umma_bf16_warp(tmem_acc, ad, bd, idesc, (it > it0 || k > 0) ? 1u : 0u);
      }
      tc::umma_commit_warp(&empty[st]);
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    tc::umma_commit_warp(acc_full);
  } else if (z != 0) {
    // ---- ranks > 0: push the fp32 partial tile into rank 0's staging area, [col][row] so that the 32 lanes
    // of a warp (32 consecutive rows) write 128 contiguous bytes per column
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t base = tcx::map_to_rank(tc::smem_u32(staging), 0) + (uint32_t)((z - 1) * BN * 128 + row) * 4u;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < BN / 16; ++ch) {
      uint32_t v[16];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 16, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        tcx::st_cluster_f32(base + (uint32_t)((ch * 16 + i) * 128) * 4u, it1 > it0 ? __uint_as_float(v[i]) : 0.f);
    }
  }
  // ---- partial tiles are in rank 0's shared memory (release / acquire at cluster scope)
  tc::tc_fence_before();
  tcx::cluster_arrive();
  tcx::cluster_wait();

  if (warp >= 2 && z == 0) {
    // ---- rank 0 epilogue: own accumulator + the peers' partials (rank order), then as tc_conv_kernel
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lw = row % p.tW, lh = (row / p.tW) % p.tH, ldp = row / (p.tW * p.tH);
    const int jd = d0 + ldp, jh = h0 + lh, jw = w0 + lw;
    const bool valid = jd < p.cD && jh < p.cH && jw < p.cW;
    const int od = jd * p.os_d + cls_d, oh = jh * p.os_h + cls_h, ow = jw * p.os_w + cls_w;
    const int64_t lin = (((int64_t)n * p.dD + od) * p.dH + oh) * p.dW + ow;
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < BN / 16; ++ch) {
      uint32_t v[16];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 16, v);
      tc::tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
      for (int r = 1; r < nsplit; ++r) {
        const float* sp = staging + ((size_t)(r - 1) * BN + ch * 16) * 128 + row;
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += sp[i * 128];
      }
      const int c0 = n0 + ch * 16;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.cout) f[i] += p.bias[c0 + i];
      }
      if (p.stats) {  // statistics of conv + bias, before the residual / accumulate addends (as tc_conv_kernel)
        float x[16], x2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          x[i] = valid ? f[i] : 0.f;
          x2[i] = x[i] * x[i];
        }
        const float a = warp_sum16(x, lane), b = warp_sum16(x2, lane);
        if ((lane & 1) == 0) {
          const int c = (lane >> 1) & 15;
          sred[(q * BN + ch * 16 + c) * 2] = a;
          sred[(q * BN + ch * 16 + c) * 2 + 1] = b;
        }
      }
      if (valid) {
        if (p.res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + lin * p.res_ld + c0);
          uint4 r0 = rp[0], r1 = rp[1];
          const __nv_bfloat162* h0_ = reinterpret_cast<const __nv_bfloat162*>(&r0);
          const __nv_bfloat162* h1_ = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 a = __bfloat1622float2(h0_[i]), b = __bfloat1622float2(h1_[i]);
            f[2 * i] += a.x; f[2 * i + 1] += a.y;
            f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
          }
        }
        uint4* op = reinterpret_cast<uint4*>(p.dst + lin * p.dst_ld + c0);
        if (p.accumulate) {
          uint4 r0 = op[0], r1 = op[1];
          const __nv_bfloat162* h0_ = reinterpret_cast<const __nv_bfloat162*>(&r0);
          const __nv_bfloat162* h1_ = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 a = __bfloat1622float2(h0_[i]), b = __bfloat1622float2(h1_[i]);
            f[2 * i] += a.x; f[2 * i + 1] += a.y;
            f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
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
        op[0] = o0;
        op[1] = o1;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (p.stats && z == 0) {
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
      const int c = i >> 1, m = i & 1;
      if (n0 + c < p.cout)
        p.stats[((int64_t)blockIdx.x * p.cout + n0 + c) * 2 + m] =
            sred[(0 * BN + c) * 2 + m] + sred[(1 * BN + c) * 2 + m] + sred[(2 * BN + c) * 2 + m] +
            sred[(3 * BN + c) * 2 + m];
    }
  }
  if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[12];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

CUtensorMapSwizzle swizzle_for_row_bytes(int b) {
  return b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 tensor map of rank `rank`; dims/strides(bytes, dims 1..)/box in elements
int make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box, int row_bytes) {
  MapKey key{};
  key.v[0] = (uint64_t)(uintptr_t)base;
  key.v[1] = (uint64_t)rank | ((uint64_t)row_bytes << 8);
  for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i] | ((uint64_t)box[i] << 40);
  for (int i = 0; i < rank - 1; ++i) key.v[7 + i] = strides[i];
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return B200SEG_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B200SEG_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(row_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return B200SEG_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return B200SEG_OK;
}

inline int round16(int c) { return (c + 15) / 16 * 16; }

struct TcGeom {
  // source (gathered) and destination tensors of the op
  int n, sD, sH, sW, dD, dH, dW, src_c, dst_c, src_ld, dst_ld;
  int k[3], s[3], p[3];
  bool transposed;  // src = (dst + p - k)/s  (else src = dst*s - p + k)
};

void tc_geom(const b200seg_conv_desc* d, int op, TcGeom& g) {
  g.n = d->n;
  g.k[0] = d->kd; g.k[1] = d->kh; g.k[2] = d->kw;
  g.s[0] = d->sd; g.s[1] = d->sh; g.s[2] = d->sw;
  g.p[0] = d->pd; g.p[1] = d->ph; g.p[2] = d->pw;
  bool src_is_x = (op == TC_CONV_FPROP || op == TC_CONVTR_FPROP);
  if (src_is_x) {
    g.sD = d->in_d; g.sH = d->in_h; g.sW = d->in_w; g.dD = d->out_d; g.dH = d->out_h; g.dW = d->out_w;
    g.src_c = d->cin; g.dst_c = d->cout; g.src_ld = d->x_ld; g.dst_ld = d->y_ld;
  } else {
    g.sD = d->out_d; g.sH = d->out_h; g.sW = d->out_w; g.dD = d->in_d; g.dH = d->in_h; g.dW = d->in_w;
    g.src_c = d->cout; g.dst_c = d->cin; g.src_ld = d->y_ld; g.dst_ld = d->x_ld;
  }
  g.transposed = (op == TC_CONV_DGRAD || op == TC_CONVTR_FPROP);
}

int kc_for(int src_pad) { return src_pad % 64 == 0 ? 64 : (src_pad % 32 == 0 ? 32 : 16); }
int bn_for(int dst_pad) { return dst_pad % 128 == 0 ? 128 : (dst_pad % 64 == 0 ? 64 : (dst_pad % 32 == 0 ? 32 : 16)); }

template <int BN, int KC>
int launch_cfg(const TcConvParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN, KC>;
  size_t smem = 1024 + (size_t)p.stages * Cfg::STAGE_BYTES + (2 * p.stages + 1) * 8 + 32 + 4 * BN * 2 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_conv_kernel<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  tc_conv_kernel<BN, KC><<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_conv");
  count_tc_launch();
  return B200SEG_OK;
}

template <int BN, int KC>
int launch_splitk(TcConvParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN, KC>;
  const size_t fixed = 1024 + (2 * 8 + 1) * 8 + 32 + 4 * BN * 2 * 4 + (size_t)(p.ksplit - 1) * BN * 128 * 4;
  int stages = (int)((200 * 1024 - fixed) / Cfg::STAGE_BYTES);
  if (stages > 8) stages = 8;
  if (stages < 2) return B200SEG_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * Cfg::STAGE_BYTES + (2 * stages + 1) * 8 + 32 + 4 * BN * 2 * 4 +
                      (size_t)(p.ksplit - 1) * BN * 128 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_conv_splitk_kernel<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid.x, grid.y, (unsigned)p.ksplit);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)p.ksplit;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_conv_splitk_kernel<BN, KC>, p);
  if (e != cudaSuccess) {
    set_error("tc_conv_splitk launch failed: %s", cudaGetErrorString(e));
    return B200SEG_ERR_CUDA;
  }
  B200SEG_CHECK_LAUNCH("tc_conv_splitk");
  count_tc_launch();
  return B200SEG_OK;
}

// ring depth: ~72 KB per CTA when the grid fills the machine several times over (2-3 CTAs per SM
// share the shared memory); the whole SM when there is at most one CTA per SM (deep layers: 8..64
// CTAs, each streaming megabytes -- bytes in flight per CTA are what bounds them)
template <int BN, int KC>
int stages_for(int64_t ctas) {
  using Cfg = TcCfg<BN, KC>;
  const int budget = ctas <= 148 ? 200 * 1024 : 72 * 1024;
  int s = budget / Cfg::STAGE_BYTES;
  if (s > 8) s = 8;
  if (s < 3) s = 3;
  return s;
}

}  // namespace

size_t tc_packed_weight_bytes(const b200seg_conv_desc* d) {
  if (d->dtype != B200SEG_BF16) return 0;
  return (size_t)d->kd * d->kh * d->kw * round16(d->cin) * round16(d->cout) * 2;
}

// Wt[((tap * kblocks + kb) * dst_pad + t) * KC + kc] = W(src = kb*KC + kc, dst = t, tap)
// and, in the same launch, the generic layout  gen[tap][src][dst]  (bf16)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out, bf16* __restrict__ gen,
                                      int taps, int src_c, int dst_c, int src_pad, int dst_pad, int KC, int cin,
                                      int cout, int kind) {
  int64_t total = (int64_t)taps * src_pad * dst_pad;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gen && idx < (int64_t)taps * src_c * dst_c) {
    int t = (int)(idx % dst_c);
    int64_t r = idx / dst_c;
    int s = (int)(r % src_c), tap = (int)(r / src_c);
    int64_t wi;
    switch (kind) {
      case B200SEG_W_CONV_FPROP:   wi = ((int64_t)t * cin + s) * taps + tap; break;
      case B200SEG_W_CONV_DGRAD:   wi = ((int64_t)s * cin + t) * taps + tap; break;
      case B200SEG_W_CONVTR_FPROP: wi = ((int64_t)s * cout + t) * taps + tap; break;
      default:                     wi = ((int64_t)t * cout + s) * taps + tap; break;
    }
    gen[idx] = __float2bfloat16_rn(w[wi]);
  }
  if (idx >= total) return;
  int kc = (int)(idx % KC);
  int64_t r = idx / KC;
  int t = (int)(r % dst_pad); r /= dst_pad;
  int kblocks = src_pad / KC;
  int kb = (int)(r % kblocks);
  int tap = (int)(r / kblocks);
  int s = kb * KC + kc;
  float v = 0.f;
  if (s < src_c && t < dst_c) {
    int64_t wi;
    switch (kind) {
      case B200SEG_W_CONV_FPROP:   wi = ((int64_t)t * cin + s) * taps + tap; break;
      case B200SEG_W_CONV_DGRAD:   wi = ((int64_t)s * cin + t) * taps + tap; break;
      case B200SEG_W_CONVTR_FPROP: wi = ((int64_t)s * cout + t) * taps + tap; break;
      default:                     wi = ((int64_t)t * cout + s) * taps + tap; break;
    }
    v = w[wi];
  }
  out[idx] = __float2bfloat16_rn(v);
}

// ---- all layers in one launch: table of b200seg_pack_entry in device memory.  The entries differ in
// size by four orders of magnitude (432 ... 1.8 M elements): the work is cut into chunks of
// PACK_CHUNK packed elements, numbered across the entries, and dealt round-robin to the blocks.
// A work item is one (dst, src) channel pair: its `taps` source values are contiguous in the PyTorch
// parameter (read once, sector-efficient through L1), its packed values are `taps` strided 2-byte
// stores that neighbouring lanes complete to full sectors (tcgen05 layout: lanes along the source
// channel; generic layout: lanes along the destination channel).  Chunks of PACK_CHUNK items.
constexpr int PACK_CHUNK = 256;
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const b200seg_pack_entry* __restrict__ table, int n_entries) {
  int64_t chunk_base = 0;
  for (int ei = 0; ei < n_entries; ++ei) {
    const b200seg_pack_entry e = table[ei];
    const int taps = e.taps, cin = e.cin, cout = e.cout, kind = e.kind & 0xff;
    const bool src_is_cin = (kind == B200SEG_W_CONV_FPROP || kind == B200SEG_W_CONVTR_FPROP);
    const int src_c = src_is_cin ? cin : cout, dst_c = src_is_cin ? cout : cin;
    const int src_pad = (src_c + 15) / 16 * 16, dst_pad = (dst_c + 15) / 16 * 16;
    const int KC = src_pad % 64 == 0 ? 64 : (src_pad % 32 == 0 ? 32 : 16);
    const int kblocks = src_pad / KC;
    const int items_tc = src_pad * dst_pad;
    const int items_gen = (e.kind & B200SEG_PACK_TC_ONLY) ? 0 : src_c * dst_c;
    const int chunks_tc = items_tc / PACK_CHUNK, chunks_gen = (items_gen + PACK_CHUNK - 1) / PACK_CHUNK;
    const int nchunks = chunks_tc + chunks_gen;
    const float* w = reinterpret_cast<const float*>(e.w);
    bf16* gen = reinterpret_cast<bf16*>(e.packed);
    bf16* out = reinterpret_cast<bf16*>(e.packed + e.tc_offset);
    // PyTorch layouts: Conv (cout, cin, taps), ConvTranspose (cin, cout, taps)
    const bool conv_layer = (kind == B200SEG_W_CONV_FPROP || kind == B200SEG_W_CONV_DGRAD);
    int c = (int)(((int64_t)blockIdx.x - chunk_base % gridDim.x + gridDim.x) % gridDim.x);
    for (; c < nchunks; c += gridDim.x) {
      if (c < chunks_tc) {
        const int item = c * PACK_CHUNK + threadIdx.x;  // = t * src_pad + sc
        const int t = item / src_pad, sc = item - t * src_pad;
        const int kb = sc / KC, kc = sc - kb * KC;
        const bool live = sc < src_c && t < dst_c;
        const int co = src_is_cin ? t : sc, ci = src_is_cin ? sc : t;
        const float* wp = w + (conv_layer ? (int64_t)co * cin + ci : (int64_t)ci * cout + co) * taps;
        bf16* op = out + ((int64_t)kb * dst_pad + t) * KC + kc;
        const int64_t tap_stride = (int64_t)kblocks * dst_pad * KC;
#pragma unroll 9
        for (int tap = 0; tap < taps; ++tap) op[tap * tap_stride] = __float2bfloat16_rn(live ? wp[tap] : 0.f);
      } else {
        const int item = (c - chunks_tc) * PACK_CHUNK + threadIdx.x;  // = sc * dst_c + t
        if (item < items_gen) {
          const int sc = item / dst_c, t = item - sc * dst_c;
          const int co = src_is_cin ? t : sc, ci = src_is_cin ? sc : t;
          const float* wp = w + (conv_layer ? (int64_t)co * cin + ci : (int64_t)ci * cout + co) * taps;
          bf16* op = gen + item;
          const int64_t tap_stride = (int64_t)items_gen;
#pragma unroll 9
          for (int tap = 0; tap < taps; ++tap) op[tap * tap_stride] = __float2bfloat16_rn(wp[tap]);
        }
      }
    }
    chunk_base += nchunks;
  }
}

int tc_pack_weights_batched(const b200seg_pack_entry* table_dev, int n_entries, cudaStream_t st) {
  pack_weights_batched_kernel<<<148 * 8, 256, 0, st>>>(table_dev, n_entries);
  B200SEG_CHECK_LAUNCH("pack_weights_batched");
  return B200SEG_OK;
}

int tc_pack_weight(const b200seg_conv_desc* d, int kind, const float* w, void* out, void* gen, cudaStream_t st) {
  bool src_is_cin = (kind == B200SEG_W_CONV_FPROP || kind == B200SEG_W_CONVTR_FPROP);
  int src_c = src_is_cin ? d->cin : d->cout, dst_c = src_is_cin ? d->cout : d->cin;
  int src_pad = round16(src_c), dst_pad = round16(dst_c);
  int taps = d->kd * d->kh * d->kw;
  int64_t total = (int64_t)taps * src_pad * dst_pad;
  pack_weight_tc_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(w, (bf16*)out, (bf16*)gen, taps, src_c, dst_c,
                                                                      src_pad, dst_pad, kc_for(src_pad), d->cin,
                                                                      d->cout, kind);
  B200SEG_CHECK_LAUNCH("pack_weight_tc");
  return B200SEG_OK;
}

bool tc_conv_supported(const b200seg_conv_desc* d, int op, const void* src, const void* dst, const void* res) {
  if (d->dtype != B200SEG_BF16 || (d->flags & B200SEG_CONV_FORCE_GENERIC)) return false;
  TcGeom g;
  tc_geom(d, op, g);
  const bool padded = (d->flags & B200SEG_CONV_PADDED_CHANNELS) != 0;
  if (g.src_c < 8) return false;  // Cin = 1 style layers: direct CUDA-core kernels
  if ((g.src_c % 16) && !(padded && g.src_ld >= round16(g.src_c))) return false;
  if ((g.dst_c % 16) && !(padded && g.dst_ld >= round16(g.dst_c))) return false;
  if ((g.src_ld % 8) || (g.dst_ld % 8)) return false;
  if (((uintptr_t)src % 16) || ((uintptr_t)dst % 16)) return false;
  if (res && (((uintptr_t)res % 16) || (d->r_ld % 8) || d->r_ld < round16(g.dst_c))) return false;
  if (g.transposed) {
    for (int i = 0; i < 3; ++i)
      if (g.s[i] == 2 && ((i == 0 ? g.dD : (i == 1 ? g.dH : g.dW)) % 2)) return false;
  }
  return encode_fn() != nullptr;
}

// blockIdx.x = (class * n + sample) * tiles + tile: the layout of the per-CTA statistic partials
void tc_conv_grid(const b200seg_conv_desc* d, int op, int* ncls_out, int64_t* tiles_out) {
  TcGeom g;
  tc_geom(d, op, g);
  const int ddim[3] = {g.dD, g.dH, g.dW};
  int ncls = 1, cdim[3];
  for (int i = 0; i < 3; ++i) {
    bool split = g.transposed && g.s[i] == 2;
    if (split) ncls *= 2;
    cdim[i] = ddim[i] / (split ? 2 : 1);
  }
  int64_t best = -1, tiles = 0;
  for (int td = 1; td <= 128; td *= 2)
    for (int th = 1; td * th <= 128; th *= 2) {
      int tw = 128 / (td * th);
      int64_t vol = (int64_t)((cdim[0] + td - 1) / td) * ((cdim[1] + th - 1) / th) * ((cdim[2] + tw - 1) / tw);
      int64_t score = vol * 1024 - (tw > 16 ? 16 : tw) * 8 - (th > 16 ? 16 : th);
      if (best < 0 || score < best) { best = score; tiles = vol; }
    }
  *ncls_out = ncls;
  *tiles_out = tiles;
}

int tc_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                const void* residual, void* dst, float* stats, cudaStream_t st) {
  TcGeom g;
  tc_geom(d, op, g);
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  const int src_pad = round16(g.src_c), dst_pad = round16(g.dst_c);
  const int KC = kc_for(src_pad);
  int BN = bn_for(dst_pad);
  p.kblocks = src_pad / KC;
  p.n = g.n;
  p.cout = g.dst_c; p.cout_pad = dst_pad;
  p.dst_ld = g.dst_ld; p.res_ld = d->r_ld;
  p.accumulate = (d->flags & B200SEG_CONV_ACCUMULATE) ? 1 : 0;
  p.bias = bias; p.res = (const bf16*)residual; p.dst = (bf16*)dst; p.stats = stats;
  p.dD = g.dD; p.dH = g.dH; p.dW = g.dW;
  const int sdim[3] = {g.sD, g.sH, g.sW}, ddim[3] = {g.dD, g.dH, g.dW};

  // ---- classes, step table, source tensor maps needed
  int ncls[3], os[3], cdim[3];
  for (int i = 0; i < 3; ++i) {
    bool split = g.transposed && g.s[i] == 2;
    ncls[i] = split ? 2 : 1;
    os[i] = split ? 2 : 1;
    cdim[i] = ddim[i] / os[i];
  }
  p.ncls_d = ncls[0]; p.ncls_h = ncls[1]; p.ncls_w = ncls[2];
  p.os_d = os[0]; p.os_h = os[1]; p.os_w = os[2];
  p.cD = cdim[0]; p.cH = cdim[1]; p.cW = cdim[2];
  bool map_used[8] = {false, false, false, false, false, false, false, false};
  int ns = 0;
  for (int cd = 0; cd < ncls[0]; ++cd)
    for (int chh = 0; chh < ncls[1]; ++chh)
      for (int cw = 0; cw < ncls[2]; ++cw) {
        const int cls = (cd * ncls[1] + chh) * ncls[2] + cw;
        const int c[3] = {cd, chh, cw};
        p.cls_begin[cls] = ns;
        for (int kd = 0; kd < g.k[0]; ++kd)
          for (int kh = 0; kh < g.k[1]; ++kh)
            for (int kw = 0; kw < g.k[2]; ++kw) {
              const int kk[3] = {kd, kh, kw};
              int off[3], par[3];
              bool ok = true;
              for (int i = 0; i < 3 && ok; ++i) {
                if (!g.transposed) {
                  int e = kk[i] - g.p[i];  // src = dst*s + e
                  if (g.s[i] == 2) { par[i] = e & 1; off[i] = (e - par[i]) / 2; }
                  else { par[i] = 0; off[i] = e; }
                } else {
                  int e = c[i] + g.p[i] - kk[i];  // src = (dst + p - k)/s, dst = s*j + c
                  par[i] = 0;
                  if (g.s[i] == 2) { if (e & 1) ok = false; else off[i] = e / 2; }
                  else off[i] = e;
                }
              }
              if (!ok) continue;
              TcStep& stp = p.steps[ns++];
              stp.map = (int8_t)((par[0] * 2 + par[1]) * 2 + par[2]);
              stp.dd = (int8_t)off[0]; stp.dh = (int8_t)off[1]; stp.dw = (int8_t)off[2];
              stp.wtile = (kd * g.k[1] + kh) * g.k[2] + kw;
              map_used[stp.map] = true;
            }
        p.cls_begin[cls + 1] = ns;
      }

  // ---- tile shape: 128 destination voxels as a box minimising padding
  {
    int64_t best = -1;
    for (int td = 1; td <= 128; td *= 2)
      for (int th = 1; td * th <= 128; th *= 2) {
        int tw = 128 / (td * th);
        int64_t vol = (int64_t)((cdim[0] + td - 1) / td) * ((cdim[1] + th - 1) / th) * ((cdim[2] + tw - 1) / tw);
        // prefer wide-in-w tiles on ties (contiguous rows), then tall-in-h
        int64_t score = vol * 1024 - (tw > 16 ? 16 : tw) * 8 - (th > 16 ? 16 : th);
        if (best < 0 || score < best) { best = score; p.tD = td; p.tH = th; p.tW = tw; }
      }
  }
  p.tilesD = (cdim[0] + p.tD - 1) / p.tD; p.tilesH = (cdim[1] + p.tH - 1) / p.tH; p.tilesW = (cdim[2] + p.tW - 1) / p.tW;

  // ---- deep layers have only a handful of 128-voxel tiles: narrower N tiles spread them over more SMs
  // (each CTA streams its A tiles from L2 at ~64 B/clk whatever BN is; B shrinks with BN)
  {
    static const int min_ctas = getenv("B200SEG_CONV_MIN_CTAS") ? atoi(getenv("B200SEG_CONV_MIN_CTAS")) : 128;
    const int64_t mtiles = (int64_t)ncls[0] * ncls[1] * ncls[2] * g.n * p.tilesD * p.tilesH * p.tilesW;
    while (BN > 32 && mtiles * (dst_pad / BN) < min_ctas) BN /= 2;
  }

  // ---- tensor maps
  const int row_bytes = KC * 2;
  for (int m = 0; m < 8; ++m) {
    if (!map_used[m]) continue;
    const int par[3] = {(m >> 2) & 1, (m >> 1) & 1, m & 1};
    const int str[3] = {(!g.transposed && g.s[0] == 2) ? 2 : 1, (!g.transposed && g.s[1] == 2) ? 2 : 1,
                        (!g.transposed && g.s[2] == 2) ? 2 : 1};
    const bf16* base = (const bf16*)src + (((int64_t)par[0] * g.sH + par[1]) * g.sW + par[2]) * g.src_ld;
    uint64_t dims[5] = {(uint64_t)src_pad, (uint64_t)((sdim[2] - par[2] + str[2] - 1) / str[2]),
                        (uint64_t)((sdim[1] - par[1] + str[1] - 1) / str[1]),
                        (uint64_t)((sdim[0] - par[0] + str[0] - 1) / str[0]), (uint64_t)g.n};
    uint64_t strides[4] = {(uint64_t)g.src_ld * 2 * str[2], (uint64_t)g.sW * g.src_ld * 2 * str[1],
                           (uint64_t)g.sH * g.sW * g.src_ld * 2 * str[0], (uint64_t)g.sD * g.sH * g.sW * g.src_ld * 2};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tD, 1};
    int rc = make_map(&p.tmA[m], base, 5, dims, strides, box, row_bytes);
    if (rc) return rc;
  }
  {
    const int taps = g.k[0] * g.k[1] * g.k[2];
    uint64_t dims[2] = {(uint64_t)KC, (uint64_t)taps * p.kblocks * dst_pad};
    uint64_t strides[1] = {(uint64_t)KC * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    int rc = make_map(&p.tmB, w_tc, 2, dims, strides, box, row_bytes);
    if (rc) return rc;
  }

  const int ncls_total = ncls[0] * ncls[1] * ncls[2];
  int64_t gx = (int64_t)ncls_total * g.n * p.tilesD * p.tilesH * p.tilesW;
  if (gx > 0x7fffffffLL) { set_error("tc_conv: grid too large"); return B200SEG_ERR_ARG; }
  dim3 grid((unsigned)gx, (unsigned)(dst_pad / BN));

  // ---- deep layers with <= 74 CTAs share an output tile between the 2 or
  // 4 CTAs of a cluster, each taking a share of the (tap, K-block) iterations (tc_conv_splitk_kernel)
  {
    // default ON since round 2 (parity-tested on a B200: tests/test_gpu_kernels.py::test_splitk_cluster_conv; per layer
    // 20 -> 14 us (128->128 @8^3), 31 -> 25 us (256->256 @8^3)); B200SEG_CONV_SPLITK=0 or the NO_SPLIT_K flag switch it off
    static const int splitk_env = getenv("B200SEG_CONV_SPLITK") ? atoi(getenv("B200SEG_CONV_SPLITK")) : 1;
    const bool splitk = (splitk_env || (d->flags & B200SEG_CONV_SPLIT_K)) && !(d->flags & B200SEG_CONV_NO_SPLIT_K);
    const int64_t ctas = gx * (dst_pad / BN);
    const int iters = (ns / (ncls_total > 0 ? ncls_total : 1)) * p.kblocks;  // smallest class has >= this / 8
    if (splitk && ctas * 2 <= 148 && iters >= 16 && (BN == 32 || BN == 64) && KC == 64) {
      p.ksplit = (ctas * 4 <= 148 && BN == 32) ? 4 : 2;
      if (BN == 32) return launch_splitk<32, 64>(p, grid, st);
      return launch_splitk<64, 64>(p, grid, st);
    }
  }

#define TC_CASE(bn, kc)                                   \
  if (BN == bn && KC == kc) {                             \
    p.stages = stages_for<bn, kc>(gx * (dst_pad / BN));    \
    return launch_cfg<bn, kc>(p, grid, st);               \
  }
  TC_CASE(16, 16) TC_CASE(16, 32) TC_CASE(16, 64)
  TC_CASE(32, 16) TC_CASE(32, 32) TC_CASE(32, 64)
  TC_CASE(64, 16) TC_CASE(64, 32) TC_CASE(64, 64)
  TC_CASE(128, 16) TC_CASE(128, 32) TC_CASE(128, 64)
#undef TC_CASE
  set_error("tc_conv: no kernel for BN=%d KC=%d", BN, KC);
  return B200SEG_ERR_UNSUPPORTED;
}

int tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, int row_bytes) {
  return make_map(out, base, rank, dims, strides, box, row_bytes);
}

}  // namespace b200seg

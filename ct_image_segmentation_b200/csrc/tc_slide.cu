// Sliding-window tcgen05 kernels for the small-channel / high-resolution 3x3x3 stride-1 layers
// (head 10->10 at full resolution, 16->16 at 1/2, 32->32 at 1/4): the layers that dominate the
// network's time and whose arithmetic intensity is too low to re-read the source once per tap.
//
// A CTA owns a column of the volume: TH=16 lines x 8 voxels in (h, w), and sweeps along d.  For
// every source slab it TMA-loads three w-shifted copies of the (TH+2) x 8 halo tile into a
// shared-memory ring (box rows of exactly 8 voxels = one swizzle atom, so a shift by whole lines
// (kh) or whole slabs (kd) is an atom-aligned descriptor offset and the w shift (kw) selects the
// copy).  Each source voxel is therefore read 3 x 18/16 = 3.4 times from L2 instead of 27 times.
//   conv (fprop / dgrad): the MMAs are organised per SOURCE slab and folded along N over the three
//     output slabs (kd taps) the slab contributes to: 9 x (KC/16) tcgen05.mma of M=128 (16 lines x 8
//     voxels), N = 3*Cout, K=16 per source slab instead of 27 of N = Cout per output slab -- these
//     small-N MMAs are bound by the shared-memory read of the 4 KB A tile, which the fold cuts 3x.
//     Output slab j lives in TMEM chunk j mod 8 (a ring of accumulators: the three live chunks are
//     adjacent columns except at the wrap, where the MMA is split); a chunk is complete after source
//     slab j+2 and is drained by the epilogue warps while later slabs accumulate.
//   wgrad: see tc_slide_wgrad_kernel below.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace b200seg {

using bf16 = __nv_bfloat16;

int tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, int row_bytes);

namespace {
constexpr int TH = 16;      // lines per tile
constexpr int TWV = 8;      // voxels per line (one swizzle atom of rows)
constexpr int RING = 4;     // source slabs in flight
inline int round16(int c) { return (c + 15) / 16 * 16; }
inline int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int x = atoi(e);
  return x >= lo && x <= hi ? x : dflt;
}
// B200SEG_SLIDE_PERSIST=1: at most `resident CTAs` CTAs that walk several items each.  Measured (r2, head layer
// 2 x 128^3): fprop 118 -> 111 us, but the variants whose epilogue reads global rows (residual addend, fused
// InstanceNorm-backward sums) 118 -> 203 and 145 -> 198 us, whole step 2.116 -> 2.167 ms: off by default, one CTA per
// item; the kernel is the same code either way (grid == items).
inline int slide_persist() { static const int v = env_int("B200SEG_SLIDE_PERSIST", 0, 0, 1); return v; }
// The plain 16 -> 16 kernel exists twice: two CTAs per SM (115 registers) and three (96 registers, 8 bytes of
// spill).  Measured (r2, graph-replayed, 2 x 128^3 head layer / 2 x 64^3): without a residual addend two are faster
// (fprop 99.6 vs 106.6 us, dgrad 101.6 vs 108.4, 22.1 vs 24.0), with one -- its rows are global loads inside the
// epilogue's per-slab loop -- three hide that latency better (137 vs 118 us).  B200SEG_SLIDE_MINB3 = 0 / 1 forces one.
inline bool slide_three_ctas(bool has_residual) {
  static const int v = env_int("B200SEG_SLIDE_MINB3", -1, -1, 1);
  return v < 0 ? has_residual : v == 1;
}
inline int sm_count() {
  static const int v = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    return n;
  }();
  return v;
}
// Work decomposition shared by the run / grid / workspace queries.  One CTA per item: enough d segments per column
// for ~4 CTAs per SM, at least 8 slabs each (r1) -- except where that lands between one and two waves of the `slots`
// CTAs resident at a time and a nearly full SINGLE wave exists: then the largest segment count that still fits one
// wave (16 -> 16 at 64^3 x 2: 384 CTAs in one wave instead of 512 in 1.15, dgrad + residual 24.7 -> 20.2 us; the
// set-up of a CTA -- TMEM allocation, 27 weight tiles, ring fill -- is worth several slabs).  Persistent mode
// (experiments): the split that minimises rounds x (slabs + 2 halo slabs) over the resident CTAs.
inline void slide_segments(int64_t cols, int D, int slots_per_sm, int& dseg, int& nseg) {
  const int64_t slots = (int64_t)sm_count() * slots_per_sm;
  if (slide_persist()) {
    int64_t best = -1;
    nseg = 1;
    for (int ns = 1; ns <= (D + 3) / 4 && ns <= 64; ++ns) {
      const int ds = (D + ns - 1) / ns;
      if ((D + ds - 1) / ds != ns) continue;
      const int64_t cost = ((cols * ns + slots - 1) / slots) * (ds + 2);
      if (best < 0 || cost < best) { best = cost; nseg = ns; }
    }
    dseg = (D + nseg - 1) / nseg;
    return;
  }
  nseg = (int)((148 * 4 + cols - 1) / cols);
  if (nseg < 1) nseg = 1;
  dseg = (D + nseg - 1) / nseg;
  if (dseg < 8) dseg = 8;
  if (dseg > D) dseg = D;
  nseg = (D + dseg - 1) / dseg;
  const int64_t items = cols * nseg;
  if (items > slots && items <= 2 * slots) {
    for (int ns = nseg - 1; ns >= 1; --ns) {
      const int ds = (D + ns - 1) / ns;
      if ((D + ds - 1) / ds != ns || cols * ns > slots) continue;
      if (4 * cols * ns >= 3 * slots) { nseg = ns; dseg = ds; }
      break;
    }
  }
}
}  // namespace

struct alignas(64) TcSlideConvParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int n, D, H, W;
  int tilesH, tilesW, dseg, nseg;
  int items;     // (sample, line tile, voxel tile, d segment) columns: CTA b works on items b, b + gridDim.x, ...
  int debug;     // B200SEG_SLIDE_DEBUG (timing experiments, results are wrong): 1 = the epilogue only drains and hands
                 // back the accumulators, 2 = no MMAs, 4 = no TMA loads of the source
  int cout, dst_ld, res_ld, accumulate, flip;
  const float* bias;
  const bf16* res;
  bf16* dst;
  float* stats;  // optional [item][epilogue warp][cout][2]: sum / sum of squares of the outputs (InstanceNorm)
  // BST variant (dgrad fused with the reduction pass of the InstanceNorm+PReLU backward of the layer whose output
  // gradient this kernel writes): nx = that layer's pre-norm tensor (same voxels as dst), its statistics and slope;
  // bstats [item][epilogue warp][BN][3] = { sum g~, sum g~ xhat, sum dy xhat [xhat <= 0] },  g~ = dy * prelu'(xhat)
  const bf16* nx;
  int nx_ld, nstat_ld;
  const float* nmean;
  const float* nrstd;
  const float* nalpha;
  float* bstats;
};

// CS (BST only): channels that really exist (10 for the head layer's 16-wide rows): the sums of the padding channels
// are identically zero and are neither computed nor kept in registers
template <int BN, int MINB>
__host__ __device__ constexpr int slide_accr() { return 8; }  // (16 where two CTAs share the 512 columns: measured, no gain)

// MINB = CTAs per SM the register allocation is held to (3 x 6 warps = 5 warps on some SM sub-partitions: 96 registers)
template <int BN, int KC, bool BST = false, int CS = BN, int MINB = (BST ? 2 : 1)>
__global__ void __launch_bounds__(192, MINB)
tc_slide_conv_kernel(const __grid_constant__ TcSlideConvParams p) {
  constexpr int PITCH = KC * 2;                          // bytes per voxel row
  constexpr int COPY_BYTES = (TH + 2) * TWV * PITCH;     // one w-shifted halo tile
  constexpr int SLAB_BYTES = 3 * COPY_BYTES;
  constexpr int WT_BYTES = BN * PITCH;                   // one weight tile (tap)
  constexpr int W_BYTES = (27 * WT_BYTES + 1023) / 1024 * 1024;
  // TMEM accumulator ring (output slabs).  The folded MMAs are split where the ring wraps (2 of ACCR slabs).  A ring
  // of 16 chunks in the two-CTAs-per-SM builds (10.1 instead of 11.25 MMAs per slab, deeper run-ahead) changed nothing:
  // head fprop 99.6 -> 99.8 us (r2b, profiles/r2b_slide_accr16.txt).
  constexpr int ACCR = slide_accr<BN, MINB>();
  constexpr uint32_t TMEM_COLS = ACCR * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RING * SLAB_BYTES);
  uint64_t* empty = full + RING;
  uint64_t* acc_full = empty + RING;      // [ACCR]
  uint64_t* acc_empty = acc_full + ACCR;  // [ACCR]
  uint64_t* wbar = acc_empty + ACCR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  // Persistent CTAs: the grid is min(items, resident CTAs) and a CTA walks its items without re-allocating TMEM,
  // re-loading the weights or draining its pipelines: the source ring (counter g) and the accumulator ring
  // (counter ob) run on across item borders.  (r2: with one item per CTA the set-up was ~8 % of a CTA's life
  // and 768 CTAs on 444 slots left the last 0.27 wave half empty.)
  struct Item { int n, h0, w0, d_begin, nd; };
  auto decode = [&](int item) {
    Item it;
    int bx = item;
    const int seg = bx % p.nseg; bx /= p.nseg;
    const int tw_i = bx % p.tilesW; bx /= p.tilesW;
    const int th_i = bx % p.tilesH; bx /= p.tilesH;
    it.n = bx;
    it.h0 = th_i * TH; it.w0 = tw_i * TWV;
    it.d_begin = seg * p.dseg;
    it.nd = min(p.dseg, p.D - it.d_begin);  // output slabs of this item
    return it;
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACCR; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_init(wbar, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmA);
    tc::prefetch_tmap(&p.tmB);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  if constexpr (BST) {  // the consumer InstanceNorm's statistics of every sample (n <= 16), read per use from shared memory
    float* nsm = reinterpret_cast<float*>(tmem_slot + 4) + 4 * BN * 3;
    for (int idx = threadIdx.x; idx < p.n * BN; idx += blockDim.x) {  // xhat = x * rstd + (-mean * rstd): one FMA
      const int c = idx % BN, nn = idx / BN;
      const float r = p.nrstd[nn * p.nstat_ld + c];
      nsm[2 * idx] = r;
      nsm[2 * idx + 1] = -p.nmean[nn * p.nstat_ld + c] * r;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  // accumulator chunks start (and are handed back by the epilogue) zeroed: every MMA accumulates, so a
  // source slab is ONE N-folded instruction per tap even for the output slab it touches first
  if (warp >= 2) {
    const uint32_t lane_base = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < ACCR * BN; c += 16) tc::tmem_st16_zero(lane_base + c);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      // weights: for every in-plane source shift (ih, iw) the three kd tiles in the order of the
      // output slabs s-2, s-1, s a source slab s feeds: slot (ih*3+iw)*3 + i holds source-shift id = 2-i.
      // Tile of source shift (id, ih, iw): tap (id, ih, iw) for fprop, the mirrored tap for dgrad.
      tc::mbar_expect_tx(wbar, 27 * WT_BYTES);
      for (int hw = 0; hw < 9; ++hw)
        for (int i = 0; i < 3; ++i) {
          const int id = 2 - i, ih = hw / 3, iw = hw % 3;
          const int tap = p.flip ? ((2 - id) * 3 + (2 - ih)) * 3 + (2 - iw) : (id * 3 + ih) * 3 + iw;
          tc::tma_load_2d(wsm + (hw * 3 + i) * WT_BYTES, &p.tmB, wbar, 0, tap * BN);
        }
      uint32_t g = 0;  // source slabs issued by this CTA
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Item it = decode(item);
        for (int s = 0; s < it.nd + 2; ++s, ++g) {
          const uint32_t slot = g % RING;
          tc::mbar_wait(&empty[slot], ((g / RING) & 1u) ^ 1u);
          if (p.debug & 4) { tc::mbar_arrive(&full[slot]); continue; }
          uint8_t* dst = ring + slot * SLAB_BYTES;
          tc::mbar_expect_tx(&full[slot], SLAB_BYTES);
          const int ds = it.d_begin - 1 + s;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            tc::tma_load_5d(dst + kw * COPY_BYTES, &p.tmA, &full[slot], 0, it.w0 + kw - 1, it.h0 - 1, ds, it.n);
        }
      }
    }
  } else if (warp == 1) {
    {  // the whole warp runs the (warp-uniform) issue loop; one elected lane issues each MMA / commit
      const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // (shadows the per-lane copy: see warp_uniform)
      constexpr uint64_t layout = tc::layout_for_row_bytes(PITCH);
      const uint32_t w_addr = tc::smem_u32(wsm), r_addr = tc::smem_u32(ring);
      const uint64_t tmpl = tc::make_smem_desc(0, 16, 8 * PITCH, layout);
      const uint64_t w_desc = tmpl + (w_addr >> 4);
      tc::mbar_wait(wbar, 0);
      uint32_t g = 0, ob = 0;  // source slabs consumed / output slabs started before this item
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int nd = decode(item).nd;
      for (int s = 0; s < nd + 2; ++s, ++g) {
        // source slab s (absolute d = d_begin - 1 + s) feeds output slabs s-2, s-1, s (clipped to [0, nd))
        const int lo = max(s - 2, 0), hi = min(s, nd - 1);
        if (s < nd) {  // output slab s receives its first contribution: its chunk must have been drained (zeroed)
          const uint32_t o = ob + (uint32_t)s;
          tc::mbar_wait(&acc_empty[o % ACCR], ((o / ACCR) & 1u) ^ 1u);
        }
        const uint32_t slot = g % RING;
        tc::mbar_wait(&full[slot], (g / RING) & 1u);
        tc::tc_fence_after();
        // The accumulating range [lo, hi] is one MMA (adjacent TMEM chunks) or two where the
        // accumulator ring wraps.
        const int c_lo = (int)((ob + (uint32_t)lo) % ACCR);
        const int len0 = min(hi - lo + 1, ACCR - c_lo), len1 = hi - lo + 1 - len0;
        const uint32_t d0 = tmem_acc + c_lo * BN, d1 = tmem_acc;  // a wrapped part starts at chunk 0
        const uint32_t i0 = tc::make_idesc_bf16(128, len0 * BN, false, false);
        const uint32_t i1 = tc::make_idesc_bf16(128, (len1 > 0 ? len1 : 1) * BN, false, false);
        const uint64_t b0 = w_desc + (((lo - (s - 2)) * WT_BYTES) >> 4);
        const uint64_t b1 = b0 + ((len0 * WT_BYTES) >> 4);
        const uint64_t slab = tmpl + ((r_addr + slot * SLAB_BYTES) >> 4);
        if (!(p.debug & 2)) {
        tc::umma_bf16_warp(d0, slab, b0, i0, 1u);
        if (len1 > 0) tc::umma_bf16_warp(d1, slab, b1, i1, 1u);
#pragma unroll
        for (int hw = 0; hw < 9; ++hw)
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            if (hw == 0 && k == 0) continue;
            const uint64_t a = slab + (((hw % 3) * COPY_BYTES + (hw / 3) * (TWV * PITCH)) >> 4) + 2 * k;
            const uint64_t bo = ((hw * 3 * WT_BYTES) >> 4) + 2 * k;
            tc::umma_bf16_warp(d0, a, b0 + bo, i0, 1u);
            if (len1 > 0) tc::umma_bf16_warp(d1, a, b1 + bo, i1, 1u);
          }
        }
        tc::umma_commit_warp(&empty[slot]);
        if (s >= 2) tc::umma_commit_warp(&acc_full[(ob + (uint32_t)(s - 2)) % ACCR]);
      }
      ob += (uint32_t)nd;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float bias[BN];           // hoisted: the epilogue runs once per slab
#pragma unroll
    for (int c = 0; c < BN; ++c) bias[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    const float nslope = BST ? p.nalpha[0] : 0.f;
    float* sred = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][BN][3]: the item's statistics, warp by warp
    uint32_t ob = 0;  // output slabs of the items before this one
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const Item it = decode(item);
    const int n = it.n, d_begin = it.d_begin, nd = it.nd;
    const int oh = it.h0 + row / TWV, ow = it.w0 + row % TWV;
    const bool valid = oh < p.H && ow < p.W;
    float ssum[BN], ssq[BN];  // per-thread partial statistics over this item's slabs (same sample n)
    float sb0[BST ? CS : 1], sb1[BST ? CS : 1], sb2[BST ? CS : 1];  // BST: the three InstanceNorm-backward sums
#pragma unroll
    for (int c = 0; c < BN; ++c) ssum[c] = ssq[c] = 0.f;
#pragma unroll
    for (int c = 0; c < (BST ? CS : 1); ++c) sb0[c] = sb1[c] = sb2[c] = 0.f;
    const float2* nsm = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(tmem_slot + 4) + 4 * BN * 3) + n * BN;
    // BST: the rows the epilogue READS from global memory (residual addend, the consumer InstanceNorm's pre-norm
    // tensor) are fetched one slab ahead, before the wait for that slab's accumulator: otherwise every slab pays a
    // global-load round trip inside the epilogue's serial per-slab loop (r2: 263 us with the loads issued after the
    // wait, 150 us with the prefetch).  The plain variants keep their code (and 96 registers = 3 CTAs per SM): the
    // same prefetch there cost more in registers / spills than it hid (dgrad + residual 115 -> 138 us).
    constexpr bool PF = BST;
    uint4 pr0 = make_uint4(0, 0, 0, 0), pr1 = pr0, px0 = pr0, px1 = pr0;
    auto prefetch = [&](int j) {
      const int64_t l = (((int64_t)n * p.D + d_begin + j) * p.H + oh) * p.W + ow;
      if (p.res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + l * p.res_ld);
        pr0 = rp[0];
        pr1 = rp[1];
      }
      if constexpr (BST) {
        const uint4* xp = reinterpret_cast<const uint4*>(p.nx + l * p.nx_ld);
        px0 = xp[0];
        px1 = xp[1];
      }
    };
    if (PF && valid && nd > 0) prefetch(0);
    for (int j = 0; j < nd; ++j) {
      const uint32_t o = ob + (uint32_t)j;
      const int buf = (int)(o % ACCR);
      const uint4 cr0 = pr0, cr1 = pr1, cx0 = px0, cx1 = px1;  // this slab's rows
      if (PF && valid && j + 1 < nd) prefetch(j + 1);
      tc::mbar_wait(&acc_full[buf], (o / ACCR) & 1u);
      tc::tc_fence_after();
      const int od = d_begin + j;
      const int64_t lin = (((int64_t)n * p.D + od) * p.H + oh) * p.W + ow;
#pragma unroll
      for (int ch = 0; ch < BN / 16; ++ch) {
        uint32_t v[16];
        tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + buf * BN + ch * 16, v);
        tc::tmem_ld_wait();
        if (valid && !(p.debug & 1)) {
          const int c0 = ch * 16;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += bias[ch * 16 + i];
          if (p.res) {
            uint4 r0 = cr0, r1 = cr1;
            if constexpr (!PF) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + lin * p.res_ld + c0);
              r0 = rp[0];
              r1 = rp[1];
            }
            const __nv_bfloat162* g0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* g1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 a = __bfloat1622float2(g0[i]), b = __bfloat1622float2(g1[i]);
              f[2 * i] += a.x; f[2 * i + 1] += a.y;
              f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
            }
          }
          uint4* op = reinterpret_cast<uint4*>(p.dst + lin * p.dst_ld + c0);
          if (p.accumulate) {
            uint4 r0 = op[0], r1 = op[1];
            const __nv_bfloat162* g0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* g1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 a = __bfloat1622float2(g0[i]), b = __bfloat1622float2(g1[i]);
              f[2 * i] += a.x; f[2 * i + 1] += a.y;
              f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
            }
          }
          if (!BST && p.stats) {
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
          op[0] = o0;
          op[1] = o1;
          if constexpr (BST) {
            // the sums of the InstanceNorm + PReLU backward this gradient feeds, from the value AS STORED (bf16)
            uint4 x0 = cx0, x1 = cx1;
            if constexpr (!PF) {
              const uint4* xp = reinterpret_cast<const uint4*>(p.nx + lin * p.nx_ld + c0);
              x0 = xp[0];
              x1 = xp[1];
            }
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&x0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&x1);
            float xv[16], gv[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 a = __bfloat1622float2(h0[i]), b = __bfloat1622float2(h1[i]);
              xv[2 * i] = a.x; xv[2 * i + 1] = a.y; xv[8 + 2 * i] = b.x; xv[8 + 2 * i + 1] = b.y;
              float2 ga = __bfloat1622float2(q0[i]), gb = __bfloat1622float2(q1[i]);
              gv[2 * i] = ga.x; gv[2 * i + 1] = ga.y; gv[8 + 2 * i] = gb.x; gv[8 + 2 * i + 1] = gb.y;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (c0 + i < CS) {  // (compile time)
                const float2 ab = nsm[c0 + i];
                const float h = fmaf(xv[i], ab.x, ab.y);
                const bool pos = h > 0.f;
                const float g = pos ? gv[i] : nslope * gv[i];
                sb0[c0 + i] += g;
                sb1[c0 + i] = fmaf(g, h, sb1[c0 + i]);
                sb2[c0 + i] += pos ? 0.f : gv[i] * h;
              }
            }
          }
        }
      }
      // hand the chunk back zeroed
#pragma unroll
      for (int ch = 0; ch < BN / 16; ++ch) tc::tmem_st16_zero(tmem_acc + ((uint32_t)(q * 32) << 16) + buf * BN + ch * 16);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
    // the item's statistics: warp tree per channel -> shared memory -> one row per item (the four epilogue warps
    // meet at a named barrier: the producer and the MMA warp are already working on the next items)
    const int et = threadIdx.x - 64;
    if constexpr (BST) {
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        float a = 0.f, b = 0.f, d3 = 0.f;
        if (c < CS) {  // (compile time; padding channels: exact zeros)
          a = warp_sum(sb0[c < CS ? c : 0]);
          b = warp_sum(sb1[c < CS ? c : 0]);
          d3 = warp_sum(sb2[c < CS ? c : 0]);
        }
        if (lane == 0) {
          sred[(q * BN + c) * 3] = a;
          sred[(q * BN + c) * 3 + 1] = b;
          sred[(q * BN + c) * 3 + 2] = d3;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < 3 * BN) {  // every one of the BN (padded) channels: the InstanceNorm kernels see them too
        const int c = et / 3, m = et % 3;
        p.bstats[((int64_t)item * BN + c) * 3 + m] =
            sred[(0 * BN + c) * 3 + m] + sred[(1 * BN + c) * 3 + m] + sred[(2 * BN + c) * 3 + m] +
            sred[(3 * BN + c) * 3 + m];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    } else if (p.stats) {
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        const float a = warp_sum(ssum[c]), b = warp_sum(ssq[c]);
        if (lane == 0) {
          sred[(q * BN + c) * 2] = a;
          sred[(q * BN + c) * 2 + 1] = b;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < 2 * BN) {
        const int c = et >> 1, m = et & 1;
        if (c < p.cout)
          p.stats[((int64_t)item * p.cout + c) * 2 + m] =
              sred[(0 * BN + c) * 2 + m] + sred[(1 * BN + c) * 2 + m] + sred[(2 * BN + c) * 2 + m] +
              sred[(3 * BN + c) * 2 + m];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ob += (uint32_t)nd;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_acc);
}

// ------------------------------------------------------------------------------------------------
namespace {

struct SlideGeom {
  int n, D, H, W, src_c, dst_c, src_ld, dst_ld;
};

bool slide_geom(const b200seg_conv_desc* d, int op, SlideGeom& g) {
  if (op != TC_CONV_FPROP && op != TC_CONV_DGRAD) return false;
  if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 1 || d->sh != 1 || d->sw != 1) return false;
  g.n = d->n; g.D = d->in_d; g.H = d->in_h; g.W = d->in_w;
  if (op == TC_CONV_FPROP) { g.src_c = d->cin; g.dst_c = d->cout; g.src_ld = d->x_ld; g.dst_ld = d->y_ld; }
  else { g.src_c = d->cout; g.dst_c = d->cin; g.src_ld = d->y_ld; g.dst_ld = d->x_ld; }
  return true;
}

template <int BN, int KC, bool BST>
constexpr size_t slide_smem() {
  constexpr int PITCH = KC * 2;
  constexpr int SLAB = 3 * (TH + 2) * TWV * PITCH;
  constexpr int WB = (27 * BN * PITCH + 1023) / 1024 * 1024;
  // ... + reduction scratch [4 warps][BN][3] + the consumer InstanceNorm's rstd / -mean * rstd of up to 16 samples (BST)
  return 1024 + WB + RING * SLAB + 8 * PITCH * 8 + (2 * RING + 2 * 16 + 1) * 8 + 64 + 4 * BN * 3 * 4 + (BST ? 16 * BN * 8 : 0);
}

// co-resident CTAs per SM of one variant (registers, shared memory, TMEM columns), asked once from the runtime
template <int BN, int KC, bool BST, int CS, int MINB = (BST ? 2 : 1)>
int slide_ctas_per_sm() {
  static const int v = [] {
    cudaFuncSetAttribute(tc_slide_conv_kernel<BN, KC, BST, CS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tc_slide_conv_kernel<BN, KC, BST, CS, MINB>, 192,
                                                      slide_smem<BN, KC, BST>()) != cudaSuccess || nb < 1)
      nb = 1;
    const int by_tmem = 512 / (slide_accr<BN, MINB>() * BN);
    return nb < by_tmem ? nb : by_tmem;
  }();
  return v;
}

template <int BN, int KC, bool BST = false, int CS = BN, int MINB = (BST ? 2 : 1)>
int launch_slide(const TcSlideConvParams& p, unsigned grid, cudaStream_t st) {
  const size_t smem = slide_smem<BN, KC, BST>();
  slide_ctas_per_sm<BN, KC, BST, CS, MINB>();  // (sets the shared-memory attribute)
  tc_slide_conv_kernel<BN, KC, BST, CS, MINB><<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH(BST ? "tc_slide_conv_bwdstats" : "tc_slide_conv");
  count_tc_launch();
  return B200SEG_OK;
}

}  // namespace

// Layers the sliding kernel takes (on top of tc_conv_supported): 3-D, 3x3x3, stride 1, source and
// destination channel counts (padded) in {16, 32}, and a volume large enough to amortise the sweep.
bool tc_slide_conv_supported(const b200seg_conv_desc* d, int op) {
  if (tc_line_conv_supported(d, op)) return true;
  SlideGeom g;
  if (!slide_geom(d, op, g)) return false;
  const int sp = round16(g.src_c), dp = round16(g.dst_c);
  if ((sp != 16 && sp != 32) || (dp != 16 && dp != 32)) return false;
  if (g.D < 8 || (int64_t)g.H * g.W < 512) return false;
  return true;
}

// co-resident CTAs per SM of the variant this layer / op runs on (the segmentation must be the same in the workspace
// queries and in the run)
static int slide_slots_per_sm(int KC, int BN, bool bst, int dst_c, bool has_residual = false) {
  if (bst) return dst_c == 10 ? slide_ctas_per_sm<16, 16, true, 10>() : slide_ctas_per_sm<16, 16, true, 16>();
  if (BN == 16 && KC == 16)
    return slide_three_ctas(has_residual) ? slide_ctas_per_sm<16, 16, false, 16, 3>() : slide_ctas_per_sm<16, 16, false, 16>();
  if (BN == 16 && KC == 32) return slide_ctas_per_sm<16, 32, false, 16>();
  if (BN == 32 && KC == 16) return slide_ctas_per_sm<32, 16, false, 32>();
  return slide_ctas_per_sm<32, 32, false, 32>();
}

// number of items (= statistic partial rows); items of one sample are contiguous
int64_t tc_slide_conv_grid(const b200seg_conv_desc* d, int op, bool bst) {
  if (tc_line_conv_supported(d, op)) return tc_line_conv_rows(d, op);
  SlideGeom g;
  slide_geom(d, op, g);
  const int tilesH = (g.H + TH - 1) / TH, tilesW = (g.W + TWV - 1) / TWV;
  const int64_t cols = (int64_t)g.n * tilesH * tilesW;
  int dseg, nseg;
  slide_segments(cols, g.D, slide_slots_per_sm(round16(g.src_c), round16(g.dst_c), bst, g.dst_c), dseg, nseg);
  return cols * nseg;
}

// dgrad whose output gradient feeds an InstanceNorm + PReLU backward: the fused variant exists for 16 (padded)
// destination channels -- the head layer 10->10 and the 16->16 layers, where that reduction pass is a full-resolution
// bandwidth pass of its own
bool tc_slide_conv_bwdstats_supported(const b200seg_conv_desc* d, int op) {
  if (op != TC_CONV_DGRAD || (d->flags & B200SEG_CONV_NO_SLIDE) || !tc_slide_conv_supported(d, op)) return false;
  return round16(d->cin) == 16 && round16(d->cout) == 16 && d->n <= 16;
}

int tc_slide_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                      const void* residual, void* dst, float* stats, cudaStream_t st, const TcBwdStats* bst) {
  if (tc_line_conv_supported(d, op)) return tc_line_conv_run(d, op, src, w_tc, bias, residual, dst, stats, st, bst);
  SlideGeom g;
  slide_geom(d, op, g);
  TcSlideConvParams p;
  memset(&p, 0, sizeof(p));
  if (bst) {
    p.nx = (const bf16*)bst->nx; p.nx_ld = bst->nx_ld; p.nstat_ld = bst->nstat_ld;
    p.nmean = bst->mean; p.nrstd = bst->rstd; p.nalpha = bst->alpha; p.bstats = bst->partials;
  }
  const int KC = round16(g.src_c), BN = round16(g.dst_c);
  p.n = g.n; p.D = g.D; p.H = g.H; p.W = g.W;
  p.tilesH = (g.H + TH - 1) / TH; p.tilesW = (g.W + TWV - 1) / TWV;
  const int64_t cols = (int64_t)g.n * p.tilesH * p.tilesW;
  // (statistics are only fused without a residual addend, so the rows tc_slide_conv_grid reports -- no residual --
  // are the rows this run writes)
  const int slots_per_sm = slide_slots_per_sm(KC, BN, bst != nullptr, g.dst_c, residual != nullptr);
  slide_segments(cols, g.D, slots_per_sm, p.dseg, p.nseg);
  p.cout = g.dst_c; p.dst_ld = g.dst_ld; p.res_ld = d->r_ld;
  static const int dbg = [] {
    const int v = env_int("B200SEG_SLIDE_DEBUG", 0, 0, 7);
    if (v) fprintf(stderr, "b200seg: B200SEG_SLIDE_DEBUG=%d switches parts of the kernel off: timing experiment, results are WRONG\n", v);
    return v;
  }();
  p.debug = dbg;
  p.accumulate = (d->flags & B200SEG_CONV_ACCUMULATE) ? 1 : 0;
  p.flip = (op == TC_CONV_DGRAD) ? 1 : 0;
  p.bias = bias; p.res = (const bf16*)residual; p.dst = (bf16*)dst; p.stats = stats;
  {
    uint64_t dims[5] = {(uint64_t)KC, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.D, (uint64_t)g.n};
    uint64_t strides[4] = {(uint64_t)g.src_ld * 2, (uint64_t)g.W * g.src_ld * 2, (uint64_t)g.H * g.W * g.src_ld * 2,
                           (uint64_t)g.D * g.H * g.W * g.src_ld * 2};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)TWV, (uint32_t)(TH + 2), 1, 1};
    int rc = tc_make_map(&p.tmA, src, 5, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)KC, (uint64_t)27 * BN};
    uint64_t strides[1] = {(uint64_t)KC * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    int rc = tc_make_map(&p.tmB, w_tc, 2, dims, strides, box, KC * 2);
    if (rc) return rc;
  }
  const int64_t items = cols * p.nseg;
  if (items > 0x3fffffffLL) { set_error("tc_slide_conv: grid too large"); return B200SEG_ERR_ARG; }
  p.items = (int)items;
  const int64_t slots = (int64_t)sm_count() * slots_per_sm;
  const int64_t grid = (slide_persist() && items > slots) ? slots : items;
  if (bst) {
    if (BN == 16 && KC == 16 && g.dst_c == 10) return launch_slide<16, 16, true, 10>(p, (unsigned)grid, st);
    if (BN == 16 && KC == 16) return launch_slide<16, 16, true>(p, (unsigned)grid, st);
    set_error("tc_slide_conv: no fused InstanceNorm-backward variant for BN=%d KC=%d", BN, KC);
    return B200SEG_ERR_UNSUPPORTED;
  }
  if (BN == 16 && KC == 16)
    return slide_three_ctas(residual != nullptr) ? launch_slide<16, 16, false, 16, 3>(p, (unsigned)grid, st)
                                                  : launch_slide<16, 16>(p, (unsigned)grid, st);
  if (BN == 16 && KC == 32) return launch_slide<16, 32>(p, (unsigned)grid, st);
  if (BN == 32 && KC == 16) return launch_slide<32, 16>(p, (unsigned)grid, st);
  if (BN == 32 && KC == 32) return launch_slide<32, 32>(p, (unsigned)grid, st);
  set_error("tc_slide_conv: no kernel for BN=%d KC=%d", BN, KC);
  return B200SEG_ERR_UNSUPPORTED;
}

}  // namespace b200seg

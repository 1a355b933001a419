// EXPERIMENTAL (off by default, see line_enabled() below for the measurements): line-tiled tcgen05 kernel for the
// 16-channel 3x3x3 stride-1 layers whose row length divides 128 voxels (head 10->10 at 128^3, 16->16 at 64^3 of the
// cfg3 net): fprop / dgrad, optional fused InstanceNorm statistics (fprop) or InstanceNorm-backward sums (dgrad), same
// contract as tc_slide_conv_kernel (tc_slide.cu), which is the production kernel.
//
// What tc_slide_conv pays for and this kernel does not (ncu r2, head layer: the tensor-core pipe 79 % busy with 9-11
// N=48 MMAs per slab that cost 46 cycles each for 24 cycles of math -- the 4 KB A-tile fetch -- and 432 TMA box rows
// of 32 bytes per slab):
//   * ONE copy of the source per slab instead of three w-shifted ones.  The three kw taps are folded along N: an
//     MMA multiplies the un-shifted source tile with the weights of all three kw taps (and all three kd taps, as
//     before), i.e. N = 3 kd x 3 kw x 16 = 144, so an accumulator row holds, for ITS source voxel, the three
//     partial sums that belong to the outputs w-1, w, w+1.  The epilogue adds them up across neighbouring rows:
//         out[w] = acc_kw0[w-1] + acc_kw1[w] + acc_kw2[w+1]          (zero beyond the ends of a line = padding)
//     3 MMAs per 128 source rows and slab instead of 9-11, each amortising the A fetch over 3x the columns.
//   * MMA rows are voxel PAIRS: a tile is 128 pairs = 256 consecutive voxels of 256 / W whole lines, a pair is one
//     64-byte row of the shared-memory tile (TMA box rows of 64 instead of 32 bytes: half the rows per byte, and
//     whole lines instead of 8-voxel segments), the voxel of a pair is selected by the K offset of the A
//     descriptor (+32 bytes inside the 64-byte swizzle span), so there are two accumulators per output slab
//     (even / odd voxels) and most of the neighbour sums above stay inside a thread: only out[even] needs the odd
//     accumulator of the previous row and out[odd] the even accumulator of the next row (one warp shuffle each;
//     rows 31|32 of a 128-voxel line go through shared memory).
//   * a tile of LPT = 256 / W lines loads LPT + 2 source lines per slab (W = 128: 2x instead of 3.4x the tensor
//     through L2), the epilogue thread stores 64 contiguous bytes.
//   * persistent CTAs (one per SM, all 512 TMEM columns: 2 parities x 5 output slabs x 48 columns) walk a static
//     list of (sample, line tile, d segment) items; weights, TMEM allocation and the barrier set-up happen once,
//     and the three roles stream across item borders without draining.
// Requires the source tensor's voxel stride to be exactly 16 elements (pairs contiguous); the destination,
// residual and InstanceNorm operands may have any 8-element-aligned stride (concatenation slices).
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
constexpr int LRING = 6;                 // source slabs in flight
constexpr int LACCR = 5;                 // output slabs resident in TMEM (3 accumulating + 2 draining)
constexpr int LCHUNK = 48;               // columns per output slab and parity: 3 kw x 16 channels
constexpr int LRSTRIDE = LACCR * LCHUNK; // column distance between the even- and the odd-voxel accumulators
constexpr uint32_t LTMEM_COLS = 512;
constexpr int LWT_BYTES = 16 * 32;       // one weight tile: 16 dst channels x 16 src channels
constexpr int LW_BYTES = (27 * LWT_BYTES + 1023) / 1024 * 1024;
}  // namespace

struct alignas(64) TcLineConvParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int n, D, H, W;
  int tilesH, dseg, nseg, items;  // item = (sample * nseg + segment) * tilesH + line tile
  int cout, dst_ld, res_ld, accumulate, flip;
  int debug;      // B200SEG_LINE_DEBUG (timing experiments, results are wrong): 1 = epilogue drains and hands back the
                  // accumulators only, 2 = no MMAs, 4 = no TMA loads of the source
  const float* bias;
  const bf16* res;
  bf16* dst;
  float* stats;   // optional [item][epilogue warp][cout][2]
  // BST (see tc_slide.cu): the InstanceNorm + PReLU backward whose output gradient this dgrad writes
  const bf16* nx;
  int nx_ld, nstat_ld;
  const float* nmean;
  const float* nrstd;
  const float* nalpha;
  float* bstats;  // [item][epilogue warp][16][3]
};

// PPL = voxel pairs per line (W / 2: 16, 32 or 64); EG = sets of 8 epilogue warps (alternating slabs);
// BST / CS as in tc_slide_conv_kernel
template <int PPL, int EG, bool BST, int CS>
__global__ void __launch_bounds__(64 + 256 * EG, 1)
tc_line_conv_kernel(const __grid_constant__ TcLineConvParams p) {
  constexpr int LPT = 128 / PPL;                         // lines per tile
  constexpr int LINE_BYTES = PPL * 64;
  constexpr int SLAB_BYTES = (LPT + 2) * LINE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + LW_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + LRING * SLAB_BYTES);
  uint64_t* empty = full + LRING;
  uint64_t* acc_full = empty + LRING;        // [LACCR]
  uint64_t* acc_empty = acc_full + LACCR;    // [LACCR]
  uint64_t* wbar = acc_empty + LACCR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* xch = reinterpret_cast<float*>(tmem_slot + 4);   // [2 slab parities][2 EG groups][4 warps][16] row exchange
  float* bias_s = xch + 2 * (2 * EG) * 4 * 16;             // [16]

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < LRING; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < LACCR; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 8);
    }
    tc::mbar_init(wbar, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmA);
    tc::prefetch_tmap(&p.tmB);
  }
  if (warp == 1) tc::tmem_alloc<LTMEM_COLS>(tmem_slot);
  if (threadIdx.x >= 64 && threadIdx.x < 80) {
    const int c = threadIdx.x - 64;
    bias_s[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // accumulators start (and are handed back by the epilogue) zeroed: every MMA accumulates
  if (warp >= 2 && warp < 6) {
    const uint32_t lane_base = *tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 2 * LRSTRIDE; c += 16) tc::tmem_st16_zero(lane_base + c);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  const int per_n = p.nseg * p.tilesH;

  if (warp == 0) {
    if (lane == 0) {
      // weights: slot (kh * 3 + i) * 3 + kw holds the tile of in-plane tap (kh, kw) and source shift kd = 2 - i,
      // i.e. for a fixed kh the 9 tiles are the B rows of ONE MMA in the column order of the accumulators:
      // output slab s-2+i (i = 0..2), then kw, then the destination channel.  dgrad: the mirrored tap.
      tc::mbar_expect_tx(wbar, 27 * LWT_BYTES);
      for (int kh = 0; kh < 3; ++kh)
        for (int i = 0; i < 3; ++i)
          for (int kw = 0; kw < 3; ++kw) {
            const int kd = 2 - i;
            const int tap = p.flip ? ((2 - kd) * 3 + (2 - kh)) * 3 + (2 - kw) : (kd * 3 + kh) * 3 + kw;
            tc::tma_load_2d(wsm + ((kh * 3 + i) * 3 + kw) * LWT_BYTES, &p.tmB, wbar, 0, tap * 16);
          }
      uint32_t g = 0;  // source slabs issued by this CTA
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / per_n, rem = item % per_n;
        const int seg = rem / p.tilesH, h0 = (rem % p.tilesH) * LPT;
        const int d_begin = seg * p.dseg;
        const int nd = min(p.dseg, p.D - d_begin);
        for (int s = 0; s < nd + 2; ++s, ++g) {
          const uint32_t slot = g % LRING;
          tc::mbar_wait(&empty[slot], ((g / LRING) & 1u) ^ 1u);
          if (p.debug & 4) { tc::mbar_arrive(&full[slot]); continue; }
          tc::mbar_expect_tx(&full[slot], SLAB_BYTES);
          tc::tma_load_5d(ring + slot * SLAB_BYTES, &p.tmA, &full[slot], 0, 0, h0 - 1, d_begin - 1 + s, n);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue: the whole warp runs the warp-uniform loop, one elected lane issues (tc::umma_bf16_warp)
    const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);
    const uint32_t w_addr = tc::smem_u32(wsm), r_addr = tc::smem_u32(ring);
    const uint64_t a_tmpl = tc::make_smem_desc(0, 16, 8 * 64, tc::LAYOUT_SW64);  // pair rows of 64 bytes
    const uint64_t b_tmpl = tc::make_smem_desc(0, 16, 8 * 32, tc::LAYOUT_SW32);  // weight rows of 32 bytes
    const uint64_t w_desc = b_tmpl + (w_addr >> 4);
    tc::mbar_wait(wbar, 0);
    uint32_t g = 0, ob = 0;  // source slabs consumed / output slabs started before this item
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int seg = (item % per_n) / p.tilesH;
      const int d_begin = seg * p.dseg;
      const int nd = min(p.dseg, p.D - d_begin);
      for (int s = 0; s < nd + 2; ++s, ++g) {
        // source slab s (absolute d = d_begin - 1 + s) feeds output slabs s-2, s-1, s (clipped to [0, nd))
        const int lo = max(s - 2, 0), hi = min(s, nd - 1);
        if (s < nd) {  // output slab s receives its first contribution: its chunk must have been drained
          const uint32_t o = ob + (uint32_t)s;
          tc::mbar_wait(&acc_empty[o % LACCR], ((o / LACCR) & 1u) ^ 1u);
        }
        const uint32_t slot = g % LRING;
        tc::mbar_wait(&full[slot], (g / LRING) & 1u);
        tc::tc_fence_after();
        const int cnt = hi - lo + 1;
        const int c_lo = (int)((ob + (uint32_t)lo) % LACCR);
        const int len0 = min(cnt, LACCR - c_lo), len1 = cnt - len0;  // the chunk ring wraps: two MMAs
        const uint32_t i0 = tc::make_idesc_bf16(128, len0 * LCHUNK, false, false);
        const uint32_t i1 = tc::make_idesc_bf16(128, (len1 > 0 ? len1 : 1) * LCHUNK, false, false);
        const uint64_t b0 = w_desc + (((lo - (s - 2)) * 3 * LWT_BYTES) >> 4);
        const uint64_t b1 = b0 + ((len0 * 3 * LWT_BYTES) >> 4);
        const uint64_t slab = a_tmpl + ((r_addr + slot * SLAB_BYTES) >> 4);
        if (!(p.debug & 2))
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint64_t a = slab + ((kh * LINE_BYTES + r * 32) >> 4);
            const uint64_t bo = (kh * 9 * LWT_BYTES) >> 4;
            const uint32_t d0 = tmem_acc + r * LRSTRIDE + c_lo * LCHUNK, d1 = tmem_acc + r * LRSTRIDE;
            tc::umma_bf16_warp(d0, a, b0 + bo, i0, 1u);
            if (len1 > 0) tc::umma_bf16_warp(d1, a, b1 + bo, i1, 1u);
          }
        tc::umma_commit_warp(&empty[slot]);
        if (s >= 2) tc::umma_commit_warp(&acc_full[(ob + (uint32_t)(s - 2)) % LACCR]);
      }
      ob += (uint32_t)nd;
    }
  } else {
    // ---- epilogue: 8 * EG warps.  Warp group (eg, half) = four warps, one per TMEM lane quadrant q; group eg of EG
    // takes every EG-th output slab and, inside it, `half` = 0 the even voxel of every pair, 1 the odd one: a thread
    // produces ONE voxel (16 channels) from three of the six 16-column blocks of the slab's two accumulators --
    //   out[even] = odd.kw0 of the previous pair + even.kw1 + odd.kw2
    //   out[odd]  = even.kw0 + odd.kw1 + even.kw2 of the next pair
    // -- so the per-slab instruction stream, which bounds this kernel (B200SEG_LINE_DEBUG: 120 us with it, 61 us with
    // the accumulators only drained and handed back), is split over twice the warps at half the registers.
    const int sub = (warp - 2) >> 2;
    const int half = sub & 1, eg = sub >> 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;          // pair index inside the tile
    const int l = row / PPL, pp = row % PPL;
    const bool edge = half == 0 ? pp == 0 : pp == PPL - 1;  // no neighbour on that side: zero padding
    const uint32_t lane_base = *tmem_slot + ((uint32_t)(q * 32) << 16);
    float nslope = 0.f;
    if constexpr (BST) nslope = p.nalpha[0];
    uint32_t ob = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int n = item / per_n, rem = item % per_n;
      const int seg = rem / p.tilesH, h0 = (rem % p.tilesH) * LPT;
      const int d_begin = seg * p.dseg;
      const int nd = min(p.dseg, p.D - d_begin);
      const int oh = h0 + l;
      const bool valid = oh < p.H;
      float ssum[BST ? 1 : 16], ssq[BST ? 1 : 16];
      float sb0[BST ? CS : 1], sb1[BST ? CS : 1], sb2[BST ? CS : 1];
#pragma unroll
      for (int c = 0; c < (BST ? 1 : 16); ++c) ssum[c] = ssq[c] = 0.f;
#pragma unroll
      for (int c = 0; c < (BST ? CS : 1); ++c) sb0[c] = sb1[c] = sb2[c] = 0.f;
      float nr[BST ? CS : 1], nb[BST ? CS : 1];  // xhat = x * nr + nb for the consumer InstanceNorm of this sample
      if constexpr (BST) {
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          nr[c] = p.nrstd[n * p.nstat_ld + c];
          nb[c] = -p.nmean[n * p.nstat_ld + c] * nr[c];
        }
      }
      // rows the epilogue READS from global memory (residual addend, the consumer InstanceNorm's pre-norm tensor)
      // are fetched one slab of this group ahead -- right after the previous slab consumed its rows, so that they
      // travel during the wait for the accumulator (see tc_slide.cu: a global-load round trip inside the serial
      // per-slab loop otherwise)
      uint4 pr0 = make_uint4(0, 0, 0, 0), pr1 = pr0, px0 = pr0, px1 = pr0;
      const int64_t lin0 = (((int64_t)n * p.D + d_begin) * p.H + oh) * p.W + 2 * pp + half;  // this thread's voxel, slab 0
      const int64_t slab_vox = (int64_t)p.H * p.W;
      auto prefetch = [&](int j) {
        const int64_t lin = lin0 + j * slab_vox;
        if (p.res) {
          const uint4* r0 = reinterpret_cast<const uint4*>(p.res + lin * p.res_ld);
          pr0 = r0[0]; pr1 = r0[1];
        }
        if constexpr (BST) {
          const uint4* x0 = reinterpret_cast<const uint4*>(p.nx + lin * p.nx_ld);
          px0 = x0[0]; px1 = x0[1];
        }
      };
      int j0 = (int)((EG - (ob % EG) + eg) % EG);  // first slab of this group inside the item
      if (valid && j0 < nd) prefetch(j0);
      for (int j = j0; j < nd; j += EG) {
        const uint32_t o = ob + (uint32_t)j;
        const uint32_t chunk = o % LACCR;
        tc::mbar_wait(&acc_full[chunk], (o / LACCR) & 1u);
        tc::tc_fence_after();
        // ctr = own voxel's kw1 block, same = the pair partner's block, nbr = the block of the neighbouring pair
        const uint32_t ce = lane_base + chunk * LCHUNK, co = ce + LRSTRIDE;
        const uint32_t a_ctr = (half ? co : ce) + 16;
        const uint32_t a_same = half ? ce : co + 32;        // odd: even.kw0      even: odd.kw2
        const uint32_t a_nbr = half ? ce + 32 : co;         // odd: even.kw2 (next pair)   even: odd.kw0 (previous pair)
        uint32_t vc[16], vs[16], vn[16];
        tc::tmem_ld16(a_ctr, vc);
        tc::tmem_ld16(a_same, vs);
        tc::tmem_ld16(a_nbr, vn);
        tc::tmem_ld_wait();
        // hand the blocks back zeroed before anything else (the chunk is free again after all 8 warps did)
        tc::tmem_st16_zero(a_ctr);
        tc::tmem_st16_zero(a_same);
        tc::tmem_st16_zero(a_nbr);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[chunk]);
        if (p.debug & 1) continue;
        float nb16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float v = __uint_as_float(vn[i]);
          const float u = __shfl_up_sync(0xffffffffu, v, 1), d2 = __shfl_down_sync(0xffffffffu, v, 1);
          nb16[i] = half ? d2 : u;
        }
        if constexpr (PPL > 32) {  // a line spans two warps: rows 31 | 32 exchange through shared memory
          // double buffered by the parity of the GROUP's slab count: a warp that is one slab ahead of its group
          // writes the other buffer, and cannot be two ahead (the barrier of the slab in between)
          float* xb = xch + ((((o / EG) & 1u) * (2 * EG) + sub) * 4) * 16;
          if (lane == (half ? 0 : 31)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xb[q * 16 + i] = __uint_as_float(vn[i]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + sub) : "memory");
          if (lane == (half ? 31 : 0) && !edge) {
            const int qq = half ? q + 1 : q - 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) nb16[i] = xb[qq * 16 + i];
          }
        }
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            f[i] = (edge ? 0.f : nb16[i]) + __uint_as_float(vc[i]) + __uint_as_float(vs[i]) + bias_s[i];
          const int64_t lin = lin0 + j * slab_vox;
          auto add_rows = [&](const uint4& a0, const uint4& a1) {
            const __nv_bfloat162* g0 = reinterpret_cast<const __nv_bfloat162*>(&a0);
            const __nv_bfloat162* g1 = reinterpret_cast<const __nv_bfloat162*>(&a1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = __bfloat1622float2(g0[i]), b = __bfloat1622float2(g1[i]);
              f[2 * i] += a.x; f[2 * i + 1] += a.y;
              f[8 + 2 * i] += b.x; f[8 + 2 * i + 1] += b.y;
            }
          };
          if (p.res) add_rows(pr0, pr1);
          uint4* op = reinterpret_cast<uint4*>(p.dst + lin * p.dst_ld);
          if (p.accumulate) {
            const uint4 a0 = op[0], a1 = op[1];
            add_rows(a0, a1);
          }
          if constexpr (!BST) {
            if (p.stats) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                ssum[i] += f[i];
                ssq[i] = fmaf(f[i], f[i], ssq[i]);
              }
            }
          }
          uint4 s0, s1;
          __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&s0);
          __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&s1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            q0[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            q1[i] = __floats2bfloat162_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
          }
          op[0] = s0;
          op[1] = s1;
          if constexpr (BST) {
            // the sums of the InstanceNorm + PReLU backward this gradient feeds, from the values AS STORED (bf16)
            const __nv_bfloat162* hx0 = reinterpret_cast<const __nv_bfloat162*>(&px0);
            const __nv_bfloat162* hx1 = reinterpret_cast<const __nv_bfloat162*>(&px1);
            float xv[16], gv[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = __bfloat1622float2(hx0[i]), b = __bfloat1622float2(hx1[i]);
              xv[2 * i] = a.x; xv[2 * i + 1] = a.y; xv[8 + 2 * i] = b.x; xv[8 + 2 * i + 1] = b.y;
              const float2 ga = __bfloat1622float2(q0[i]), gb = __bfloat1622float2(q1[i]);
              gv[2 * i] = ga.x; gv[2 * i + 1] = ga.y; gv[8 + 2 * i] = gb.x; gv[8 + 2 * i + 1] = gb.y;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < CS) {  // (compile time)
                const float h = fmaf(xv[i], nr[i < CS ? i : 0], nb[i < CS ? i : 0]);
                const bool pos = h > 0.f;
                const float g = pos ? gv[i] : nslope * gv[i];
                sb0[i < CS ? i : 0] += g;
                sb1[i < CS ? i : 0] = fmaf(g, h, sb1[i < CS ? i : 0]);
                sb2[i < CS ? i : 0] += pos ? 0.f : gv[i] * h;
              }
            }
          }
          if (j + EG < nd) prefetch(j + EG);
        }
      }
      // per-warp partial statistics of this item (rows of one sample are contiguous: item order is sample-major)
      if constexpr (BST) {
        float* out = p.bstats + (((int64_t)item * (2 * EG) + sub) * 4 + q) * 16 * 3;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float a = 0.f, b = 0.f, d3 = 0.f;
          if (c < CS) {  // (compile time; padding channels: exact zeros)
            a = warp_sum(sb0[c < CS ? c : 0]);
            b = warp_sum(sb1[c < CS ? c : 0]);
            d3 = warp_sum(sb2[c < CS ? c : 0]);
          }
          if (lane == 0) {
            out[c * 3] = a;
            out[c * 3 + 1] = b;
            out[c * 3 + 2] = d3;
          }
        }
      } else if (p.stats) {
        float* out = p.stats + (((int64_t)item * (2 * EG) + sub) * 4 + q) * p.cout * 2;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float a = warp_sum(ssum[c]), b = warp_sum(ssq[c]);
          if (lane == 0 && c < p.cout) {
            out[c * 2] = a;
            out[c * 2 + 1] = b;
          }
        }
      }
      ob += (uint32_t)nd;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<LTMEM_COLS>(*tmem_slot);
}

// ------------------------------------------------------------------------------------------------
namespace {

int sm_count() {
  static const int v = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    return n;
  }();
  return v;
}

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int x = atoi(e);
  return x >= lo && x <= hi ? x : dflt;
}

// OFF by default (B200SEG_LINE_CONV=1 enables it; B200SEG_LINE_W128=1 adds rows of 128 voxels).  Validated on the GPU
// (the -m gpu kernel / network / guard tests pass with it on) but not faster than tc_slide_conv where it counts --
// r2, head layer 10->10 @128^3 x2:  fprop 100-118 us (slide) vs 106 us (line), dgrad + residual 118 vs 202 us, whole
// step 2.096 ms (slide) vs 2.265 ms (line everywhere) / 2.106 ms (line for rows of 32 / 64 voxels only).
// Where its time goes (B200SEG_LINE_DEBUG, fprop): everything 106 us; accumulators only drained and handed back 61 us;
// additionally no MMAs 38 us; additionally no TMA loads 31 us.  So the folded MMAs are cheap (~25 us: N=144 costs
// 72 cycles against 46 for N=48, scripts/ubench/tmem_drain.cu) and what remains is (a) the hand-back round trip of an
// accumulator chunk -- commit -> mbarrier -> tcgen05.ld of three blocks (each wait::ld ~90 cycles while MMAs run) ->
// tcgen05.st zeros -> arrive -- with only 5 chunks of 96 columns fitting the 512 TMEM columns, i.e. two slabs of
// run-ahead, and (b) ~45 us of per-voxel epilogue instructions on 8 warps.  The residual variant additionally keeps
// global loads in flight across the releasing mbarrier arrive.  tc_slide_conv hides the same costs behind three
// co-resident CTAs per SM.  Kept as a tested experiment; the next step would be first-touch MMAs instead of the zero
// fill and a third TMEM parity-free layout (48 instead of 96 columns per slab).
int line_enabled() { static const int v = env_int("B200SEG_LINE_CONV", 0, 0, 1); return v; }
int line_eg() { static const int v = env_int("B200SEG_LINE_EG", 1, 1, 2); return v; }
int line_w128() { static const int v = env_int("B200SEG_LINE_W128", 0, 0, 1); return v; }

struct LineGeom {
  int n, D, H, W, src_c, dst_c, src_ld, dst_ld;
  int lpt, tilesH, dseg, nseg, items;
};

bool line_geom(const b200seg_conv_desc* d, int op, LineGeom& g) {
  if (op != TC_CONV_FPROP && op != TC_CONV_DGRAD) return false;
  if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 1 || d->sh != 1 || d->sw != 1) return false;
  g.n = d->n; g.D = d->in_d; g.H = d->in_h; g.W = d->in_w;
  if (op == TC_CONV_FPROP) { g.src_c = d->cin; g.dst_c = d->cout; g.src_ld = d->x_ld; g.dst_ld = d->y_ld; }
  else { g.src_c = d->cout; g.dst_c = d->cin; g.src_ld = d->y_ld; g.dst_ld = d->x_ld; }
  // rows of 128 voxels run (B200SEG_LINE_W128=1, tests) but are slower than tc_slide_conv there: draining three kw
  // accumulators per voxel through tcgen05.ld costs more than the folded MMAs save
  if (g.W != 32 && g.W != 64 && !(g.W == 128 && line_w128())) return false;
  if (g.src_c > 16 || g.dst_c > 16 || g.src_ld != 16) return false;
  if (g.D < 4 || (int64_t)g.D * g.H < 64) return false;
  g.lpt = 256 / g.W;
  g.tilesH = (g.H + g.lpt - 1) / g.lpt;
  // d segments: the split that minimises rounds x slabs per item over the persistent CTAs (fewest segments on ties:
  // every segment re-reads two halo slabs)
  const int sms = sm_count();
  int64_t best = -1;
  g.nseg = 1;
  for (int ns = 1; ns <= g.D / 4 && ns <= 64; ++ns) {
    const int ds = (g.D + ns - 1) / ns;
    const int real = (g.D + ds - 1) / ds;
    if (real != ns) continue;
    const int64_t items = (int64_t)g.n * g.tilesH * ns;
    const int64_t cost = ((items + sms - 1) / sms) * (ds + 2);
    if (best < 0 || cost < best) { best = cost; g.nseg = ns; }
  }
  g.dseg = (g.D + g.nseg - 1) / g.nseg;
  const int64_t items = (int64_t)g.n * g.tilesH * g.nseg;
  if (items > 0x3fffffff) return false;
  g.items = (int)items;
  return true;
}

template <int PPL, int EG, bool BST, int CS>
int launch_line(const TcLineConvParams& p, cudaStream_t st) {
  constexpr int SLAB = (128 / PPL + 2) * PPL * 64;
  size_t smem = 1024 + LW_BYTES + (size_t)LRING * SLAB + (2 * LRING + 2 * LACCR + 1) * 8 + 16 + 2 * EG * 4 * 32 * 4 + 64 + 64;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA of this kernel per SM: it owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_line_conv_kernel<PPL, EG, BST, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  const int grid = p.items < sm_count() ? p.items : sm_count();
  tc_line_conv_kernel<PPL, EG, BST, CS><<<grid, 64 + 256 * EG, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH(BST ? "tc_line_conv_bwdstats" : "tc_line_conv");
  count_tc_launch();
  return B200SEG_OK;
}

template <int PPL, int EG>
int launch_line_variant(const TcLineConvParams& p, bool bst, int cs, cudaStream_t st) {
  if (!bst) return launch_line<PPL, EG, false, 16>(p, st);
  if (cs == 10) return launch_line<PPL, EG, true, 10>(p, st);
  return launch_line<PPL, EG, true, 16>(p, st);
}

}  // namespace

bool tc_line_conv_supported(const b200seg_conv_desc* d, int op) {
  LineGeom g;
  return line_enabled() && line_geom(d, op, g);
}

// rows of per-warp partial statistics (rows of one sample contiguous)
int64_t tc_line_conv_rows(const b200seg_conv_desc* d, int op) {
  LineGeom g;
  if (!line_geom(d, op, g)) return 0;
  return (int64_t)g.items * line_eg() * 8;
}

int tc_line_conv_run(const b200seg_conv_desc* d, int op, const void* src, const void* w_tc, const float* bias,
                     const void* residual, void* dst, float* stats, cudaStream_t st, const TcBwdStats* bst) {
  LineGeom g;
  if (!line_geom(d, op, g)) { set_error("tc_line_conv: unsupported layer"); return B200SEG_ERR_UNSUPPORTED; }
  TcLineConvParams p;
  memset(&p, 0, sizeof(p));
  if (bst) {
    p.nx = (const bf16*)bst->nx; p.nx_ld = bst->nx_ld; p.nstat_ld = bst->nstat_ld;
    p.nmean = bst->mean; p.nrstd = bst->rstd; p.nalpha = bst->alpha; p.bstats = bst->partials;
  }
  p.n = g.n; p.D = g.D; p.H = g.H; p.W = g.W;
  p.tilesH = g.tilesH; p.dseg = g.dseg; p.nseg = g.nseg; p.items = g.items;
  p.cout = g.dst_c; p.dst_ld = g.dst_ld; p.res_ld = d->r_ld;
  p.accumulate = (d->flags & B200SEG_CONV_ACCUMULATE) ? 1 : 0;
  p.flip = (op == TC_CONV_DGRAD) ? 1 : 0;
  p.bias = bias; p.res = (const bf16*)residual; p.dst = (bf16*)dst; p.stats = stats;
  static const int dbg = [] {
    const int v = env_int("B200SEG_LINE_DEBUG", 0, 0, 7);
    if (v) fprintf(stderr, "b200seg: B200SEG_LINE_DEBUG=%d switches parts of the kernel off: timing experiment, results are WRONG\n", v);
    return v;
  }();
  p.debug = dbg;
  {
    // the source as (N, D, H, W/2) voxel pairs of 32 bf16 (pairs are contiguous: voxel stride = 16 elements)
    uint64_t dims[5] = {32, (uint64_t)g.W / 2, (uint64_t)g.H, (uint64_t)g.D, (uint64_t)g.n};
    uint64_t strides[4] = {64, (uint64_t)g.W * 32, (uint64_t)g.H * g.W * 32, (uint64_t)g.D * g.H * g.W * 32};
    uint32_t box[5] = {32, (uint32_t)(g.W / 2), (uint32_t)(g.lpt + 2), 1, 1};
    int rc = tc_make_map(&p.tmA, src, 5, dims, strides, box, 64);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {16, (uint64_t)27 * 16};
    uint64_t strides[1] = {32};
    uint32_t box[2] = {16, 16};
    int rc = tc_make_map(&p.tmB, w_tc, 2, dims, strides, box, 32);
    if (rc) return rc;
  }
  const bool b = bst != nullptr;
  const int cs = g.dst_c == 10 ? 10 : 16;
  const int eg = line_eg();
  if (g.W == 128) return eg == 2 ? launch_line_variant<64, 2>(p, b, cs, st) : launch_line_variant<64, 1>(p, b, cs, st);
  if (g.W == 64) return eg == 2 ? launch_line_variant<32, 2>(p, b, cs, st) : launch_line_variant<32, 1>(p, b, cs, st);
  return eg == 2 ? launch_line_variant<16, 2>(p, b, cs, st) : launch_line_variant<16, 1>(p, b, cs, st);
}

}  // namespace b200seg

// Sliding-window tcgen05 weight gradient for the small-channel / high-resolution 3x3x3 stride-1
// conv layers:   G[tap][a][b] = sum_v S[v + tap - 1][a] * T[v][b]      (S = x, T = dy)
//
// Same column sweep and 3-copy source-slab ring as tc_slide.cu (each S voxel is read 3.4x from L2
// instead of 27x); T is loaded once per slab.  Voxels are the K dimension, so both operands are
// MN-major exactly as TMA delivers them.  The M rows of a tcgen05.mma are "slots" = consecutive LINE
// offsets into an S copy (leading-dimension byte offset = one line = one swizzle atom): the first
// three slots are the taps kh = 0, 1, 2, the other rows are never read back.  Its N columns are
// folded over kd: an S slab s pairs with the T slabs s-2, s-1, s (taps kd = 2, 1, 0), which lie in
// adjacent slots of the T ring (N-chunk stride = one T slab; split in two where the ring wraps), so
// one instruction of N = 3*CB replaces three of N = CB -- these small MMAs are bound by the
// shared-memory read of their A tile.  Three accumulators (one per kw, one issuing warp each) of
// 3*CB columns stay in TMEM for the whole sweep (zeroed with tcgen05.st, every MMA accumulates);
// per-CTA partial tiles are summed by the unpack kernel.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace b200seg {

using bf16 = __nv_bfloat16;

int tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, int row_bytes);

namespace {
constexpr int TH = 16;
constexpr int TWV = 8;
constexpr int RT = 6;
inline int round16(int c) { return (c + 15) / 16 * 16; }
}  // namespace

struct alignas(64) TcSlideWgradParams {
  CUtensorMap tmS;
  CUtensorMap tmT;
  int n, D, H, W;
  int tilesH, tilesW, dseg, nseg;
  int a_pad, b_pad;
  float* out;  // [CTA][27][a_pad][b_pad] fp32 partial sums
};

template <int CA, int CB>
__global__ void __launch_bounds__(192)
tc_slide_wgrad_kernel(const __grid_constant__ TcSlideWgradParams p) {
  constexpr int RING = 5;  // source slabs in flight: deep enough to hide the TMA latency
  constexpr int PA = CA * 2, PB = CB * 2;
  constexpr int COPY_BYTES = (TH + 2) * TWV * PA;
  constexpr int SLAB_BYTES = 3 * COPY_BYTES;
  constexpr int TSLAB_BYTES = TH * TWV * PB;
  constexpr uint32_t TMEM_COLS = 9 * CB <= 256 ? 256 : 512;
  // CA = 16: an M = 64 instruction holds 4 line slots (3 taps) and reads half the A bytes of M = 128
  // from shared memory -- these small-N MMAs are bound by the A-operand read, not by the math.
  // M = 64 accumulator rows live in lanes 0..15 of each 32-lane TMEM quarter: quarter = slot (kh).
  constexpr int MM = CA == 16 ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* tring = smem + RING * SLAB_BYTES;  // doubles as the read slack behind the last S copy
  uint64_t* fullS = reinterpret_cast<uint64_t*>(tring + RT * TSLAB_BYTES);
  uint64_t* emptyS = fullS + RING;
  uint64_t* fullT = emptyS + RING;
  uint64_t* emptyT = fullT + RT;
  uint64_t* acc_full = emptyT + RT;
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
    // three MMA-issuing threads (one per kd): a stage is free / the result complete after all three
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&fullS[i], 1); tc::mbar_init(&emptyS[i], 3); }
    for (int i = 0; i < RT; ++i) { tc::mbar_init(&fullT[i], 1); tc::mbar_init(&emptyT[i], 3); }
    tc::mbar_init(acc_full, 3);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmS);
    tc::prefetch_tmap(&p.tmT);
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  // ---- accumulators start at zero: every MMA accumulates (the N-folded instructions touch several
  // kd chunks whose first contributions come at different slabs)
  if (warp >= 2) {
    const uint32_t lane_base = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 9 * CB; c += 16) tc::tmem_st16_zero(lane_base + c);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < nd + 2; ++s) {
        const int slot = s % RING;
        tc::mbar_wait(&emptyS[slot], (((uint32_t)(s / RING)) & 1u) ^ 1u);
        uint8_t* dst = ring + slot * SLAB_BYTES;
        tc::mbar_expect_tx(&fullS[slot], SLAB_BYTES);
        const int ds = d_begin - 1 + s;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          tc::tma_load_5d(dst + kw * COPY_BYTES, &p.tmS, &fullS[slot], 0, w0 + kw - 1, h0 - 1, ds, n);
        if (s < nd) {  // T slab j = s is first needed together with S slab s (tap kd = 0)
          const int j = s, ts = j % RT;
          tc::mbar_wait(&emptyT[ts], (((uint32_t)(j / RT)) & 1u) ^ 1u);
          tc::mbar_expect_tx(&fullT[ts], TSLAB_BYTES);
          tc::tma_load_5d(tring + ts * TSLAB_BYTES, &p.tmT, &fullT[ts], 0, w0, h0, d_begin + j, n);
        }
      }
    }
  }
  // ---- MMA issue: warps 1, 2, 3 each own one kw (one accumulator of 3*CB columns); warp-uniform loop,
  // one elected lane issues
  else if (warp <= 3) {  // (`else`, not a second `if`: see tc::warp_index)
    constexpr uint64_t layA = tc::layout_for_row_bytes(PA), layB = tc::layout_for_row_bytes(PB);
    const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform registers: see tc::warp_uniform
    const int iw = (int)tc::warp_uniform((uint32_t)warp) - 1;
    const uint32_t r_addr = tc::smem_u32(ring), t_addr = tc::smem_u32(tring);
    const uint64_t a_tmpl = tc::make_smem_desc(0, TWV * PA, TWV * PA, layA);    // slots one line apart
    const uint64_t b_tmpl = tc::make_smem_desc(0, TSLAB_BYTES, TWV * PB, layB); // N chunks one T slab apart
    const uint32_t acc = tmem_acc + iw * 3 * CB;
    int waitedT = 0;
    for (int s = 0; s < nd + 2; ++s) {
      tc::mbar_wait(&fullS[s % RING], ((uint32_t)(s / RING)) & 1u);
      const int t_last = min(s, nd - 1);
      while (waitedT <= t_last) {
        tc::mbar_wait(&fullT[waitedT % RT], ((uint32_t)(waitedT / RT)) & 1u);
        ++waitedT;
      }
      tc::tc_fence_after();
      // chunk i of the accumulator <-> T slab j = s - 2 + i <-> tap kd = 2 - i; valid chunks are contiguous
      const int i_lo = max(0, 2 - s), i_hi = min(2, nd + 1 - s);
      const int cnt = i_hi - i_lo + 1;
      const int slot_lo = (s - 2 + i_lo) % RT;
      const int len0 = min(cnt, RT - slot_lo), len1 = cnt - len0;
      const uint32_t i0 = tc::make_idesc_bf16(MM, len0 * CB, true, true);
      const uint32_t i1 = tc::make_idesc_bf16(MM, (len1 > 0 ? len1 : 1) * CB, true, true);
      const uint32_t d0 = acc + i_lo * CB, d1 = d0 + len0 * CB;
      const uint64_t sa = a_tmpl + ((r_addr + (s % RING) * SLAB_BYTES + iw * COPY_BYTES) >> 4);
      const uint64_t tb0 = b_tmpl + ((t_addr + slot_lo * TSLAB_BYTES) >> 4);
      const uint64_t tb1 = b_tmpl + (t_addr >> 4);  // a wrapped part starts at ring slot 0
#pragma unroll
      for (int t = 0; t < TH / 2; ++t) {
        // K step t = output lines 2t, 2t+1; slot i of A starts at S line 2t + i
        const uint64_t a = sa + (((2 * t) * (TWV * PA)) >> 4);
        const uint32_t bo = ((2 * t) * (TWV * PB)) >> 4;
        tc::umma_bf16_warp(d0, a, tb0 + bo, i0, 1u);
        if (len1 > 0) tc::umma_bf16_warp(d1, a, tb1 + bo, i1, 1u);
      }
      tc::umma_commit_warp(&emptyS[s % RING]);
      if (s >= 2) tc::umma_commit_warp(&emptyT[(s - 2) % RT]);
    }
    tc::umma_commit_warp(acc_full);
  }
  if (warp >= 2) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ih = MM == 64 ? q : row / CA;
    const int a = MM == 64 ? (lane & 15) : row % CA;
    const bool valid = MM == 64 ? (q < 3 && lane < 16) : (ih < 3);
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    if (MM == 64 ? (q < 3) : (q * 32 < 3 * CA)) {  // warp-uniform: only rows that are taps
      for (int acc = 0; acc < 9; ++acc) {
        const int iw = acc / 3, id = 2 - acc % 3;  // column chunk (iw, i) holds tap kd = 2 - i
        const int tap = (id * 3 + (valid ? ih : 0)) * 3 + iw;
        float* orow = p.out + (int64_t)blockIdx.x * 27 * p.a_pad * p.b_pad + ((int64_t)tap * p.a_pad + a) * p.b_pad;
#pragma unroll
        for (int ch = 0; ch < CB / 16; ++ch) {
          uint32_t v[16];
          tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + acc * CB + ch * 16, v);
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

// gw[b][a][tap] = sum_cta G[cta][tap][a][b]: the block-per-row unpack of tc_wgrad.cu (coalesced along b,
// partial tiles dealt over split lanes, fixed summation order)
int tc_wgrad_unpack(const float* G, float* gw, int taps, int a_c, int b_c, int a_pad, int b_pad, int parts,
                    const char* what, cudaStream_t st);

// host launcher shared with the hi/lo sliding wgrad (tc_convtr.cu)
int tc_slide_wgrad_unpack(const float* G, float* gw, int taps, int a_c, int b_c, int a_pad, int b_pad, int nparts,
                          cudaStream_t st) {
  return tc_wgrad_unpack(G, gw, taps, a_c, b_c, a_pad, b_pad, nparts, "tc_slide_wgrad_unpack", st);
}

namespace {
template <int CA, int CB>
int launch_slide_wgrad(const TcSlideWgradParams& p, unsigned grid, cudaStream_t st) {
  constexpr int RING = 5;
  constexpr int SLAB = 3 * (TH + 2) * TWV * CA * 2;
  constexpr int TSLAB = TH * TWV * CB * 2;
  const size_t smem = 1024 + RING * SLAB + RT * TSLAB + 40 * 8 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_slide_wgrad_kernel<CA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  tc_slide_wgrad_kernel<CA, CB><<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_slide_wgrad");
  count_tc_launch();
  return B200SEG_OK;
}
}  // namespace

namespace {
// column / d-segment decomposition shared by the launcher and the workspace query
void slide_wgrad_grid(const b200seg_conv_desc* d, int& tilesH, int& tilesW, int& dseg, int& nseg, int64_t& grid) {
  tilesH = (d->in_h + TH - 1) / TH;
  tilesW = (d->in_w + TWV - 1) / TWV;
  const int64_t cols = (int64_t)d->n * tilesH * tilesW;
  // one wave: 2 resident CTAs per SM for 16/16 channels (TMEM 256 columns, ~95 KB), else 1
  const int per_sm = (round16(d->cin) == 16 && round16(d->cout) == 16) ? 2 : 1;
  int ns = (int)((148 * per_sm) / cols);
  if (ns < 1) ns = 1;
  dseg = (d->in_d + ns - 1) / ns;
  if (dseg < 8) dseg = 8;
  if (dseg > d->in_d) dseg = d->in_d;
  nseg = (d->in_d + dseg - 1) / dseg;
  grid = cols * nseg;
}
}  // namespace

size_t tc_slide_wgrad_workspace(const b200seg_conv_desc* d) {
  int th, tw, dseg, nseg;
  int64_t grid;
  slide_wgrad_grid(d, th, tw, dseg, nseg, grid);
  return (size_t)grid * 27 * round16(d->cin) * round16(d->cout) * sizeof(float);
}

// conv layers only (S = x, T = dy): 3-D 3x3x3 stride 1, padded channel counts in {16, 32}
bool tc_slide_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer) {
  if (transposed_layer || (d->flags & B200SEG_CONV_NO_SLIDE)) return false;
  if (d->kd != 3 || d->kh != 3 || d->kw != 3 || d->sd != 1 || d->sh != 1 || d->sw != 1) return false;
  const int ap = round16(d->cin), bp = round16(d->cout);
  if ((ap != 16 && ap != 32) || (bp != 16 && bp != 32)) return false;
  if (d->in_d < 8 || (int64_t)d->in_h * d->in_w < 512) return false;
  if (tc_slide_wgrad_workspace(d) > (size_t)256 << 20) return false;  // partial tiles: one per CTA
  return true;
}

int tc_slide_wgrad_run(const b200seg_conv_desc* d, const void* x, const void* dy, float* gw, float* G32,
                       cudaStream_t st) {
  TcSlideWgradParams p;
  memset(&p, 0, sizeof(p));
  const int CA = round16(d->cin), CB = round16(d->cout);
  p.n = d->n; p.D = d->in_d; p.H = d->in_h; p.W = d->in_w;
  int64_t grid;
  slide_wgrad_grid(d, p.tilesH, p.tilesW, p.dseg, p.nseg, grid);
  p.a_pad = CA; p.b_pad = CB; p.out = G32;
  {
    uint64_t dims[5] = {(uint64_t)CA, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.n};
    uint64_t strides[4] = {(uint64_t)d->x_ld * 2, (uint64_t)p.W * d->x_ld * 2, (uint64_t)p.H * p.W * d->x_ld * 2,
                           (uint64_t)p.D * p.H * p.W * d->x_ld * 2};
    uint32_t box[5] = {(uint32_t)CA, (uint32_t)TWV, (uint32_t)(TH + 2), 1, 1};
    int rc = tc_make_map(&p.tmS, x, 5, dims, strides, box, CA * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)CB, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.n};
    uint64_t strides[4] = {(uint64_t)d->y_ld * 2, (uint64_t)p.W * d->y_ld * 2, (uint64_t)p.H * p.W * d->y_ld * 2,
                           (uint64_t)p.D * p.H * p.W * d->y_ld * 2};
    uint32_t box[5] = {(uint32_t)CB, (uint32_t)TWV, (uint32_t)TH, 1, 1};
    int rc = tc_make_map(&p.tmT, dy, 5, dims, strides, box, CB * 2);
    if (rc) return rc;
  }

  int rc;
  if (CA == 16 && CB == 16) rc = launch_slide_wgrad<16, 16>(p, (unsigned)grid, st);
  else if (CA == 16 && CB == 32) rc = launch_slide_wgrad<16, 32>(p, (unsigned)grid, st);
  else if (CA == 32 && CB == 16) rc = launch_slide_wgrad<32, 16>(p, (unsigned)grid, st);
  else rc = launch_slide_wgrad<32, 32>(p, (unsigned)grid, st);
  if (rc) return rc;
  return tc_slide_wgrad_unpack(G32, gw, 27, d->cin, d->cout, CA, CB, (int)grid, st);
}

}  // namespace b200seg

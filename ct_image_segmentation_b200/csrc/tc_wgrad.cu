// tcgen05 / TMEM / TMA weight-gradient kernel (bf16 operands, fp32 accumulation in TMEM).
//
//     G[tap][a][b] = sum_{n,o} S[n, o*stride - pad + tap, a] * T[n, o, b]
//
// (conv: S = x, T = dy; ConvTranspose: S = dy, T = x).  The reduction dimension of the GEMM is the
// VOXEL index, so both operands are consumed exactly as they lie in memory -- channels-last tiles
// [voxel][channel] loaded by TMA are "MN-major" UMMA operands (channel = M/N, voxel = K); no
// transposes, no im2col.  The M dimension of one tcgen05.mma (128 rows) is filled with several
// (tap, channel-chunk) "slots": each slot is one shifted TMA box of S, consecutive slots sit at the
// descriptor's leading-dimension byte offset.  N = channels of T.  All taps that fit into the 512
// TMEM columns accumulate concurrently ("pass"); voxel tiles are split across CTAs (split-K) and
// the fp32 partial sums are combined with red.global.add.f32 into a zeroed workspace, which a
// small kernel then transposes into the PyTorch parameter layout.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace b200seg {

using bf16 = __nv_bfloat16;

int tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, int row_bytes);

struct TcWSlot {
  int8_t map, dd, dh, dw;
  int16_t chan;  // first S channel of the slot
  int16_t tap;
};

constexpr int kMaxSlots = 224;

struct alignas(64) TcWgradParams {
  CUtensorMap tmS[8];
  CUtensorMap tmT;
  TcWSlot slots[kMaxSlots];
  int nslots;            // real slots
  int spm;               // slots per 128-row M tile (= 128 / CA)
  int mt_per_pass;       // M tiles per pass (TMEM budget)
  int CA, CB;            // channels per S slot / per T chunk (16, 32 or 64)
  int nb;                // T chunks per N tile
  int N;                 // N of the MMA (= nb * CB)
  int kv;                // voxels per stage
  int n;
  int tD, tH, tW;        // voxel tile box
  int tilesD, tilesH, tilesW;
  int64_t total_tiles;
  int tiles_per_cta;
  int stages;
  int a_pad, b_pad;
  float* out;            // [splits][taps][a_pad][b_pad] fp32 partial sums (plain stores, no atomics)
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __launch_bounds__(192, 1)
tc_wgrad_kernel(const __grid_constant__ TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int pitchA = p.CA * 2, pitchB = p.CB * 2;
  const int tileA = p.kv * pitchA, tileB = p.kv * pitchB;
  const int slots_pass_cap = p.mt_per_pass * p.spm;
  const int stage_bytes = (p.nb * tileB + slots_pass_cap * tileA + 1023) / 1024 * 1024;
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = tc::warp_index(), lane = threadIdx.x & 31;
  const int pass = blockIdx.z;
  const int n0 = blockIdx.y * p.N;
  const int slot_begin = pass * slots_pass_cap;
  const int slot_end = min(slot_begin + slots_pass_cap, p.nslots);
  const int nsl = slot_end - slot_begin;
  const int mtiles = (nsl + p.spm - 1) / p.spm;
  const int64_t tile_begin = (int64_t)blockIdx.x * p.tiles_per_cta;
  const int64_t tile_end = min(tile_begin + p.tiles_per_cta, p.total_tiles);
  const int ntiles = (int)(tile_end - tile_begin);

  // up to three MMA-issuing threads (warps 1..3), each owning the M tiles mt = i, i+3, ... (distinct
  // accumulators): one thread cannot issue these small MMAs fast enough to keep the tensor pipe busy
  const int nissue = mtiles < 3 ? mtiles : 3;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], nissue);
    }
    tc::mbar_init(acc_full, nissue);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&p.tmT);
  }
  if (warp == 1) tmem_alloc_dyn(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (ntiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t tx = (uint32_t)(p.nb * tileB + nsl * tileA);
        int st = 0;
        uint32_t ph = 0;
        for (int64_t t = tile_begin; t < tile_end; ++t) {
          int64_t r = t;
          const int w0 = (int)(r % p.tilesW) * p.tW; r /= p.tilesW;
          const int h0 = (int)(r % p.tilesH) * p.tH; r /= p.tilesH;
          const int d0 = (int)(r % p.tilesD) * p.tD; r /= p.tilesD;
          const int n = (int)r;
          tc::mbar_wait(&empty[st], ph ^ 1u);
          uint8_t* base = smem + (size_t)st * stage_bytes;
          tc::mbar_expect_tx(&full[st], tx);
          for (int c = 0; c < p.nb; ++c)
            tc::tma_load_5d(base + c * tileB, &p.tmT, &full[st], n0 + c * p.CB, w0, h0, d0, n);
          uint8_t* sa = base + p.nb * tileB;
          for (int s = 0; s < nsl; ++s) {
            const TcWSlot sl = p.slots[slot_begin + s];
            tc::tma_load_5d(sa + s * tileA, &p.tmS[sl.map], &full[st], sl.chan, w0 + sl.dw, h0 + sl.dh,
                            d0 + sl.dd, n);
          }
          if (++st == stages) { st = 0; ph ^= 1u; }
        }
      }
    } else if (warp <= nissue) {  // warp-uniform issue loop, one elected lane issues (`else`: see tc::warp_index)
      const uint32_t tmem_acc = tc::warp_uniform(*tmem_slot);  // uniform registers: see tc::warp_uniform
      const int warp_u = (int)tc::warp_uniform((uint32_t)warp);
      const uint32_t idesc = tc::make_idesc_bf16(128, p.N, true, true);
      const uint64_t layA = tc::layout_for_row_bytes(pitchA), layB = tc::layout_for_row_bytes(pitchB);
      const uint64_t a_tmpl = tc::make_smem_desc(0, tileA, 8 * pitchA, layA);
      const uint64_t b_tmpl = tc::make_smem_desc(0, tileB, 8 * pitchB, layB);
      const int ksteps = p.kv / 16;
      const uint32_t a_kstep = (16 * pitchA) >> 4, b_kstep = (16 * pitchB) >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < ntiles; ++it) {
        tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        const uint32_t b_addr = tc::smem_u32(smem + (size_t)st * stage_bytes);
        const uint64_t bd0 = b_tmpl + (b_addr >> 4);
        for (int mt = warp_u - 1; mt < mtiles; mt += nissue) {
          const uint64_t ad0 = a_tmpl + ((b_addr + p.nb * tileB + mt * p.spm * tileA) >> 4);
          const uint32_t acc = tmem_acc + mt * p.N;
          for (int j = 0; j < ksteps; ++j)
            tc::umma_bf16_warp(acc, ad0 + j * a_kstep, bd0 + j * b_kstep, idesc, (it > 0 || j > 0) ? 1u : 0u);
        }
        tc::umma_commit_warp(&empty[st]);
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
      tc::umma_commit_warp(acc_full);
    }
    if (warp >= 2) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      tc::mbar_wait(acc_full, 0);
      tc::tc_fence_after();
      for (int mt = 0; mt < mtiles; ++mt) {
        const int sidx = slot_begin + mt * p.spm + row / p.CA;
        const bool valid = sidx < slot_end;
        int tap = 0, a = 0;
        if (valid) {
          const TcWSlot sl = p.slots[sidx];
          tap = sl.tap;
          a = sl.chan + row % p.CA;
        }
        const int64_t g_elems = (int64_t)(p.nslots / (p.a_pad / p.CA)) * p.a_pad * p.b_pad;
        float* orow = p.out + (int64_t)blockIdx.x * g_elems + ((int64_t)tap * p.a_pad + a) * p.b_pad + n0;
        for (int ch = 0; ch < p.N / 16; ++ch) {
          uint32_t v[16];
          tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + mt * p.N + ch * 16, v);
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
  if (warp == 1) tmem_dealloc_dyn(tmem_acc, 512);
}

// gw[b][a][tap] = sum_split G[split][tap][a][b] for the real channels.  One block per (tap, a) row:
// threads = (split lane, b); coalesced reads along b, fixed summation order (deterministic).
__global__ void __launch_bounds__(256)
tc_wgrad_unpack_kernel(const float* __restrict__ G, float* __restrict__ gw, int taps, int a_c, int b_c, int a_pad,
                       int b_pad, int splits) {
  __shared__ float red[256];
  const int tap = blockIdx.x / a_c, a = blockIdx.x % a_c;
  int bw = 16;
  while (bw < b_c && bw < 256) bw <<= 1;  // threads along b (power of two)
  const int sy_n = 256 / bw;              // split lanes
  const int bx = threadIdx.x % bw, sy = threadIdx.x / bw;
  const int64_t g_elems = (int64_t)taps * a_pad * b_pad;
  const float* row = G + ((int64_t)tap * a_pad + a) * b_pad;
  for (int b0 = 0; b0 < b_c; b0 += bw) {
    const int b = b0 + bx;
    float s = 0.f;
    if (b < b_c)
      for (int i = sy; i < splits; i += sy_n) s += row[(int64_t)i * g_elems + b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = sy_n / 2; o >= 1; o >>= 1) {
      if (sy < o) red[threadIdx.x] += red[threadIdx.x + o * bw];
      __syncthreads();
    }
    if (sy == 0 && b < b_c) gw[((int64_t)b * a_c + a) * taps + tap] = red[bx];
    __syncthreads();
  }
}

// gw (PyTorch layout) = sum over `parts` partial tiles G[part][tap][a_pad][b_pad]; shared by the sliding kernels
int tc_wgrad_unpack(const float* G, float* gw, int taps, int a_c, int b_c, int a_pad, int b_pad, int parts,
                    const char* what, cudaStream_t st) {
  tc_wgrad_unpack_kernel<<<(unsigned)(taps * a_c), 256, 0, st>>>(G, gw, taps, a_c, b_c, a_pad, b_pad, parts);
  B200SEG_CHECK_LAUNCH(what);
  return B200SEG_OK;
}

namespace {

inline int round16(int c) { return (c + 15) / 16 * 16; }
inline int chunk_for(int pad) { return pad % 64 == 0 ? 64 : (pad % 32 == 0 ? 32 : 16); }

struct WGeom {
  int n, sD, sH, sW, tD, tH, tW, a_c, b_c, s_ld, t_ld;
  int k[3], s[3], p[3];
  const void* S;
  const void* T;
};

void wgeom(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy, WGeom& g) {
  g.n = d->n;
  g.k[0] = d->kd; g.k[1] = d->kh; g.k[2] = d->kw;
  g.s[0] = d->sd; g.s[1] = d->sh; g.s[2] = d->sw;
  g.p[0] = d->pd; g.p[1] = d->ph; g.p[2] = d->pw;
  if (!transposed_layer) {  // S = x (gathered), T = dy
    g.sD = d->in_d; g.sH = d->in_h; g.sW = d->in_w; g.tD = d->out_d; g.tH = d->out_h; g.tW = d->out_w;
    g.a_c = d->cin; g.b_c = d->cout; g.s_ld = d->x_ld; g.t_ld = d->y_ld; g.S = x; g.T = dy;
  } else {  // S = dy (gathered, the larger grid), T = x
    g.sD = d->out_d; g.sH = d->out_h; g.sW = d->out_w; g.tD = d->in_d; g.tH = d->in_h; g.tW = d->in_w;
    g.a_c = d->cout; g.b_c = d->cin; g.s_ld = d->y_ld; g.t_ld = d->x_ld; g.S = dy; g.T = x;
  }
}

int n_tile_for(int b_pad) {
  if (b_pad <= 256) return b_pad;
  if (b_pad % 256 == 0) return 256;
  if (b_pad % 192 == 0) return 192;
  if (b_pad % 128 == 0) return 128;
  if (b_pad % 64 == 0) return 64;
  return 16;
}

}  // namespace

// split-K partial tiles: at most ~4 M fp32 per launch (see the split heuristic) plus one tile; the
// sliding-window kernels need one tile per CTA (<= 2048 CTAs of 27*32*32)
size_t tc_wgrad_extra_workspace(const b200seg_conv_desc* d) {
  if (d->dtype != B200SEG_BF16) return 0;
  const size_t g = (size_t)d->kd * d->kh * d->kw * round16(d->cin) * round16(d->cout);
  size_t streaming = (g + (4u << 20)) * sizeof(float);
  size_t sliding = tc_slide_wgrad_supported(d, false) ? tc_slide_wgrad_workspace(d) : 0;
  for (int tl = 0; tl < 2; ++tl)
    if (tc_convtr_wgrad_supported(d, tl != 0) && tc_convtr_wgrad_workspace(d, tl != 0) > sliding)
      sliding = tc_convtr_wgrad_workspace(d, tl != 0);
  return align_up(streaming > sliding ? streaming : sliding, 256);
}

bool tc_wgrad_supported(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy) {
  if (d->dtype != B200SEG_BF16 || (d->flags & B200SEG_CONV_FORCE_GENERIC)) return false;
  WGeom g;
  wgeom(d, transposed_layer, x, dy, g);
  const bool padded = (d->flags & B200SEG_CONV_PADDED_CHANNELS) != 0;
  if (g.a_c < 8 || g.b_c < 8) return false;
  if ((g.a_c % 16) && !(padded && g.s_ld >= round16(g.a_c))) return false;
  if ((g.b_c % 16) && !(padded && g.t_ld >= round16(g.b_c))) return false;
  if ((g.s_ld % 8) || (g.t_ld % 8)) return false;
  if (((uintptr_t)g.S % 16) || ((uintptr_t)g.T % 16)) return false;
  const int a_pad = round16(g.a_c), CA = chunk_for(a_pad);
  const int taps = g.k[0] * g.k[1] * g.k[2];
  if (taps * (a_pad / CA) > kMaxSlots) return false;
  return true;
}

int tc_wgrad_run(const b200seg_conv_desc* d, bool transposed_layer, const void* x, const void* dy, float* gw,
                 float* G32, cudaStream_t st) {
  WGeom g;
  wgeom(d, transposed_layer, x, dy, g);
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  const int a_pad = round16(g.a_c), b_pad = round16(g.b_c);
  const int CA = chunk_for(a_pad);
  const int N = n_tile_for(b_pad);
  const int CB = chunk_for(N);
  const int taps = g.k[0] * g.k[1] * g.k[2];
  p.CA = CA; p.CB = CB; p.N = N; p.nb = N / CB;
  p.spm = 128 / CA;
  p.a_pad = a_pad; p.b_pad = b_pad;
  p.n = g.n;
  p.out = G32;
  int gcap = 512 / N;  // M tiles that fit in TMEM
  if (gcap < 1) gcap = 1;

  // ---- slots
  bool map_used[8] = {false, false, false, false, false, false, false, false};
  int ns = 0;
  const int achunks = a_pad / CA;
  for (int kd = 0; kd < g.k[0]; ++kd)
    for (int kh = 0; kh < g.k[1]; ++kh)
      for (int kw = 0; kw < g.k[2]; ++kw) {
        const int kk[3] = {kd, kh, kw};
        int off[3], par[3];
        for (int i = 0; i < 3; ++i) {
          int e = kk[i] - g.p[i];
          if (g.s[i] == 2) { par[i] = e & 1; off[i] = (e - par[i]) / 2; }
          else { par[i] = 0; off[i] = e; }
        }
        for (int c = 0; c < achunks; ++c) {
          TcWSlot& sl = p.slots[ns++];
          sl.map = (int8_t)((par[0] * 2 + par[1]) * 2 + par[2]);
          sl.dd = (int8_t)off[0]; sl.dh = (int8_t)off[1]; sl.dw = (int8_t)off[2];
          sl.chan = (int16_t)(c * CA);
          sl.tap = (int16_t)((kd * g.k[1] + kh) * g.k[2] + kw);
          map_used[sl.map] = true;
        }
      }
  p.nslots = ns;
  const int mtiles_total = (ns + p.spm - 1) / p.spm;
  p.mt_per_pass = mtiles_total < gcap ? mtiles_total : gcap;
  // Voxels per stage (a 4-deep ring in ~190 KB) shrink with the number of M tiles staged together;
  // 16-voxel stages are a string of 2 KB TMA boxes and one K step per MMA tile.  Prefer fewer M tiles
  // per pass (more passes = more CTAs, T is re-read from L2) until a stage holds >= 64 voxels.
  auto kv_for = [&](int mt) {
    const int pv = p.nb * CB * 2 + mt * p.spm * CA * 2;
    int k = 128;
    while (k > 16 && (size_t)k * pv * 4 > 190 * 1024) k /= 2;
    return k;
  };
  static const int kv_target = getenv("B200SEG_WGRAD_KV") ? atoi(getenv("B200SEG_WGRAD_KV")) : 64;
  while (p.mt_per_pass > 1 && kv_for(p.mt_per_pass) < kv_target) p.mt_per_pass = (p.mt_per_pass + 1) / 2;
  const int passes = (mtiles_total + p.mt_per_pass - 1) / p.mt_per_pass;

  // ---- voxels per stage and ring depth (shared memory budget ~ 190 KB)
  const int slots_pass_cap = p.mt_per_pass * p.spm;
  const int per_vox = p.nb * CB * 2 + slots_pass_cap * CA * 2;
  int kv = 128;
  while (kv > 16 && (size_t)kv * per_vox * 4 > 190 * 1024) kv /= 2;  // aim for a 4-deep ring
  if ((size_t)kv * per_vox > 190 * 1024) {
    set_error("tc_wgrad: stage does not fit shared memory");
    return B200SEG_ERR_UNSUPPORTED;
  }
  p.kv = kv;
  const int stage_bytes = (kv * per_vox + 1023) / 1024 * 1024;
  int stages = (190 * 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 1) stages = 1;
  p.stages = stages;

  // ---- voxel tile box (kv voxels) minimising padding of the T grid
  const int tdim[3] = {g.tD, g.tH, g.tW};
  {
    int64_t best = -1;
    for (int td = 1; td <= kv; td *= 2)
      for (int th = 1; td * th <= kv; th *= 2) {
        int tw = kv / (td * th);
        int64_t vol = (int64_t)((tdim[0] + td - 1) / td) * ((tdim[1] + th - 1) / th) * ((tdim[2] + tw - 1) / tw);
        int64_t score = vol * 1024 - (tw > 16 ? 16 : tw) * 8 - (th > 16 ? 16 : th);
        if (best < 0 || score < best) { best = score; p.tD = td; p.tH = th; p.tW = tw; }
      }
  }
  p.tilesD = (tdim[0] + p.tD - 1) / p.tD; p.tilesH = (tdim[1] + p.tH - 1) / p.tH; p.tilesW = (tdim[2] + p.tW - 1) / p.tW;
  p.total_tiles = (int64_t)g.n * p.tilesD * p.tilesH * p.tilesW;

  // ---- tensor maps
  const int sdim[3] = {g.sD, g.sH, g.sW};
  for (int m = 0; m < 8; ++m) {
    if (!map_used[m]) continue;
    const int par[3] = {(m >> 2) & 1, (m >> 1) & 1, m & 1};
    const int str[3] = {g.s[0] == 2 ? 2 : 1, g.s[1] == 2 ? 2 : 1, g.s[2] == 2 ? 2 : 1};
    const bf16* base = (const bf16*)g.S + (((int64_t)par[0] * g.sH + par[1]) * g.sW + par[2]) * g.s_ld;
    uint64_t dims[5] = {(uint64_t)a_pad, (uint64_t)((sdim[2] - par[2] + str[2] - 1) / str[2]),
                        (uint64_t)((sdim[1] - par[1] + str[1] - 1) / str[1]),
                        (uint64_t)((sdim[0] - par[0] + str[0] - 1) / str[0]), (uint64_t)g.n};
    uint64_t strides[4] = {(uint64_t)g.s_ld * 2 * str[2], (uint64_t)g.sW * g.s_ld * 2 * str[1],
                           (uint64_t)g.sH * g.sW * g.s_ld * 2 * str[0], (uint64_t)g.sD * g.sH * g.sW * g.s_ld * 2};
    uint32_t box[5] = {(uint32_t)CA, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tD, 1};
    int rc = tc_make_map(&p.tmS[m], base, 5, dims, strides, box, CA * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)b_pad, (uint64_t)g.tW, (uint64_t)g.tH, (uint64_t)g.tD, (uint64_t)g.n};
    uint64_t strides[4] = {(uint64_t)g.t_ld * 2, (uint64_t)g.tW * g.t_ld * 2, (uint64_t)g.tH * g.tW * g.t_ld * 2,
                           (uint64_t)g.tD * g.tH * g.tW * g.t_ld * 2};
    uint32_t box[5] = {(uint32_t)CB, (uint32_t)p.tW, (uint32_t)p.tH, (uint32_t)p.tD, 1};
    int rc = tc_make_map(&p.tmT, g.T, 5, dims, strides, box, CB * 2);
    if (rc) return rc;
  }

  // ---- split-K: enough CTAs for ~2 waves, at least 4 voxel tiles each
  const int ntile_n = b_pad / N;
  int64_t want = (2 * 148 + passes * ntile_n - 1) / (passes * ntile_n);
  int64_t max_split = (p.total_tiles + 3) / 4;
  if (max_split < 1) max_split = 1;
  int64_t splits = want < max_split ? want : max_split;
  // every split adds taps*a_pad*b_pad fp32 red.global.add: keep that below ~4 M per launch
  const int64_t g_elems = (int64_t)taps * a_pad * b_pad;
  int64_t atomic_cap = (4 << 20) / g_elems;
  if (atomic_cap < 1) atomic_cap = 1;
  if (splits > atomic_cap) splits = atomic_cap;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (int)((p.total_tiles + splits - 1) / splits);
  splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16;
  dim3 grid((unsigned)splits, (unsigned)ntile_n, (unsigned)passes);
  tc_wgrad_kernel<<<grid, 192, smem, st>>>(p);
  B200SEG_CHECK_LAUNCH("tc_wgrad");
  count_tc_launch();
  const int64_t total = (int64_t)taps * g.a_c * g.b_c;
  (void)total;
  return tc_wgrad_unpack(G32, gw, taps, g.a_c, g.b_c, a_pad, b_pad, (int)splits, "tc_wgrad_unpack", st);
}

}  // namespace b200seg

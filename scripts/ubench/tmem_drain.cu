// Micro-benchmark: do tcgen05.mma and tcgen05.ld / tcgen05.st share a pipe?  One thread issues N-column MMAs back to
// back (as mma_rate.cu) while four other warps (one per TMEM lane quadrant) drain LD columns per round with
// tcgen05.ld (and optionally zero them again with tcgen05.st), as the sliding conv kernels' epilogue does.
// Reports cycles per MMA with the drain warps idle / running, and cycles per drained 16-column block alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_drain tmem_drain.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ct_image_segmentation_b200/csrc/tc_common.cuh"
using namespace b200seg;

// MODE bit 0: MMAs run; bit 1: drain warps run; bit 2: drain warps also zero what they read (tcgen05.st)
template <int N, int LD, int MODE>
__global__ void __launch_bounds__(160) k(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  for (int i = threadIdx.x; i < (4 * 128 * 32 + N * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); done = 0; }
  if (threadIdx.x < 32) tc::tmem_alloc<512>(&slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t acc = slot;
  const int warp = threadIdx.x >> 5;
  long long t_mma = 0, n_ld = 0, t_ld = 0;
  if (threadIdx.x == 0) {
    if (MODE & 1) {
      const uint32_t idesc = tc::make_idesc_bf16(128, N, false, false);
      const uint32_t a_addr = tc::smem_u32(smem), b_addr = a_addr + 4 * 128 * 32;
      const uint64_t ad = tc::make_smem_desc(a_addr, 16, 256, tc::LAYOUT_SW32);
      const uint64_t bd = tc::make_smem_desc(b_addr, 16, 256, tc::LAYOUT_SW32);
      for (int i = 0; i < 16; ++i) tc::umma_bf16(acc, ad, bd, idesc, 1u);
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) tc::umma_bf16(acc, ad + (uint64_t)((i & 3) * (128 * 32 / 16)), bd, idesc, 1u);
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 1);
      t_mma = clock64() - t0;
    }
    done = 1;
  } else if (warp >= 1 && (MODE & 2)) {
    // drain columns [256, 256 + LD) of this warp's lane quadrant, over and over (disjoint from the accumulator)
    const uint32_t base = acc + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    const long long t0 = clock64();
    const int fixed = (MODE & 1) ? (1 << 30) : iters;   // without MMAs: a fixed number of rounds
    for (int r = 0; r < fixed; ++r) {
      if ((MODE & 1) && done) break;
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < LD; c += 16) {
        tc::tmem_ld16(base + c, v);
        tc::tmem_ld_wait();
        if (MODE & 4) tc::tmem_st16_zero(base + c);
      }
      if (MODE & 4) tc::tmem_st_wait();
      asm volatile("" ::"r"(v[0]), "r"(v[15]));
      ++n_ld;
    }
    t_ld = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) out[0] = t_mma;
    if (threadIdx.x == 32) { out[1] = n_ld; out[2] = t_ld; }
  }
  if (threadIdx.x < 32) tc::tmem_dealloc<512>(acc);
}

template <int N, int LD, int MODE>
void run(long long* d, int grid) {
  const int iters = 2000;
  const int smem = 4 * 128 * 32 + N * 32 + 2048;
  cudaFuncSetAttribute(k<N, LD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d, 0, 24);
  k<N, LD, MODE><<<grid, 160, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[3] = {0, 0, 0};
  cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("N=%3d drain=%3d cols mode=%d ctas/SM=%d : %7.1f cycles/MMA   drain rounds %6lld  %8.1f cycles/round (%5.1f per 16-col block)  (%s)\n",
         N, LD, MODE, grid / 148, (MODE & 1) ? (double)h[0] / iters : 0.0, h[1], h[1] ? (double)h[2] / h[1] : 0.0,
         h[1] ? (double)h[2] / h[1] / (LD / 16) : 0.0, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 24);
  for (int g : {148}) {
    run<48, 16, 1>(d, g);    // MMAs alone
    run<48, 16, 2>(d, g);    // drain alone (ld)
    run<48, 16, 6>(d, g);    // drain alone (ld + st)
    run<48, 16, 3>(d, g);    // MMAs + ld
    run<48, 16, 7>(d, g);    // MMAs + ld + st
    run<48, 96, 2>(d, g);
    run<48, 96, 3>(d, g);
    run<144, 96, 1>(d, g);
    run<144, 96, 3>(d, g);
    run<144, 96, 7>(d, g);
    run<256, 16, 1>(d, g);
    run<256, 16, 3>(d, g);
  }
  return 0;
}

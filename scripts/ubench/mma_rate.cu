// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, SS operands) as a function of
// M, N and the shared-memory layout (row bytes 32 / 64 / 128), issued back to back by one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ct_image_segmentation_b200/csrc/tc_common.cuh"
using namespace b200seg;

template <int M, int N, int ROWB, int DISTINCT_A, int NACC, int NA = 4, int ASTR = 0, int NB = 1>
__global__ void __launch_bounds__(128) k(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (NA * M * ROWB + NB * N * ROWB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc<128>(&slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t acc = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(M, N, false, false);
    const uint64_t lay = tc::layout_for_row_bytes(ROWB);
    const uint32_t a_addr = tc::smem_u32(smem), b_addr = a_addr + NA * M * ROWB;
    const uint64_t ad = tc::make_smem_desc(a_addr, 16, 8 * ROWB, lay);
    const uint64_t bd = tc::make_smem_desc(b_addr, 16, 8 * ROWB, lay);
    // warm
    for (int i = 0; i < 16; ++i) tc::umma_bf16(acc, ad, bd, idesc, 1u);
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // DISTINCT_A: walk over 4 different A tiles (as the conv kernels do), else the same tile
      const uint64_t a = DISTINCT_A ? ad + (uint64_t)((i % NA) * ((ASTR ? ASTR : M * ROWB) / 16)) : ad;
      const uint64_t b = bd + (uint64_t)((i % NB) * (N * ROWB / 16));
      tc::umma_bf16(acc + (NACC > 1 ? (i % NACC) * N : 0), a, b, idesc, 1u);
    }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 1);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc<128>(acc);
}

template <int M, int N, int ROWB, int DA, int NACC = 1, int NA = 4, int ASTR = 0, int NB = 1>
void run(long long* d, int grid) {
  const int iters = 1998;
  const int smem = NA * M * ROWB + NB * N * ROWB + 2048;
  cudaFuncSetAttribute(k<M, N, ROWB, DA, NACC, NA, ASTR, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<M, N, ROWB, DA, NACC, NA, ASTR, NB><<<grid, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d rowbytes=%3d nA=%d aStride=%d nB=%d nacc=%d ctas/SM=%d : %.1f cycles/MMA  (%s)\n", M, N, ROWB, DA ? NA : 1, ASTR ? ASTR : M * ROWB, NB, NACC, grid / 148,
         (double)h / iters, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int g : {148, 296, 444, 592}) {
    run<128, 48, 32, 1, 1, 4, 0, 1>(d, g);
    run<128, 16, 32, 1, 1, 4, 0, 1>(d, g);
    run<128, 96, 32, 1, 1, 4, 0, 1>(d, g);
  }
  return 0;
}

// Launch-gap microbenchmark: a chain of N dependent small kernels captured into a CUDA graph, with plain
// stream-order dependencies vs. programmatic dependent launch (griddepcontrol.wait at the top of every kernel,
// launch_dependents right after).  Prints us per node.   nvcc -arch=sm_100a -O3 -o pdl_gap pdl_gap.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <bool PDL>
__global__ void node(float* buf, int n, int work) {
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = buf[i];
    for (int k = 0; k < work; ++k) v = v * 1.0001f + 0.5f;
    buf[i] = v;
  }
}

template <bool PDL>
float run(int nodes, int blocks, int work, float* buf, int n) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < nodes; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = PDL ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, node<PDL>, buf, n, work));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGraphExecDestroy(ge));
  CK(cudaGraphDestroy(g));
  CK(cudaStreamDestroy(st));
  return ms * 1e3f / (reps * nodes);
}

int main() {
  const int n = 148 * 8 * 256;
  float* buf;
  CK(cudaMalloc(&buf, n * sizeof(float)));
  CK(cudaMemset(buf, 0, n * sizeof(float)));
  const int nodes = 200;
  for (int blocks : {1, 148, 148 * 8}) {
    for (int work : {0, 2000}) {
      float a = run<false>(nodes, blocks, work, buf, n);
      float b = run<true>(nodes, blocks, work, buf, n);
      printf("blocks %5d work %5d: plain %.3f us/node   PDL %.3f us/node   (saves %.3f)\n", blocks, work, a, b, a - b);
    }
  }
  return 0;
}

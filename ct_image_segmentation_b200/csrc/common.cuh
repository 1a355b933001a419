// Shared helpers for the b200seg kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200seg.h"

namespace b200seg {

void set_error(const char* fmt, ...);
void count_launch();
void note_launch(const char* what);
void count_tc_launch();

#define B200SEG_CHECK_ARG(cond, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      b200seg::set_error(__VA_ARGS__);    \
      return B200SEG_ERR_ARG;             \
    }                                     \
  } while (0)

#define B200SEG_CHECK_LAUNCH(what)                                               \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    b200seg::count_launch();                                                     \
    b200seg::note_launch(what);                                                  \
    if (e__ != cudaSuccess) {                                                    \
      b200seg::set_error("%s: CUDA error: %s", what, cudaGetErrorString(e__));   \
      return B200SEG_ERR_CUDA;                                                   \
    }                                                                            \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element access as float --------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Vector of V consecutive elements of T held as floats.  V*sizeof(T) must be 4, 8 or 16 bytes
// when the vector path is taken; the caller guarantees alignment.
template <typename T, int V>
struct Vec {
  float v[V];
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (sizeof(T) * V == 16) {
      uint4 raw = *reinterpret_cast<const uint4*>(p);
      unpack(reinterpret_cast<const T*>(&raw));
    } else if constexpr (sizeof(T) * V == 8) {
      uint2 raw = *reinterpret_cast<const uint2*>(p);
      unpack(reinterpret_cast<const T*>(&raw));
    } else if constexpr (sizeof(T) * V == 4) {
      uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
      unpack(reinterpret_cast<const T*>(&raw));
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = to_f<T>(p[i]);
    }
  }
  __device__ __forceinline__ void store(T* p) const {
    if constexpr (sizeof(T) * V == 16) {
      uint4 raw;
      pack(reinterpret_cast<T*>(&raw));
      *reinterpret_cast<uint4*>(p) = raw;
    } else if constexpr (sizeof(T) * V == 8) {
      uint2 raw;
      pack(reinterpret_cast<T*>(&raw));
      *reinterpret_cast<uint2*>(p) = raw;
    } else if constexpr (sizeof(T) * V == 4) {
      uint32_t raw;
      pack(reinterpret_cast<T*>(&raw));
      *reinterpret_cast<uint32_t*>(p) = raw;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) p[i] = from_f<T>(v[i]);
    }
  }
  __device__ __forceinline__ void unpack(const T* t) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = to_f<T>(t[i]);
  }
  __device__ __forceinline__ void pack(T* t) const {
#pragma unroll
    for (int i = 0; i < V; ++i) t[i] = from_f<T>(v[i]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200seg

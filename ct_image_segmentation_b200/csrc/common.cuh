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
// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (recursive halving: at every
// step a lane keeps half of its values and receives the partner's partial sums for them) instead of
// 16 x 5.  On return EVERY lane holds the complete sum of channel (lane >> 1) & 15.
__device__ __forceinline__ float warp_sum16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = b4 ? v[8 + i] : v[i], give = b4 ? v[i] : v[8 + i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = b3 ? a8[4 + i] : a8[i], give = b3 ? a8[i] : a8[4 + i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b2 ? a4[2 + i] : a4[i], give = b2 ? a4[i] : a4[2 + i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
  }
  const float keep = b1 ? a2[1] : a2[0], give = b1 ? a2[0] : a2[1];
  const float a1 = keep + __shfl_xor_sync(0xffffffffu, give, 2);
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200seg

// Shared device/host helpers for libhdf_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hdf_b200.h"

void hdf_set_error(const char* fmt, ...);
extern unsigned long long g_hdf_launches;   // kernels launched by this library (bench.py "gpu_launches")

#define HDF_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      hdf_set_error(__VA_ARGS__);              \
      return HDF_ERR_ARG;                      \
    }                                          \
  } while (0)

#define HDF_LAUNCH_CHECK(name)                                                     \
  do {                                                                             \
    ++g_hdf_launches;                                                              \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      hdf_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return HDF_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float[4]; p must be aligned to 4 elements when vec=true
template <typename T> __device__ __forceinline__ void load4(const T* p, float* o);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* o) {
  float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float* o) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float* o);
template <> __device__ __forceinline__ void store4<float>(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, const float* o) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = v;
}

// 8 consecutive elements (16 B for bf16, 32 B for f32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* o) {
  load4<T>(p, o);
  load4<T>(p + 4, o + 4);
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float* o) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* o) {
  store4<T>(p, o);
  store4<T>(p + 4, o + 4);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float* o) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

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

// Counter-based dropout RNG: one 32-bit draw per (seed, call-site id, element).
// Masks are recomputed in backward from the same triple, never stored.
__device__ __forceinline__ uint32_t hdf_rng(uint64_t seed, uint32_t call_id, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(call_id + 1) + idx * 0xD1342543DE82EF95ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// returns multiplier (0 or 1/(1-p))
__device__ __forceinline__ float hdf_dropout_scale(uint64_t seed, uint32_t call_id, uint64_t idx, float p) {
  if (p <= 0.f) return 1.f;
  float u = (float)(hdf_rng(seed, call_id, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? 1.0f / (1.0f - p) : 0.f;
}

#define HDF_DISPATCH_DTYPE(dtype, T, ...)                         \
  if ((dtype) == HDF_F32) { typedef float T; __VA_ARGS__; }       \
  else if ((dtype) == HDF_BF16) { typedef bf16 T; __VA_ARGS__; }  \
  else { hdf_set_error("bad dtype %d", (int)(dtype)); return HDF_ERR_ARG; }

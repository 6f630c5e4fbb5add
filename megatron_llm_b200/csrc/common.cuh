// Shared helpers for the SIMT (memory-bound) kernels: 16-byte vector access for bf16/fp16/fp32,
// block reductions.  All kernels accumulate in fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlb {

enum DType : int { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };

template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec<__half> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec<float> {
  static constexpr int N = 8;  // two 16-byte transactions so every dtype moves 8 elements per step
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// explicit scalar conversions (torch builds with __CUDA_NO_HALF_CONVERSIONS__ / __CUDA_NO_BFLOAT16_CONVERSIONS__)
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float warp_reduce_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_reduce_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum broadcast to every thread; `scratch` holds >= 32 floats
__device__ __forceinline__ float block_reduce_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_reduce_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_reduce_sum(r);
  return r;
}
__device__ __forceinline__ float block_reduce_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_reduce_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_reduce_max(r);
  return r;
}

#define MLB_DISPATCH_DTYPE(dt, T, ...)                                   \
  switch (dt) {                                                          \
    case mlb::DT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }  \
    case mlb::DT_F16: { using T = __half; __VA_ARGS__; break; }          \
    case mlb::DT_F32: { using T = float; __VA_ARGS__; break; }           \
    default: return -100;                                                \
  }

}  // namespace mlb

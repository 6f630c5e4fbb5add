// Minimal CUDA execution model on CPU threads, enough to RUN the repo's kernels in the CPU test suite (overview:
// docs/guide/testing.md): one std::thread per CUDA thread of a block, blocks one after the other,
//   __shared__        -> a static (one block is alive at a time)
//   __syncthreads()   -> std::barrier over the block          __syncwarp() -> std::barrier over the warp
//   __shfl_xor_sync   -> exchange through a per-warp buffer between two warp barriers
// A thread that returns from the kernel drops out of its barriers, so early exits of whole warps do not dead-lock.
// The fake <cuda_runtime.h> / <cuda_bf16.h> / <cuda_fp16.h> next to this file all include it.  This header alone serves
// the SIMT kernels; the tcgen05 / TMA / mbarrier kernels additionally run on tcgen05_model.h + ptx.cuh.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

using std::max;
using std::min;
using std::isfinite;
using std::isinf;
using std::isnan;

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
// (natural alignment only: kernels reinterpret register arrays as vectors, which the host must not turn into aligned
// SSE moves; the 16-byte alignment of global / shared addresses is a property of the device code under test)
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }

// ---- 16-bit floating point types (round-to-nearest-even conversions)
struct __nv_bfloat16 { uint16_t bits; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
struct __half { _Float16 v; };
struct __half2 { __half x, y; };
inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = uint32_t(h.bits) << 16; float f; std::memcpy(&f, &u, 4); return f;
}
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u; std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{uint16_t((u >> 16) | 0x40)};   // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{uint16_t(u >> 16)};
}
inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{__bfloat162float(h.x), __bfloat162float(h.y)}; }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
  return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)};
}
inline float __half2float(__half h) { return float(h.v); }
inline __half __float2half_rn(float f) { return __half{_Float16(f)}; }
inline float2 __half22float2(__half2 h) { return float2{float(h.x.v), float(h.y.v)}; }
inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }

// ---- execution context
namespace cuda_emu {
struct Block {
  unsigned n_threads, n_warps;
  std::barrier<> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint32_t> shfl;     // [n_warps][32]
  std::map<int, std::unique_ptr<uint8_t[]>> statics;     // scalar __shared__ variables of kernels whose CTAs overlap
  std::mutex statics_mu;
  explicit Block(unsigned n) : n_threads(n), n_warps((n + 31) / 32), block_bar(n), shfl(((n + 31) / 32) * 32) {
    for (unsigned w = 0; w < n_warps; ++w)
      warp_bar.emplace_back(new std::barrier<>(std::min(32u, n - w * 32)));
  }
};
inline thread_local Block* blk = nullptr;
}  // namespace cuda_emu

namespace cuda_emu {
// a scalar `__shared__ T name;` of a kernel that runs with several CTAs alive at once (clusters, concurrent blocks):
// one instance per block instead of the `static` the plain __shared__ macro gives (host_build.py rewrites those)
template <class T> inline T& block_static(int key) {
  std::lock_guard<std::mutex> g(blk->statics_mu);
  auto& slot = blk->statics[key];
  if (!slot) slot.reset(new uint8_t[sizeof(T)]());
  return *reinterpret_cast<T*>(slot.get());
}
}  // namespace cuda_emu

inline thread_local uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { cuda_emu::blk->block_bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::blk->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
inline uint32_t emu_shfl_bits(uint32_t v, int src_lane) {
  auto* b = cuda_emu::blk;
  const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  b->shfl[w * 32 + lane] = v;
  b->warp_bar[w]->arrive_and_wait();
  const uint32_t r = b->shfl[w * 32 + (unsigned(src_lane) & 31u)];
  b->warp_bar[w]->arrive_and_wait();
  return r;
}
inline float __shfl_xor_sync(unsigned, float v, int o) {
  uint32_t u; std::memcpy(&u, &v, 4);
  u = emu_shfl_bits(u, int(threadIdx.x & 31) ^ o);
  float r; std::memcpy(&r, &u, 4); return r;
}
inline int __shfl_xor_sync(unsigned, int v, int o) { return int(emu_shfl_bits(uint32_t(v), int(threadIdx.x & 31) ^ o)); }

namespace cuda_emu {
// run `kernel()` for every thread of every block of a 1-D-thread-block grid.  The OS threads are created once per
// launch and walk the blocks together (an outer barrier separates consecutive blocks), which keeps launches with
// thousands of small blocks cheap.
template <class F>
void launch(dim3 grid, unsigned threads, F kernel) {
  const unsigned nb = grid.x * grid.y * grid.z;
  if (nb == 0) return;
  std::vector<std::unique_ptr<Block>> blocks(nb);
  for (auto& b : blocks) b.reset(new Block(threads));
  std::barrier<> between(threads);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      threadIdx = uint3{t, 0, 0};
      blockDim = dim3(threads);
      gridDim = grid;
      for (unsigned b = 0; b < nb; ++b) {
        Block& block = *blocks[b];
        blk = &block;
        blockIdx = uint3{b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
        kernel();
        block.warp_bar[t >> 5]->arrive_and_drop();
        block.block_bar.arrive_and_drop();
        between.arrive_and_wait();                 // the next block starts when every thread has left this one
      }
    });
  for (auto& th : pool) th.join();
}
}  // namespace cuda_emu

// ---- the little of the runtime API the launchers touch
using cudaStream_t = void*;
using cudaError_t = int;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
#define CUDART_VERSION 12090
inline cudaError_t cudaFree(void*) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& symbol, size_t bytes) {
  std::memcpy(dst, &symbol, bytes);
  return cudaSuccess;
}
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __fdividef(float a, float b) { return a / b; }
template <class T> inline T __ldg(const T* p) { return *p; }

namespace cuda_emu {
// red.global.add.v4.f32: four relaxed float atomics (blocks run one after the other, threads of a block concurrently)
inline void atomic_add4(float* addr, float a, float b, float c, float d) {
  const float v[4] = {a, b, c, d};
  for (int i = 0; i < 4; ++i) std::atomic_ref<float>(addr[i]).fetch_add(v[i], std::memory_order_relaxed);
}
}  // namespace cuda_emu

// Host stand-in for csrc/ptx.cuh, limited to what the peer-memory kernels of csrc/comm.cu use (system-scope loads /
// stores and fences become C++ atomics; "peers" are other PROCESSES that map the same shared memory, see
// tests/test_kernel_emulation.py).  The tcgen05 / TMA / mbarrier wrappers of the real header have no host meaning and
// are not provided: a translation unit that needs them cannot be built for the host.
#pragma once
#include "cuda_emu.h"

inline void __nanosleep(unsigned) { std::this_thread::yield(); }   // (keeps the bounded spins of the kernels time-like)
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_acq_rel); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
template <class T> inline T __ldcs(const T* p) { return *p; }

namespace mlb {
inline int ld_acquire_sys(const int* p) { return std::atomic_ref<const int>(*p).load(std::memory_order_acquire); }
inline int ld_relaxed_sys(const int* p) { return std::atomic_ref<const int>(*p).load(std::memory_order_relaxed); }
inline void st_release_sys(int* p, int v) { std::atomic_ref<int>(*p).store(v, std::memory_order_release); }
inline uint4 ld_v4_relaxed_sys(const void* p) { uint4 r; std::memcpy(&r, p, 16); return r; }
inline void st_v4(void* p, const uint4& v) { std::memcpy(p, &v, 16); }
}  // namespace mlb

namespace cuda_emu {
// An NVSwitch multicast object: one "multicast address" range that stands for the same offset in every rank's buffer.
// multimem.ld_reduce.add returns the sum over the copies, multimem.st writes all of them.
struct Multicast { const float* mc_base; float* copy[8]; int world; long long n; };
inline Multicast& multicast() { static Multicast m{}; return m; }
inline float4 multimem_ld_reduce_add_v4(const float* mc_addr) {
  const Multicast& m = multicast();
  const long long off = mc_addr - m.mc_base;
  float4 s{0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < m.world; ++r) {
    float4 v; std::memcpy(&v, m.copy[r] + off, 16);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  return s;
}
inline void multimem_st_v4(float* mc_addr, float a, float b, float c, float d) {
  const Multicast& m = multicast();
  const long long off = mc_addr - m.mc_base;
  const float v[4] = {a, b, c, d};
  for (int r = 0; r < m.world; ++r) std::memcpy(m.copy[r] + off, v, 16);
}
}  // namespace cuda_emu

// test hook: describe the multicast object to this process
extern "C" __attribute__((used, visibility("default"))) inline void emu_register_multicast(const float* mc_base, const long long* copies, int world, long long n) {
  auto& m = cuda_emu::multicast();
  m.mc_base = mc_base; m.world = world; m.n = n;
  for (int r = 0; r < world; ++r) m.copy[r] = reinterpret_cast<float*>(copies[r]);
}

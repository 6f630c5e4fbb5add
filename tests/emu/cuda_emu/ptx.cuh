// Host stand-in for csrc/ptx.cuh (see cuda_emu.h and tcgen05_model.h):
//   * system-scope loads / stores / fences of the peer-memory kernels (csrc/comm.cu) become C++ atomics; "peers" are other
//     PROCESSES that map the same files (tests/emu/comm_rank.py);
//   * mbarrier / TMA / tensor memory / tcgen05.mma wrappers run on the functional model of tcgen05_model.h;
//   * the descriptor builders (make_smem_desc, make_idesc_f16) are NOT re-implemented: host_build.py copies their source
//     out of the real header into "ptx_real_extract.h", so the kernels' descriptors are decoded by the model as the
//     real code encodes them.
#pragma once
#include <chrono>

#include "tcgen05_model.h"

inline void __nanosleep(unsigned) { std::this_thread::yield(); }   // (keeps the bounded spins of the kernels time-like)
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return std::atomic_ref<unsigned long long>(*p).fetch_add(v, std::memory_order_acq_rel);
}
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_acq_rel); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }

namespace mlb {
// ---- system-scope flags
inline int ld_acquire_sys(const int* p) { return std::atomic_ref<const int>(*p).load(std::memory_order_acquire); }
inline int ld_relaxed_sys(const int* p) { return std::atomic_ref<const int>(*p).load(std::memory_order_relaxed); }
inline void st_release_sys(int* p, int v) { std::atomic_ref<int>(*p).store(v, std::memory_order_release); }
inline uint4 ld_v4_relaxed_sys(const void* p) { uint4 r; std::memcpy(&r, p, 16); return r; }
inline void st_v4(void* p, const uint4& v) { std::memcpy(p, &v, 16); }

// ---- shared-memory addresses, misc
inline uint32_t smem_u32(const void* p) {
  return uint32_t(static_cast<const uint8_t*>(p) - cuda_emu::bm->smem) | (cuda_emu::bm->cluster_rank << 24);
}
// the same shared-memory object in CTA `rank` of the cluster
template <class T> inline T* cluster_peer_ptr(T* p, uint32_t rank) {
  return reinterpret_cast<T*>(cuda_emu::bm->cluster[rank]->smem + (reinterpret_cast<uint8_t*>(p) - cuda_emu::bm->smem));
}
inline uint32_t cluster_ctarank() { return cuda_emu::bm->cluster_rank; }
inline void cluster_sync() { cuda_emu::bm->cluster_bar->arrive_and_wait(); }
inline uint32_t lane_id() { return threadIdx.x & 31; }
inline uint32_t pack_bf16x2(float a, float b) {
  return uint32_t(__float2bfloat16_rn(a).bits) | (uint32_t(__float2bfloat16_rn(b).bits) << 16);
}
inline uint32_t pack_f16x2(float a, float b) {
  _Float16 x = _Float16(a), y = _Float16(b); uint16_t lo, hi; std::memcpy(&lo, &x, 2); std::memcpy(&hi, &y, 2);
  return uint32_t(lo) | (uint32_t(hi) << 16);
}
inline float2 unpack_bf16x2(uint32_t u) { return float2{cuda_emu::to_float16bits(uint16_t(u), true), cuda_emu::to_float16bits(uint16_t(u >> 16), true)}; }
inline float2 unpack_f16x2(uint32_t u) { return float2{cuda_emu::to_float16bits(uint16_t(u), false), cuda_emu::to_float16bits(uint16_t(u >> 16), false)}; }
inline float warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
inline float warp_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

// ---- mbarrier
inline void mbar_init(uint64_t* bar, uint32_t count) {
  auto* b = reinterpret_cast<cuda_emu::MBar*>(bar);
  std::lock_guard<std::mutex> g(cuda_emu::mbar_mutex());
  *b = cuda_emu::MBar{uint8_t(count), uint8_t(count), 0, 0, 0};
}
inline void fence_barrier_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  cuda_emu::chaos();
  auto* b = reinterpret_cast<cuda_emu::MBar*>(bar);
  for (unsigned spins = 0;; ++spins) {
    {
      std::lock_guard<std::mutex> g(cuda_emu::mbar_mutex());
      if (b->phase != (parity & 1)) return;
    }
    // hundreds of waiters per block: yield a few times, then sleep so that the threads doing the work get the cores
    if (spins < 16) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(spins < 256 ? 20 : 200));
  }
}
inline void mbar_arrive(uint64_t* bar) {
  cuda_emu::chaos();
  auto* b = reinterpret_cast<cuda_emu::MBar*>(bar);
  std::lock_guard<std::mutex> g(cuda_emu::mbar_mutex());
  b->pending -= 1;
  cuda_emu::mbar_check(b);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  cuda_emu::chaos();
  auto* b = reinterpret_cast<cuda_emu::MBar*>(bar);
  std::lock_guard<std::mutex> g(cuda_emu::mbar_mutex());
  b->tx += int32_t(bytes);
  b->pending -= 1;
  cuda_emu::mbar_check(b);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  auto* b = reinterpret_cast<cuda_emu::MBar*>(bar);
  std::lock_guard<std::mutex> g(cuda_emu::mbar_mutex());
  b->tx -= int32_t(bytes);
  cuda_emu::mbar_check(b);
}

// ---- TMA: 4-D tiled load through the 128-byte swizzle
inline void tma_prefetch_desc(const void*) {}
inline void tma_load_tile(void* smem_dst, const void* desc, uint64_t* bar, const int (&c)[4]) {
  cuda_emu::chaos();
  const CUtensorMap& tm = *static_cast<const CUtensorMap*>(desc);
  const uint32_t dst = smem_u32(smem_dst);
  const uint32_t row_bytes = tm.box[0] * tm.elem_bytes;
  uint32_t box[4] = {tm.box[0], 1, 1, 1};
  uint64_t dims[4] = {tm.dims[0], 1, 1, 1}, strides[4] = {tm.elem_bytes, 0, 0, 0};
  for (uint32_t i = 1; i < tm.rank; ++i) { box[i] = tm.box[i]; dims[i] = tm.dims[i]; strides[i] = tm.strides[i]; }
  uint32_t r = 0;
  for (uint32_t i3 = 0; i3 < box[3]; ++i3)
    for (uint32_t i2 = 0; i2 < box[2]; ++i2)
      for (uint32_t i1 = 0; i1 < box[1]; ++i1, ++r) {
        const int64_t x1 = int64_t(c[1]) + i1, x2 = int64_t(c[2]) + i2, x3 = int64_t(c[3]) + i3;
        const bool row_ok = x1 >= 0 && x2 >= 0 && x3 >= 0 && uint64_t(x1) < dims[1] && uint64_t(x2) < dims[2] && uint64_t(x3) < dims[3];
        for (uint32_t byte = 0; byte < row_bytes; byte += tm.elem_bytes) {
          const int64_t x0 = int64_t(c[0]) + byte / tm.elem_bytes;
          uint32_t a = dst + r * row_bytes + byte;
          if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B) a = cuda_emu::swizzle128(a);
          if (row_ok && x0 >= 0 && uint64_t(x0) < dims[0])
            std::memcpy(cuda_emu::smem_ptr(a), tm.base + x0 * tm.elem_bytes + x1 * strides[1] + x2 * strides[2] +
                                                 x3 * strides[3], tm.elem_bytes);
          else
            std::memset(cuda_emu::smem_ptr(a), 0, tm.elem_bytes);          // out-of-bounds elements are zero-filled
        }
      }
  mbar_complete_tx(bar, r * row_bytes);                                      // (the full box counts, in or out of bounds)
}
inline void tma_load_4d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3) {
  tma_load_tile(smem_dst, desc, bar, {c0, c1, c2, c3});
}
inline void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  tma_load_tile(smem_dst, desc, bar, {c0, c1, 0, 0});
}
inline void mbar_arrive_cluster(uint64_t* bar, uint32_t cta_rank) { mbar_arrive(cluster_peer_ptr(bar, cta_rank)); }
// cta_group::2 load: the data lands in THIS CTA's shared memory, the bytes complete on the LEADER CTA's barrier
inline void tma_load_2d_2sm(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  tma_load_tile(smem_dst, desc, cluster_peer_ptr(bar, 0), {c0, c1, 0, 0});
}
// shared -> global tile store (optionally an element-wise add), clipped to the tensor
inline void tma_store_tile_2d(const void* desc, const void* smem_src, int c0, int c1, bool add) {
  const CUtensorMap& tm = *static_cast<const CUtensorMap*>(desc);
  const uint32_t src = smem_u32(smem_src) & 0xFFFFFFu, row_bytes = tm.box[0] * tm.elem_bytes;
  for (uint32_t r = 0; r < tm.box[1]; ++r)
    for (uint32_t byte = 0; byte < row_bytes; byte += tm.elem_bytes) {
      const int64_t x0 = int64_t(c0) + byte / tm.elem_bytes, x1 = int64_t(c1) + r;
      if (x0 < 0 || x1 < 0 || uint64_t(x0) >= tm.dims[0] || uint64_t(x1) >= tm.dims[1]) continue;
      uint32_t a = src + r * row_bytes + byte;
      if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B) a = cuda_emu::swizzle128(a);
      uint8_t* g = const_cast<uint8_t*>(tm.base) + x0 * tm.elem_bytes + x1 * tm.strides[1];
      const uint8_t* sp = cuda_emu::bm->smem + a;
      if (add && tm.elem_bytes == 4) {
        float v; std::memcpy(&v, sp, 4);
        std::atomic_ref<float>(*reinterpret_cast<float*>(g)).fetch_add(v, std::memory_order_relaxed);
      } else {
        std::memcpy(g, sp, tm.elem_bytes);
      }
    }
}
inline void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) { tma_store_tile_2d(desc, smem_src, c0, c1, false); }
inline void tma_reduce_add_2d(const void* desc, const void* smem_src, int c0, int c1) { tma_store_tile_2d(desc, smem_src, c0, c1, true); }
inline void fence_proxy_async_global() {}
inline void fence_proxy_async_all() {}
inline void tma_store_commit() {}
template <int N> inline void tma_store_wait_read() {}
template <int N> inline void tma_store_wait() {}
inline void red_add_release_sys(int* p, int v) { std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_release); }
inline int atom_add_acqrel_sys(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_acq_rel); }
inline unsigned long long globaltimer_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- tensor memory
template <int G = 1> inline void tmem_alloc(uint32_t* smem_out, uint32_t) { if ((threadIdx.x & 31) == 0) *smem_out = 0; }   // (warp-collective)
template <int G = 1> inline void tmem_relinquish() {}
template <int G = 1> inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tmem_ld_wait() {}
inline void tmem_st_wait() {}
template <int N> inline void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]) {
  cuda_emu::chaos();
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < N; ++i) r[i] = cuda_emu::tmem_at(lane, col + i);
}
template <int N> inline void tmem_st_n(uint32_t taddr, const uint32_t (&r)[N]) {
  cuda_emu::chaos();
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < N; ++i) cuda_emu::tmem_at(lane, col + i) = r[i];
}
inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_n<32>(taddr, r); }
inline void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_n<16>(taddr, r); }
inline void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st_n<32>(taddr, r); }

}  // namespace mlb

#define __host__
#include "ptx_real_extract.h"      // enum kSwizzle*, make_smem_desc, make_idesc_f16: copied from the real header

namespace mlb {
// ---- tcgen05.mma kind::f16, cta_group::1, M = 128: D[128 x N] (+)= A[128 x 16] * B[16 x N], executed at issue
struct SmemOperand {
  uint32_t start, lbo, sbo, swizzle;
  explicit SmemOperand(uint64_t d)
      : start(uint32_t(d & 0x3FFF) << 4), lbo(uint32_t((d >> 16) & 0x3FFF) << 4), sbo(uint32_t((d >> 32) & 0x3FFF) << 4),
        swizzle(uint32_t(d >> 61) & 7) {}
  // element (mn, k) of an operand tile, k in [0, 16).  K-major: rows of 128 bytes hold the reduction dim, 8-row atoms are
  // SBO apart.  MN-major: rows of 128 bytes hold 64 mn-elements, the next 64 are LBO apart, 8 k-rows are SBO apart.
  uint16_t at(uint32_t mn, uint32_t k, bool mn_major, const cuda_emu::BlockModel* cta = nullptr) const {
    uint32_t a = mn_major ? start + (mn / 64) * lbo + (k / 8) * sbo + (k % 8) * 128 + (mn % 64) * 2
                          : start + (mn / 8) * sbo + (mn % 8) * 128 + k * 2;
    if (swizzle == kSwizzle128B) a = cuda_emu::swizzle128(a);
    uint16_t v; std::memcpy(&v, (cta ? cta : cuda_emu::bm)->smem + a, 2);     // descriptors hold CTA-local addresses
    return v;
  }
};
struct IDesc {
  uint32_t M, N; bool a_bf16, b_bf16, a_mn, b_mn;
  explicit IDesc(uint32_t d)
      : M(((d >> 24) & 0x1F) << 4), N(((d >> 17) & 0x3F) << 3), a_bf16(((d >> 7) & 7) == 1), b_bf16(((d >> 10) & 7) == 1),
        a_mn((d >> 15) & 1), b_mn((d >> 16) & 1) {}
};
inline void umma_store(uint32_t tmem_d, const IDesc& id, const std::vector<float>& acc, uint32_t accumulate) {
  const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFF;
  for (uint32_t m = 0; m < id.M; ++m)
    for (uint32_t n = 0; n < id.N; ++n) {
      uint32_t& slot = cuda_emu::tmem_at(lane0 + m, col0 + n);
      slot = __float_as_uint((accumulate ? __uint_as_float(slot) : 0.f) + acc[m * id.N + n]);
    }
}
// cta_group::2 (issued by the leader): M = 256 rows, CTA r holds rows [128 r, 128 r + 128) of A and columns
// [N/2 r, N/2 r + N/2) of B in its own shared memory at the descriptors' addresses, and receives its 128 rows of D
// (all N columns) in its own tensor memory.
inline void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  cuda_emu::chaos();
  const IDesc id(idesc);
  const SmemOperand A(desc_a), B(desc_b);
  const uint32_t half_n = id.N / 2, col0 = tmem_d & 0xFFFF;
  std::vector<float> b(16 * id.N);
  for (uint32_t n = 0; n < id.N; ++n)
    for (uint32_t k = 0; k < 16; ++k)
      b[k * id.N + n] = cuda_emu::to_float16bits(B.at(n % half_n, k, id.b_mn, cuda_emu::bm->cluster[n / half_n]), id.b_bf16);
  for (uint32_t r = 0; r < 2; ++r) {
    cuda_emu::BlockModel* cta = cuda_emu::bm->cluster[r];
    for (uint32_t m = 0; m < 128; ++m) {
      std::vector<float> acc(id.N, 0.f);
      for (uint32_t k = 0; k < 16; ++k) {
        const float a = cuda_emu::to_float16bits(A.at(m, k, id.a_mn, cta), id.a_bf16);
        for (uint32_t n = 0; n < id.N; ++n) acc[n] += a * b[k * id.N + n];
      }
      for (uint32_t n = 0; n < id.N; ++n) {
        uint32_t& slot = cta->tmem[m * 512 + ((col0 + n) & 511)];
        slot = __float_as_uint((accumulate ? __uint_as_float(slot) : 0.f) + acc[n]);
      }
    }
  }
}
template <int G = 1>
inline void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (G == 2) { umma_f16_ss_pair(tmem_d, desc_a, desc_b, idesc, accumulate); return; }
  cuda_emu::chaos();
  const IDesc id(idesc);
  const SmemOperand A(desc_a), B(desc_b);
  std::vector<float> acc(id.M * id.N, 0.f), b(16 * id.N);
  for (uint32_t n = 0; n < id.N; ++n)
    for (uint32_t k = 0; k < 16; ++k) b[k * id.N + n] = cuda_emu::to_float16bits(B.at(n, k, id.b_mn), id.b_bf16);
  for (uint32_t m = 0; m < id.M; ++m)
    for (uint32_t k = 0; k < 16; ++k) {
      const float a = cuda_emu::to_float16bits(A.at(m, k, id.a_mn), id.a_bf16);
      for (uint32_t n = 0; n < id.N; ++n) acc[m * id.N + n] += a * b[k * id.N + n];
    }
  umma_store(tmem_d, id, acc, accumulate);
}
// A from tensor memory: lane = row, 16 k-elements as 8 columns of 16-bit pairs (even element in the low half)
inline void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  cuda_emu::chaos();
  const IDesc id(idesc);
  const SmemOperand B(desc_b);
  std::vector<float> acc(id.M * id.N, 0.f), b(16 * id.N);
  for (uint32_t n = 0; n < id.N; ++n)
    for (uint32_t k = 0; k < 16; ++k) b[k * id.N + n] = cuda_emu::to_float16bits(B.at(n, k, id.b_mn), id.b_bf16);
  const uint32_t alane = tmem_a >> 16, acol = tmem_a & 0xFFFF;
  for (uint32_t m = 0; m < id.M; ++m)
    for (uint32_t k = 0; k < 16; ++k) {
      const uint32_t word = cuda_emu::tmem_at(alane + m, acol + k / 2);
      const float a = cuda_emu::to_float16bits(uint16_t(k & 1 ? word >> 16 : word), id.a_bf16);
      for (uint32_t n = 0; n < id.N; ++n) acc[m * id.N + n] += a * b[k * id.N + n];
    }
  umma_store(tmem_d, id, acc, accumulate);
}
template <int G = 1> inline void umma_commit(uint64_t* bar) {     // the MMAs above already retired
  if constexpr (G == 2) { mbar_arrive(cluster_peer_ptr(bar, 0)); mbar_arrive(cluster_peer_ptr(bar, 1)); }   // multicast
  else mbar_arrive(bar);
}

// run-time 16-bit format helpers of the real header
inline uint32_t pack_16x2(int fp16, float a, float b) { return fp16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); }
inline float2 unpack_16x2(int fp16, uint32_t u) { return fp16 ? unpack_f16x2(u) : unpack_bf16x2(u); }
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

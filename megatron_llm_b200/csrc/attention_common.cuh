// Shared pieces of the sm_100a attention kernels (forward / dK,dV / dQ).
#pragma once
#include "attention_dropout.cuh"
#include "gemm_types.h"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <math.h>
#include <string.h>

namespace mlb {

constexpr int AT_M = 128;       // query rows per tile
constexpr int AT_N = 128;       // kv rows per tile
constexpr int AT_D = 128;       // head dim
constexpr int AT_THREADS = 256;
constexpr int AT_TILE_BYTES = AT_M * AT_D * 2;     // 32 KB (two 64-column swizzle boxes of 16 KB)
constexpr int AT_HALF_BYTES = AT_M * 64 * 2;       // 16 KB

static inline PFN_cuTensorMapEncodeTiled_v12000 at_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 4-D map over a [seq, batch, heads, hn]-indexable bf16 tensor: dims (hn, heads, seq, batch), box (64, 1, rows, 1)
static inline int make_tmap_heads(CUtensorMap* tm, const void* base, int hn, int heads, int seq, int batch,
                                  long long head_stride, long long seq_stride, long long batch_stride, int box_rows) {
  auto fn = at_encode_fn();
  if (!fn) return -1;
  mlb_bind_context();
  cuuint64_t dims[4] = {(cuuint64_t)hn, (cuuint64_t)heads, (cuuint64_t)seq, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)head_stride * 2, (cuuint64_t)seq_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// load a [128 rows x D hn] tile as D/64 swizzled boxes of 64 columns.  A tile slot in shared memory is always
// AT_TILE_BYTES (the head_dim-64 instantiations use the first box of a slot only), so the descriptor helpers below and
// every smem offset are the same for both head dims.
template <int D = AT_D>
__device__ __forceinline__ void load_tile(uint8_t* smem_dst, const CUtensorMap* tm, uint64_t* bar, int head, int row0,
                                          int b) {
  static_assert(D == 64 || D == 128, "head_dim 64 or 128");
  tma_load_4d(smem_dst, tm, bar, 0, head, row0, b);
  if constexpr (D == 128) tma_load_4d(smem_dst + AT_HALF_BYTES, tm, bar, 64, head, row0, b);
}
// bytes one load_tile<D> delivers (the mbarrier's expect_tx)
template <int D>
constexpr int at_tile_tx() { return AT_M * D * 2; }

// The same smem tile ([128 rows][2 boxes of 64 hn]) can feed an MMA two ways:
//  * K-major  : rows = the MMA's M or N index, hn = the reduction dim; k-step `k` covers 16 hn elements
//  * MN-major : rows = the reduction dim, hn = the MMA's N index (D: one or two 64-wide atoms, LBO = the box stride);
//               k-step `k` covers 16 rows
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int k) {
  const uint32_t a = tile_addr + (k >> 2) * AT_HALF_BYTES + (k & 3) * 32;
  return make_smem_desc(a, 16, 1024, kSwizzle128B);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int k) {
  return make_smem_desc(tile_addr + k * 2048, AT_HALF_BYTES, 1024, kSwizzle128B);
}

// Kernel variants (template flags): 16-bit element type of Q / K / V / O and the gradients, and attention dropout.
// Flag 0 (bf16, no dropout) is the training hot path; the other instantiations add code only under `if constexpr`.
constexpr int AF_FP16 = 1, AF_DROPOUT = 2;
template <int F>
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  if constexpr ((F & AF_FP16) != 0) return pack_f16x2(a, b);
  else return pack_bf16x2(a, b);
}
template <int F>
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  if constexpr ((F & AF_FP16) != 0) return unpack_f16x2(u);
  else return unpack_bf16x2(u);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// head -> coordinate in a tensor map's "heads" dimension (handles separate tensors and the packed QKV buffer)
struct HeadMap {
  int q_group_stride, q_off, k_group_stride, k_off, v_group_stride, v_off;
  int q_per_kv;
  __device__ __forceinline__ int q(int h) const { return (h / q_per_kv) * q_group_stride + (h % q_per_kv) + q_off; }
  __device__ __forceinline__ int k(int kvh) const { return kvh * k_group_stride + k_off; }
  __device__ __forceinline__ int v(int kvh) const { return kvh * v_group_stride + v_off; }
};

// strided row pointer of a [seq, batch, heads, 128] tensor: (head, seq, batch) strides in elements
struct RowAddr {
  void* base;
  long long head_stride, seq_stride, batch_stride;
  __device__ __forceinline__ __nv_bfloat16* row(int s, int b, int head) const {
    return reinterpret_cast<__nv_bfloat16*>(base) + (long long)s * seq_stride + (long long)b * batch_stride +
           (long long)head * head_stride;
  }
};

}  // namespace mlb

// NVLink transfer micro-benchmarks (tools/profiling/nvlink_bench.py): how fast can a few CTAs move data between peers, and
// with which instruction path?  These decide the design of the fused GEMM+collective kernels' communication side:
//   mode 0  ld/st      : every thread LDG.128 x4 -> STG.128 x4 (pull when src is the peer, push when dst is)
//   mode 1  bulk       : one thread per CTA, cp.async.bulk global->smem->global pipeline (piece_bytes, stages)
//   mode 2  bulk rows  : as mode 1, but every smem piece leaves as separate bulk stores of `row_bytes` to rows that are
//                        `dst_stride` apart (the access pattern of a 2-D TMA store box into a row-major peer tensor)
//   mode 3  multimem   : LDG.128 -> multimem.st.v4 to a multicast address (one store reaches every peer, NVLS)
// Not used by the training path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mlb {

__device__ __forceinline__ void pb_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pb_bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(256) p2p_ldst_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                       long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) v[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) dst[i] = v[u];
    }
  }
}

__global__ void __launch_bounds__(256) p2p_multimem_kernel(const uint4* __restrict__ src, uint4* mc_dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) v[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i),
                     "f"(__uint_as_float(v[u].x)), "f"(__uint_as_float(v[u].y)), "f"(__uint_as_float(v[u].z)),
                     "f"(__uint_as_float(v[u].w))
                     : "memory");
    }
  }
}

// one thread per CTA; CTA b moves the pieces b, b + grid, ... of the buffer
__global__ void __launch_bounds__(32) p2p_bulk_kernel(const uint8_t* src, uint8_t* dst, long long bytes, int piece_bytes,
                                                      int stages, int row_bytes, long long dst_stride) {
  extern __shared__ __align__(128) uint8_t pb_smem[];
  if (threadIdx.x != 0) return;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pb_smem + (size_t)stages * piece_bytes);
  for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
  fence_barrier_init();
  fence_proxy_async_smem();
  const long long pieces = (bytes + piece_bytes - 1) / piece_bytes;
  const long long n_mine = pieces > blockIdx.x ? (pieces - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  uint32_t phase_bits = 0;
  long long issued = 0, stored = 0;
  auto issue = [&](long long k) {
    const long long off = (blockIdx.x + k * gridDim.x) * (long long)piece_bytes;
    const uint32_t nb = (uint32_t)min((long long)piece_bytes, bytes - off);
    const int stage = (int)(k % stages);
    mbar_arrive_expect_tx(&bars[stage], nb);
    pb_bulk_g2s(pb_smem + (size_t)stage * piece_bytes, src + off, nb, &bars[stage]);
  };
  while (issued < n_mine && issued < stages - 1) issue(issued++);
  while (stored < n_mine) {
    const int stage = (int)(stored % stages);
    mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
    phase_bits ^= (1u << stage);
    const long long off = (blockIdx.x + stored * gridDim.x) * (long long)piece_bytes;
    const uint32_t nb = (uint32_t)min((long long)piece_bytes, bytes - off);
    const uint8_t* s = pb_smem + (size_t)stage * piece_bytes;
    if (row_bytes <= 0) {
      pb_bulk_s2g(dst + off, s, nb);
    } else {
      // the piece holds nb / row_bytes rows of a row-major tensor whose rows are dst_stride bytes apart
      const long long first_row = off / row_bytes;
      for (uint32_t r = 0; r * (uint32_t)row_bytes < nb; ++r)
        pb_bulk_s2g(dst + (first_row + r) * dst_stride, s + (size_t)r * row_bytes, (uint32_t)row_bytes);
    }
    tma_store_commit();
    ++stored;
    if (issued < n_mine) {
      tma_store_wait_read<1>();
      issue(issued++);
    }
  }
  tma_store_wait<0>();
}

}  // namespace mlb

extern "C" int mlb_p2p_bench(int mode, const void* src, void* dst, long long bytes, int ctas, int piece_bytes,
                             int stages, int row_bytes, long long dst_stride, cudaStream_t stream) {
  using namespace mlb;
  if (mode == 0) {
    p2p_ldst_kernel<<<ctas, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                              bytes / 16);
  } else if (mode == 3) {
    p2p_multimem_kernel<<<ctas, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                  bytes / 16);
  } else {
    const size_t smem = (size_t)stages * piece_bytes + 8 * stages + 128;
    cudaError_t e = cudaFuncSetAttribute(p2p_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    p2p_bulk_kernel<<<ctas, 32, smem, stream>>>(reinterpret_cast<const uint8_t*>(src), reinterpret_cast<uint8_t*>(dst),
                                                bytes, piece_bytes, stages, mode == 2 ? row_bytes : 0, dst_stride);
  }
  return (int)cudaGetLastError();
}

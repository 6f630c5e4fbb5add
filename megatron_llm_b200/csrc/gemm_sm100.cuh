// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring -> tcgen05.mma (single issuing thread)
//   -> fp32 accumulators in TMEM (double buffered) -> tcgen05.ld epilogue.
//
//   D[M,N] = A[M,K] * B[N,K]^T      (A and B each either K-major or MN-major in global memory)
//
// It replaces every cuBLAS call on the tensor-parallel hot path of the reference
// (megatron/core/tensor_parallel/layers.py:240 fwd, :267 dgrad, :298-307 wgrad) and the orphaned
// fused_weight_gradient_dense.cu (cublasGemmEx beta=1 into fp32 main_grad).
//
// The same kernel is the compute half of the fused GEMM+collective kernels (comm.cuh hooks):
//   * producer side: wait on per-chunk arrival flags before TMA-loading A rows (all-gather -> GEMM),
//     with the m-block order rotated so a rank starts on its local shard;
//   * epilogue side: scatter output row-chunks straight into peer (NVLink-mapped) buffers and
//     publish per-destination tile counters with release.sys (GEMM -> reduce-scatter).
#pragma once
#include "ptx.cuh"

namespace mlb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = one 128B swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_MAX_PEERS = 8;

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // C(bf16) = acc
  EPI_F32_ACCUM = 1,   // C(fp32) += acc      (wgrad into main_grad)
  EPI_F32 = 2,         // C(fp32) = acc
  EPI_BF16_ACCUM = 3,  // C(bf16) += acc
};

struct GemmComm {
  // all-gather -> GEMM: A rows [c*a_chunk_rows, (c+1)*a_chunk_rows) are valid once
  // a_ready_flags[c] >= a_ready_epoch (written with release.sys by the producer of that chunk).
  const int* a_ready_flags;
  int a_chunk_rows;
  int a_ready_epoch;
  int m_rotate_blocks;  // first m-block processed (local shard first)
  // GEMM -> scatter: output rows of chunk c go to out_ptrs[c] (row index relative to the chunk).
  void* out_ptrs[GEMM_MAX_PEERS];
  int out_chunk_rows;  // 0 = disabled
  // after each finished output tile: red.release.sys.add(tile_counters[c], 1) on the destination
  int* tile_counters[GEMM_MAX_PEERS];
};

struct GemmParams {
  void* C;
  int M, N, K;
  int ldc;  // elements
  GemmComm comm;
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (power of two: 256 or 512)
  constexpr uint32_t IDESC = make_idesc_f16(GEMM_BLOCK_M, BLOCK_N, A_MN, B_MN, true);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_m = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int total_tiles = num_m * num_n;
  constexpr int GROUP_M = 8;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = GROUP_M * num_n;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(GROUP_M, num_m - first_m);
    const int in_group = tile - group * per_group;
    m_blk = first_m + in_group % gsize;
    n_blk = in_group / gsize;
    m_blk += p.comm.m_rotate_blocks;
    if (m_blk >= num_m) m_blk -= num_m;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * GEMM_BLOCK_M, n0 = n_blk * BLOCK_N;
        if (p.comm.a_ready_flags != nullptr) {
          const int c_lo = m0 / p.comm.a_chunk_rows;
          const int c_hi = (min(m0 + GEMM_BLOCK_M, p.M) - 1) / p.comm.a_chunk_rows;
          for (int c = c_lo; c <= c_hi; ++c) {
            while (ld_acquire_sys(p.comm.a_ready_flags + c) < p.comm.a_ready_epoch) __nanosleep(64);
          }
          fence_proxy_async_global();  // generic-proxy acquire -> async-proxy (TMA) reads
        }
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int k0 = kb * GEMM_BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < GEMM_BLOCK_M / 64; ++i)
              tma_load_2d(sA + i * (GEMM_BLOCK_K * 128), &tmA, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i)
              tma_load_2d(sB + i * (GEMM_BLOCK_K * 128), &tmB, &full_bar[stage], n0 + i * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sB = sA + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // K-major SW128: 8-row groups 1024B apart, advance 32B per UMMA_K inside the swizzle row.
            // MN-major SW128: 64-element MN chunks BLOCK_K*128B apart (LBO), 8-row K groups 1024B apart
            // (SBO), advance 16 K-rows = 2048B per UMMA_K.
            const uint64_t da = A_MN ? make_smem_desc(sA + k * 2048, GEMM_BLOCK_K * 128, 1024, kSwizzle128B)
                                     : make_smem_desc(sA + k * 32, 0, 1024, kSwizzle128B);
            const uint64_t db = B_MN ? make_smem_desc(sB + k * 2048, GEMM_BLOCK_K * 128, 1024, kSwizzle128B)
                                     : make_smem_desc(sB + k * 32, 0, 1024, kSwizzle128B);
            umma_f16_ss<1>(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<1>(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (4 warps, one TMEM lane quadrant each) ============
    const int q = warp - 4;  // == warp % 4 : the TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int m0 = m_blk * GEMM_BLOCK_M, n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      // destination row pointer (optionally scattered per m-chunk into peer buffers)
      uint8_t* crow;
      int dst_chunk = 0;
      if (p.comm.out_chunk_rows > 0) {
        const int rr = row_ok ? row : m0;
        dst_chunk = rr / p.comm.out_chunk_rows;
        const int lr = rr - dst_chunk * p.comm.out_chunk_rows;
        crow = reinterpret_cast<uint8_t*>(p.comm.out_ptrs[dst_chunk]) +
               (size_t)lr * p.ldc * ((EPI == EPI_BF16 || EPI == EPI_BF16_ACCUM) ? 2 : 4);
      } else {
        crow = reinterpret_cast<uint8_t*>(p.C) +
               (size_t)(row_ok ? row : 0) * p.ldc * ((EPI == EPI_BF16 || EPI == EPI_BF16_ACCUM) ? 2 : 4);
      }
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col = n0 + c * 32;
        if (row_ok && col < p.N) {
          if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_ACCUM) {
            uint4* dst = reinterpret_cast<uint4*>(crow + (size_t)col * 2);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (col + v * 8 < p.N) {
                uint4 o;
                if constexpr (EPI == EPI_BF16_ACCUM) {
                  const uint4 old = dst[v];
                  float2 a0 = unpack_bf16x2(old.x), a1 = unpack_bf16x2(old.y), a2 = unpack_bf16x2(old.z),
                         a3 = unpack_bf16x2(old.w);
                  o.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]) + a0.x, __uint_as_float(r[v * 8 + 1]) + a0.y);
                  o.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]) + a1.x, __uint_as_float(r[v * 8 + 3]) + a1.y);
                  o.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]) + a2.x, __uint_as_float(r[v * 8 + 5]) + a2.y);
                  o.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]) + a3.x, __uint_as_float(r[v * 8 + 7]) + a3.y);
                } else {
                  o.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
                  o.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
                  o.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
                  o.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
                }
                dst[v] = o;
              }
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(crow + (size_t)col * 4);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              if (col + v * 4 < p.N) {
                float4 o = make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                                       __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
                if constexpr (EPI == EPI_F32_ACCUM) {
                  const float4 old = dst[v];
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                dst[v] = o;
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (p.comm.out_chunk_rows > 0 && p.comm.tile_counters[0] != nullptr) {
        // all 4 epilogue warps' stores of this tile must be visible before the tile is published
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q == 0 && lane == 0) {
          const int c0 = m0 / p.comm.out_chunk_rows;
          __threadfence_system();
          red_add_release_sys(p.comm.tile_counters[c0], 1);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

}  // namespace mlb

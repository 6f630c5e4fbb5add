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
// MODE selects the fused collective (one kernel = math + NVLink transfer, tile by tile):
//   MODE_PLAIN   : GEMM only.
//   MODE_AG_GEMM : all-gather -> GEMM (ColumnParallelLinear fwd under sequence parallelism, RowParallel dgrad).
//                  The last `num_comm_ctas` CTAs of the grid are "puller" CTAs: they stream every peer's
//                  activation shard out of NVLink-mapped symmetric memory with cp.async.bulk (peer global ->
//                  smem -> local gathered buffer) and publish a flag per 128-row chunk; the GEMM CTAs' TMA
//                  producer warp acquires the flag of the chunk a tile needs, and walks the m-blocks starting at
//                  the local shard, so math on early chunks overlaps the arrival of later ones.
//   MODE_GEMM_RS : GEMM -> reduce-scatter (RowParallelLinear fwd, ColumnParallel dgrad).  The epilogue stores
//                  each output tile straight into the destination rank's receive slot over NVLink (remote chunks
//                  first, local chunk last) and bumps that rank's arrival counter with red.release.sys; when a
//                  CTA runs out of tiles it joins the reduction of the local chunk (sum of the `world` slots in
//                  fp32) as soon as all sources have delivered.
#pragma once
#include "gemm_types.h"
#include "ptx.cuh"

namespace mlb {

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

// bounded spin: a lost peer must not hang the GPU box; after ~2^26 polls mark the pad and carry on
__device__ __forceinline__ void spin_until_ge(const int* flag, int value, int* pad_local) {
  long long polls = 0;
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(40);
    if (++polls > (1LL << 26)) {
      if (pad_local) st_release_sys(pad_local + PAD_ERROR, 1);
      break;
    }
  }
}

// spin_until_ge that also returns how long it blocked (0 if the flag was already there)
__device__ __forceinline__ unsigned long long spin_until_ge_timed(const int* flag, int value, int* pad_local) {
  if (ld_acquire_sys(flag) >= value) return 0;
  const unsigned long long t0 = globaltimer_ns();
  spin_until_ge(flag, value, pad_local);
  return globaltimer_ns() - t0;
}

// ------------------------------------------------------------------------------------------------
// puller CTA: all-gather peer shards into the local gathered buffer with bulk async copies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

constexpr int AG_PIECE_BYTES = 32 * 1024;
constexpr int AG_STAGES = 6;  // 192 KB of smem in flight per puller CTA

// live epoch / arrival target of this call = launch argument + device-resident offset (see GemmComm::state)
enum CommStateSlot : int { STATE_AG_EPOCH = 0, STATE_RS_EPOCH = 1, STATE_RS_TOTAL = 2 };
__device__ __forceinline__ int comm_epoch(const GemmComm& c, int slot) {
  return c.epoch + (c.state ? *reinterpret_cast<const volatile int*>(c.state + slot) : 0);
}
__device__ __forceinline__ int comm_rs_expected(const GemmComm& c) {
  return c.rs_expected_total + (c.state ? *reinterpret_cast<const volatile int*>(c.state + STATE_RS_TOTAL) : 0);
}

// ------------------------------------------------------------------------------------------------
// NVLS pusher CTA (all 256 threads): local 16-byte loads -> multimem.st into every rank's gather buffer; one flag per
// 128-row chunk, released at every destination (unicast) after a system fence.  The buffer of this parity was last
// read two calls ago: wait until every rank's GEMM of that call retired (PAD_AG_ACK).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}

// called by one thread of every CTA of the launch when it no longer reads the gather buffer
static __device__ void ag_nvls_finish(const GemmComm& c) {
  __threadfence();
  if (atomicAdd(c.ag_done_counter, 1) + 1 != (int)gridDim.x) return;
  *c.ag_done_counter = 0;
  __threadfence_system();
  const int epoch = comm_epoch(c, STATE_AG_EPOCH);
  for (int d = 0; d < c.world; ++d)
    if (d != c.rank) st_release_sys(c.pad_peer[d] + PAD_AG_ACK + c.rank, epoch);
}

static __device__ void ag_pusher_nvls(const GemmComm& c, int comm_id) {
  const int epoch = comm_epoch(c, STATE_AG_EPOCH);
  if (threadIdx.x == 0) {
    for (int p = 0; p < c.world; ++p)
      if (p != c.rank) spin_until_ge(c.pad_local + PAD_AG_ACK + p, epoch - 2, c.pad_local);
  }
  __syncthreads();
  const int cpr = c.ag_rows_per_rank / GEMM_BLOCK_M;
  const long long chunk_vec = (long long)GEMM_BLOCK_M * c.ag_row_bytes / 16;
  constexpr int U = 8;
  for (int chunk = comm_id; chunk < cpr; chunk += c.num_comm_ctas) {
    const uint4* src = reinterpret_cast<const uint4*>(c.ag_local_src) + (long long)chunk * chunk_vec;
    uint4* dst = reinterpret_cast<uint4*>(c.ag_mc_dst) + ((long long)c.rank * cpr + chunk) * chunk_vec;
    for (long long i0 = threadIdx.x; i0 < chunk_vec; i0 += (long long)U * blockDim.x) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        if (i < chunk_vec) v[u] = __ldg(src + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        if (i < chunk_vec) multimem_st_v4(dst + i, v[u]);
      }
    }
    __syncthreads();                       // every thread's stores of this chunk are issued ...
    if (threadIdx.x == 0) {
      __threadfence_system();              // ... and ordered before the flags at system scope
      for (int k = 0; k < c.world; ++k) {
        const int d = (c.rank + 1 + k) % c.world;                  // own flag last
        st_release_sys(c.ag_flag_peer[d] + c.rank * cpr + chunk, epoch);
      }
    }
  }
  if (threadIdx.x == 0) ag_nvls_finish(c);
}

static __device__ void ag_puller(const GemmComm& c, uint8_t* smem, int comm_id) {
  if (c.ag_nvls) {
    ag_pusher_nvls(c, comm_id);
    return;
  }
  // one thread drives the whole copy pipeline (bulk copies are issued by a single thread anyway)
  if (threadIdx.x != 0) return;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AG_STAGES * AG_PIECE_BYTES);
  for (int i = 0; i < AG_STAGES; ++i) mbar_init(&bars[i], 1);
  fence_barrier_init();
  fence_proxy_async_smem();

  const int chunks_per_rank = c.ag_rows_per_rank / GEMM_BLOCK_M;
  const long long chunk_bytes = (long long)GEMM_BLOCK_M * c.ag_row_bytes;
  const int pieces_per_chunk = (int)((chunk_bytes + AG_PIECE_BYTES - 1) / AG_PIECE_BYTES);
  uint32_t phase_bits = 0;  // per-stage mbarrier parity
  long long piece_seq = 0;  // pieces issued so far by this CTA (stage = seq % AG_STAGES)

  if (comm_id == 0) {
    // my shard was written by earlier stream-ordered work: publish it to every peer
    __threadfence_system();
    for (int p = 0; p < c.world; ++p)
      if (p != c.rank) st_release_sys(c.pad_peer[p] + PAD_AG_READY + c.rank, comm_epoch(c, STATE_AG_EPOCH));
  }

  // (the own shard is placed into ag_dst by the host-side copy that also publishes it)
  for (int i = 1; i < c.world; ++i) {
    const int p = (c.rank + i) % c.world;
    bool waited = false;
    for (int j = 0; j < chunks_per_rank; ++j) {
      const int g = (i - 1) * chunks_per_rank + j;
      if (g % c.num_comm_ctas != comm_id) continue;
      if (!waited && p != c.rank) {
        spin_until_ge(c.pad_local + PAD_AG_READY + p, comm_epoch(c, STATE_AG_EPOCH), c.pad_local);
        fence_proxy_async_global();
        waited = true;
      }
      const uint8_t* src = reinterpret_cast<const uint8_t*>(c.ag_src[p]) + (long long)j * chunk_bytes;
      uint8_t* dst = reinterpret_cast<uint8_t*>(c.ag_dst) + ((long long)p * chunks_per_rank + j) * chunk_bytes;
      // software pipeline: keep up to AG_STAGES-1 loads in flight
      int issued = 0, stored = 0;
      auto issue = [&](int k) {
        const int stage = (int)((piece_seq + k) % AG_STAGES);
        const long long off = (long long)k * AG_PIECE_BYTES;
        const uint32_t bytes = (uint32_t)min((long long)AG_PIECE_BYTES, chunk_bytes - off);
        mbar_arrive_expect_tx(&bars[stage], bytes);
        bulk_g2s(smem + stage * AG_PIECE_BYTES, src + off, bytes, &bars[stage]);
      };
      while (issued < pieces_per_chunk && issued < AG_STAGES - 1) issue(issued++);
      while (stored < pieces_per_chunk) {
        const int stage = (int)((piece_seq + stored) % AG_STAGES);
        mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
        phase_bits ^= (1u << stage);
        const long long off = (long long)stored * AG_PIECE_BYTES;
        const uint32_t bytes = (uint32_t)min((long long)AG_PIECE_BYTES, chunk_bytes - off);
        bulk_s2g(dst + off, smem + stage * AG_PIECE_BYTES, bytes);
        tma_store_commit();
        ++stored;
        if (issued < pieces_per_chunk) {
          // the stage about to be refilled is the one whose store was committed one iteration ago
          tma_store_wait_read<1>();
          issue(issued++);
        }
      }
      piece_seq += pieces_per_chunk;
      tma_store_wait<0>();          // all bytes of the chunk are in local HBM
      fence_proxy_async_global();
      __threadfence();
      st_release_sys(c.ag_chunk_flags + p * chunks_per_rank + j, comm_epoch(c, STATE_AG_EPOCH));
    }
    if (p != c.rank) {
      // tell the owner when ALL of this rank's pullers are done with its shard
      __threadfence();
      const int done = atomicAdd(c.ag_read_counters + p, 1) + 1;
      if (done == c.num_comm_ctas) {
        c.ag_read_counters[p] = 0;
        __threadfence_system();
        st_release_sys(c.pad_peer[p] + PAD_AG_ACK + c.rank, comm_epoch(c, STATE_AG_EPOCH));
      }
    }
  }
  if (comm_id == 0) {
    // my published shard may be overwritten by the next call only once every peer has read it
    for (int p = 0; p < c.world; ++p)
      if (p != c.rank) spin_until_ge(c.pad_local + PAD_AG_ACK + p, comm_epoch(c, STATE_AG_EPOCH), c.pad_local);
  }
}

// ------------------------------------------------------------------------------------------------
// reduce-scatter tail (all CTAs of the grid, after the GEMM part): wait until every source delivered its tiles, sum the
// world receive slots in fp32 into the output rows, and hand the slot back to the senders
// ------------------------------------------------------------------------------------------------
static __device__ void rs_reduce_phase(const GemmParams& p) {
  const GemmComm& c = p.comm;
  __shared__ int s_last;
  const unsigned long long t_tail = (c.stats && threadIdx.x == 0) ? globaltimer_ns() : 0;
  if (threadIdx.x == 0) {
    const int expected = comm_rs_expected(c);
    for (int s = 0; s < c.world; ++s) spin_until_ge(c.pad_local + PAD_RS_ARRIVED + s, expected, c.pad_local);
  }
  __syncthreads();
  // rows are contiguous (ldc == N): treat the slot as a flat array of 16-byte vectors.  Four vectors x two slots are
  // loaded before the first add, so every thread keeps 8 independent 16-byte loads in flight (HBM latency bound
  // otherwise: this phase is pure streaming of world+1 x [rows, N] bf16)
  const long long total_vec = (long long)c.rs_rows_per_rank * p.N / 8;
  const long long slot_vec = (long long)c.rs_rows_per_rank * p.ldc / 8;
  const uint4* slots = reinterpret_cast<const uint4*>(c.rs_slots);
  uint4* out = reinterpret_cast<uint4*>(c.rs_out);
  const bool all_reduce = c.ar_dst[0] != nullptr;          // reduced slice goes to every rank's [M, N] output buffer
  const long long ar_off = (long long)c.rank * slot_vec;   // my rows inside that buffer
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total_vec; i0 += U * stride) {
    float acc[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
    for (int s = 0; s < c.world; s += 2) {
      uint4 v[2][U];
      const bool two = s + 1 < c.world;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        v[0][u] = i < total_vec ? __ldcg(slots + s * slot_vec + i) : make_uint4(0, 0, 0, 0);
        v[1][u] = (two && i < total_vec) ? __ldcg(slots + (s + 1) * slot_vec + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float2 f0 = unpack_bf16x2(v[h][u].x), f1 = unpack_bf16x2(v[h][u].y), f2 = unpack_bf16x2(v[h][u].z),
                       f3 = unpack_bf16x2(v[h][u].w);
          acc[u][0] += f0.x; acc[u][1] += f0.y; acc[u][2] += f1.x; acc[u][3] += f1.y;
          acc[u][4] += f2.x; acc[u][5] += f2.y; acc[u][6] += f3.x; acc[u][7] += f3.y;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total_vec) {
        uint4 o;
        o.x = pack_bf16x2(acc[u][0], acc[u][1]); o.y = pack_bf16x2(acc[u][2], acc[u][3]);
        o.z = pack_bf16x2(acc[u][4], acc[u][5]); o.w = pack_bf16x2(acc[u][6], acc[u][7]);
        if (!all_reduce) {
          out[i] = o;
        } else {
          for (int k = 0; k < c.world; ++k) {                // own copy first, then the peers (posted NVLink stores)
            const int d = (c.rank + k) % c.world;
            st_v4(reinterpret_cast<uint4*>(c.ar_dst[d]) + ar_off + i, o);
          }
        }
      }
    }
  }
  // last CTA out tells every peer that this rank's receive slot (this parity) is free again
  __syncthreads();
  if (threadIdx.x == 0) {
    if (all_reduce) __threadfence_system(); else __threadfence();     // (peer stores of this CTA precede the count)
    s_last = (atomicAdd(c.rs_reduce_counter, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    *c.rs_reduce_counter = 0;
    __threadfence_system();
    const int epoch = comm_epoch(c, STATE_RS_EPOCH);
    for (int d = 0; d < c.world; ++d)
      if (d != c.rank) st_release_sys(c.pad_peer[d] + PAD_RS_FREE + c.rank, epoch);
    if (all_reduce) {
      // all-gather handshake: my slice is in everybody's buffer; the launch retires once theirs are in mine
      for (int d = 0; d < c.world; ++d)
        if (d != c.rank) st_release_sys(c.pad_peer[d] + PAD_AR_DONE + c.rank, epoch);
      for (int s = 0; s < c.world; ++s)
        if (s != c.rank) spin_until_ge(c.pad_local + PAD_AR_DONE + s, epoch, c.pad_local);
    }
  }
  if (c.stats && threadIdx.x == 0) atomicAdd(c.stats + 1, (globaltimer_ns() - t_tail) / gridDim.x);
}

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (power of two: 256 or 512)
  const uint32_t IDESC = make_idesc_f16(GEMM_BLOCK_M, BLOCK_N, A_MN, B_MN, !p.fp16);
  constexpr int OUT_ELEM = (EPI == EPI_BF16 || EPI == EPI_BF16_ACCUM) ? 2 : 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  int num_compute_ctas = gridDim.x;
  if constexpr (MODE == MODE_AG_GEMM) {
    num_compute_ctas = gridDim.x - p.comm.num_comm_ctas;
    if ((int)blockIdx.x >= num_compute_ctas) {
      ag_puller(p.comm, smem, blockIdx.x - num_compute_ctas);
      return;
    }
  }

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
  // plain GEMM: groups of 8 m-blocks share B tiles through L2.  fused modes: one m-block row at a time in
  // rotated order so work follows the arrival (AG) / departure (RS) order of the chunks.
  constexpr int GROUP_M = (MODE == MODE_PLAIN) ? 8 : 1;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = GROUP_M * num_n;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(GROUP_M, num_m - first_m);
    const int in_group = tile - group * per_group;
    m_blk = first_m + in_group % gsize;
    n_blk = in_group / gsize;
    if constexpr (MODE != MODE_PLAIN) {
      m_blk += p.comm.m_rotate_blocks;
      if (m_blk >= num_m) m_blk -= num_m;
    }
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long ag_wait_ns = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += num_compute_ctas) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * GEMM_BLOCK_M, n0 = n_blk * BLOCK_N;
        if constexpr (MODE == MODE_AG_GEMM) {
          if (m0 / p.comm.ag_rows_per_rank != p.comm.rank)     // (the own shard was placed before the launch)
            ag_wait_ns += spin_until_ge_timed(p.comm.ag_chunk_flags + m_blk, comm_epoch(p.comm, STATE_AG_EPOCH),
                                              p.comm.pad_local);
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
      if constexpr (MODE == MODE_AG_GEMM) {
        if (p.comm.stats) atomicAdd(p.comm.stats + 0, ag_wait_ns / (unsigned long long)num_compute_ctas);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += num_compute_ctas) {
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
                                     : make_smem_desc(sA + k * 32, 16, 1024, kSwizzle128B);
            const uint64_t db = B_MN ? make_smem_desc(sB + k * 2048, GEMM_BLOCK_K * 128, 1024, kSwizzle128B)
                                     : make_smem_desc(sB + k * 32, 16, 1024, kSwizzle128B);
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
    uint32_t free_checked = 0;  // RS: destinations whose receive slot we already know to be reusable
    for (int tile = blockIdx.x; tile < total_tiles; tile += num_compute_ctas) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int m0 = m_blk * GEMM_BLOCK_M, n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      uint8_t* crow;
      int dst = 0;
      if constexpr (MODE == MODE_GEMM_RS) {
        dst = m0 / p.comm.rs_rows_per_rank;
        if (!((free_checked >> dst) & 1u)) {
          // the receive slot of this parity on `dst` was last used two calls ago: wait until dst reduced it
          if (dst != p.comm.rank) spin_until_ge(p.comm.pad_local + PAD_RS_FREE + dst, comm_epoch(p.comm, STATE_RS_EPOCH) - 2, p.comm.pad_local);
          free_checked |= (1u << dst);
        }
        const int lr = (row_ok ? row : m0) - dst * p.comm.rs_rows_per_rank;
        crow = reinterpret_cast<uint8_t*>(p.comm.rs_dst[dst]) + (size_t)lr * p.ldc * OUT_ELEM;
      } else {
        crow = reinterpret_cast<uint8_t*>(p.C) + (size_t)(row_ok ? row : 0) * p.ldc * OUT_ELEM;
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
            uint4* dptr = reinterpret_cast<uint4*>(crow + (size_t)col * 2);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (col + v * 8 < p.N) {
                uint4 o;
                if constexpr (EPI == EPI_BF16_ACCUM) {
                  const uint4 old = dptr[v];
                  float2 a0 = unpack_16x2(p.fp16, old.x), a1 = unpack_16x2(p.fp16, old.y), a2 = unpack_16x2(p.fp16, old.z),
                         a3 = unpack_16x2(p.fp16, old.w);
                  o.x = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 0]) + a0.x, __uint_as_float(r[v * 8 + 1]) + a0.y);
                  o.y = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 2]) + a1.x, __uint_as_float(r[v * 8 + 3]) + a1.y);
                  o.z = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 4]) + a2.x, __uint_as_float(r[v * 8 + 5]) + a2.y);
                  o.w = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 6]) + a3.x, __uint_as_float(r[v * 8 + 7]) + a3.y);
                } else {
                  o.x = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
                  o.y = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
                  o.z = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
                  o.w = pack_16x2(p.fp16, __uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
                }
                dptr[v] = o;
              }
            }
          } else {
            float4* dptr = reinterpret_cast<float4*>(crow + (size_t)col * 4);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              if (col + v * 4 < p.N) {
                float4 o = make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                                       __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
                if constexpr (EPI == EPI_F32_ACCUM) {
                  const float4 old = dptr[v];
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                dptr[v] = o;
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if constexpr (MODE == MODE_GEMM_RS) {
        // every epilogue warp's stores of this tile must be visible at the destination before it is counted
        __threadfence_system();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q == 0 && lane == 0) {
          int* counter = (dst == p.comm.rank ? p.comm.pad_local : p.comm.pad_peer[dst]) + PAD_RS_ARRIVED + p.comm.rank;
          red_add_release_sys(counter, 1);
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

  if constexpr (MODE == MODE_GEMM_RS) rs_reduce_phase(p);
}

}  // namespace mlb

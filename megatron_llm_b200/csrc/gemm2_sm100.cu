// 2-CTA (cta_group::2) variant of the persistent bf16 GEMM: a pair of SMs in one cluster works on a 256x256 output
// tile.  Each CTA TMA-loads its 128 rows of A and its 128-row half of B (so B's shared-memory traffic per SM is
// halved and the smem ring is 6 stages deep instead of 4), the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (UMMA 256x256x16), accumulators live in both CTAs' TMEM (128 lanes x 256 columns each,
// double buffered), and each CTA's four epilogue warps drain their own half.
#include "gemm_sm100.cuh"   // GemmParams, ptx helpers, spin_until_ge, ag_puller (shared with the 1-CTA kernel)

#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

namespace mlb {

int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);
int make_tmap_2d_out(CUtensorMap* tm, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_outer);

constexpr int G2_BLOCK_M = 256;      // per pair
constexpr int G2_BLOCK_N = 256;
constexpr int G2_HALF = 128;         // rows of A / rows of B held by one CTA
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = G2_HALF * GEMM_BLOCK_K * 2;   // 16 KB
constexpr int G2_B_BYTES = G2_HALF * GEMM_BLOCK_K * 2;   // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;  // 32 KB per CTA per stage
// epilogue staging for the TMA stores: per warp 2 buffers of [32 rows][128 B] (128B-swizzled), 32 KB per CTA
constexpr int G2_EPI_OFFSET = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_EPI_BYTES = 4 * 2 * 4096;
constexpr int G2_BAR_OFFSET = G2_EPI_OFFSET + G2_EPI_BYTES;
constexpr int G2_SMEM_TOTAL = G2_BAR_OFFSET + 256 + 1024;

// Optional timeline counters (MLB200_GEMM2_DEBUG=1): per CTA 8 x u64 =
// [0] MMA cycles waiting on full_bar   [1] MMA cycles waiting on tmem_empty   [2] MMA total cycles
// [3] producer cycles waiting on empty_bar   [4] producer total cycles
// [5] epilogue(warp 4) cycles waiting on tmem_full   [6] epilogue total cycles   [7] tiles
__device__ unsigned long long g2_dbg[512 * 8];
__device__ __forceinline__ unsigned long long g2_clock() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}

// MODE_GEMM_RS: one bf16 output map per destination rank (this rank's receive slot [rows_per_rank, N] on that rank,
// NVLink-mapped): the epilogue's TMA stores go straight into peer memory in 32-row x 128-byte boxes
struct RsMaps {
  CUtensorMap m[GEMM_MAX_PEERS];
};

template <bool A_MN, bool B_MN, int EPI, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ RsMaps rsm,
                      const GemmParams p, const int dbg) {
  constexpr uint32_t TMEM_COLS = 2 * G2_BLOCK_N;  // two accumulator stages
  const uint32_t IDESC = make_idesc_f16(G2_BLOCK_M, G2_BLOCK_N, A_MN, B_MN, !p.fp16);
  constexpr int OUT_ELEM = (EPI == EPI_BF16 || EPI == EPI_BF16_ACCUM) ? 2 : 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // MODE_AG_GEMM: the trailing clusters of the grid are puller CTAs (all-gather of the peer shards into the local
  // gathered buffer with bulk async copies over NVLink, one flag per 128-row chunk); they never join the GEMM
  int num_pairs = gridDim.x >> 1;
  if constexpr (MODE == MODE_AG_GEMM) {
    num_pairs = ((int)gridDim.x - p.comm.num_comm_ctas) >> 1;
    if ((int)blockIdx.x >= 2 * num_pairs) {
      ag_puller(p.comm, smem, (int)blockIdx.x - 2 * num_pairs);
      return;
    }
  }
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full_bar = empty_bar + G2_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);    // the leader's arrive.expect_tx covers the bytes of BOTH CTAs' loads; the peer
                                     // never arrives (a remote mbarrier.arrive.release.cluster costs ~1.4k cycles
                                     // and serialised the peer's producer: measured with the timeline counters)
      mbar_init(&empty_bar[i], 1);   // tcgen05.commit multicast arrives on both CTAs' copies
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);  // 4 epilogue warps x 2 CTAs (leader's copy)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_m = (p.M + G2_BLOCK_M - 1) / G2_BLOCK_M;
  const int num_n = (p.N + G2_BLOCK_N - 1) / G2_BLOCK_N;
  const int num_k = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int total_tiles = num_m * num_n;
  const int pair = blockIdx.x >> 1;
  constexpr int GROUP_M = 8;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    if constexpr (MODE != MODE_PLAIN) {
      // fused modes: groups of (up to) 4 row blocks in the order their data becomes available -- all-gather: own shard
      // first; reduce-scatter: the next rank's rows first, own rows last -- n-major inside a group so that a B
      // panel is reused by the whole group while it is hot in L2
      int G = p.comm.m_group_blocks;
      if (G <= 0 || num_m % G != 0) G = (num_m % 4 == 0) ? 4 : ((num_m % 2 == 0) ? 2 : 1);
      const int per_group = G * num_n;
      int group = tile / per_group;
      const int in_group = tile - group * per_group;
      if (p.comm.m_stripe) {
        // NVLS all-gather: every source pushes its shard front to back at the same time, so the j-th tile group of ALL
        // sources arrives together: visit (j, source) with the own rank's group first in every round
        const int gpr = (num_m / G) / p.comm.world;          // tile groups per source rank
        const int j = group / p.comm.world, s_idx = group - j * p.comm.world;
        const int src = (p.comm.rank + s_idx) % p.comm.world;
        m_blk = (src * gpr + j) * G + in_group % G;
        n_blk = in_group / G;
        return;
      }
      if (p.comm.m_interleave) {          // rotated order is [remote groups | own groups]: take them alternately
        const int half = (num_m / G) >> 1;
        group = (group & 1) ? half + (group >> 1) : (group >> 1);
      }
      m_blk = group * G + in_group % G + (p.comm.m_rotate_blocks >> 1);
      if (m_blk >= num_m) m_blk -= num_m;
      n_blk = in_group / G;
      return;
    }
    const int per_group = GROUP_M * num_n;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(GROUP_M, num_m - first_m);
    const int in_group = tile - group * per_group;
    m_blk = first_m + in_group % gsize;
    n_blk = in_group / gsize;
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long w_empty = 0, t_begin = dbg ? g2_clock() : 0, ag_wait_ns = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * G2_BLOCK_M + cta_rank * G2_HALF;   // this CTA's rows of A
        const int n0 = n_blk * G2_BLOCK_N + cta_rank * G2_HALF;   // this CTA's half of B
        if constexpr (MODE == MODE_AG_GEMM) {
          if (p.comm.ag_nvls) {                                 // every chunk (the own rows too) arrives by multicast
            ag_wait_ns += spin_until_ge_timed(p.comm.ag_flag_peer[p.comm.rank] + (m0 >> 7),
                                              comm_epoch(p.comm, STATE_AG_EPOCH), p.comm.pad_local);
            fence_proxy_async_global();
          } else if (m0 / p.comm.ag_rows_per_rank != p.comm.rank) {   // (the own shard was placed before the launch)
            ag_wait_ns += spin_until_ge_timed(p.comm.ag_chunk_flags + (m0 >> 7), comm_epoch(p.comm, STATE_AG_EPOCH),
                                              p.comm.pad_local);
            fence_proxy_async_global();  // generic-proxy acquire -> async-proxy (TMA) reads
          }
        }
        for (int kb = 0; kb < num_k; ++kb) {
          const unsigned long long t0 = dbg ? g2_clock() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (dbg) w_empty += g2_clock() - t0;
          uint8_t* sA = smem + stage * G2_STAGE_BYTES;
          uint8_t* sB = sA + G2_A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          const int k0 = kb * GEMM_BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sA, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < G2_HALF / 64; ++i)
              tma_load_2d_2sm(sA + i * (GEMM_BLOCK_K * 128), &tmA, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(sB, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < G2_HALF / 64; ++i)
              tma_load_2d_2sm(sB + i * (GEMM_BLOCK_K * 128), &tmB, &full_bar[stage], n0 + i * 64, k0);
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) { g2_dbg[blockIdx.x * 8 + 3] = w_empty; g2_dbg[blockIdx.x * 8 + 4] = g2_clock() - t_begin; }
      if constexpr (MODE == MODE_AG_GEMM) {
        if (p.comm.stats) atomicAdd(p.comm.stats + 0, ag_wait_ns / (unsigned long long)(2 * num_pairs));
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      unsigned long long w_full = 0, w_tmem = 0, n_tiles = 0, t_begin = dbg ? g2_clock() : 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        unsigned long long t0 = dbg ? g2_clock() : 0;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        if (dbg) { w_tmem += g2_clock() - t0; ++n_tiles; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          t0 = dbg ? g2_clock() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (dbg) w_full += g2_clock() - t0;
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t sB = sA + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sA + k * 2048, GEMM_BLOCK_K * 128, 1024, kSwizzle128B)
                                     : make_smem_desc(sA + k * 32, 16, 1024, kSwizzle128B);
            const uint64_t db = B_MN ? make_smem_desc(sB + k * 2048, GEMM_BLOCK_K * 128, 1024, kSwizzle128B)
                                     : make_smem_desc(sB + k * 32, 16, 1024, kSwizzle128B);
            umma_f16_ss<2>(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<2>(&empty_bar[stage]);     // frees the stage in BOTH CTAs
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<2>(&tmem_full_bar[acc]);     // accumulators ready in BOTH CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg) {
        g2_dbg[blockIdx.x * 8 + 0] = w_full; g2_dbg[blockIdx.x * 8 + 1] = w_tmem;
        g2_dbg[blockIdx.x * 8 + 2] = g2_clock() - t_begin; g2_dbg[blockIdx.x * 8 + 7] = n_tiles;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs) ================================
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    unsigned long long w_tf = 0, t_begin = dbg ? g2_clock() : 0;
    uint32_t free_checked = 0;   // RS: destinations whose receive slot is known to be reusable
    int rs_prev_dst = -1;        // RS: destination of the previous tile (its arrival is signalled one tile late)
    // this CTA's half tile counts as delivered once the stores of all four epilogue warps have completed at `dst`
    auto rs_signal = [&](int dst, bool lagged) {
      if (lane == 0) {
        if (lagged) tma_store_wait<G2_BLOCK_N / 64>(); else tma_store_wait<0>();
        fence_proxy_async_global();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (q == 0 && lane == 0) {
        __threadfence_system();
        int* counter = (dst == p.comm.rank ? p.comm.pad_local : p.comm.pad_peer[dst]) + PAD_RS_ARRIVED + p.comm.rank;
        red_add_release_sys(counter, 1);
      }
    };
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int m0 = m_blk * G2_BLOCK_M + cta_rank * G2_HALF, n0 = n_blk * G2_BLOCK_N;
      const unsigned long long t0 = dbg ? g2_clock() : 0;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      if (dbg) w_tf += g2_clock() - t0;
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      uint8_t* crow = reinterpret_cast<uint8_t*>(p.C) + (size_t)(row_ok ? row : 0) * p.ldc * OUT_ELEM;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * G2_BLOCK_N;
      int rs_dst = 0;
      if constexpr (MODE == MODE_GEMM_RS) {
        rs_dst = (m_blk * G2_BLOCK_M) / p.comm.rs_rows_per_rank;
        if (!((free_checked >> rs_dst) & 1u)) {
          // the receive slot of this parity on `rs_dst` was last used two calls ago: wait until it was reduced there
          if (rs_dst != p.comm.rank)
            spin_until_ge(p.comm.pad_local + PAD_RS_FREE + rs_dst, comm_epoch(p.comm, STATE_RS_EPOCH) - 2, p.comm.pad_local);
          free_checked |= (1u << rs_dst);
        }
      }
      if constexpr (EPI != EPI_BF16_ACCUM) {
        // TMEM -> registers -> swizzled smem rows of 128 B -> one TMA store (or fp32 reduce-add) per 32-row x 128-B
        // box: full-line global writes, no read-modify-write traffic through the SM for the wgrad accumulation,
        // out-of-range rows / columns clipped by the tensor map
        uint8_t* ebase = smem + G2_EPI_OFFSET + q * 8192;
        constexpr int COLS = 128 / OUT_ELEM;                 // columns per box: 32 fp32 or 64 bf16
#pragma unroll 1
        for (int c = 0; c < G2_BLOCK_N / COLS; ++c) {
          uint32_t w[32];                                    // this row's 128 bytes
          if constexpr (OUT_ELEM == 4) {
            tmem_ld_32x32(taddr + c * 32, w);
            tmem_ld_wait();
          } else {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr + c * 64, r0);
            tmem_ld_32x32(taddr + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              w[i] = pack_16x2(p.fp16, __uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
              w[16 + i] = pack_16x2(p.fp16, __uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1]));
            }
          }
          uint8_t* buf = ebase + (c & 1) * 4096;
          if (lane == 0) tma_store_wait_read<1>();           // the store from two chunks ago has drained this buffer
          __syncwarp();
          uint8_t* rowp = buf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) =
                make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int c0 = n0 + c * COLS, c1 = m0 + q * 32;
            if (c0 < p.N && c1 < p.M) {
              if constexpr (MODE == MODE_GEMM_RS) tma_store_2d(&rsm.m[rs_dst], buf, c0, c1 - rs_dst * p.comm.rs_rows_per_rank);
              else if constexpr (EPI == EPI_F32_ACCUM) tma_reduce_add_2d(&tmC, buf, c0, c1);
              else tma_store_2d(&tmC, buf, c0, c1);
            }
            tma_store_commit();
          }
        }
      } else
#pragma unroll 1
      for (int c = 0; c < G2_BLOCK_N / 64; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr + c * 64, r0);       // two loads in flight before the wait
        tmem_ld_32x32(taddr + c * 64 + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t* r = h ? r1 : r0;
          const int col = n0 + c * 64 + h * 32;
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);   // the leader's MMA thread is the only waiter
      if constexpr (MODE == MODE_GEMM_RS) {
        // Arrival signalling runs one tile behind: the stores of the PREVIOUS tile have had a whole tile time to reach
        // their destination, so waiting for them (all but this tile's 4 bulk groups per warp) does not stall the
        // epilogue for an NVLink round trip.
        if (rs_prev_dst >= 0) rs_signal(rs_prev_dst, /*pending_groups_allowed=*/true);
        rs_prev_dst = rs_dst;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (MODE == MODE_GEMM_RS) {
      if (rs_prev_dst >= 0) rs_signal(rs_prev_dst, false);
    }
    if (lane == 0) tma_store_wait<0>();     // all epilogue stores of this warp have completed
    if (dbg && warp == 4 && lane == 0) {
      g2_dbg[blockIdx.x * 8 + 5] = w_tf; g2_dbg[blockIdx.x * 8 + 6] = g2_clock() - t_begin;
    }
  }

  tc_fence_before();
  cluster_sync();   // nobody may exit (or free TMEM) while the peer can still signal / read it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, TMEM_COLS);
  }

  if constexpr (MODE == MODE_GEMM_RS) rs_reduce_phase(p);
  if constexpr (MODE == MODE_AG_GEMM) {
    // (after the cluster barrier: every TMA load of this CTA has been consumed) tell the pushers of call e+2
    if (p.comm.ag_nvls && threadIdx.x == 0) ag_nvls_finish(p.comm);
  }
}

template <bool A_MN, bool B_MN, int EPI, int MODE = MODE_PLAIN>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                   int num_sms, cudaStream_t stream, const RsMaps* rsm = nullptr) {
  static RsMaps no_maps;
  if (!rsm) rsm = &no_maps;
  auto kern = gemm_bf16_2cta_kernel<A_MN, B_MN, EPI, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int tiles = ((p.M + G2_BLOCK_M - 1) / G2_BLOCK_M) * ((p.N + G2_BLOCK_N - 1) / G2_BLOCK_N);
  int pairs = num_sms / 2;
  int extra = 0;
  if constexpr (MODE == MODE_AG_GEMM) {
    extra = p.comm.num_comm_ctas;                  // even: whole clusters of pullers
    pairs = (num_sms - extra) / 2;
  }
  if (pairs > tiles) pairs = tiles;
  static const int dbg = getenv("MLB200_GEMM2_DEBUG") != nullptr;
  kern<<<2 * pairs + extra, GEMM_THREADS, G2_SMEM_TOTAL, stream>>>(tmA, tmB, tmC, *rsm, p, dbg);
  return (int)cudaGetLastError();
}

template <bool A_MN, bool B_MN>
static int dispatch2_epi(int epi, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                         const GemmParams& p, int sms, cudaStream_t st) {
  switch (epi) {
    case EPI_BF16: return launch2<A_MN, B_MN, EPI_BF16>(a, b, c, p, sms, st);
    case EPI_F32_ACCUM: return launch2<A_MN, B_MN, EPI_F32_ACCUM>(a, b, c, p, sms, st);
    case EPI_F32: return launch2<A_MN, B_MN, EPI_F32>(a, b, c, p, sms, st);
    case EPI_BF16_ACCUM: return launch2<A_MN, B_MN, EPI_BF16_ACCUM>(a, b, c, p, sms, st);
  }
  return -2;
}

}  // namespace mlb

// all-gather -> GEMM with the 2-CTA kernel: A = local gathered buffer [world * rows_per_rank, K] (filled by the puller
// CTAs while the tiles of already-arrived row blocks are computed), C[M, N] = A @ B^T (or A @ B), bf16 out
extern "C" int mlb_gemm_bf16_2cta_ag(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                                     int ldc, int b_mn_major, const mlb::GemmComm* comm, int num_sms,
                                     cudaStream_t stream) {
  using namespace mlb;
  if ((comm->ag_rows_per_rank % G2_BLOCK_M) != 0 || (comm->num_comm_ctas & 1)) return -3;
  CUtensorMap tmA, tmB, tmC;
  int r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, G2_HALF);
  if (r) return 1000 + r;
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, G2_HALF);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return 2000 + r;
  r = make_tmap_2d_out(&tmC, C, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32);
  if (r) return 3000 + r;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.comm = *comm;
  if (!b_mn_major) return launch2<false, false, EPI_BF16, MODE_AG_GEMM>(tmA, tmB, tmC, p, num_sms, stream);
  return launch2<false, true, EPI_BF16, MODE_AG_GEMM>(tmA, tmB, tmC, p, num_sms, stream);
}

// GEMM -> reduce-scatter with the 2-CTA kernel: partial C[M, N] = A @ B^T (or A @ B); the 256-row blocks that belong
// to rank d are TMA-stored into this rank's receive slot on d, every rank then reduces its world slots in fp32.
// Returns the number of half-tile arrivals every source contributes per destination (so the caller can keep the
// cumulative expected-arrivals counter), or a negative / >= 1000 error code.
extern "C" int mlb_gemm_bf16_2cta_rs(const void* A, const void* B, int M, int N, int K, int lda, int ldb,
                                     int b_mn_major, mlb::GemmComm* comm, int prev_total, int num_sms,
                                     cudaStream_t stream) {
  using namespace mlb;
  const int m = comm->rs_rows_per_rank;
  if (m % G2_BLOCK_M != 0 || M != m * comm->world || (N * 2) % 16 != 0) return -3;
  CUtensorMap tmA, tmB, tmC;
  int r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, G2_HALF);
  if (r) return -(1000 + r);
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, G2_HALF);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return -(2000 + r);
  RsMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int d = 0; d < comm->world; ++d) {
    r = make_tmap_2d_out(&maps.m[d], comm->rs_dst[d], 2, (uint64_t)N, (uint64_t)m, (uint64_t)N, 32);
    if (r) return -(3000 + r);
  }
  tmC = maps.m[comm->rank];
  const int arrivals = (m / G2_HALF) * ((N + G2_BLOCK_N - 1) / G2_BLOCK_N);   // one per CTA per tile
  comm->rs_expected_total = prev_total + arrivals;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = nullptr; p.M = M; p.N = N; p.K = K; p.ldc = N;
  p.comm = *comm;
  int e = b_mn_major ? launch2<false, true, EPI_BF16, MODE_GEMM_RS>(tmA, tmB, tmC, p, num_sms, stream, &maps)
                     : launch2<false, false, EPI_BF16, MODE_GEMM_RS>(tmA, tmB, tmC, p, num_sms, stream, &maps);
  return e ? -(4000 + e) : arrivals;
}

extern "C" int mlb_gemm2_debug_read(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, mlb::g2_dbg, sizeof(unsigned long long) * n);
}

extern "C" int mlb_gemm_bf16_2cta(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                                  int ldc, int a_mn_major, int b_mn_major, int epilogue, int fp16, int num_sms,
                                  cudaStream_t stream) {
  using namespace mlb;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap tmA, tmB;
  int r;
  if (!a_mn_major) r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, G2_HALF);
  else r = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BLOCK_K);
  if (r) return 1000 + r;
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, G2_HALF);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return 2000 + r;
  CUtensorMap tmC;
  const int out_bytes = (epilogue == EPI_BF16 || epilogue == EPI_BF16_ACCUM) ? 2 : 4;
  r = make_tmap_2d_out(&tmC, C, out_bytes, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32);
  if (r) return 3000 + r;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.fp16 = fp16;
  if (!a_mn_major && !b_mn_major) return dispatch2_epi<false, false>(epilogue, tmA, tmB, tmC, p, num_sms, stream);
  if (!a_mn_major && b_mn_major) return dispatch2_epi<false, true>(epilogue, tmA, tmB, tmC, p, num_sms, stream);
  if (a_mn_major && b_mn_major) return dispatch2_epi<true, true>(epilogue, tmA, tmB, tmC, p, num_sms, stream);
  return dispatch2_epi<true, false>(epilogue, tmA, tmB, tmC, p, num_sms, stream);
}

// Backward of the sm_100a attention (see attention_sm100.cu for the overall design).
//
//   delta   : D[b,h,s] = sum_d dO * O                                   (row-wise, memory bound)
//   dK/dV   : one CTA per (kv tile, kv head, batch).  For every query tile i >= j (causal) of every query head in the
//             GQA group:   S^T = K Q^T, dP^T = V dO^T  (kv rows on the TMEM lanes, queries on the columns)
//                          P^T = exp2(S^T c - lse[q]),  dS^T = P^T o (dP^T - D[q]) * scale       (bf16, written in place
//                          over the fp32 tiles they were computed from)
//                          dV += P^T dO,  dK += dS^T Q                   (A operand from TMEM, B = MN-major view of the
//                                                                         same smem tile that fed the first two MMAs)
//   dQ      : one CTA per (query tile, query head, batch).  For every kv tile j <= i:
//                          S = Q K^T, dP = dO V^T, dS = P o (dP - D) * scale (in place), dQ += dS K
// Every accumulator lives in TMEM; nothing is accumulated with atomics, so the result is deterministic.
#include "attention_common.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace mlb {

struct AttnBwdParams {
  HeadMap hm;
  int seq, batch, heads, kv_heads;
  int window;
  float scale, scale_log2;
  const float* lse;     // [b, heads, seq]
  const float* delta;   // [b, heads, seq]
  RowAddr dq, dk, dv;   // strided outputs ([seq, batch, heads, 128])
  DropoutParams drop;   // (read by the AF_DROPOUT instantiations only; same p and seed as the forward)
};

// ------------------------------------------------------------------------------------------------
// delta = rowsum(dO * O)
// ------------------------------------------------------------------------------------------------
template <int D, int F>
__global__ void attn_delta_kernel(RowAddr o, RowAddr dout, float* __restrict__ delta, int seq, int batch, int heads) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row_id = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const long long total = (long long)seq * batch * heads;
  if (row_id >= total) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(row_id % seq);
  const int h = (int)((row_id / seq) % heads);
  const int b = (int)(row_id / ((long long)seq * heads));
  float v;
  if constexpr (D == 128) {          // 4 elements per lane
    const uint2 a = *reinterpret_cast<const uint2*>(o.row(s, b, h) + lane * 4);
    const uint2 g = *reinterpret_cast<const uint2*>(dout.row(s, b, h) + lane * 4);
    const float2 a0 = unpack_h2<F>(a.x), a1 = unpack_h2<F>(a.y), g0 = unpack_h2<F>(g.x), g1 = unpack_h2<F>(g.y);
    v = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y;
  } else {                           // 2 elements per lane
    const float2 a0 = unpack_h2<F>(*reinterpret_cast<const uint32_t*>(o.row(s, b, h) + lane * 2));
    const float2 g0 = unpack_h2<F>(*reinterpret_cast<const uint32_t*>(dout.row(s, b, h) + lane * 2));
    v = a0.x * g0.x + a0.y * g0.y;
  }
  v = warp_sum(v);
  if (lane == 0) delta[((long long)b * heads + h) * seq + s] = v;
}

// Both backward kernels split every 128-wide score tile into two 64-column halves.  Each half has its own
// elementwise warpgroup (the math is purely elementwise: P = exp2(S c - lse), dS = P o (dP - D) c) and its own
// barriers, and the MMA warp ping-pongs between them:
//        [S,dP]_hi(t) and acc_lo(t) run while the warpgroups work on lo(t) / hi(t); [S,dP]_lo(t+1) follows acc_lo(t).
// The bf16 P / dS tiles overwrite the first half of the fp32 columns of their own 64-column half in place.
constexpr int BW_THREADS = 384;
constexpr int AT_ROWS64 = 64 * 128;   // byte offset of tile row 64 inside a 64-column swizzle box

// TMEM column map (dK/dV kernel)
constexpr uint32_t KV_ST = 0, KV_DPT = 128, KV_DV = 256, KV_DK = 384;
constexpr int BWD_SMEM = 6 * AT_TILE_BYTES + 2 * 2 * 128 * 4 + 256 + 1024;   // K, V, 2 x (Q, dO), 2 x (lse, D)
constexpr int BWD_SMEM_DROPOUT = BWD_SMEM + 2 * 128 * 4;                     // + 2 x per-query dropout row keys

template <int D, int F>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                     const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AT_TILE_BYTES;
  uint8_t* sQ = smem + 2 * AT_TILE_BYTES;     // 2 stages
  uint8_t* sDO = smem + 4 * AT_TILE_BYTES;    // 2 stages
  float* sLse = reinterpret_cast<float*>(smem + 6 * AT_TILE_BYTES);   // [2][128]  (log2 units)
  float* sD = sLse + 2 * 128;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * 128);
  uint64_t* kv_full = bars;        // K and V landed
  uint64_t* in_full = bars + 1;    // [2] Q + dO of stage s landed
  uint64_t* in_empty = bars + 3;   // [2] dV/dK MMAs of the tile finished reading stage s
  uint64_t* sdp_full = bars + 5;   // [2] S^T and dP^T of half x computed
  uint64_t* pds_full = bars + 7;   // [2] P^T and dS^T of half x written (4 warps)
  uint64_t* acc_done = bars + 9;   // all MMAs retired (epilogue)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);
  uint32_t* sRowKey = reinterpret_cast<uint32_t*>(bars) + 64;    // [2][128], behind the 256-byte barrier block (dropout)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int kv0 = j * AT_N;
  const int n_q_tiles = p.seq / AT_M;
  int i_hi = n_q_tiles - 1;
  if (p.window > 0) i_hi = min(i_hi, (kv0 + AT_N - 1 + p.window) / AT_M);
  const int tiles_per_head = i_hi - j + 1;
  const int g = p.hm.q_per_kv;
  const int n_iter = tiles_per_head * g;     // iteration it -> (head hq = it / tiles_per_head, tile i = j + it % tiles)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1);
      mbar_init(&sdp_full[s], 1); mbar_init(&pds_full[s], 4);
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc<1>(tmem_ptr_smem, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  constexpr bool BF16 = (F & AF_FP16) == 0;
  constexpr uint32_t ID_KK = make_idesc_f16(AT_N, 64, false, false, BF16);     // [128 kv] x [64 q], both K-major smem
  constexpr uint32_t ID_TS = make_idesc_f16(AT_N, D, false, true, BF16);    // A from TMEM, B MN-major smem

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * at_tile_tx<D>());
      load_tile<D>(sK, &tmK, kv_full, p.hm.k(kvh), kv0, b);
      load_tile<D>(sV, &tmV, kv_full, p.hm.v(kvh), kv0, b);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const int hq = kvh * g + it / tiles_per_head;
        const int q0 = (j + it % tiles_per_head) * AT_M;
        mbar_wait(&in_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&in_full[st], 2 * at_tile_tx<D>());
        load_tile<D>(sQ + st * AT_TILE_BYTES, &tmQ, &in_full[st], p.hm.q(hq), q0, b);
        load_tile<D>(sDO + st * AT_TILE_BYTES, &tmDO, &in_full[st], hq, q0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      // S^T_x = K Q_x^T ; dP^T_x = V dO_x^T          (x = query-column half)
      auto issue_sdp = [&](int x, int it) {
        const int st = it & 1;
        const uint32_t aQ = smem_u32(sQ + st * AT_TILE_BYTES) + x * AT_ROWS64;
        const uint32_t aDO = smem_u32(sDO + st * AT_TILE_BYTES) + x * AT_ROWS64;
        mbar_wait(&in_full[st], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + KV_ST + x * 64, desc_kmajor(aK, k), desc_kmajor(aQ, k), ID_KK, k != 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + KV_DPT + x * 64, desc_kmajor(aV, k), desc_kmajor(aDO, k), ID_KK, k != 0);
        umma_commit<1>(&sdp_full[x]);
      };
      // dV += P^T_x dO_x ; dK += dS^T_x Q_x
      auto issue_acc = [&](int x, int it) {
        const int st = it & 1;
        const uint32_t aQ = smem_u32(sQ + st * AT_TILE_BYTES), aDO = smem_u32(sDO + st * AT_TILE_BYTES);
        mbar_wait(&pds_full[x], it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tmem + KV_DV, tmem + KV_ST + x * 64 + k * 8, desc_mnmajor(aDO, x * 4 + k), ID_TS,
                      (it | x | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tmem + KV_DK, tmem + KV_DPT + x * 64 + k * 8, desc_mnmajor(aQ, x * 4 + k), ID_TS,
                      (it | x | k) != 0 ? 1u : 0u);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0, 0);
      issue_sdp(1, 0);
      for (int it = 0; it < n_iter; ++it) {
        issue_acc(0, it);
        if (it + 1 < n_iter) issue_sdp(0, it + 1);
        issue_acc(1, it);
        umma_commit<1>(&in_empty[it & 1]);
        if (it + 1 < n_iter) issue_sdp(1, it + 1);
      }
      umma_commit<1>(acc_done);
    }
  } else if (warp >= 4) {
    const int x = (warp - 4) >> 2;        // query-column half of this warpgroup
    const int q = warp & 3;               // TMEM lane quarter
    const int r = q * 32 + lane;          // kv row inside the tile (TMEM lane)
    const int kv = kv0 + r;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t st_addr = tmem + lane_addr + KV_ST + x * 64, dp_addr = tmem + lane_addr + KV_DPT + x * 64;
    for (int it = 0; it < n_iter; ++it) {
      const int st = it & 1;
      const int hq = kvh * g + it / tiles_per_head;
      const int q0 = (j + it % tiles_per_head) * AT_M;
      // per-query statistics of this half -> smem: threads 0..63 fetch lse, 64..127 fetch D
      {
        const int c = r & 63;
        const long long stat = ((long long)b * p.heads + hq) * p.seq + q0 + x * 64 + c;
        if (r < 64) sLse[st * 128 + x * 64 + c] = p.lse[stat] * 1.4426950408889634f;
        else sD[st * 128 + x * 64 + c] = p.delta[stat];
        if constexpr ((F & AF_DROPOUT) != 0) {
          if (r < 64)
            sRowKey[st * 128 + x * 64 + c] =
                drop_row_key(p.drop.seed_lo, drop_head_key(p.drop.seed_hi, uint32_t(b * p.heads + hq)),
                             uint32_t(q0 + x * 64 + c));
        }
      }
      if (x == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
      else asm volatile("bar.sync 3, 128;" ::: "memory");
      mbar_wait(&sdp_full[x], it & 1);
      tc_fence_after();
      // warp-uniform boundary-tile test; the masked path uses selects against per-chunk column limits
      const bool need_mask = (q0 < kv0 + AT_N - 1) || (p.window > 0 && q0 + AT_M - 1 > kv0 + p.window);
      const int qlo = kv - q0 - x * 64;                                       // query columns < qlo precede this key
      const int qhi = (p.window > 0) ? kv + p.window - q0 - x * 64 : (1 << 30);   // columns > qhi lost it from the window
      const float* lse_s = sLse + st * 128 + x * 64;
      const float* d_s = sD + st * 128 + x * 64;
      [[maybe_unused]] const uint32_t* rk_s = sRowKey + st * 128 + x * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32(st_addr + c * 32, s);
        tmem_ld_32x32(dp_addr + c * 32, dp);
        tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c * 32 + 2 * i + e;     // query index inside the half
            float pe = fast_exp2(fmaf(__uint_as_float(s[2 * i + e]), p.scale_log2, -lse_s[col]));
            if (need_mask) pe = (col < qlo || col > qhi) ? 0.f : pe;
            if constexpr ((F & AF_DROPOUT) != 0) {
              // O = (P o Z) V with Z = keep / (1 - p):  dV += (P o Z)^T dO,  dS = P o (Z o dP - D)
              const bool dropped = drop_is_dropped(drop_bytes(rk_s[col], uint32_t(kv) >> 2), uint32_t(kv),
                                                   p.drop.threshold);
              const float dpz = dropped ? 0.f : __uint_as_float(dp[2 * i + e]) * p.drop.inv_keep;
              pv[e] = dropped ? 0.f : pe * p.drop.inv_keep;
              dv[e] = pe * (dpz - d_s[col]) * p.scale;
            } else {
              pv[e] = pe;
              dv[e] = pe * (__uint_as_float(dp[2 * i + e]) - d_s[col]) * p.scale;
            }
          }
          pk[i] = pack_h2<F>(pv[0], pv[1]);
          dk[i] = pack_h2<F>(dv[0], dv[1]);
        }
        tmem_st_32x16(st_addr + c * 16, pk);     // in place: only already-consumed columns of this half
        tmem_st_32x16(dp_addr + c * 16, dk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[x]);
    }
    // epilogue: dV, dK -> bf16 (each warpgroup writes one half of the head-dim columns)
    mbar_wait(acc_done, 0);
    tc_fence_after();
    if (kv < p.seq) {
      // gradients use the same head -> coordinate map as the inputs (separate tensors or one packed QKV buffer)
      __nv_bfloat16* dvrow = p.dv.row(kv, b, p.hm.v(kvh)) + x * (D / 2);
      __nv_bfloat16* dkrow = p.dk.row(kv, b, p.hm.k(kvh)) + x * (D / 2);
#pragma unroll 1
      for (int c = 0; c < D / 64; ++c) {
        uint32_t a[32], bb[32];
        tmem_ld_32x32(tmem + lane_addr + KV_DV + x * (D / 2) + c * 32, a);
        tmem_ld_32x32(tmem + lane_addr + KV_DK + x * (D / 2) + c * 32, bb);
        tmem_ld_wait();
        uint4* d0 = reinterpret_cast<uint4*>(dvrow + c * 32);
        uint4* d1 = reinterpret_cast<uint4*>(dkrow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_h2<F>(__uint_as_float(a[i * 8 + 0]), __uint_as_float(a[i * 8 + 1]));
          o.y = pack_h2<F>(__uint_as_float(a[i * 8 + 2]), __uint_as_float(a[i * 8 + 3]));
          o.z = pack_h2<F>(__uint_as_float(a[i * 8 + 4]), __uint_as_float(a[i * 8 + 5]));
          o.w = pack_h2<F>(__uint_as_float(a[i * 8 + 6]), __uint_as_float(a[i * 8 + 7]));
          d0[i] = o;
          o.x = pack_h2<F>(__uint_as_float(bb[i * 8 + 0]), __uint_as_float(bb[i * 8 + 1]));
          o.y = pack_h2<F>(__uint_as_float(bb[i * 8 + 2]), __uint_as_float(bb[i * 8 + 3]));
          o.z = pack_h2<F>(__uint_as_float(bb[i * 8 + 4]), __uint_as_float(bb[i * 8 + 5]));
          o.w = pack_h2<F>(__uint_as_float(bb[i * 8 + 6]), __uint_as_float(bb[i * 8 + 7]));
          d1[i] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

// TMEM column map (dQ kernel)
constexpr uint32_t Q_S = 0, Q_DP = 128, Q_DQ = 256;
constexpr int BWD_DQ_SMEM = 6 * AT_TILE_BYTES + 256 + 1024;   // Q, dO, 2 x (K, V)

template <int D, int F>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + AT_TILE_BYTES;
  uint8_t* sK = smem + 2 * AT_TILE_BYTES;    // 2 stages
  uint8_t* sV = smem + 4 * AT_TILE_BYTES;    // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * AT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* in_full = bars + 1;    // [2]
  uint64_t* in_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;   // [2] per kv-column half
  uint64_t* ds_full = bars + 7;    // [2] per kv-column half
  uint64_t* acc_done = bars + 9;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = (int)gridDim.x - 1 - (int)blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = i * AT_M;
  const int kvh = h / p.hm.q_per_kv;
  int j_lo = 0;
  if (p.window > 0) j_lo = max(0, (q0 - p.window) / AT_N);
  const int n_iter = i - j_lo + 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1);
      mbar_init(&sdp_full[s], 1); mbar_init(&ds_full[s], 4);
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc<1>(tmem_ptr_smem, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  constexpr bool BF16 = (F & AF_FP16) == 0;
  constexpr uint32_t ID_KK = make_idesc_f16(AT_M, 64, false, false, BF16);
  constexpr uint32_t ID_TS = make_idesc_f16(AT_M, D, false, true, BF16);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * at_tile_tx<D>());
      load_tile<D>(sQ, &tmQ, q_full, p.hm.q(h), q0, b);
      load_tile<D>(sDO, &tmDO, q_full, h, q0, b);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const int kv0 = (j_lo + it) * AT_N;
        mbar_wait(&in_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&in_full[st], 2 * at_tile_tx<D>());
        load_tile<D>(sK + st * AT_TILE_BYTES, &tmK, &in_full[st], p.hm.k(kvh), kv0, b);
        load_tile<D>(sV + st * AT_TILE_BYTES, &tmV, &in_full[st], p.hm.v(kvh), kv0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO);
      auto issue_sdp = [&](int x, int it) {      // S_x = Q K_x^T ; dP_x = dO V_x^T
        const int st = it & 1;
        const uint32_t aK = smem_u32(sK + st * AT_TILE_BYTES) + x * AT_ROWS64;
        const uint32_t aV = smem_u32(sV + st * AT_TILE_BYTES) + x * AT_ROWS64;
        mbar_wait(&in_full[st], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + Q_S + x * 64, desc_kmajor(aQ, k), desc_kmajor(aK, k), ID_KK, k != 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + Q_DP + x * 64, desc_kmajor(aDO, k), desc_kmajor(aV, k), ID_KK, k != 0);
        umma_commit<1>(&sdp_full[x]);
      };
      auto issue_acc = [&](int x, int it) {      // dQ += dS_x K_x
        const uint32_t aK = smem_u32(sK + (it & 1) * AT_TILE_BYTES);
        mbar_wait(&ds_full[x], it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tmem + Q_DQ, tmem + Q_DP + x * 64 + k * 8, desc_mnmajor(aK, x * 4 + k), ID_TS,
                      (it | x | k) != 0 ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      issue_sdp(0, 0);
      issue_sdp(1, 0);
      for (int it = 0; it < n_iter; ++it) {
        issue_acc(0, it);
        if (it + 1 < n_iter) issue_sdp(0, it + 1);
        issue_acc(1, it);
        umma_commit<1>(&in_empty[it & 1]);
        if (it + 1 < n_iter) issue_sdp(1, it + 1);
      }
      umma_commit<1>(acc_done);
    }
  } else if (warp >= 4) {
    const int x = (warp - 4) >> 2;        // kv-column half of this warpgroup
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem + lane_addr + Q_S + x * 64, dp_addr = tmem + lane_addr + Q_DP + x * 64;
    const long long stat = ((long long)b * p.heads + h) * p.seq + row;
    const float lse2 = p.lse[stat] * 1.4426950408889634f;
    const float dlt = p.delta[stat];
    uint32_t row_key = 0;
    if constexpr ((F & AF_DROPOUT) != 0)
      row_key = drop_row_key(p.drop.seed_lo, drop_head_key(p.drop.seed_hi, uint32_t(b * p.heads + h)), uint32_t(row));
    for (int it = 0; it < n_iter; ++it) {
      const int kv0 = (j_lo + it) * AT_N + x * 64;
      mbar_wait(&sdp_full[x], it & 1);
      tc_fence_after();
      const bool need_mask = __any_sync(0xffffffffu, (kv0 + 63 > row) || (p.window > 0 && kv0 < row - p.window));
      const int hi = row - kv0;                                             // columns > hi are in the future
      const int lo = (p.window > 0) ? row - p.window - kv0 : -(1 << 30);    // columns < lo fell out of the window
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32(s_addr + c * 32, s);
        tmem_ld_32x32(dp_addr + c * 32, dp);
        tmem_ld_wait();
        uint32_t dk[16];
#pragma unroll
        for (int e2 = 0; e2 < 16; ++e2) {
          float dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c * 32 + 2 * e2 + e;
            float pe = fast_exp2(fmaf(__uint_as_float(s[2 * e2 + e]), p.scale_log2, -lse2));
            if (need_mask) pe = (col > hi || col < lo) ? 0.f : pe;
            float dpe = __uint_as_float(dp[2 * e2 + e]);
            if constexpr ((F & AF_DROPOUT) != 0) {
              const uint32_t key = uint32_t(kv0 + col);
              dpe = drop_is_dropped(drop_bytes(row_key, key >> 2), key, p.drop.threshold) ? 0.f
                                                                                          : dpe * p.drop.inv_keep;
            }
            dv[e] = pe * (dpe - dlt) * p.scale;
          }
          dk[e2] = pack_h2<F>(dv[0], dv[1]);
        }
        tmem_st_32x16(dp_addr + c * 16, dk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[x]);
    }
    mbar_wait(acc_done, 0);
    tc_fence_after();
    if (row < p.seq) {
      __nv_bfloat16* dqrow = p.dq.row(row, b, p.hm.q(h)) + x * (D / 2);
#pragma unroll 1
      for (int c = 0; c < D / 64; ++c) {
        uint32_t a[32];
        tmem_ld_32x32(tmem + lane_addr + Q_DQ + x * (D / 2) + c * 32, a);
        tmem_ld_wait();
        uint4* d0 = reinterpret_cast<uint4*>(dqrow + c * 32);
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          uint4 o;
          o.x = pack_h2<F>(__uint_as_float(a[i2 * 8 + 0]), __uint_as_float(a[i2 * 8 + 1]));
          o.y = pack_h2<F>(__uint_as_float(a[i2 * 8 + 2]), __uint_as_float(a[i2 * 8 + 3]));
          o.z = pack_h2<F>(__uint_as_float(a[i2 * 8 + 4]), __uint_as_float(a[i2 * 8 + 5]));
          o.w = pack_h2<F>(__uint_as_float(a[i2 * 8 + 6]), __uint_as_float(a[i2 * 8 + 7]));
          d0[i2] = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

template <int D, int F>
static int launch_attn_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                           const AttnBwdParams& p, const RowAddr& ro, const RowAddr& rdo, float* delta, int q_per_kv,
                           cudaStream_t stream) {
  constexpr int KV_SMEM = (F & AF_DROPOUT) != 0 ? BWD_SMEM_DROPOUT : BWD_SMEM;
  const long long rows = (long long)p.seq * p.batch * p.heads;
  attn_delta_kernel<D, F & AF_FP16><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(ro, rdo, delta, p.seq, p.batch,
                                                                                    p.heads);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkdv_kernel<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, KV_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_bwd_dq_kernel<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_DQ_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 g1(p.seq / AT_N, p.heads / q_per_kv, p.batch);
  attn_bwd_dkdv_kernel<D, F><<<g1, BW_THREADS, KV_SMEM, stream>>>(tq, tk, tv, tdo, p);
  dim3 g2(p.seq / AT_M, p.heads, p.batch);
  attn_bwd_dq_kernel<D, F><<<g2, BW_THREADS, BWD_DQ_SMEM, stream>>>(tq, tk, tv, tdo, p);
  return (int)cudaGetLastError();
}

template <int D>
static int launch_attn_bwd_flags(int flags, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                 const CUtensorMap& tdo, const AttnBwdParams& p, const RowAddr& ro, const RowAddr& rdo,
                                 float* delta, int q_per_kv, cudaStream_t stream) {
  switch (flags) {
    case 0: return launch_attn_bwd<D, 0>(tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream);
    case AF_FP16: return launch_attn_bwd<D, AF_FP16>(tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream);
    case AF_DROPOUT: return launch_attn_bwd<D, AF_DROPOUT>(tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream);
    default: return launch_attn_bwd<D, AF_FP16 | AF_DROPOUT>(tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream);
  }
}

}  // namespace mlb

// ``fp16`` / ``dropout_p`` / ``seed``: as in mlb_attn_fwd_ex (the mask is regenerated from the same p and seed).
extern "C" int mlb_attn_bwd_ex(const void* q, const void* k, const void* v, const void* o, const void* dout,
                               const long long* q_str, const long long* k_str, const long long* v_str,
                               const long long* o_str, const long long* do_str, int q_map_heads, int k_map_heads,
                               int v_map_heads, const int* head_map, int q_per_kv, int seq, int batch, int heads,
                               int window, float softmax_scale, const float* lse, float* delta, void* dq, void* dk,
                               void* dv, const long long* dq_str, const long long* dk_str, const long long* dv_str,
                               int head_dim, int fp16, float dropout_p, unsigned long long seed, cudaStream_t stream) {
  using namespace mlb;
  if (seq % AT_M != 0 || (head_dim != 64 && head_dim != 128) || dropout_p < 0.f || dropout_p >= 1.f) return -2;
  CUtensorMap tq, tk, tv, tdo;
  int r = make_tmap_heads(&tq, q, head_dim, q_map_heads, seq, batch, q_str[0], q_str[1], q_str[2], AT_M);
  if (r) return 1000 + r;
  r = make_tmap_heads(&tk, k, head_dim, k_map_heads, seq, batch, k_str[0], k_str[1], k_str[2], AT_N);
  if (r) return 2000 + r;
  r = make_tmap_heads(&tv, v, head_dim, v_map_heads, seq, batch, v_str[0], v_str[1], v_str[2], AT_N);
  if (r) return 3000 + r;
  r = make_tmap_heads(&tdo, dout, head_dim, heads, seq, batch, do_str[0], do_str[1], do_str[2], AT_M);
  if (r) return 4000 + r;
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.hm.q_group_stride = head_map[0]; p.hm.q_off = head_map[1]; p.hm.k_group_stride = head_map[2];
  p.hm.k_off = head_map[3]; p.hm.v_group_stride = head_map[4]; p.hm.v_off = head_map[5]; p.hm.q_per_kv = q_per_kv;
  p.seq = seq; p.batch = batch; p.heads = heads; p.kv_heads = heads / q_per_kv; p.window = window;
  p.scale = softmax_scale; p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta;
  p.dq = RowAddr{dq, dq_str[0], dq_str[1], dq_str[2]};
  p.dk = RowAddr{dk, dk_str[0], dk_str[1], dk_str[2]};
  p.dv = RowAddr{dv, dv_str[0], dv_str[1], dv_str[2]};
  const RowAddr ro{const_cast<void*>(o), o_str[0], o_str[1], o_str[2]};
  const RowAddr rdo{const_cast<void*>(dout), do_str[0], do_str[1], do_str[2]};
  p.drop = make_dropout_params(dropout_p, seed);
  const int flags = (fp16 ? AF_FP16 : 0) | (p.drop.threshold > 0 ? AF_DROPOUT : 0);
  return head_dim == 128 ? launch_attn_bwd_flags<128>(flags, tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream)
                         : launch_attn_bwd_flags<64>(flags, tq, tk, tv, tdo, p, ro, rdo, delta, q_per_kv, stream);
}

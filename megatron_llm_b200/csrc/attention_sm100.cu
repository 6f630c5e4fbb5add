// FlashAttention-style causal attention for sm_100a on the 5th-gen tensor cores (head_dim 128, bf16).
//
// Replaces the reference's call into the FA-2 library (megatron/model/transformer.py:538-553; Ampere-style
// mma.sync + cp.async kernels, K/V first broadcast to the query head count :458-465).  Here:
//   * S = Q K^T and O += P V run on tcgen05.mma with the S / P / O tiles resident in TMEM (256 KB per SM):
//       TMEM columns [0,256)   two S buffers  (128 lanes x 128 fp32)      -- S_{j+1} is computed while softmax_j runs
//       TMEM columns [256,384) two P buffers  (128 lanes x 128 bf16)      -- P is the A operand of the PV MMA
//       TMEM columns [384,512) O accumulator  (128 lanes x 128 fp32)
//   * Q / K / V tiles arrive by TMA (4-D tensor maps over the strided [s, b, heads, hn] layout, so the packed QKV GEMM
//     output is consumed in place: no head-major copies, native GQA/MQA through the head -> coordinate map);
//   * one thread per query row does the online softmax (exp2 with the scale folded in), with lazy rescaling of O
//     (only when the running max grows by more than 2^8), causal / sliding-window masking on the boundary tiles only.
// Backward = two kernels built from the same blocks (no atomics, deterministic):
//   * dK/dV: one CTA per KV tile, loops over the query tiles (and the query heads of the GQA group) with the
//     transposed score tile S^T = K Q^T so that P^T / dS^T are directly the A operands of dV += P^T dO, dK += dS^T Q;
//   * dQ:    one CTA per query tile, recomputes S and dP and accumulates dQ += dS K.
#include "attention_common.cuh"

#include <stdlib.h>

namespace mlb {

// TMEM column map (forward)
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_P0 = 256, TM_P1 = 320, TM_O = 384;

struct AttnParams {
  // head -> coordinate in the tensor maps' "heads" dimension:  q: (h / g) * q_gs + (h % g) + q_off ; kv: (h / g) * kv_gs + off
  int q_group_stride, q_off, k_group_stride, k_off, v_group_stride, v_off;
  int q_per_kv;
  int seq, batch, heads;          // heads = query heads
  int window;                     // <= 0: none; else keys in [row - window, row]
  float scale_log2;               // softmax_scale * log2(e)
  void* out;                      // [.., hn] rows: out + ((s * out_s_stride) + b * out_b_stride + h * 128)
  long long out_s_stride, out_b_stride;   // in elements
  float* lse;                     // [batch, heads, seq] natural-log logsumexp of the scaled scores
  DropoutParams drop;             // (read by the AF_DROPOUT instantiations only)
};

// =================================================================================================
// forward
// =================================================================================================
template <int D, int F>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE_BYTES;          // 2 stages
  uint8_t* sV = smem + 3 * AT_TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * AT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* o_done = bars + 13;   // one completion per PV
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = (int)gridDim.x - 1 - (int)blockIdx.x;   // heavy (late) tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * AT_M;
  const int kvh = h / p.q_per_kv;
  const int q_coord = kvh * p.q_group_stride + (h % p.q_per_kv) + p.q_off;
  const int k_coord = kvh * p.k_group_stride + p.k_off;
  const int v_coord = kvh * p.v_group_stride + p.v_off;
  // kv tile range [j_lo, j_hi]
  const int j_hi = q_tile;
  int j_lo = 0;
  if (p.window > 0) j_lo = max(0, (q0 - p.window) / AT_N);
  const int n_tiles = j_hi - j_lo + 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  constexpr bool BF16 = (F & AF_FP16) == 0;
  constexpr uint32_t IDESC_S = make_idesc_f16(AT_M, AT_N, false, false, BF16);   // Q K^T : both K-major
  constexpr uint32_t IDESC_PV = make_idesc_f16(AT_M, D, false, true, BF16);   // P (TMEM) x V (MN-major)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, at_tile_tx<D>());
      load_tile<D>(sQ, &tmQ, q_full, q_coord, q0, b);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int kv0 = (j_lo + t) * AT_N;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], at_tile_tx<D>());
        load_tile<D>(sK + st * AT_TILE_BYTES, &tmK, &k_full[st], k_coord, kv0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], at_tile_tx<D>());
        load_tile<D>(sV + st * AT_TILE_BYTES, &tmV, &v_full[st], v_coord, kv0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ);
      auto issue_S = [&](int t) {
        const int st = t & 1;
        mbar_wait(&k_full[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + (st ? TM_S1 : TM_S0), desc_kmajor(aQ, k), desc_kmajor(aK, k), IDESC_S, k != 0);
        umma_commit<1>(&k_empty[st]);
        umma_commit<1>(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      if (n_tiles > 1) issue_S(1);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(&p_full[st], ph);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < AT_N / 16; ++k)
          umma_f16_ts(tmem + TM_O, tmem + (st ? TM_P1 : TM_P0) + k * 8, desc_mnmajor(aV, k), IDESC_PV,
                      (t | k) != 0 ? 1u : 0u);
        umma_commit<1>(&v_empty[st]);
        umma_commit<1>(o_done);
        if (t + 2 < n_tiles) issue_S(t + 2);   // S buffer `st` was drained by softmax_t before it signalled p_full
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax / correction / epilogue: one thread per query row ------------------
    const int q = warp - 4;
    const int r = q * 32 + lane;            // row inside the tile == TMEM lane
    const int row = q0 + r;                 // global query position
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    uint32_t row_key = 0;
    if constexpr ((F & AF_DROPOUT) != 0)
      row_key = drop_row_key(p.drop.seed_lo, drop_head_key(p.drop.seed_hi, uint32_t(b * p.heads + h)), uint32_t(row));
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t & 1;
      const int kv0 = (j_lo + t) * AT_N;
      mbar_wait(&s_full[st], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem + lane_addr + (st ? TM_S1 : TM_S0);
      const bool need_mask = (kv0 + AT_N - 1 > row) || (p.window > 0 && kv0 < row - p.window);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (need_mask) {
            const int col = kv0 + c * 32 + i;
            if (col > row || (p.window > 0 && col < row - p.window)) s = -INFINITY;
          }
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      // lazy rescale: only when the running max grew by more than 2^8 (warp-uniform decision)
      const bool grow = (m_new > m_used + 8.0f) || (m_used == -INFINITY && m_new != -INFINITY);
      bool saw_pv = false;                 // (warp-uniform) this iteration already waited for PV_{t-1}
      if (__any_sync(0xffffffffu, grow && t > 0)) {
        mbar_wait(o_done, (t - 1) & 1);    // PV_{t-1} must have landed in O
        saw_pv = true;
        tc_fence_after();
        const float alpha = grow ? ((m_used == -INFINITY) ? 0.f : fast_exp2(m_used - m_new)) : 1.f;
        if (grow) { m_used = m_new; l *= alpha; }
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(tmem + lane_addr + TM_O + c * 32, v);
        }
        tmem_st_wait();
      } else if (grow) {
        m_used = m_new;   // t == 0: nothing accumulated yet
      }
      // pass 2: p = exp2(s * c - m_used), packed to bf16 into the P buffer
      const float moff = (m_used == -INFINITY) ? 0.f : m_used;
      const uint32_t p_addr = tmem + lane_addr + (st ? TM_P1 : TM_P0);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
          if (need_mask) {
            const int col = kv0 + c * 32 + 2 * i;
            if (col > row || (p.window > 0 && col < row - p.window)) s0 = -INFINITY;
            if (col + 1 > row || (p.window > 0 && col + 1 < row - p.window)) s1 = -INFINITY;
          }
          float p0 = fast_exp2(s0 * p.scale_log2 - moff), p1 = fast_exp2(s1 * p.scale_log2 - moff);
          l += p0 + p1;
          if constexpr ((F & AF_DROPOUT) != 0) {     // the row sum keeps the undropped probabilities
            const uint32_t key = uint32_t(kv0 + c * 32 + 2 * i);
            const uint32_t bytes = drop_bytes(row_key, key >> 2);
            p0 = drop_is_dropped(bytes, key, p.drop.threshold) ? 0.f : p0;
            p1 = drop_is_dropped(bytes, key + 1, p.drop.threshold) ? 0.f : p1;
          }
          pk[i] = pack_h2<F>(p0, p1);
        }
        tmem_st_32x16(p_addr + c * 16, pk);
      }
      tmem_st_wait();
      // Every warp observes EVERY phase of o_done, one per PV: a parity wait can only tell two consecutive phases
      // apart, and with two S buffers a warp may otherwise finish its last tile while PV_{n-2} is still pending and
      // take the epilogue's wait for PV_{n-1} as already satisfied (found by running this kernel on the functional
      // tcgen05 model under ThreadSanitizer, tests/test_attention_kernel_model.py).  PV_t cannot complete before this
      // warp arrives on p_full below, so the barrier is never more than one phase ahead of the waiter.
      if (t > 0 && !saw_pv) mbar_wait(o_done, (t - 1) & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
    // ------------------------------ epilogue: O / l -> bf16, plus the log-sum-exp ------------------------------
    mbar_wait(o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    if constexpr ((F & AF_DROPOUT) != 0) inv_l *= p.drop.inv_keep;
    if (row < p.seq) {
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.out_s_stride +
                            (long long)b * p.out_b_stride + (long long)h * D;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + lane_addr + TM_O + c * 32, v);
        tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_h2<F>(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
          o.y = pack_h2<F>(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
          o.z = pack_h2<F>(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
          o.w = pack_h2<F>(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
          dst[i] = o;
        }
      }
      if (p.lse) p.lse[((long long)b * p.heads + h) * p.seq + row] = m_used * 0.6931471805599453f + logf(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

constexpr int AT_FWD_SMEM = 5 * AT_TILE_BYTES + 256 + 1024;

// =================================================================================================
// forward, two query tiles per CTA (256 query rows)
// =================================================================================================
// The single-tile kernel above is bound by its one softmax warpgroup (one warp per SM sub-partition: no latency
// hiding, and the tensor core idles while it works).  Here a CTA owns two adjacent query tiles `a` and `b`, each with
// its own softmax warpgroup, and the MMA warp ping-pongs between them:
//        S_a(t+1) / PV_b(t) run while softmax_b(t) / softmax_a(t+1) are busy.
// TMEM: [0,128) S_a, [128,256) S_b, [256,384) O_a, [384,512) O_b; P_x (bf16) overwrites the first 64 columns of S_x in
// place (a thread has consumed columns [0, 32c+32) of its row before it writes P columns [16c, 16c+16)).
constexpr int AT2_THREADS = 384;
constexpr uint32_t T2_S0 = 0, T2_S1 = 128, T2_O0 = 256, T2_O1 = 384;
constexpr int AT_FWD2_SMEM = 6 * AT_TILE_BYTES + 256 + 1024;

template <int D, int F>
__global__ void __launch_bounds__(AT2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 2 query tiles
  uint8_t* sK = smem + 2 * AT_TILE_BYTES;      // 2 stages
  uint8_t* sV = smem + 4 * AT_TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * AT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  uint64_t* p_full = bars + 11;   // [2] per query tile
  uint64_t* o_done = bars + 13;   // [2] per query tile
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = (int)gridDim.x - 1 - (int)blockIdx.x;   // heavy (late) tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = pair * 2 * AT_M;
  const int kvh = h / p.q_per_kv;
  const int q_coord = kvh * p.q_group_stride + (h % p.q_per_kv) + p.q_off;
  const int k_coord = kvh * p.k_group_stride + p.k_off;
  const int v_coord = kvh * p.v_group_stride + p.v_off;
  // kv tiles [j_lo, j_hi]: tile b ends on its diagonal (2*pair+1), tile a one tile earlier
  const int j_hi = 2 * pair + 1;
  int j_lo = 0;
  if (p.window > 0) j_lo = max(0, (q0 - p.window) / AT_N);
  const int n_b = j_hi - j_lo + 1;
  const int n_a = n_b - 1;          // >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  constexpr bool BF16 = (F & AF_FP16) == 0;
  constexpr uint32_t IDESC_S = make_idesc_f16(AT_M, AT_N, false, false, BF16);
  constexpr uint32_t IDESC_PV = make_idesc_f16(AT_M, D, false, true, BF16);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * at_tile_tx<D>());
      load_tile<D>(sQ, &tmQ, q_full, q_coord, q0, b);
      load_tile<D>(sQ + AT_TILE_BYTES, &tmQ, q_full, q_coord, q0 + AT_M, b);
      for (int t = 0; t < n_b; ++t) {
        const int st = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int kv0 = (j_lo + t) * AT_N;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], at_tile_tx<D>());
        load_tile<D>(sK + st * AT_TILE_BYTES, &tmK, &k_full[st], k_coord, kv0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], at_tile_tx<D>());
        load_tile<D>(sV + st * AT_TILE_BYTES, &tmV, &v_full[st], v_coord, kv0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      auto issue_S = [&](int x, int t) {
        const int st = t & 1;
        mbar_wait(&k_full[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + x * AT_TILE_BYTES);
        const uint32_t aK = smem_u32(sK + st * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16_ss<1>(tmem + (x ? T2_S1 : T2_S0), desc_kmajor(aQ, k), desc_kmajor(aK, k), IDESC_S, k != 0);
        umma_commit<1>(&s_full[x]);
      };
      auto issue_PV = [&](int x, int t) {
        const int st = t & 1;
        mbar_wait(&p_full[x], t & 1);
        mbar_wait(&v_full[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < AT_N / 16; ++k)
          umma_f16_ts(tmem + (x ? T2_O1 : T2_O0), tmem + (x ? T2_S1 : T2_S0) + k * 8, desc_mnmajor(aV, k), IDESC_PV,
                      (t | k) != 0 ? 1u : 0u);
        umma_commit<1>(&o_done[x]);
      };
      mbar_wait(q_full, 0);
      issue_S(0, 0);
      issue_S(1, 0);
      umma_commit<1>(&k_empty[0]);
      for (int t = 0; t < n_b; ++t) {
        if (t < n_a) {
          issue_PV(0, t);
          if (t + 1 < n_a) issue_S(0, t + 1);
        }
        issue_PV(1, t);
        umma_commit<1>(&v_empty[t & 1]);
        if (t + 1 < n_b) {
          issue_S(1, t + 1);
          umma_commit<1>(&k_empty[(t + 1) & 1]);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax / correction / epilogue: one thread per query row ------------------
    const int x = (warp - 4) >> 2;          // query tile of this warpgroup
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int row = q0 + x * AT_M + r;
    const int n_x = x ? n_b : n_a;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tmem + lane_addr + (x ? T2_S1 : T2_S0);
    const uint32_t o_addr = tmem + lane_addr + (x ? T2_O1 : T2_O0);
    const float sc = p.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    uint32_t row_key = 0;
    if constexpr ((F & AF_DROPOUT) != 0)
      row_key = drop_row_key(p.drop.seed_lo, drop_head_key(p.drop.seed_hi, uint32_t(b * p.heads + h)), uint32_t(row));
    for (int t = 0; t < n_x; ++t) {
      const int kv0 = (j_lo + t) * AT_N;
      mbar_wait(&s_full[x], t & 1);
      tc_fence_after();
      // masking only touches the boundary tiles; the decision is made warp-uniform so the common path carries no
      // per-element branches, and the masked path uses selects against per-chunk column limits
      const bool need_mask = __any_sync(0xffffffffu, (kv0 + AT_N - 1 > row) || (p.window > 0 && kv0 < row - p.window));
      const int hi = row - kv0;                                        // columns  > hi are in the future
      const int lo = (p.window > 0) ? row - p.window - kv0 : -(1 << 30);   // columns < lo fell out of the window
      // pass 1: row max (two chunks in flight)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(s_addr + c * 32, v0);
        tmem_ld_32x32(s_addr + c * 32 + 32, v1);
        tmem_ld_wait();
        if (need_mask) {
          const int h0 = hi - c * 32, l0 = lo - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = (i > h0 || i < l0) ? 0xff800000u : v0[i];
            v1[i] = (i + 32 > h0 || i + 32 < l0) ? 0xff800000u : v1[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(v0[i]));
          mx1 = fmaxf(mx1, __uint_as_float(v1[i]));
        }
      }
      const float m_new = fmaxf(m_used, fmaxf(mx0, mx1) * sc);
      const bool grow = (m_new > m_used + 8.0f) || (m_used == -INFINITY && m_new != -INFINITY);
      if (__any_sync(0xffffffffu, grow && t > 0)) {
        mbar_wait(&o_done[x], (t - 1) & 1);    // PV_x(t-1) must have landed in O_x
        tc_fence_after();
        const float alpha = grow ? ((m_used == -INFINITY) ? 0.f : fast_exp2(m_used - m_new)) : 1.f;
        if (grow) { m_used = m_new; l *= alpha; }
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(o_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(o_addr + c * 32, v);
        }
        tmem_st_wait();
      } else if (grow) {
        m_used = m_new;   // t == 0: nothing accumulated yet
      }
      // pass 2: p = exp2(s * c - m_used) packed to bf16 over the consumed part of S
      const float moff = (m_used == -INFINITY) ? 0.f : m_used;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + c * 32, v);
        tmem_ld_wait();
        if (need_mask) {
          const int h0 = hi - c * 32, l0c = lo - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (i > h0 || i < l0c) ? 0xff800000u : v[i];
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), sc, -moff));
          float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), sc, -moff));
          l0 += p0; l1 += p1;
          if constexpr ((F & AF_DROPOUT) != 0) {     // the row sum keeps the undropped probabilities
            const uint32_t key = uint32_t(kv0 + c * 32 + 2 * i);
            const uint32_t bytes = drop_bytes(row_key, key >> 2);
            p0 = drop_is_dropped(bytes, key, p.drop.threshold) ? 0.f : p0;
            p1 = drop_is_dropped(bytes, key + 1, p.drop.threshold) ? 0.f : p1;
          }
          pk[i] = pack_h2<F>(p0, p1);
        }
        tmem_st_32x16(s_addr + c * 16, pk);
      }
      l += l0 + l1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x]);
    }
    // ------------------------------ epilogue ------------------------------------------------------------------
    mbar_wait(&o_done[x], (n_x - 1) & 1);
    tc_fence_after();
    float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    if constexpr ((F & AF_DROPOUT) != 0) inv_l *= p.drop.inv_keep;
    if (row < p.seq) {
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.out_s_stride +
                            (long long)b * p.out_b_stride + (long long)h * D;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(o_addr + c * 32, v);
        tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_h2<F>(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
          o.y = pack_h2<F>(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
          o.z = pack_h2<F>(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
          o.w = pack_h2<F>(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
          dst[i] = o;
        }
      }
      if (p.lse) p.lse[((long long)b * p.heads + h) * p.seq + row] = m_used * 0.6931471805599453f + logf(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

template <int D, int F>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                           cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_fwd2_kernel<D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD2_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  static const int force_single = getenv("MLB200_ATTN_FWD1") != nullptr;
  if (p.seq % (2 * AT_M) == 0 && !force_single) {
    dim3 grid(p.seq / (2 * AT_M), p.heads, p.batch);
    attn_fwd2_kernel<D, F><<<grid, AT2_THREADS, AT_FWD2_SMEM, stream>>>(tq, tk, tv, p);
  } else {
    dim3 grid(p.seq / AT_M, p.heads, p.batch);
    attn_fwd_kernel<D, F><<<grid, AT_THREADS, AT_FWD_SMEM, stream>>>(tq, tk, tv, p);
  }
  return (int)cudaGetLastError();
}

template <int D>
static int launch_attn_fwd_flags(int flags, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                 const AttnParams& p, cudaStream_t stream) {
  switch (flags) {
    case 0: return launch_attn_fwd<D, 0>(tq, tk, tv, p, stream);
    case AF_FP16: return launch_attn_fwd<D, AF_FP16>(tq, tk, tv, p, stream);
    case AF_DROPOUT: return launch_attn_fwd<D, AF_DROPOUT>(tq, tk, tv, p, stream);
    default: return launch_attn_fwd<D, AF_FP16 | AF_DROPOUT>(tq, tk, tv, p, stream);
  }
}

}  // namespace mlb

// q/k/v described by (base pointer, head stride, seq stride, batch stride) in elements + number of heads in the map;
// head coordinates come from the group-stride / offset triple (see AttnParams).  head_dim = 128 or 64.
// ``fp16``: the tensors hold IEEE half instead of bf16.  ``dropout_p`` > 0: attention-probability dropout with the
// counter-based mask of attention_dropout.cuh (the backward must be called with the same p and seed).
extern "C" int mlb_attn_fwd_ex(const void* q, const void* k, const void* v, const long long* q_str,
                               const long long* k_str, const long long* v_str, int q_map_heads, int k_map_heads,
                               int v_map_heads, const int* head_map /* 6 ints */, int q_per_kv, int seq, int batch,
                               int heads, int window, float softmax_scale, void* out, long long out_s_stride,
                               long long out_b_stride, float* lse, int head_dim, int fp16, float dropout_p,
                               unsigned long long seed, cudaStream_t stream) {
  using namespace mlb;
  if (seq % AT_M != 0 || (head_dim != 64 && head_dim != 128) || dropout_p < 0.f || dropout_p >= 1.f) return -2;
  CUtensorMap tq, tk, tv;
  int r = make_tmap_heads(&tq, q, head_dim, q_map_heads, seq, batch, q_str[0], q_str[1], q_str[2], AT_M);
  if (r) return 1000 + r;
  r = make_tmap_heads(&tk, k, head_dim, k_map_heads, seq, batch, k_str[0], k_str[1], k_str[2], AT_N);
  if (r) return 2000 + r;
  r = make_tmap_heads(&tv, v, head_dim, v_map_heads, seq, batch, v_str[0], v_str[1], v_str[2], AT_N);
  if (r) return 3000 + r;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q_group_stride = head_map[0]; p.q_off = head_map[1]; p.k_group_stride = head_map[2]; p.k_off = head_map[3];
  p.v_group_stride = head_map[4]; p.v_off = head_map[5];
  p.q_per_kv = q_per_kv; p.seq = seq; p.batch = batch; p.heads = heads; p.window = window;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.out = out; p.out_s_stride = out_s_stride; p.out_b_stride = out_b_stride; p.lse = lse;
  p.drop = make_dropout_params(dropout_p, seed);
  const int flags = (fp16 ? AF_FP16 : 0) | (p.drop.threshold > 0 ? AF_DROPOUT : 0);
  return head_dim == 128 ? launch_attn_fwd_flags<128>(flags, tq, tk, tv, p, stream)
                         : launch_attn_fwd_flags<64>(flags, tq, tk, tv, p, stream);
}

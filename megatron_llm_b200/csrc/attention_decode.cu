// KV-cache decode attention for sm_100a: a few query positions (the incremental step of text generation) against a
// long key/value cache.  The reference reaches this shape through the FA-2 library (megatron/model/transformer.py:
// 538-553 with inference_params) or the unfused bmm + softmax path (:353-446).
//
// The step is bandwidth bound (every K and V element is read once, ~1 FLOP per byte), so it is a SIMT split-KV kernel,
// not a tensor-core one:
//   * grid (split, kv head, batch): a CTA owns one slice of the cache for ALL query rows that share the kv head
//     (g query heads of the GQA group x sq positions) -- K / V are read once per group, not once per query head;
//   * K tiles (32 keys) are staged through shared memory: coalesced 16-byte global loads (two cache rows per warp
//     instruction), rows padded by 16 bytes so that the per-lane row reads are bank-conflict free;
//   * scores: one key per lane (the query rows are broadcast from shared memory), so a score needs no shuffle;
//   * online softmax per 32-key tile (two warp reductions per row and tile), probabilities parked in shared memory;
//   * P V: the lanes split the head dimension, 32 coalesced V rows per tile;
//   * the four warps of a CTA, then the splits, are merged with the usual (max, sum, weighted accumulator) rule.
// Masking follows flash-attention's bottom-right alignment: query i (of sq) sits at position sk - sq + i.
#include "common.cuh"

namespace mlb {

constexpr int DEC_WARPS = 4;
constexpr int DEC_THREADS = DEC_WARPS * 32;
constexpr int DEC_RC = 8;          // query rows handled per pass over the slice

struct DecodeParams {
  const void* q;                   // [b, sq, nq, D]   strides in elements (D contiguous)
  const void* k;                   // [b, sk, nkv, D]
  const void* v;
  long long q_b, q_s, q_h;
  long long k_b, k_s, k_h;
  long long v_b, v_s, v_h;
  int batch, sq, sk, nq, nkv, g;
  int window;                      // <= 0: none; else keys in [pos - window, pos]
  int n_splits, keys_per_split;    // keys_per_split is a multiple of 32
  float scale_log2;
  float* part_o;                   // [b, nkv, n_splits, R, D]   R = sq * g, row r = qi * g + hg
  float* part_ml;                  // [b, nkv, n_splits, R, 2]   (max in log2 units, sum)
  void* out;                       // [b, sq, nq, D] contiguous
};

template <typename T> struct Pair16;
template <> struct Pair16<__nv_bfloat16> {
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  }
};
template <> struct Pair16<__half> {
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
};

template <int D, typename T>
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_split_kernel(const DecodeParams p) {
  constexpr int E = D / 32;                       // head-dim elements per lane in the P V phase (4 or 2)
  constexpr int CH = D / 8;                       // 16-byte chunks per K row
  constexpr int KROW = D * 2 + 16;                // bytes per staged K row (+16: conflict-free 16-byte column reads)
  __shared__ __align__(16) float sQ[DEC_RC][D];   // query rows of this pass, pre-scaled
  __shared__ float sP[DEC_WARPS][DEC_RC][32];     // probabilities of the warp's current tile
  // per warp: the staged K tile; after the warp's last tile the same bytes carry its accumulators to the merge
  __shared__ __align__(16) uint8_t sK[DEC_WARPS][32 * KROW];
  __shared__ float sM[DEC_WARPS][DEC_RC], sL[DEC_WARPS][DEC_RC];
  static_assert(32 * KROW >= DEC_RC * D * 4, "the accumulator hand-over must fit into the K staging area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int R = p.sq * p.g;
  const int k_begin = split * p.keys_per_split;
  const int k_end = min(p.sk, k_begin + p.keys_per_split);
  const int n_tiles = (k_end > k_begin) ? (k_end - k_begin + 31) / 32 : 0;
  const T* kbase = reinterpret_cast<const T*>(p.k) + (long long)b * p.k_b + (long long)kvh * p.k_h;
  const T* vbase = reinterpret_cast<const T*>(p.v) + (long long)b * p.v_b + (long long)kvh * p.v_h;
  const long long part_row0 = (((long long)b * p.nkv + kvh) * p.n_splits + split) * R;

  for (int r0 = 0; r0 < R; r0 += DEC_RC) {
    const int rc = min(DEC_RC, R - r0);
    // ---- stage the query rows (fp32, softmax scale and log2(e) folded in); rows >= rc are zero
    __syncthreads();                               // previous pass finished with sQ and the merge buffers
    for (int idx = threadIdx.x; idx < DEC_RC * D; idx += DEC_THREADS) {
      const int r = idx / D, d = idx % D;
      float val = 0.f;
      if (r < rc) {
        const int row = r0 + r, qi = row / p.g, hg = row % p.g;
        const T* qp = reinterpret_cast<const T*>(p.q) + (long long)b * p.q_b + (long long)qi * p.q_s +
                      (long long)(kvh * p.g + hg) * p.q_h;
        val = to_f(qp[d]) * p.scale_log2;
      }
      sQ[r][d] = val;
    }
    __syncthreads();

    float m[DEC_RC], l[DEC_RC], acc[DEC_RC][E];
    int qpos[DEC_RC];
#pragma unroll
    for (int r = 0; r < DEC_RC; ++r) {
      m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) acc[r][e] = 0.f;
      qpos[r] = p.sk - p.sq + (r0 + r) / p.g;      // (rows >= rc are never written out)
    }

    for (int t = warp; t < n_tiles; t += DEC_WARPS) {
      const int key0 = k_begin + t * 32;
      const int key = key0 + lane;
      const bool key_ok = key < k_end;
      // ---- stage the K tile (rows past the slice repeat its last key; they are masked below)
      uint8_t* sKw = sK[warp];
#pragma unroll 4
      for (int idx = lane; idx < 32 * CH; idx += 32) {
        const int r = idx / CH, c = idx % CH;
        const T* src = kbase + (long long)min(key0 + r, k_end - 1) * p.k_s + c * 8;
        *reinterpret_cast<uint4*>(sKw + r * KROW + c * 16) = *reinterpret_cast<const uint4*>(src);
      }
      __syncwarp();
      // ---- scores of this lane's key against every row
      float s[DEC_RC];
#pragma unroll
      for (int r = 0; r < DEC_RC; ++r) s[r] = 0.f;
      const uint8_t* krow = sKw + lane * KROW;
#pragma unroll 2
      for (int d0 = 0; d0 < D; d0 += 8) {
        const uint4 kv4 = *reinterpret_cast<const uint4*>(krow + d0 * 2);
        const float2 k01 = Pair16<T>::unpack(kv4.x), k23 = Pair16<T>::unpack(kv4.y);
        const float2 k45 = Pair16<T>::unpack(kv4.z), k67 = Pair16<T>::unpack(kv4.w);
#pragma unroll
        for (int r = 0; r < DEC_RC; ++r) {
          const float4 qa = *reinterpret_cast<const float4*>(&sQ[r][d0]);
          const float4 qb = *reinterpret_cast<const float4*>(&sQ[r][d0 + 4]);
          s[r] += qa.x * k01.x + qa.y * k01.y + qa.z * k23.x + qa.w * k23.y + qb.x * k45.x + qb.y * k45.y +
                  qb.z * k67.x + qb.w * k67.y;
        }
      }
      // ---- online softmax per row (m, l are warp-uniform)
#pragma unroll
      for (int r = 0; r < DEC_RC; ++r) {
        const bool allowed = key_ok && key <= qpos[r] && (p.window <= 0 || key >= qpos[r] - p.window);
        const float sv = allowed ? s[r] : -INFINITY;
        const float m_new = fmaxf(m[r], warp_reduce_max(sv));
        float pe = 0.f, alpha = 1.f;
        if (m_new != -INFINITY) {
          pe = allowed ? exp2f(sv - m_new) : 0.f;
          alpha = (m[r] == -INFINITY) ? 0.f : exp2f(m[r] - m_new);
        }
        l[r] = l[r] * alpha + warp_reduce_sum(pe);
        m[r] = m_new;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[r][e] *= alpha;
        sP[warp][r][lane] = pe;
      }
      __syncwarp();
      // ---- acc += P V : lanes own E consecutive head-dim elements, 32 keys per tile
      const int n_keys = min(32, k_end - key0);
      for (int kk = 0; kk < n_keys; ++kk) {
        const T* vrow = vbase + (long long)(key0 + kk) * p.v_s + lane * E;
        float vf[E];
        if constexpr (E == 4) {
          const uint2 u = *reinterpret_cast<const uint2*>(vrow);
          const float2 a = Pair16<T>::unpack(u.x), c = Pair16<T>::unpack(u.y);
          vf[0] = a.x; vf[1] = a.y; vf[2] = c.x; vf[3] = c.y;
        } else {
          const float2 a = Pair16<T>::unpack(*reinterpret_cast<const uint32_t*>(vrow));
          vf[0] = a.x; vf[1] = a.y;
        }
#pragma unroll
        for (int r = 0; r < DEC_RC; ++r) {
          const float pr = sP[warp][r][kk];
#pragma unroll
          for (int e = 0; e < E; ++e) acc[r][e] += pr * vf[e];
        }
      }
      __syncwarp();                                // sP and the K stage are rewritten by the next tile
    }

    // ---- merge the warps of the CTA and publish the slice's partial result
    float* sAccW = reinterpret_cast<float*>(sK[warp]);          // [DEC_RC][D], this warp's own staging bytes
#pragma unroll
    for (int r = 0; r < DEC_RC; ++r) {
#pragma unroll
      for (int e = 0; e < E; ++e) sAccW[r * D + lane * E + e] = acc[r][e];
      if (lane == 0) { sM[warp][r] = m[r]; sL[warp][r] = l[r]; }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rc * D; idx += DEC_THREADS) {
      const int r = idx / D, d = idx % D;
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < DEC_WARPS; ++w) mm = fmaxf(mm, sM[w][r]);
      float o = 0.f, ll = 0.f;
#pragma unroll
      for (int w = 0; w < DEC_WARPS; ++w) {
        const float wgt = (sM[w][r] == -INFINITY) ? 0.f : exp2f(sM[w][r] - mm);
        o += reinterpret_cast<const float*>(sK[w])[r * D + d] * wgt;
        ll += sL[w][r] * wgt;
      }
      p.part_o[(part_row0 + r0 + r) * D + d] = o;
      if (d == 0) {
        p.part_ml[(part_row0 + r0 + r) * 2 + 0] = mm;
        p.part_ml[(part_row0 + r0 + r) * 2 + 1] = ll;
      }
    }
  }
}

// one warp per (batch, kv head, row): merge the splits, normalise, write the 16-bit output
template <int D, typename T>
__global__ void attn_decode_merge_kernel(const DecodeParams p) {
  constexpr int E = D / 32;
  const int R = p.sq * p.g;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)p.batch * p.nkv * R;
  if (gw >= total) return;
  const int lane = threadIdx.x & 31;
  const int r = (int)(gw % R);
  const long long bk = gw / R;                   // b * nkv + kvh
  const int kvh = (int)(bk % p.nkv), b = (int)(bk / p.nkv);
  float mm = -INFINITY;
  for (int s = 0; s < p.n_splits; ++s) mm = fmaxf(mm, p.part_ml[((bk * p.n_splits + s) * R + r) * 2]);
  float o[E], ll = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) o[e] = 0.f;
  for (int s = 0; s < p.n_splits; ++s) {
    const long long row = (bk * p.n_splits + s) * R + r;
    const float ms = p.part_ml[row * 2];
    const float wgt = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
    ll += p.part_ml[row * 2 + 1] * wgt;
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] += p.part_o[row * D + lane * E + e] * wgt;
  }
  const float inv = ll > 0.f ? 1.f / ll : 0.f;
  const int qi = r / p.g, hg = r % p.g;
  T* dst = reinterpret_cast<T*>(p.out) + (((long long)b * p.sq + qi) * p.nq + kvh * p.g + hg) * D + lane * E;
#pragma unroll
  for (int e = 0; e < E; ++e) dst[e] = from_f<T>(o[e] * inv);
}

// validates the problem and fills the parameter block; strides are (batch, seq, head) in elements
static inline int fill_decode_params(DecodeParams& p, const void* q, const void* k, const void* v, const long long* q_str,
                                     const long long* k_str, const long long* v_str, int batch, int sq, int sk, int nq,
                                     int nkv, int head_dim, int window, float softmax_scale, int n_splits,
                                     int keys_per_split, float* part_o, float* part_ml, void* out) {
  if ((head_dim != 64 && head_dim != 128) || nkv < 1 || nq % nkv != 0 || batch < 1 || sq < 1 || sk < sq ||
      n_splits < 1 || keys_per_split < 32 || keys_per_split % 32 != 0 || (long long)n_splits * keys_per_split < sk)
    return -2;
  p.q = q; p.k = k; p.v = v;
  p.q_b = q_str[0]; p.q_s = q_str[1]; p.q_h = q_str[2];
  p.k_b = k_str[0]; p.k_s = k_str[1]; p.k_h = k_str[2];
  p.v_b = v_str[0]; p.v_s = v_str[1]; p.v_h = v_str[2];
  p.batch = batch; p.sq = sq; p.sk = sk; p.nq = nq; p.nkv = nkv; p.g = nq / nkv;
  p.window = window; p.n_splits = n_splits; p.keys_per_split = keys_per_split;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.part_o = part_o; p.part_ml = part_ml; p.out = out;
  return 0;
}

template <int D, typename T>
static int launch_decode(const DecodeParams& p, cudaStream_t stream) {
  dim3 grid(p.n_splits, p.nkv, p.batch);
  attn_decode_split_kernel<D, T><<<grid, DEC_THREADS, 0, stream>>>(p);
  const long long rows = (long long)p.batch * p.nkv * p.sq * p.g;
  attn_decode_merge_kernel<D, T><<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace mlb

// part_o: fp32 [b * nkv * n_splits * sq * g * hn], part_ml: 2 floats per such row.
extern "C" int mlb_attn_decode(int dtype, const void* q, const void* k, const void* v, const long long* q_str,
                               const long long* k_str, const long long* v_str, int batch, int sq, int sk, int nq,
                               int nkv, int head_dim, int window, float softmax_scale, int n_splits,
                               int keys_per_split, float* part_o, float* part_ml, void* out, cudaStream_t stream) {
  using namespace mlb;
  DecodeParams p;
  const int r = fill_decode_params(p, q, k, v, q_str, k_str, v_str, batch, sq, sk, nq, nkv, head_dim, window,
                                   softmax_scale, n_splits, keys_per_split, part_o, part_ml, out);
  if (r) return r;
  if (dtype == DT_BF16)
    return head_dim == 128 ? launch_decode<128, __nv_bfloat16>(p, stream) : launch_decode<64, __nv_bfloat16>(p, stream);
  if (dtype == DT_F16)
    return head_dim == 128 ? launch_decode<128, __half>(p, stream) : launch_decode<64, __half>(p, stream);
  return -100;
}

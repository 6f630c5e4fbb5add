// Fused vocab-parallel cross-entropy for sm_100a.
//
// Replaces megatron/core/tensor_parallel/cross_entropy.py:14-127 (~15 torch ops, three TP all-reduces, two fp32
// [s,b,V/t] temporaries).  Pass 1 streams each logits row ONCE with an online softmax and emits the four per-token
// statistics (local max, local sum-exp, owned target logit, sum of logits); after one packed cross-rank all-gather,
// pass 2 rewrites the logits buffer in place with d(loss)/d(logits).
#include "common.cuh"

namespace mlb {

template <typename T>
__global__ void __launch_bounds__(256)
ce_stats_kernel(const T* __restrict__ logits, const long long* __restrict__ target, float4* __restrict__ stats,
                int rows, int Vp, int vocab_start, long long row_stride) {
  __shared__ float scratch[32];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = logits + (long long)row * row_stride;
    float m = -INFINITY, s = 0.f, xs = 0.f;
    const int nvec = Vp / 8;
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      Vec<T> a;
      float x[8];
      a.load(xr + vi * 8);
      a.to_float(x);
      float lm = x[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) lm = fmaxf(lm, x[j]);
      const float nm = fmaxf(m, lm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc += __expf(x[j] - nm); xs += x[j]; }
      s = s * __expf(m - nm) + acc;
      m = nm;
    }
    for (int i = nvec * 8 + threadIdx.x; i < Vp; i += blockDim.x) {  // tail (Vp % 8)
      const float x = to_f(xr[i]);
      const float nm = fmaxf(m, x);
      s = s * __expf(m - nm) + __expf(x - nm);
      m = nm;
      xs += x;
    }
    const float M = block_reduce_max(m, scratch);
    const float sc = (m == -INFINITY) ? 0.f : s * __expf(m - M);
    const float S = block_reduce_sum(sc, scratch);
    const float XS = block_reduce_sum(xs, scratch);
    if (threadIdx.x == 0) {
      const long long t = target[row] - vocab_start;
      const float tl = (t >= 0 && t < Vp) ? to_f(xr[t]) : 0.f;
      stats[row] = make_float4(M, S, tl, XS);
    }
  }
}

// grad = (softmax - (1-sm)*onehot - sm/V) * g   written to `out` (may alias logits)
template <typename T>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const T* __restrict__ logits, T* __restrict__ out, const long long* __restrict__ target,
              const float* __restrict__ M, const float* __restrict__ logS, const float* __restrict__ g, int rows,
              int Vp, int vocab_start, float smoothing, int vocab_size, long long row_stride) {
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = logits + (long long)row * row_stride;
    T* orow = out + (long long)row * row_stride;
    const float lse = M[row] + logS[row];
    const float gr = g[row];
    const long long t = target[row] - vocab_start;
    const float sm_u = smoothing / (float)vocab_size;
    const int nvec = Vp / 8;
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      Vec<T> a;
      float x[8], o[8];
      a.load(xr + vi * 8);
      a.to_float(x);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p = __expf(x[j] - lse) - sm_u;
        if ((long long)(vi * 8 + j) == t) p -= (1.f - smoothing);
        o[j] = p * gr;
      }
      a.from_float(o);
      a.store(orow + vi * 8);
    }
    for (int i = nvec * 8 + threadIdx.x; i < Vp; i += blockDim.x) {
      float p = __expf(to_f(xr[i]) - lse) - sm_u;
      if ((long long)i == t) p -= (1.f - smoothing);
      orow[i] = from_f<T>(p * gr);
    }
  }
}

}  // namespace mlb

extern "C" int mlb_ce_stats(int dtype, const void* logits, const long long* target, float* stats, int rows, int Vp,
                            int vocab_start, long long row_stride, cudaStream_t st) {
  if (row_stride % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::ce_stats_kernel<T><<<rows < 148 * 8 ? rows : 148 * 8, 256, 0, st>>>(
                         (const T*)logits, target, (float4*)stats, rows, Vp, vocab_start, row_stride));
  return (int)cudaGetLastError();
}

extern "C" int mlb_ce_bwd(int dtype, const void* logits, void* out, const long long* target, const float* M,
                          const float* logS, const float* g, int rows, int Vp, int vocab_start, float smoothing,
                          int vocab_size, long long row_stride, cudaStream_t st) {
  if (row_stride % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::ce_bwd_kernel<T><<<rows < 148 * 8 ? rows : 148 * 8, 256, 0, st>>>(
                         (const T*)logits, (T*)out, target, M, logS, g, rows, Vp, vocab_start, smoothing, vocab_size,
                         row_stride));
  return (int)cudaGetLastError();
}

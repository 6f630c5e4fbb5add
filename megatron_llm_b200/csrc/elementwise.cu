// Element-wise hot ops for sm_100a: RoPE on the packed QKV buffer, GLU family (swiglu/geglu/reglu/liglu),
// GeLU / bias-GeLU, bias+dropout+add.  All are pure bandwidth kernels: 16-byte vectors, one pass.
//
// Replaces: positional_embeddings.py:24-51 (complex fp32 multiply + 2 casts + gather),
// glu_activations.py:7-15 (chunk + act + mul), fused_bias_gelu.py (nvFuser JIT, gone in torch 2.11),
// transformer.py:563-609 (bias-dropout-add).
#include "common.cuh"

namespace mlb {

// ------------------------------------------------------------------------------------------------
// RoPE (interleaved-pair / Meta convention), in place on qkv laid out [tokens, n_groups, (q_per_kv+2), hn]
// where the last two "heads" of every group are K and V (reference layout, transformer.py:458-461).
// Q heads and the K head are rotated, V is untouched.  token t -> (s = t / b, bi = t % b) for [s,b] order.
// freqs: float2 (cos, sin) table [max_pos, hn/2].  inverse=1 applies the transpose (backward).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void rope_qkv_kernel(T* __restrict__ qkv, const float2* __restrict__ freqs,
                                const long long* __restrict__ position_ids, int tokens, int batch, int n_groups,
                                int heads_per_group /* q_per_kv + 2 */, int hn, int pos_offset, int inverse,
                                long long token_stride) {
  const int vec_per_head = hn / 8;
  const int rot_heads = heads_per_group - 1;  // q heads + k
  const long long total = (long long)tokens * n_groups * rot_heads * vec_per_head;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_head);
    long long r = i / vec_per_head;
    const int hh = (int)(r % rot_heads);
    r /= rot_heads;
    const int g = (int)(r % n_groups);
    const int t = (int)(r / n_groups);
    const int s = t / batch, bi = t % batch;
    const long long pos = position_ids ? position_ids[(long long)bi * (tokens / batch) + s] : (long long)(s + pos_offset);
    T* p = qkv + (long long)t * token_stride + ((long long)g * heads_per_group + hh) * hn + v * 8;
    Vec<T> a;
    float x[8], o[8];
    a.load(p);
    a.to_float(x);
    const float2* f = freqs + pos * (hn / 2) + v * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 cs = f[j];
      const float sn = inverse ? -cs.y : cs.y;
      o[2 * j] = x[2 * j] * cs.x - x[2 * j + 1] * sn;
      o[2 * j + 1] = x[2 * j] * sn + x[2 * j + 1] * cs.x;
    }
    a.from_float(o);
    a.store(p);
  }
}

// generic variant: x [tokens, heads, hn] contiguous (used for separate q / k tensors and the KV-cache path)
template <typename T>
__global__ void rope_heads_kernel(T* __restrict__ x, const float2* __restrict__ freqs,
                                  const long long* __restrict__ position_ids, int tokens, int batch, int heads, int hn,
                                  int pos_offset, int inverse, long long token_stride, long long head_stride) {
  const int vec_per_head = hn / 8;
  const long long total = (long long)tokens * heads * vec_per_head;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_head);
    long long r = i / vec_per_head;
    const int h = (int)(r % heads);
    const int t = (int)(r / heads);
    const int s = t / batch, bi = t % batch;
    const long long pos = position_ids ? position_ids[(long long)bi * (tokens / batch) + s] : (long long)(s + pos_offset);
    T* p = x + (long long)t * token_stride + (long long)h * head_stride + v * 8;
    Vec<T> a;
    float xi[8], o[8];
    a.load(p);
    a.to_float(xi);
    const float2* f = freqs + pos * (hn / 2) + v * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 cs = f[j];
      const float sn = inverse ? -cs.y : cs.y;
      o[2 * j] = xi[2 * j] * cs.x - xi[2 * j + 1] * sn;
      o[2 * j + 1] = xi[2 * j] * sn + xi[2 * j + 1] * cs.x;
    }
    a.from_float(o);
    a.store(p);
  }
}

// ------------------------------------------------------------------------------------------------
// GLU family: x [rows, 2F] = [x1 | x2];  y = x1 * act(x2)   (reference: first half = up, second = gate)
// kind: 0 liglu (identity) 1 geglu (erf gelu) 2 reglu 3 swiglu
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int kind, float z) {
  switch (kind) {
    case 0: return z;
    case 1: return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
    case 2: return fmaxf(z, 0.f);
    default: return z / (1.f + __expf(-z));
  }
}
__device__ __forceinline__ float act_bwd(int kind, float z) {
  switch (kind) {
    case 0: return 1.f;
    case 1: {
      const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case 2: return z > 0.f ? 1.f : 0.f;
    default: {
      const float s = 1.f / (1.f + __expf(-z));
      return s * (1.f + z * (1.f - s));
    }
  }
}

template <typename T>
__global__ void glu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long rows, int F, int kind) {
  const int vpr = F / 8;
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int v = (int)(i % vpr);
    Vec<T> a, b;
    float x1[8], x2[8], o[8];
    a.load(x + r * 2 * F + v * 8);
    b.load(x + r * 2 * F + F + v * 8);
    a.to_float(x1);
    b.to_float(x2);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = x1[j] * act_fwd(kind, x2[j]);
    Vec<T> ov;
    ov.from_float(o);
    ov.store(y + r * F + v * 8);
  }
}

template <typename T>
__global__ void glu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, long long rows,
                               int F, int kind) {
  const int vpr = F / 8;
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int v = (int)(i % vpr);
    Vec<T> a, b, d;
    float x1[8], x2[8], g[8], o1[8], o2[8];
    a.load(x + r * 2 * F + v * 8);
    b.load(x + r * 2 * F + F + v * 8);
    d.load(dy + r * F + v * 8);
    a.to_float(x1);
    b.to_float(x2);
    d.to_float(g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o1[j] = g[j] * act_fwd(kind, x2[j]);
      o2[j] = g[j] * x1[j] * act_bwd(kind, x2[j]);
    }
    Vec<T> ov;
    ov.from_float(o1);
    ov.store(dx + r * 2 * F + v * 8);
    ov.from_float(o2);
    ov.store(dx + r * 2 * F + F + v * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// GeLU: y = gelu(x + bias); approx=1 -> tanh form (bias_gelu of the reference), approx=0 -> erf (F.gelu)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float z, int approx) {
  if (approx) return 0.5f * z * (1.f + tanhf(0.79788456f * z * (1.f + 0.044715f * z * z)));
  return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_df(float z, int approx) {
  if (approx) {
    const float t = tanhf(0.79788456f * z * (1.f + 0.044715f * z * z));
    return 0.5f * z * ((1.f - t * t) * (0.79788456f + 0.1070322243f * z * z)) + 0.5f * (1.f + t);
  }
  const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
  return cdf + z * 0.39894228040143267794f * __expf(-0.5f * z * z);
}

template <typename T, bool BWD>
__global__ void gelu_kernel(const T* __restrict__ x, const T* __restrict__ bias, const T* __restrict__ dy,
                            T* __restrict__ out, long long rows, int F, int approx) {
  const int vpr = F / 8;
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    Vec<T> a;
    float xf[8], bf[8], o[8];
    a.load(x + i * 8);
    a.to_float(xf);
    if (bias) {
      Vec<T> bv;
      bv.load(bias + v * 8);
      bv.to_float(bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) xf[j] += bf[j];
    }
    if constexpr (BWD) {
      Vec<T> d;
      float g[8];
      d.load(dy + i * 8);
      d.to_float(g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j] * gelu_df(xf[j], approx);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = gelu_f(xf[j], approx);
    }
    Vec<T> ov;
    ov.from_float(o);
    ov.store(out + i * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// out = residual + dropout(x + bias, p)   (Philox counter RNG; mask regenerated in backward from seed)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t idx) {
  // splitmix64-style counter hash; statistically adequate for dropout masks and replayable
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

template <typename T, bool BWD>
__global__ void bias_dropout_add_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                        const T* __restrict__ residual, T* __restrict__ out, long long rows, int F,
                                        float p, unsigned long long seed) {
  const int vpr = F / 8;
  const long long total = rows * vpr;
  const float scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  const uint32_t thresh = (uint32_t)(p * 4294967295.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    Vec<T> a;
    float xf[8], o[8];
    a.load(x + i * 8);
    a.to_float(xf);
    if (!BWD && bias) {
      Vec<T> bv;
      float bf[8];
      bv.load(bias + v * 8);
      bv.to_float(bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) xf[j] += bf[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool keep = (p <= 0.f) || (mix32(seed, (uint64_t)i * 8 + j) >= thresh);
      o[j] = keep ? xf[j] * scale : 0.f;
      if (p <= 0.f) o[j] = xf[j];
    }
    if (!BWD && residual) {
      Vec<T> rv;
      float rf[8];
      rv.load(residual + i * 8);
      rv.to_float(rf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rf[j];
    }
    Vec<T> ov;
    ov.from_float(o);
    ov.store(out + i * 8);
  }
}

static inline int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace mlb

extern "C" int mlb_rope_qkv(int dtype, void* qkv, const void* freqs, const long long* position_ids, int tokens,
                            int batch, int n_groups, int heads_per_group, int hn, int pos_offset, int inverse,
                            long long token_stride, cudaStream_t st) {
  if (hn % 8) return -2;
  const long long total = (long long)tokens * n_groups * (heads_per_group - 1) * (hn / 8);
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::rope_qkv_kernel<T><<<mlb::grid_for(total, 256), 256, 0, st>>>(
                         (T*)qkv, (const float2*)freqs, position_ids, tokens, batch, n_groups, heads_per_group, hn,
                         pos_offset, inverse, token_stride));
  return (int)cudaGetLastError();
}

extern "C" int mlb_rope_heads(int dtype, void* x, const void* freqs, const long long* position_ids, int tokens,
                              int batch, int heads, int hn, int pos_offset, int inverse, long long token_stride,
                              long long head_stride, cudaStream_t st) {
  if (hn % 8) return -2;
  const long long total = (long long)tokens * heads * (hn / 8);
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::rope_heads_kernel<T><<<mlb::grid_for(total, 256), 256, 0, st>>>(
                         (T*)x, (const float2*)freqs, position_ids, tokens, batch, heads, hn, pos_offset, inverse,
                         token_stride, head_stride));
  return (int)cudaGetLastError();
}

extern "C" int mlb_glu_fwd(int dtype, const void* x, void* y, long long rows, int F, int kind, cudaStream_t st) {
  if (F % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::glu_fwd_kernel<T><<<mlb::grid_for(rows * (F / 8), 256), 256, 0, st>>>((const T*)x, (T*)y,
                                                                                             rows, F, kind));
  return (int)cudaGetLastError();
}

extern "C" int mlb_glu_bwd(int dtype, const void* dy, const void* x, void* dx, long long rows, int F, int kind,
                           cudaStream_t st) {
  if (F % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::glu_bwd_kernel<T><<<mlb::grid_for(rows * (F / 8), 256), 256, 0, st>>>(
                         (const T*)dy, (const T*)x, (T*)dx, rows, F, kind));
  return (int)cudaGetLastError();
}

extern "C" int mlb_gelu(int dtype, const void* x, const void* bias, const void* dy, void* out, long long rows, int F,
                        int approx, int backward, cudaStream_t st) {
  if (F % 8) return -2;
  const int grid = mlb::grid_for(rows * (F / 8), 256);
  MLB_DISPATCH_DTYPE(dtype, T, {
    if (backward) mlb::gelu_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, (const T*)bias, (const T*)dy, (T*)out, rows, F, approx);
    else mlb::gelu_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, (const T*)bias, (const T*)dy, (T*)out, rows, F, approx);
  });
  return (int)cudaGetLastError();
}

extern "C" int mlb_bias_dropout_add(int dtype, const void* x, const void* bias, const void* residual, void* out,
                                    long long rows, int F, float p, unsigned long long seed, int backward,
                                    cudaStream_t st) {
  if (F % 8) return -2;
  const int grid = mlb::grid_for(rows * (F / 8), 256);
  MLB_DISPATCH_DTYPE(dtype, T, {
    if (backward) mlb::bias_dropout_add_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, (const T*)bias, (const T*)residual, (T*)out, rows, F, p, seed);
    else mlb::bias_dropout_add_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, (const T*)bias, (const T*)residual, (T*)out, rows, F, p, seed);
  });
  return (int)cudaGetLastError();
}

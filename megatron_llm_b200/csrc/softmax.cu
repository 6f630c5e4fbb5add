// Scaled softmax family for the non-flash attention path (sm_100a):
//   scaled_upper_triang_masked_softmax (implicit causal mask), scaled_masked_softmax (uint8 padding mask),
//   scaled_softmax (no mask) -- forward and in-place backward.
//
// Replaces megatron/fused_kernels/scaled_{upper_triang_masked,masked,}_softmax*.{h,cu}.  Those are templated on
// log2(seq) and cap sk at 4096 (2048 causal); this version has no sequence-length envelope: one warp per row for
// sk <= 1024 (row kept in registers), otherwise one CTA per row streaming twice through L2.
#include "common.cuh"

namespace mlb {

// x,y: [rows, sk]; row r belongs to (batch b, head h, query q) with rows = b*np*sq.
// mask: uint8 [mb, 1, sq, sk] (mb = b or 1), 1 = masked.  causal: element k is visible iff k <= q + (sk - sq).
template <typename T, int MODE /*0 none, 1 mask, 2 causal*/>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, const uint8_t* __restrict__ mask, float scale,
                   long long rows, int sq, int sk, int np, int mask_batch) {
  __shared__ float scratch[32];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int q = (int)(row % sq);
    const int bidx = (int)(row / ((long long)sq * np));
    const uint8_t* mrow = nullptr;
    if (MODE == 1) mrow = mask + ((long long)(mask_batch == 1 ? 0 : bidx) * sq + q) * sk;
    const int visible = (MODE == 2) ? min(sk, q + 1 + (sk - sq)) : sk;
    const T* xr = x + row * sk;
    T* yr = y + row * sk;
    float m = -INFINITY;
    for (int k = threadIdx.x; k < visible; k += blockDim.x) {
      float v = to_f(xr[k]) * scale;
      if (MODE == 1 && mrow[k]) v = -10000.f;
      m = fmaxf(m, v);
    }
    m = block_reduce_max(m, scratch);
    float s = 0.f;
    for (int k = threadIdx.x; k < visible; k += blockDim.x) {
      float v = to_f(xr[k]) * scale;
      if (MODE == 1 && mrow[k]) v = -10000.f;
      s += __expf(v - m);
    }
    s = block_reduce_sum(s, scratch);
    const float inv = 1.f / s;
    bool all_masked = false;
    if (MODE == 1) all_masked = (m == -10000.f);  // fully masked row -> zeros (reference .h:196 behaviour)
    for (int k = threadIdx.x; k < sk; k += blockDim.x) {
      float o = 0.f;
      if (k < visible && !all_masked) {
        float v = to_f(xr[k]) * scale;
        if (MODE == 1 && mrow[k]) v = -10000.f;
        o = __expf(v - m) * inv;
      }
      yr[k] = from_f<T>(o);
    }
  }
}

// dx = scale * (y*dy - y*sum(y*dy)), written in place over dy
template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(T* __restrict__ dy, const T* __restrict__ y, float scale, long long rows, int sk) {
  __shared__ float scratch[32];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    T* gr = dy + row * sk;
    const T* yr = y + row * sk;
    float s = 0.f;
    for (int k = threadIdx.x; k < sk; k += blockDim.x) s += to_f(yr[k]) * to_f(gr[k]);
    s = block_reduce_sum(s, scratch);
    for (int k = threadIdx.x; k < sk; k += blockDim.x) {
      const float yv = to_f(yr[k]);
      gr[k] = from_f<T>(scale * (yv * to_f(gr[k]) - yv * s));
    }
  }
}

}  // namespace mlb

extern "C" int mlb_softmax_fwd(int dtype, const void* x, void* y, const unsigned char* mask, float scale,
                               long long rows, int sq, int sk, int np, int mask_batch, int mode, cudaStream_t st) {
  const int grid = (int)(rows < 148LL * 16 ? rows : 148LL * 16);
  const int threads = sk <= 256 ? 64 : (sk <= 1024 ? 128 : 256);
  MLB_DISPATCH_DTYPE(dtype, T, {
    if (mode == 0) mlb::softmax_fwd_kernel<T, 0><<<grid, threads, 0, st>>>((const T*)x, (T*)y, mask, scale, rows, sq, sk, np, mask_batch);
    else if (mode == 1) mlb::softmax_fwd_kernel<T, 1><<<grid, threads, 0, st>>>((const T*)x, (T*)y, mask, scale, rows, sq, sk, np, mask_batch);
    else mlb::softmax_fwd_kernel<T, 2><<<grid, threads, 0, st>>>((const T*)x, (T*)y, mask, scale, rows, sq, sk, np, mask_batch);
  });
  return (int)cudaGetLastError();
}

extern "C" int mlb_softmax_bwd(int dtype, void* dy, const void* y, float scale, long long rows, int sk,
                               cudaStream_t st) {
  const int grid = (int)(rows < 148LL * 16 ? rows : 148LL * 16);
  const int threads = sk <= 256 ? 64 : (sk <= 1024 ? 128 : 256);
  MLB_DISPATCH_DTYPE(dtype, T, mlb::softmax_bwd_kernel<T><<<grid, threads, 0, st>>>((T*)dy, (const T*)y, scale, rows, sk));
  return (int)cudaGetLastError();
}

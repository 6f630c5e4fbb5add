// Vocab-parallel embedding lookup and its backward (sm_100a, memory-bound SIMT kernels).
//
// Reference: megatron/core/tensor_parallel/layers.py:187-210 builds a mask, subtracts the vocab offset, calls
// F.embedding, zeroes the masked rows (4 elementwise kernels) and transposes [b, s, h] -> [s, b, h] afterwards; its
// backward goes through torch's sort-based dense embedding gradient.  Here:
//   forward : one gather kernel that reads ids [b, s], writes the [s, b, h] layout directly and emits zeros for ids
//             that live on another tensor-parallel rank (so the partial results can be reduce-scattered / all-reduced);
//   backward: one scatter kernel that adds every token's output gradient row into the fp32 ``main_grad`` of the table
//             (red.global.add.v4.f32: the read-modify-write happens in L2) -- no [vocab, h] temporary, no sort.
#include "common.cuh"

namespace mlb {

// one warp per token row: H / 8 vectors of 16 bytes
template <typename T>
__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const long long* __restrict__ ids, const T* __restrict__ weight, T* __restrict__ out,
                     int batch, int seq, int H, long long vocab_start, long long rows_local, int sbh) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_tokens = (long long)batch * seq;
  const int vpr = H / 8;
  for (long long t = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < n_tokens;
       t += (long long)gridDim.x * warps_per_block) {
    // t indexes the OUTPUT row: [s, b] order when sbh, else [b, s]
    long long src_tok = t;
    if (sbh) {
      const long long s = t / batch, b = t - s * batch;
      src_tok = b * seq + s;
    }
    const long long id = ids[src_tok] - vocab_start;
    const bool mine = id >= 0 && id < rows_local;
    T* orow = out + t * H;
    if (mine) {
      const T* wrow = weight + id * H;
      for (int v = lane; v < vpr; v += 32) {
        Vec<T> x;
        x.load(wrow + v * 8);
        x.store(orow + v * 8);
      }
    } else {
      Vec<T> z;
      const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      z.from_float(zero);
      for (int v = lane; v < vpr; v += 32) z.store(orow + v * 8);
    }
  }
}

__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const long long* __restrict__ ids, const T* __restrict__ dout, float* __restrict__ dweight,
                     int batch, int seq, int H, long long vocab_start, long long rows_local, int sbh) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_tokens = (long long)batch * seq;
  const int vpr = H / 8;
  for (long long t = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < n_tokens;
       t += (long long)gridDim.x * warps_per_block) {
    long long src_tok = t;
    if (sbh) {
      const long long s = t / batch, b = t - s * batch;
      src_tok = b * seq + s;
    }
    const long long id = ids[src_tok] - vocab_start;
    if (id < 0 || id >= rows_local) continue;          // warp-uniform
    const T* grow = dout + t * H;
    float* wrow = dweight + id * H;
    for (int v = lane; v < vpr; v += 32) {
      Vec<T> g;
      float f[8];
      g.load(grow + v * 8);
      g.to_float(f);
      red_add_v4_f32(wrow + v * 8, f[0], f[1], f[2], f[3]);
      red_add_v4_f32(wrow + v * 8 + 4, f[4], f[5], f[6], f[7]);
    }
  }
}

}  // namespace mlb

extern "C" int mlb_embedding_fwd(int dtype, const long long* ids, const void* weight, void* out, int batch, int seq,
                                 int H, long long vocab_start, long long rows_local, int sbh, cudaStream_t st) {
  if (H % 8) return -2;
  const long long n = (long long)batch * seq;
  const int grid = (int)((n + 7) / 8 < 148 * 8 ? (n + 7) / 8 : 148 * 8);
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::embedding_fwd_kernel<T><<<grid > 0 ? grid : 1, 256, 0, st>>>(
                         ids, (const T*)weight, (T*)out, batch, seq, H, vocab_start, rows_local, sbh));
  return (int)cudaGetLastError();
}

extern "C" int mlb_embedding_bwd(int dtype, const long long* ids, const void* dout, float* dweight, int batch, int seq,
                                 int H, long long vocab_start, long long rows_local, int sbh, cudaStream_t st) {
  if (H % 8) return -2;
  const long long n = (long long)batch * seq;
  const int grid = (int)((n + 7) / 8 < 148 * 8 ? (n + 7) / 8 : 148 * 8);
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::embedding_bwd_kernel<T><<<grid > 0 ? grid : 1, 256, 0, st>>>(
                         ids, (const T*)dout, dweight, batch, seq, H, vocab_start, rows_local, sbh));
  return (int)cudaGetLastError();
}

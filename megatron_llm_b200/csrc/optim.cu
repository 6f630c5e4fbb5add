// Optimizer kernels over FLAT contiguous buffers (sm_100a).
//
// The reference drives apex multi-tensor kernels over Python lists of per-parameter tensors (FusedAdam,
// amp_C.multi_tensor_l2norm / multi_tensor_scale; optimizer/__init__.py:75-85, clip_grads.py:75-105) with host
// syncs for the grad-norm and inf checks.  Here the fp32 master params, Adam moments, fp32 main_grads and the
// bf16 model weights each live in ONE contiguous buffer, so:
//   * grad-norm  = one streaming reduction (per-segment weights drop TP-duplicated params),
//   * AdamW      = one streaming kernel that reads g/p/m/v once, applies the clip coefficient read from DEVICE
//                  memory (no host sync), and writes p/m/v plus the bf16 model copy,
//   * a `skip` flag in device memory turns the step into a no-op on inf/nan (fp16 loss scaling).
#include "common.cuh"

namespace mlb {

// Segment table: param k occupies [seg_start[k], seg_start[k+1]) of the flat buffer (sorted, nseg+1 entries).
__device__ __forceinline__ int find_segment(const long long* __restrict__ seg_start, int nseg, long long idx) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_start[mid] <= idx) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int OPT_CHUNK = 2048;  // elements per CTA iteration (256 threads x 8)

// ZeRO-1: the updated 16-bit weights of this rank's shard are stored into EVERY data-parallel peer's parameter buffer
// (NVLink-mapped symmetric memory) by the optimizer kernel itself: the fp32 -> bf16 cast and the parameter all-gather
// (reference: optimizer/distrib_optimizer.py:592-608, one NCCL all-gather + per-tensor copies) are the same pass.
struct P16Peers {
  void* ptr[8];   // peer d's buffer at the start of this shard (own buffer included)
  int n;          // 0: write only `p16`
};

template <typename TP16>
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  TP16* __restrict__ p16, long long n, long long global_offset,
                  const long long* __restrict__ seg_start, const float* __restrict__ seg_wd,
                  const float* __restrict__ seg_lr_mult, int nseg, float lr, float beta1, float beta2, float eps,
                  float bc1, float bc2, const float* __restrict__ grad_scale_ptr, const int* __restrict__ skip_flag,
                  const P16Peers peers) {
  if (skip_flag != nullptr && *skip_flag != 0) return;
  const float gscale = grad_scale_ptr ? *grad_scale_ptr : 1.f;
  const long long nchunks = (n + OPT_CHUNK - 1) / OPT_CHUNK;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long long base = c * OPT_CHUNK + (long long)threadIdx.x * 8;
    if (base >= n) continue;
    int seg = find_segment(seg_start, nseg, base + global_offset);
    long long seg_end = seg_start[seg + 1] - global_offset;
    float wd = seg_wd[seg], lrm = seg_lr_mult ? seg_lr_mult[seg] : 1.f;
    float pv[8], gv[8], mv[8], vv[8];
    const bool full = base + 8 <= n;
    if (full) {
      *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + base);
      *reinterpret_cast<float4*>(pv + 4) = *reinterpret_cast<const float4*>(p + base + 4);
      *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + base);
      *reinterpret_cast<float4*>(gv + 4) = *reinterpret_cast<const float4*>(g + base + 4);
      *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + base);
      *reinterpret_cast<float4*>(mv + 4) = *reinterpret_cast<const float4*>(m + base + 4);
      *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + base);
      *reinterpret_cast<float4*>(vv + 4) = *reinterpret_cast<const float4*>(v + base + 4);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = base + j < n;
        pv[j] = ok ? p[base + j] : 0.f; gv[j] = ok ? g[base + j] : 0.f;
        mv[j] = ok ? m[base + j] : 0.f; vv[j] = ok ? v[base + j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (base + j >= seg_end) {  // crossed into the next parameter
        seg = find_segment(seg_start, nseg, base + j + global_offset);
        seg_end = seg_start[seg + 1] - global_offset;
        wd = seg_wd[seg];
        lrm = seg_lr_mult ? seg_lr_mult[seg] : 1.f;
      }
      const float gg = gv[j] * gscale;
      mv[j] = beta1 * mv[j] + (1.f - beta1) * gg;
      vv[j] = beta2 * vv[j] + (1.f - beta2) * gg * gg;
      const float mhat = mv[j] / bc1;
      const float vhat = vv[j] / bc2;
      const float upd = mhat / (sqrtf(vhat) + eps) + wd * pv[j];
      pv[j] -= lr * lrm * upd;
    }
    if (full) {
      *reinterpret_cast<float4*>(p + base) = *reinterpret_cast<float4*>(pv);
      *reinterpret_cast<float4*>(p + base + 4) = *reinterpret_cast<float4*>(pv + 4);
      *reinterpret_cast<float4*>(m + base) = *reinterpret_cast<float4*>(mv);
      *reinterpret_cast<float4*>(m + base + 4) = *reinterpret_cast<float4*>(mv + 4);
      *reinterpret_cast<float4*>(v + base) = *reinterpret_cast<float4*>(vv);
      *reinterpret_cast<float4*>(v + base + 4) = *reinterpret_cast<float4*>(vv + 4);
      if (peers.n > 0) {
        Vec<TP16> o;
        o.from_float(pv);
#pragma unroll
        for (int d = 0; d < 8; ++d)          // (constant indices: the by-value table stays in the parameter bank)
          if (d < peers.n) o.store(reinterpret_cast<TP16*>(peers.ptr[d]) + base);
      } else if (p16 != nullptr) {
        Vec<TP16> o;
        o.from_float(pv);
        o.store(p16 + base);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j < n) {
          p[base + j] = pv[j]; m[base + j] = mv[j]; v[base + j] = vv[j];
          if (peers.n > 0) {
#pragma unroll
            for (int d = 0; d < 8; ++d)
              if (d < peers.n) reinterpret_cast<TP16*>(peers.ptr[d])[base + j] = from_f<TP16>(pv[j]);
          } else if (p16 != nullptr) {
            p16[base + j] = from_f<TP16>(pv[j]);
          }
        }
    }
  }
}

// SGD with momentum over the same flat layout
template <typename TP16>
__global__ void __launch_bounds__(256)
sgd_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, TP16* __restrict__ p16,
                long long n, long long global_offset, const long long* __restrict__ seg_start,
                const float* __restrict__ seg_wd, const float* __restrict__ seg_lr_mult, int nseg, float lr,
                float momentum, int first_step, const float* __restrict__ grad_scale_ptr,
                const int* __restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0) return;
  const float gscale = grad_scale_ptr ? *grad_scale_ptr : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int seg = find_segment(seg_start, nseg, i + global_offset);
    float gg = g[i] * gscale + seg_wd[seg] * p[i];
    float b = first_step ? gg : momentum * mom[i] + gg;
    if (momentum != 0.f) { mom[i] = b; gg = b; }
    const float np_ = p[i] - lr * (seg_lr_mult ? seg_lr_mult[seg] : 1.f) * gg;
    p[i] = np_;
    if (p16 != nullptr) p16[i] = from_f<TP16>(np_);
  }
}

// partial[b] = sum over this CTA's elements of w_seg * x^2
template <typename T>
__global__ void __launch_bounds__(256)
sqnorm_flat_kernel(const T* __restrict__ x, long long n, long long global_offset,
                   const long long* __restrict__ seg_start, const float* __restrict__ seg_weight, int nseg,
                   float* __restrict__ partial) {
  __shared__ float scratch[32];
  float acc = 0.f;
  const long long nchunks = (n + OPT_CHUNK - 1) / OPT_CHUNK;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long long base = c * OPT_CHUNK + (long long)threadIdx.x * 8;
    if (base >= n) continue;
    int seg = 0;
    long long seg_end = n + 1;
    float w = 1.f;
    if (seg_start != nullptr) {
      seg = find_segment(seg_start, nseg, base + global_offset);
      seg_end = seg_start[seg + 1] - global_offset;
      w = seg_weight[seg];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (base + j < n) {
        if (base + j >= seg_end) {
          seg = find_segment(seg_start, nseg, base + j + global_offset);
          seg_end = seg_start[seg + 1] - global_offset;
          w = seg_weight[seg];
        }
        const float xv = to_f(x[base + j]);
        acc += w * xv * xv;
      }
    }
  }
  const float tot = block_reduce_sum(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void finalize_sum_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out,
                                    int accumulate) {
  __shared__ float scratch[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += partial[i];
  const float t = block_reduce_sum(a, scratch);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + t : t;
}

// nrm = sqrt(total_sq) * norm_scale (norm_scale = 1/loss_scale for fp16, else 1)
// coef_out = min(1, max_norm / (nrm + 1e-6)) * norm_scale  -> the single multiplier the optimizer applies to grads
__global__ void clip_coef_kernel(const float* __restrict__ total_sq, float max_norm, float* __restrict__ norm_out,
                                 float* __restrict__ coef_out, int* __restrict__ found_inf, float norm_scale) {
  const float nrm = sqrtf(total_sq[0]) * norm_scale;
  norm_out[0] = nrm;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (nrm + 1.0e-6f));
  coef_out[0] = coef * norm_scale;
  if (found_inf) found_inf[0] = isfinite(nrm) ? 0 : 1;
}

template <typename TI, typename TO>
__global__ void scale_cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n, float scale,
                                  const float* __restrict__ scale_ptr) {
  const float s = scale * (scale_ptr ? *scale_ptr : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(to_f(x[i]) * s);
}

// y(fp32) += x(T)   (DDP grad accumulation hook: main_grad += param.grad)
template <typename T>
__global__ void accumulate_kernel(const T* __restrict__ x, float* __restrict__ y, long long n) {
  const long long nv = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    Vec<T> a;
    float f[8];
    a.load(x + i * 8);
    a.to_float(f);
    float4* yp = reinterpret_cast<float4*>(y + i * 8);
    float4 y0 = yp[0], y1 = yp[1];
    y0.x += f[0]; y0.y += f[1]; y0.z += f[2]; y0.w += f[3];
    y1.x += f[4]; y1.y += f[5]; y1.z += f[6]; y1.w += f[7];
    yp[0] = y0; yp[1] = y1;
  }
  for (long long i = nv * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += to_f(x[i]);
}

static inline int opt_grid(long long n) {
  long long g = (n + OPT_CHUNK - 1) / OPT_CHUNK;
  const long long cap = 148LL * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace mlb

extern "C" int mlb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, int p16_dtype, long long n,
                              long long global_offset, const long long* seg_start, const float* seg_wd,
                              const float* seg_lr_mult, int nseg, float lr, float beta1, float beta2, float eps,
                              float bc1, float bc2, const float* grad_scale_ptr, const int* skip_flag,
                              const long long* p16_peers, int n_peers, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n_peers > 8) return -2;
  mlb::P16Peers peers;
  peers.n = n_peers;
  for (int i = 0; i < 8; ++i) peers.ptr[i] = i < n_peers ? reinterpret_cast<void*>(p16_peers[i]) : nullptr;
  const int grid = mlb::opt_grid(n);
  if (p16 == nullptr || p16_dtype == mlb::DT_BF16)
    mlb::adamw_flat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)p16, n, global_offset, seg_start, seg_wd, seg_lr_mult, nseg, lr, beta1, beta2, eps, bc1, bc2, grad_scale_ptr, skip_flag, peers);
  else if (p16_dtype == mlb::DT_F16)
    mlb::adamw_flat_kernel<__half><<<grid, 256, 0, st>>>(p, g, m, v, (__half*)p16, n, global_offset, seg_start, seg_wd, seg_lr_mult, nseg, lr, beta1, beta2, eps, bc1, bc2, grad_scale_ptr, skip_flag, peers);
  else return -100;
  return (int)cudaGetLastError();
}

extern "C" int mlb_sgd_flat(float* p, const float* g, float* mom, void* p16, int p16_dtype, long long n,
                            long long global_offset, const long long* seg_start, const float* seg_wd,
                            const float* seg_lr_mult, int nseg, float lr, float momentum, int first_step,
                            const float* grad_scale_ptr, const int* skip_flag, cudaStream_t st) {
  if (n <= 0) return 0;
  const int grid = mlb::opt_grid(n);
  if (p16 == nullptr || p16_dtype == mlb::DT_BF16)
    mlb::sgd_flat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, g, mom, (__nv_bfloat16*)p16, n, global_offset, seg_start, seg_wd, seg_lr_mult, nseg, lr, momentum, first_step, grad_scale_ptr, skip_flag);
  else if (p16_dtype == mlb::DT_F16)
    mlb::sgd_flat_kernel<__half><<<grid, 256, 0, st>>>(p, g, mom, (__half*)p16, n, global_offset, seg_start, seg_wd, seg_lr_mult, nseg, lr, momentum, first_step, grad_scale_ptr, skip_flag);
  else return -100;
  return (int)cudaGetLastError();
}

// workspace: >= 148*8 floats.  out[0] (+)= sum w * x^2
extern "C" int mlb_sqnorm_flat(int dtype, const void* x, long long n, long long global_offset,
                               const long long* seg_start, const float* seg_weight, int nseg, float* workspace,
                               float* out, int accumulate, cudaStream_t st) {
  const int grid = mlb::opt_grid(n > 0 ? n : 1);
  MLB_DISPATCH_DTYPE(dtype, T,
                     mlb::sqnorm_flat_kernel<T><<<grid, 256, 0, st>>>((const T*)x, n, global_offset, seg_start,
                                                                     seg_weight, nseg, workspace));
  mlb::finalize_sum_kernel<<<1, 256, 0, st>>>(workspace, grid, out, accumulate);
  return (int)cudaGetLastError();
}

extern "C" int mlb_clip_coef(const float* total_sq, float max_norm, float* norm_out, float* coef_out, int* found_inf,
                             float extra_scale, cudaStream_t st) {
  mlb::clip_coef_kernel<<<1, 1, 0, st>>>(total_sq, max_norm, norm_out, coef_out, found_inf, extra_scale);
  return (int)cudaGetLastError();
}

extern "C" int mlb_scale_cast(int in_dtype, int out_dtype, const void* x, void* y, long long n, float scale,
                              const float* scale_ptr, cudaStream_t st) {
  const int grid = mlb::opt_grid(n > 0 ? n : 1);
#define SC(TI, TO) mlb::scale_cast_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x, (TO*)y, n, scale, scale_ptr)
  if (in_dtype == mlb::DT_F32 && out_dtype == mlb::DT_F32) SC(float, float);
  else if (in_dtype == mlb::DT_F32 && out_dtype == mlb::DT_BF16) SC(float, __nv_bfloat16);
  else if (in_dtype == mlb::DT_F32 && out_dtype == mlb::DT_F16) SC(float, __half);
  else if (in_dtype == mlb::DT_BF16 && out_dtype == mlb::DT_F32) SC(__nv_bfloat16, float);
  else if (in_dtype == mlb::DT_F16 && out_dtype == mlb::DT_F32) SC(__half, float);
  else if (in_dtype == mlb::DT_BF16 && out_dtype == mlb::DT_BF16) SC(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == mlb::DT_F16 && out_dtype == mlb::DT_F16) SC(__half, __half);
  else return -100;
#undef SC
  return (int)cudaGetLastError();
}

extern "C" int mlb_accumulate(int dtype, const void* x, float* y, long long n, cudaStream_t st) {
  const int grid = mlb::opt_grid(n > 0 ? n : 1);
  MLB_DISPATCH_DTYPE(dtype, T, mlb::accumulate_kernel<T><<<grid, 256, 0, st>>>((const T*)x, y, n));
  return (int)cudaGetLastError();
}

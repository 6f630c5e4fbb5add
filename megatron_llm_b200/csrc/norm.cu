// RMSNorm / LayerNorm forward+backward for sm_100a, fp32 statistics, optional fused residual add.
//
// Replaces: the reference's torch RMSNorm (5-6 unfused elementwise kernels per call,
// megatron/model/fused_layer_norm.py:125-139) and the apex/in-tree mixed-precision LayerNorm
// (megatron/fused_kernels/layer_norm_cuda_kernel.cu).  One CTA owns one row at a time and keeps it
// in registers between the statistics pass and the normalise pass, so every element is read from
// HBM exactly once per direction; 16-byte vector accesses throughout.
#include "common.cuh"

namespace mlb {

constexpr int NORM_MAXV = 4;      // vectors (of 8 elements) cached per thread, forward
constexpr int NORM_BWD_MAXV = 2;  // backward keeps 5 such arrays live -> fewer vectors per thread

template <typename T, bool RMS, bool ADD_RES>
__global__ void __launch_bounds__(512)
norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res_in, const T* __restrict__ w,
                const T* __restrict__ b, T* __restrict__ y, T* __restrict__ res_out,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H, float eps) {
  __shared__ float scratch[32];
  const int nvec = H / 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + (size_t)row * H;
    float v[NORM_MAXV][8];
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (int it = 0; it < NORM_MAXV; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        Vec<T> a;
        a.load(xr + vi * 8);
        a.to_float(v[it]);
        if constexpr (ADD_RES) {
          Vec<T> r;
          float rf[8];
          r.load(res_in + (size_t)row * H + vi * 8);
          r.to_float(rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[it][j] += rf[j];
          // the residual stream is stored in T: normalise the *rounded* value so fwd/bwd agree
          Vec<T> o;
          o.from_float(v[it]);
          o.store(res_out + (size_t)row * H + vi * 8);
          o.to_float(v[it]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum += v[it][j]; sumsq += v[it][j] * v[it][j]; }
      }
    }
    float mean = 0.f, rstd;
    if constexpr (RMS) {
      const float ss = block_reduce_sum(sumsq, scratch);
      rstd = rsqrtf(ss / H + eps);
    } else {
      mean = block_reduce_sum(sum, scratch) / H;
      float var = 0.f;
#pragma unroll
      for (int it = 0; it < NORM_MAXV; ++it) {
        const int vi = threadIdx.x + it * blockDim.x;
        if (vi < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = v[it][j] - mean; var += d * d; }
        }
      }
      var = block_reduce_sum(var, scratch) / H;
      rstd = rsqrtf(var + eps);
    }
    if (threadIdx.x == 0) {
      if (rstd_out) rstd_out[row] = rstd;
      if (!RMS && mean_out) mean_out[row] = mean;
    }
#pragma unroll
    for (int it = 0; it < NORM_MAXV; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        Vec<T> wv;
        float wf[8], bf[8], o[8];
        wv.load(w + vi * 8);
        wv.to_float(wf);
        if (!RMS && b != nullptr) {
          Vec<T> bv;
          bv.load(b + vi * 8);
          bv.to_float(bf);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) bf[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mean) * rstd * wf[j] + bf[j];
        Vec<T> ov;
        ov.from_float(o);
        ov.store(y + (size_t)row * H + vi * 8);
      }
    }
  }
}

// dx = rstd * (g - c1 - xhat * c2) [+ dres], g = dy*w; partial dw/db per CTA into workspace [grid, H]
template <typename T, bool RMS, bool ADD_DRES>
__global__ void __launch_bounds__(512)
norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ w,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dw_part,
                float* __restrict__ db_part, int rows, int H) {
  __shared__ float scratch[32];
  const int nvec = H / 8;
  float dw_acc[NORM_BWD_MAXV][8], db_acc[NORM_BWD_MAXV][8], wf[NORM_BWD_MAXV][8];
#pragma unroll
  for (int it = 0; it < NORM_BWD_MAXV; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dw_acc[it][j] = 0.f; db_acc[it][j] = 0.f; wf[it][j] = 0.f; }
    if (vi < nvec) {
      Vec<T> wv;
      wv.load(w + vi * 8);
      wv.to_float(wf[it]);
    }
  }
  // software pipeline over the rows of this CTA: the 16-byte loads of row r+1 (x, dy, dres) are in flight while row r
  // goes through its block reductions, so a CTA always has a full row of requests outstanding (the single-row version
  // was latency bound: ~2 TB/s, tools/profiling/ew_drive.py)
  // (the incoming residual gradient is prefetched too where the register budget allows: 16-bit RMSNorm, the hot case;
  // the other instantiations load it at its use so that nothing spills under the 128-register cap of 512 threads)
  constexpr bool PREFETCH_DRES = ADD_DRES && RMS && sizeof(T) == 2;
  constexpr bool PREFETCH = sizeof(T) == 2;
  Vec<T> xa[NORM_BWD_MAXV], da[NORM_BWD_MAXV], ra[NORM_BWD_MAXV];
  auto fetch_one = [&](int row, int it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      xa[it].load(x + (size_t)row * H + vi * 8);
      da[it].load(dy + (size_t)row * H + vi * 8);
      if constexpr (PREFETCH_DRES) ra[it].load(dres + (size_t)row * H + vi * 8);
    }
  };
  auto fetch = [&](int row) {
#pragma unroll
    for (int it = 0; it < NORM_BWD_MAXV; ++it) fetch_one(row, it);
  };
  if (PREFETCH && (int)blockIdx.x < rows) fetch(blockIdx.x);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float rstd = rstd_in[row];
    const float mean = RMS ? 0.f : mean_in[row];
    float g[NORM_BWD_MAXV][8], xh[NORM_BWD_MAXV][8], rf[NORM_BWD_MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < NORM_BWD_MAXV; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if constexpr (!PREFETCH) fetch_one(row, it);
      if (vi < nvec) {
        float xf[8], df[8];
        xa[it].to_float(xf);
        da[it].to_float(df);
        if constexpr (PREFETCH_DRES) ra[it].to_float(rf[it]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[it][j] = (xf[j] - mean) * rstd;
          g[it][j] = df[j] * wf[it][j];
          dw_acc[it][j] += df[j] * xh[it][j];
          if (!RMS) db_acc[it][j] += df[j];
          s1 += g[it][j];
          s2 += g[it][j] * xh[it][j];
        }
      }
    }
    if (PREFETCH && row + (int)gridDim.x < rows) fetch(row + gridDim.x);   // next row's loads fly during the reductions
    const float c2 = block_reduce_sum(s2, scratch) / H;
    float c1 = 0.f;
    if constexpr (!RMS) c1 = block_reduce_sum(s1, scratch) / H;
#pragma unroll
    for (int it = 0; it < NORM_BWD_MAXV; ++it) {
      const int vi = threadIdx.x + it * blockDim.x;
      if (vi < nvec) {
        float o[8];
        if constexpr (ADD_DRES && !PREFETCH_DRES) {
          Vec<T> r;
          r.load(dres + (size_t)row * H + vi * 8);
          r.to_float(rf[it]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = rstd * (g[it][j] - c1 - xh[it][j] * c2);
          if constexpr (ADD_DRES) o[j] += rf[it][j];
        }
        Vec<T> ov;
        ov.from_float(o);
        ov.store(dx + (size_t)row * H + vi * 8);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < NORM_BWD_MAXV; ++it) {
    const int vi = threadIdx.x + it * blockDim.x;
    if (vi < nvec) {
      float* p = dw_part + (size_t)blockIdx.x * H + vi * 8;
      *reinterpret_cast<float4*>(p) = make_float4(dw_acc[it][0], dw_acc[it][1], dw_acc[it][2], dw_acc[it][3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(dw_acc[it][4], dw_acc[it][5], dw_acc[it][6], dw_acc[it][7]);
      if (!RMS && db_part != nullptr) {
        float* q = db_part + (size_t)blockIdx.x * H + vi * 8;
        *reinterpret_cast<float4*>(q) = make_float4(db_acc[it][0], db_acc[it][1], db_acc[it][2], db_acc[it][3]);
        *reinterpret_cast<float4*>(q + 4) = make_float4(db_acc[it][4], db_acc[it][5], db_acc[it][6], db_acc[it][7]);
      }
    }
  }
}

// out[c] = sum_p part[p, c]   (deterministic column reduction of the per-CTA partials)
template <typename T>
__global__ void colsum_kernel(const float* __restrict__ part, T* __restrict__ out, int parts, int H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += part[(size_t)p * H + c];
  out[c] = from_f<T>(s);
}

static int pick_threads(int H, int maxv = NORM_MAXV) {
  int nvec = H / 8;
  int t = (nvec + maxv - 1) / maxv;
  t = ((t + 31) / 32) * 32;
  if (t < 64) t = 64;
  // prefer >=128 threads for latency hiding when the row is long enough
  if (t < 128 && nvec >= 128) t = 128;
  return t;
}

template <typename T>
static int norm_fwd_t(const void* x, const void* res_in, const void* w, const void* b, void* y, void* res_out,
                      float* mean, float* rstd, int rows, int H, float eps, int rms, cudaStream_t st) {
  const int threads = pick_threads(H);
  if (threads > 512) return -3;
  const int grid = rows;
  const T* xp = (const T*)x; const T* rp = (const T*)res_in; const T* wp = (const T*)w; const T* bp = (const T*)b;
  T* yp = (T*)y; T* rop = (T*)res_out;
  if (rms) {
    if (res_in) norm_fwd_kernel<T, true, true><<<grid, threads, 0, st>>>(xp, rp, wp, bp, yp, rop, mean, rstd, rows, H, eps);
    else norm_fwd_kernel<T, true, false><<<grid, threads, 0, st>>>(xp, rp, wp, bp, yp, rop, mean, rstd, rows, H, eps);
  } else {
    if (res_in) norm_fwd_kernel<T, false, true><<<grid, threads, 0, st>>>(xp, rp, wp, bp, yp, rop, mean, rstd, rows, H, eps);
    else norm_fwd_kernel<T, false, false><<<grid, threads, 0, st>>>(xp, rp, wp, bp, yp, rop, mean, rstd, rows, H, eps);
  }
  return (int)cudaGetLastError();
}

template <typename T>
static int norm_bwd_t(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                      const void* dres, void* dx, void* dw, void* db, float* workspace, int parts, int rows, int H,
                      int rms, cudaStream_t st) {
  const int threads = pick_threads(H, NORM_BWD_MAXV);
  if (threads > 512) return -3;
  float* dw_part = workspace;
  float* db_part = workspace + (size_t)parts * H;
  const T* dyp = (const T*)dy; const T* xp = (const T*)x; const T* wp = (const T*)w; const T* drp = (const T*)dres;
  T* dxp = (T*)dx;
  if (rms) {
    if (dres) norm_bwd_kernel<T, true, true><<<parts, threads, 0, st>>>(dyp, xp, wp, mean, rstd, drp, dxp, dw_part, db_part, rows, H);
    else norm_bwd_kernel<T, true, false><<<parts, threads, 0, st>>>(dyp, xp, wp, mean, rstd, drp, dxp, dw_part, db_part, rows, H);
  } else {
    if (dres) norm_bwd_kernel<T, false, true><<<parts, threads, 0, st>>>(dyp, xp, wp, mean, rstd, drp, dxp, dw_part, db_part, rows, H);
    else norm_bwd_kernel<T, false, false><<<parts, threads, 0, st>>>(dyp, xp, wp, mean, rstd, drp, dxp, dw_part, db_part, rows, H);
  }
  colsum_kernel<T><<<(H + 255) / 256, 256, 0, st>>>(dw_part, (T*)dw, parts, H);
  if (!rms && db) colsum_kernel<T><<<(H + 255) / 256, 256, 0, st>>>(db_part, (T*)db, parts, H);
  return (int)cudaGetLastError();
}

}  // namespace mlb

extern "C" int mlb_norm_fwd(int dtype, const void* x, const void* res_in, const void* w, const void* b, void* y,
                            void* res_out, float* mean, float* rstd, int rows, int H, float eps, int rms,
                            cudaStream_t st) {
  if (H % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T, return mlb::norm_fwd_t<T>(x, res_in, w, b, y, res_out, mean, rstd, rows, H, eps, rms, st));
  return 0;
}

// workspace: 2 * parts * H floats
extern "C" int mlb_norm_bwd(int dtype, const void* dy, const void* x, const void* w, const float* mean,
                            const float* rstd, const void* dres, void* dx, void* dw, void* db, float* workspace,
                            int parts, int rows, int H, int rms, cudaStream_t st) {
  if (H % 8) return -2;
  MLB_DISPATCH_DTYPE(dtype, T, return mlb::norm_bwd_t<T>(dy, x, w, mean, rstd, dres, dx, dw, db, workspace, parts, rows, H, rms, st));
  return 0;
}

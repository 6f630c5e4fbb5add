// Bindings for the sm_100a attention kernels (attention_sm100.cu).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdio>
#include <stdexcept>

#include <cuda_runtime.h>

extern "C" {
int mlb_attn_fwd(const void* q, const void* k, const void* v, const long long* q_str, const long long* k_str,
                 const long long* v_str, int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map,
                 int q_per_kv, int seq, int batch, int heads, int window, float softmax_scale, void* out,
                 long long out_s_stride, long long out_b_stride, float* lse, int head_dim, cudaStream_t stream);
int mlb_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const long long* q_str,
                 const long long* k_str, const long long* v_str, const long long* o_str, const long long* do_str,
                 int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map, int q_per_kv, int seq,
                 int batch, int heads, int window, float softmax_scale, const float* lse, float* delta, void* dq,
                 void* dk, void* dv, const long long* dq_str, const long long* dk_str, const long long* dv_str,
                 int head_dim, cudaStream_t stream);
}

static cudaStream_t cur() { return at::cuda::getCurrentCUDAStream().stream(); }
#define CHK(call)                                                                                     \
  do {                                                                                                \
    int _e = (call);                                                                                  \
    if (_e != 0) {                                                                                    \
      char _buf[256];                                                                                 \
      snprintf(_buf, sizeof(_buf), "%s failed with code %d", #call, _e);                              \
      throw std::runtime_error(_buf);                                                                 \
    }                                                                                                 \
  } while (0)

// tensors are [b, s, n, hn] views (any strides, hn contiguous, hn = 128 or 64); strides passed as (head, seq, batch)
static void strides_of(const torch::Tensor& t, long long* s) {
  TORCH_CHECK(t.dim() == 4 && t.stride(3) == 1 && (t.size(3) == 128 || t.size(3) == 64),
              "attention: expected [b, s, n, hn] with contiguous hn = 128 or 64");
  s[0] = t.stride(2); s[1] = t.stride(1); s[2] = t.stride(0);
}

// Separate q/k/v tensors.  Returns (out as a [b, s, n, hn] view over [s, b, n, hn] storage, lse [b, n, s]).
static std::vector<torch::Tensor> attn_fwd(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                                           bool causal, int64_t window, double scale) {
  TORCH_CHECK(causal, "attn_fwd: causal only");
  c10::cuda::CUDAGuard guard(q.device());
  const int b = q.size(0), s = q.size(1), n = q.size(2), nkv = k.size(2), hn = q.size(3);
  TORCH_CHECK(k.size(3) == hn && v.size(3) == hn);
  long long qs[3], ks[3], vs[3];
  strides_of(q, qs); strides_of(k, ks); strides_of(v, vs);
  auto out = torch::empty({s, b, n, hn}, q.options());
  auto lse = torch::empty({b, n, s}, q.options().dtype(torch::kFloat32));
  const int g = n / nkv;
  int head_map[6] = {g, 0, 1, 0, 1, 0};
  CHK(mlb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, n, nkv, nkv, head_map, g, s, b, n,
                   (int)window, (float)scale, out.data_ptr(), (long long)b * n * hn, (long long)n * hn,
                   lse.data_ptr<float>(), hn, cur()));
  return {out.permute({1, 0, 2, 3}), lse};
}

// Returns (dq, dk, dv) as [b, s, n, hn] views over [s, b, n, hn] storage.
static std::vector<torch::Tensor> attn_bwd(const torch::Tensor& dout, const torch::Tensor& q, const torch::Tensor& k,
                                           const torch::Tensor& v, const torch::Tensor& out, const torch::Tensor& lse,
                                           bool causal, int64_t window, double scale) {
  TORCH_CHECK(causal, "attn_bwd: causal only");
  c10::cuda::CUDAGuard guard(q.device());
  const int b = q.size(0), s = q.size(1), n = q.size(2), nkv = k.size(2), hn = q.size(3);
  long long qs[3], ks[3], vs[3], os[3], ds[3], dqs[3], dks[3], dvs[3];
  strides_of(q, qs); strides_of(k, ks); strides_of(v, vs); strides_of(out, os); strides_of(dout, ds);
  auto dq = torch::empty({s, b, n, hn}, q.options()).permute({1, 0, 2, 3});
  auto dk = torch::empty({s, b, nkv, hn}, q.options()).permute({1, 0, 2, 3});
  auto dv = torch::empty({s, b, nkv, hn}, q.options()).permute({1, 0, 2, 3});
  strides_of(dq, dqs); strides_of(dk, dks); strides_of(dv, dvs);
  auto delta = torch::empty({b, n, s}, q.options().dtype(torch::kFloat32));
  const int g = n / nkv;
  int head_map[6] = {g, 0, 1, 0, 1, 0};
  CHK(mlb_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), qs, ks, vs, os, ds, n,
                   nkv, nkv, head_map, g, s, b, n, (int)window, (float)scale, lse.data_ptr<float>(),
                   delta.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dqs, dks, dvs, hn, cur()));
  return {dq, dk, dv};
}

// Packed path: ``mixed`` is the QKV projection output [s, b, nkv * (g + 2) * hn] (per KV group: g query heads, k, v),
// already rotated in place.  The kernels address Q/K/V inside it through the head map: no splits, no transposes, and
// the backward writes dQ/dK/dV straight into one ``dmixed`` buffer of the same layout.
static std::vector<torch::Tensor> attn_fwd_packed(const torch::Tensor& mixed, int64_t nkv, int64_t g, int64_t window,
                                                  double scale, int64_t hn) {
  TORCH_CHECK((hn == 128 || hn == 64) && mixed.dim() == 3 && mixed.stride(2) == 1 &&
                  mixed.size(2) == nkv * (g + 2) * hn,
              "attn_fwd_packed: expected [s, b, nkv * (g + 2) * hn], hn = 128 or 64");
  c10::cuda::CUDAGuard guard(mixed.device());
  const int s = mixed.size(0), b = mixed.size(1), n = nkv * g, mh = nkv * (g + 2);
  long long ms[3] = {hn, (long long)mixed.stride(0), (long long)mixed.stride(1)};
  auto out = torch::empty({s, b, (int64_t)n * hn}, mixed.options());
  auto lse = torch::empty({b, n, s}, mixed.options().dtype(torch::kFloat32));
  int head_map[6] = {(int)g + 2, 0, (int)g + 2, (int)g, (int)g + 2, (int)g + 1};
  CHK(mlb_attn_fwd(mixed.data_ptr(), mixed.data_ptr(), mixed.data_ptr(), ms, ms, ms, mh, mh, mh, head_map, (int)g, s,
                   b, n, (int)window, (float)scale, out.data_ptr(), (long long)b * n * hn, (long long)n * hn,
                   lse.data_ptr<float>(), (int)hn, cur()));
  return {out, lse};
}

static torch::Tensor attn_bwd_packed(const torch::Tensor& dout, const torch::Tensor& mixed, const torch::Tensor& out,
                                     const torch::Tensor& lse, int64_t nkv, int64_t g, int64_t window, double scale,
                                     int64_t hn) {
  TORCH_CHECK(dout.dim() == 3 && dout.stride(2) == 1 && out.stride(2) == 1, "attn_bwd_packed: contiguous hn expected");
  c10::cuda::CUDAGuard guard(mixed.device());
  const int s = mixed.size(0), b = mixed.size(1), n = nkv * g, mh = nkv * (g + 2);
  long long ms[3] = {hn, (long long)mixed.stride(0), (long long)mixed.stride(1)};
  long long os[3] = {hn, (long long)out.stride(0), (long long)out.stride(1)};
  long long ds[3] = {hn, (long long)dout.stride(0), (long long)dout.stride(1)};
  auto dmixed = torch::empty({s, b, mixed.size(2)}, mixed.options());
  long long dms[3] = {hn, (long long)dmixed.stride(0), (long long)dmixed.stride(1)};
  auto delta = torch::empty({b, n, s}, mixed.options().dtype(torch::kFloat32));
  int head_map[6] = {(int)g + 2, 0, (int)g + 2, (int)g, (int)g + 2, (int)g + 1};
  CHK(mlb_attn_bwd(mixed.data_ptr(), mixed.data_ptr(), mixed.data_ptr(), out.data_ptr(), dout.data_ptr(), ms, ms, ms,
                   os, ds, mh, mh, mh, head_map, (int)g, s, b, n, (int)window, (float)scale, lse.data_ptr<float>(),
                   delta.data_ptr<float>(), dmixed.data_ptr(), dmixed.data_ptr(), dmixed.data_ptr(), dms, dms, dms,
                   (int)hn, cur()));
  return dmixed;
}

void register_attention(pybind11::module_& m) {
  m.def("attn_fwd_packed", &attn_fwd_packed);
  m.def("attn_bwd_packed", &attn_bwd_packed);
  m.def("attn_fwd", &attn_fwd);
  m.def("attn_bwd", &attn_bwd);
}

// Bindings for the sm_100a attention kernels (attention_sm100.cu).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <algorithm>
#include <cstdio>
#include <stdexcept>

#include <cuda_runtime.h>

extern "C" {
int mlb_attn_fwd_ex(const void* q, const void* k, const void* v, const long long* q_str, const long long* k_str,
                    const long long* v_str, int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map,
                    int q_per_kv, int seq, int batch, int heads, int window, float softmax_scale, void* out,
                    long long out_s_stride, long long out_b_stride, float* lse, int head_dim, int fp16, float dropout_p,
                    unsigned long long seed, cudaStream_t stream);
int mlb_attn_bwd_ex(const void* q, const void* k, const void* v, const void* o, const void* dout,
                    const long long* q_str, const long long* k_str, const long long* v_str, const long long* o_str,
                    const long long* do_str, int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map,
                    int q_per_kv, int seq, int batch, int heads, int window, float softmax_scale, const float* lse,
                    float* delta, void* dq, void* dk, void* dv, const long long* dq_str, const long long* dk_str,
                    const long long* dv_str, int head_dim, int fp16, float dropout_p, unsigned long long seed,
                    cudaStream_t stream);
int mlb_attn_decode(int dtype, const void* q, const void* k, const void* v, const long long* q_str,
                    const long long* k_str, const long long* v_str, int batch, int sq, int sk, int nq, int nkv,
                    int head_dim, int window, float softmax_scale, int n_splits, int keys_per_split, float* part_o,
                    float* part_ml, void* out, cudaStream_t stream);
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
// element type of the attention tensors: 0 = bf16, 1 = fp16 (one instantiation of the kernels each)
static int fp16_flag(const torch::Tensor& t) {
  TORCH_CHECK(t.scalar_type() == torch::kBFloat16 || t.scalar_type() == torch::kFloat16,
              "attention: bf16 or fp16 tensors expected");
  return t.scalar_type() == torch::kFloat16 ? 1 : 0;
}

static void strides_of(const torch::Tensor& t, long long* s) {
  TORCH_CHECK(t.dim() == 4 && t.stride(3) == 1 && (t.size(3) == 128 || t.size(3) == 64),
              "attention: expected [b, s, n, hn] with contiguous hn = 128 or 64");
  s[0] = t.stride(2); s[1] = t.stride(1); s[2] = t.stride(0);
}

// Separate q/k/v tensors.  Returns (out as a [b, s, n, hn] view over [s, b, n, hn] storage, lse [b, n, s]).
static std::vector<torch::Tensor> attn_fwd(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                                           bool causal, int64_t window, double scale, double dropout_p,
                                           int64_t seed) {
  TORCH_CHECK(causal, "attn_fwd: causal only");
  TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type(), "attn_fwd: one dtype");
  c10::cuda::CUDAGuard guard(q.device());
  const int b = q.size(0), s = q.size(1), n = q.size(2), nkv = k.size(2), hn = q.size(3);
  TORCH_CHECK(k.size(3) == hn && v.size(3) == hn);
  long long qs[3], ks[3], vs[3];
  strides_of(q, qs); strides_of(k, ks); strides_of(v, vs);
  auto out = torch::empty({s, b, n, hn}, q.options());
  auto lse = torch::empty({b, n, s}, q.options().dtype(torch::kFloat32));
  const int g = n / nkv;
  int head_map[6] = {g, 0, 1, 0, 1, 0};
  CHK(mlb_attn_fwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, n, nkv, nkv, head_map, g, s, b, n,
                      (int)window, (float)scale, out.data_ptr(), (long long)b * n * hn, (long long)n * hn,
                      lse.data_ptr<float>(), hn, fp16_flag(q), (float)dropout_p, (unsigned long long)seed, cur()));
  return {out.permute({1, 0, 2, 3}), lse};
}

// Returns (dq, dk, dv) as [b, s, n, hn] views over [s, b, n, hn] storage.
static std::vector<torch::Tensor> attn_bwd(const torch::Tensor& dout, const torch::Tensor& q, const torch::Tensor& k,
                                           const torch::Tensor& v, const torch::Tensor& out, const torch::Tensor& lse,
                                           bool causal, int64_t window, double scale, double dropout_p,
                                           int64_t seed) {
  TORCH_CHECK(causal, "attn_bwd: causal only");
  TORCH_CHECK(dout.scalar_type() == q.scalar_type() && out.scalar_type() == q.scalar_type(), "attn_bwd: one dtype");
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
  CHK(mlb_attn_bwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), qs, ks, vs, os, ds,
                      n, nkv, nkv, head_map, g, s, b, n, (int)window, (float)scale, lse.data_ptr<float>(),
                      delta.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dqs, dks, dvs, hn,
                      fp16_flag(q), (float)dropout_p, (unsigned long long)seed, cur()));
  return {dq, dk, dv};
}

// Packed path: ``mixed`` is the QKV projection output [s, b, nkv * (g + 2) * hn] (per KV group: g query heads, k, v),
// already rotated in place.  The kernels address Q/K/V inside it through the head map: no splits, no transposes, and
// the backward writes dQ/dK/dV straight into one ``dmixed`` buffer of the same layout.
static std::vector<torch::Tensor> attn_fwd_packed(const torch::Tensor& mixed, int64_t nkv, int64_t g, int64_t window,
                                                  double scale, int64_t hn, double dropout_p, int64_t seed) {
  TORCH_CHECK((hn == 128 || hn == 64) && mixed.dim() == 3 && mixed.stride(2) == 1 &&
                  mixed.size(2) == nkv * (g + 2) * hn,
              "attn_fwd_packed: expected [s, b, nkv * (g + 2) * hn], hn = 128 or 64");
  c10::cuda::CUDAGuard guard(mixed.device());
  const int s = mixed.size(0), b = mixed.size(1), n = nkv * g, mh = nkv * (g + 2);
  long long ms[3] = {hn, (long long)mixed.stride(0), (long long)mixed.stride(1)};
  auto out = torch::empty({s, b, (int64_t)n * hn}, mixed.options());
  auto lse = torch::empty({b, n, s}, mixed.options().dtype(torch::kFloat32));
  int head_map[6] = {(int)g + 2, 0, (int)g + 2, (int)g, (int)g + 2, (int)g + 1};
  CHK(mlb_attn_fwd_ex(mixed.data_ptr(), mixed.data_ptr(), mixed.data_ptr(), ms, ms, ms, mh, mh, mh, head_map, (int)g,
                      s, b, n, (int)window, (float)scale, out.data_ptr(), (long long)b * n * hn, (long long)n * hn,
                      lse.data_ptr<float>(), (int)hn, fp16_flag(mixed), (float)dropout_p, (unsigned long long)seed,
                      cur()));
  return {out, lse};
}

static torch::Tensor attn_bwd_packed(const torch::Tensor& dout, const torch::Tensor& mixed, const torch::Tensor& out,
                                     const torch::Tensor& lse, int64_t nkv, int64_t g, int64_t window, double scale,
                                     int64_t hn, double dropout_p, int64_t seed) {
  TORCH_CHECK(dout.dim() == 3 && dout.stride(2) == 1 && out.stride(2) == 1, "attn_bwd_packed: contiguous hn expected");
  TORCH_CHECK(dout.scalar_type() == mixed.scalar_type() && out.scalar_type() == mixed.scalar_type(),
              "attn_bwd_packed: one dtype");
  c10::cuda::CUDAGuard guard(mixed.device());
  const int s = mixed.size(0), b = mixed.size(1), n = nkv * g, mh = nkv * (g + 2);
  long long ms[3] = {hn, (long long)mixed.stride(0), (long long)mixed.stride(1)};
  long long os[3] = {hn, (long long)out.stride(0), (long long)out.stride(1)};
  long long ds[3] = {hn, (long long)dout.stride(0), (long long)dout.stride(1)};
  auto dmixed = torch::empty({s, b, mixed.size(2)}, mixed.options());
  long long dms[3] = {hn, (long long)dmixed.stride(0), (long long)dmixed.stride(1)};
  auto delta = torch::empty({b, n, s}, mixed.options().dtype(torch::kFloat32));
  int head_map[6] = {(int)g + 2, 0, (int)g + 2, (int)g, (int)g + 2, (int)g + 1};
  CHK(mlb_attn_bwd_ex(mixed.data_ptr(), mixed.data_ptr(), mixed.data_ptr(), out.data_ptr(), dout.data_ptr(), ms, ms,
                      ms, os, ds, mh, mh, mh, head_map, (int)g, s, b, n, (int)window, (float)scale,
                      lse.data_ptr<float>(), delta.data_ptr<float>(), dmixed.data_ptr(), dmixed.data_ptr(),
                      dmixed.data_ptr(), dms, dms, dms, (int)hn, fp16_flag(mixed), (float)dropout_p,
                      (unsigned long long)seed, cur()));
  return dmixed;
}

// KV-cache decode step (attention_decode.cu): q [b, sq, n, hn] (a few positions), k / v [b, sk, nkv, hn] cache views
// (any strides that keep 16-byte row alignment, hn contiguous), bf16 or fp16, causal with bottom-right alignment.
// Returns out [b, sq, n, hn] (contiguous).  ``splits`` <= 0 picks the split count from the problem size.
static torch::Tensor attn_decode(const torch::Tensor& q, const torch::Tensor& k, const torch::Tensor& v,
                                 int64_t window, double scale, int64_t splits) {
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "attn_decode: CUDA tensors expected");
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "attn_decode: [b, s, n, hn] tensors expected");
  const auto dt = q.scalar_type();
  TORCH_CHECK((dt == torch::kBFloat16 || dt == torch::kFloat16) && k.scalar_type() == dt && v.scalar_type() == dt,
              "attn_decode: bf16 or fp16 q / k / v of one dtype");
  const int b = q.size(0), sq = q.size(1), n = q.size(2), hn = q.size(3), sk = k.size(1), nkv = k.size(2);
  TORCH_CHECK((hn == 128 || hn == 64) && k.size(3) == hn && v.size(3) == hn && k.size(0) == b && v.size(0) == b &&
                  v.size(1) == sk && v.size(2) == nkv && n % nkv == 0 && sq >= 1 && sk >= sq,
              "attn_decode: shape mismatch");
  for (const torch::Tensor* t : {&q, &k, &v}) {
    TORCH_CHECK(t->stride(3) == 1 && t->stride(0) % 8 == 0 && t->stride(1) % 8 == 0 && t->stride(2) % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(t->data_ptr()) % 16 == 0,
                "attn_decode: rows must be contiguous and 16-byte aligned");
  }
  c10::cuda::CUDAGuard guard(q.device());
  int n_splits = (int)splits;
  if (n_splits <= 0) {                                  // ~2 CTAs per SM, at least 256 keys per slice
    const int ctas = b * nkv;
    n_splits = std::max(1, std::min((296 + ctas - 1) / ctas, (sk + 255) / 256));
  }
  int keys_per_split = (((sk + n_splits - 1) / n_splits) + 31) / 32 * 32;
  n_splits = (sk + keys_per_split - 1) / keys_per_split;
  const int g = n / nkv;
  const int64_t rows = (int64_t)b * nkv * n_splits * sq * g;
  auto f32 = q.options().dtype(torch::kFloat32);
  auto part_o = torch::empty({rows, hn}, f32);
  auto part_ml = torch::empty({rows, 2}, f32);
  auto out = torch::empty({b, sq, n, hn}, q.options());
  long long qs[3] = {q.stride(0), q.stride(1), q.stride(2)};
  long long ks[3] = {k.stride(0), k.stride(1), k.stride(2)};
  long long vs[3] = {v.stride(0), v.stride(1), v.stride(2)};
  CHK(mlb_attn_decode(dt == torch::kBFloat16 ? 0 : 1, q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, b, sq, sk,
                      n, nkv, hn, (int)window, (float)scale, n_splits, keys_per_split, part_o.data_ptr<float>(),
                      part_ml.data_ptr<float>(), out.data_ptr(), cur()));
  return out;
}

void register_attention(pybind11::module_& m) {
  m.def("attn_decode", &attn_decode);
  namespace py = pybind11;
  // dropout_p / seed are optional trailing arguments: 0 keeps the (bf16 or fp16) no-dropout kernels
  m.def("attn_fwd_packed", &attn_fwd_packed, py::arg("mixed"), py::arg("nkv"), py::arg("g"), py::arg("window"),
        py::arg("scale"), py::arg("hn"), py::arg("dropout_p") = 0.0, py::arg("seed") = 0);
  m.def("attn_bwd_packed", &attn_bwd_packed, py::arg("dout"), py::arg("mixed"), py::arg("out"), py::arg("lse"),
        py::arg("nkv"), py::arg("g"), py::arg("window"), py::arg("scale"), py::arg("hn"), py::arg("dropout_p") = 0.0,
        py::arg("seed") = 0);
  m.def("attn_fwd", &attn_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("causal"), py::arg("window"),
        py::arg("scale"), py::arg("dropout_p") = 0.0, py::arg("seed") = 0);
  m.def("attn_bwd", &attn_bwd, py::arg("dout"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("out"),
        py::arg("lse"), py::arg("causal"), py::arg("window"), py::arg("scale"), py::arg("dropout_p") = 0.0,
        py::arg("seed") = 0);
}

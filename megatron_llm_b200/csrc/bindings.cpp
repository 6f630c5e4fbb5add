// Python bindings (torch tensors -> raw pointers + current CUDA stream) for the sm_100a kernels.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdio>
#include <stdexcept>

#include <cuda_runtime.h>

namespace mlb { struct GemmComm; }

extern "C" {
int mlb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                  int a_mn_major, int b_mn_major, int epilogue, int block_n, int fp16, int num_sms,
                  cudaStream_t stream);
int mlb_gemm2_debug_read(unsigned long long* host, int n);
int mlb_gemm_bf16_2cta(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                       int a_mn_major, int b_mn_major, int epilogue, int fp16, int num_sms, cudaStream_t stream);
int mlb_norm_fwd(int dtype, const void* x, const void* res_in, const void* w, const void* b, void* y, void* res_out,
                 float* mean, float* rstd, int rows, int H, float eps, int rms, cudaStream_t st);
int mlb_norm_bwd(int dtype, const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                 const void* dres, void* dx, void* dw, void* db, float* workspace, int parts, int rows, int H, int rms,
                 cudaStream_t st);
int mlb_rope_qkv(int dtype, void* qkv, const void* freqs, const long long* position_ids, int tokens, int batch,
                 int n_groups, int heads_per_group, int hn, int pos_offset, int inverse, long long token_stride,
                 cudaStream_t st);
int mlb_rope_heads(int dtype, void* x, const void* freqs, const long long* position_ids, int tokens, int batch,
                   int heads, int hn, int pos_offset, int inverse, long long token_stride, long long head_stride,
                   cudaStream_t st);
int mlb_glu_fwd(int dtype, const void* x, void* y, long long rows, int F, int kind, cudaStream_t st);
int mlb_glu_bwd(int dtype, const void* dy, const void* x, void* dx, long long rows, int F, int kind, cudaStream_t st);
int mlb_gelu(int dtype, const void* x, const void* bias, const void* dy, void* out, long long rows, int F, int approx,
             int backward, cudaStream_t st);
int mlb_bias_dropout_add(int dtype, const void* x, const void* bias, const void* residual, void* out, long long rows,
                         int F, float p, unsigned long long seed, int backward, cudaStream_t st);
int mlb_embedding_fwd(int dtype, const long long* ids, const void* weight, void* out, int batch, int seq, int H,
                      long long vocab_start, long long rows_local, int sbh, cudaStream_t st);
int mlb_embedding_bwd(int dtype, const long long* ids, const void* dout, float* dweight, int batch, int seq, int H,
                      long long vocab_start, long long rows_local, int sbh, cudaStream_t st);
int mlb_ce_stats(int dtype, const void* logits, const long long* target, float* stats, int rows, int Vp,
                 int vocab_start, long long row_stride, cudaStream_t st);
int mlb_ce_bwd(int dtype, const void* logits, void* out, const long long* target, const float* M, const float* logS,
               const float* g, int rows, int Vp, int vocab_start, float smoothing, int vocab_size,
               long long row_stride, cudaStream_t st);
int mlb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, int p16_dtype, long long n,
                   long long global_offset, const long long* seg_start, const float* seg_wd, const float* seg_lr_mult,
                   int nseg, float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                   const float* grad_scale_ptr, const int* skip_flag, const long long* p16_peers, int n_peers,
                   cudaStream_t st);
int mlb_sgd_flat(float* p, const float* g, float* mom, void* p16, int p16_dtype, long long n, long long global_offset,
                 const long long* seg_start, const float* seg_wd, const float* seg_lr_mult, int nseg, float lr,
                 float momentum, int first_step, const float* grad_scale_ptr, const int* skip_flag, cudaStream_t st);
int mlb_sqnorm_flat(int dtype, const void* x, long long n, long long global_offset, const long long* seg_start,
                    const float* seg_weight, int nseg, float* workspace, float* out, int accumulate, cudaStream_t st);
int mlb_clip_coef(const float* total_sq, float max_norm, float* norm_out, float* coef_out, int* found_inf,
                  float extra_scale, cudaStream_t st);
int mlb_scale_cast(int in_dtype, int out_dtype, const void* x, void* y, long long n, float scale,
                   const float* scale_ptr, cudaStream_t st);
int mlb_accumulate(int dtype, const void* x, float* y, long long n, cudaStream_t st);
int mlb_softmax_fwd(int dtype, const void* x, void* y, const unsigned char* mask, float scale, long long rows, int sq,
                    int sk, int np, int mask_batch, int mode, cudaStream_t st);
int mlb_softmax_bwd(int dtype, void* dy, const void* y, float scale, long long rows, int sk, cudaStream_t st);
}

static int dt(const torch::Tensor& t) {
  switch (t.scalar_type()) {
    case torch::kBFloat16: return 0;
    case torch::kFloat16: return 1;
    case torch::kFloat32: return 2;
    default: TORCH_CHECK(false, "unsupported dtype ", t.scalar_type());
  }
  return -1;
}
static cudaStream_t cur() { return at::cuda::getCurrentCUDAStream().stream(); }
static const void* optp(const c10::optional<torch::Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }
#define CHK(call)                                                                                     \
  do {                                                                                                \
    int _e = (call);                                                                                  \
    if (_e != 0) {                                                                                    \
      char _buf[256];                                                                                 \
      snprintf(_buf, sizeof(_buf), "%s failed with code %d", #call, _e);                              \
      throw std::runtime_error(_buf);                                                                 \
    }                                                                                                 \
  } while (0)

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

// C[M,N] (+)= A * B^T ; see mlb_gemm_bf16 for the major flags.  comm: optional int64 tensor (host) holding a packed
// GemmComm struct (built by parallel/fused_tp.py).
static void gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C, int64_t M, int64_t N, int64_t K,
                 int64_t lda, int64_t ldb, int64_t ldc, bool a_mn, bool b_mn, int64_t epilogue, int64_t block_n,
                 const c10::optional<torch::Tensor>& comm, int64_t sms) {
  const bool fp16 = A.scalar_type() == torch::kFloat16;
  TORCH_CHECK((fp16 || A.scalar_type() == torch::kBFloat16) && B.scalar_type() == A.scalar_type(),
              "gemm: bf16 x bf16 or fp16 x fp16 operands");
  TORCH_CHECK(C.scalar_type() == torch::kFloat32 || C.scalar_type() == A.scalar_type(), "gemm: C is fp32 or the operand type");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(A.data_ptr()) % 16) == 0 && (reinterpret_cast<uintptr_t>(B.data_ptr()) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(C.data_ptr()) % 16) == 0, "gemm: 16-byte aligned operands");
  TORCH_CHECK(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && N % 8 == 0, "gemm: leading dims / N must be multiples of 8");
  c10::cuda::CUDAGuard guard(A.device());
  // 2-CTA (cta_group::2) 256x256 tiles; its TMA-store epilogue needs 16-byte multiples of the output row pitch
  const int64_t out_bytes = (epilogue == 0 || epilogue == 3) ? 2 : 4;
  if (block_n == 512 && (ldc * out_bytes) % 16 == 0) {
    CHK(mlb_gemm_bf16_2cta(A.data_ptr(), B.data_ptr(), C.data_ptr(), (int)M, (int)N, (int)K, (int)lda, (int)ldb,
                           (int)ldc, a_mn, b_mn, (int)epilogue, fp16, sms > 0 ? (int)sms : num_sms(), cur()));
    return;
  }
  if (block_n == 512) block_n = 0;
  CHK(mlb_gemm_bf16(A.data_ptr(), B.data_ptr(), C.data_ptr(), (int)M, (int)N, (int)K, (int)lda, (int)ldb, (int)ldc,
                    a_mn, b_mn, (int)epilogue, (int)block_n, fp16, sms > 0 ? (int)sms : num_sms(), cur()));
}

static void norm_fwd(const torch::Tensor& x, const c10::optional<torch::Tensor>& res_in, const torch::Tensor& w,
                     const c10::optional<torch::Tensor>& b, torch::Tensor& y, const c10::optional<torch::Tensor>& res_out,
                     const c10::optional<torch::Tensor>& mean, torch::Tensor& rstd, double eps, bool rms) {
  c10::cuda::CUDAGuard guard(x.device());
  const int H = x.size(-1);
  const int rows = x.numel() / H;
  CHK(mlb_norm_fwd(dt(x), x.data_ptr(), optp(res_in), w.data_ptr(), optp(b), y.data_ptr(),
                   const_cast<void*>(optp(res_out)), mean.has_value() ? mean->data_ptr<float>() : nullptr,
                   rstd.data_ptr<float>(), rows, H, (float)eps, rms, cur()));
}

static void norm_bwd(const torch::Tensor& dy, const torch::Tensor& x, const torch::Tensor& w,
                     const c10::optional<torch::Tensor>& mean, const torch::Tensor& rstd,
                     const c10::optional<torch::Tensor>& dres, torch::Tensor& dx, torch::Tensor& dw,
                     const c10::optional<torch::Tensor>& db, torch::Tensor& workspace, int64_t parts, bool rms) {
  c10::cuda::CUDAGuard guard(x.device());
  const int H = x.size(-1);
  const int rows = x.numel() / H;
  CHK(mlb_norm_bwd(dt(x), dy.data_ptr(), x.data_ptr(), w.data_ptr(), mean.has_value() ? mean->data_ptr<float>() : nullptr,
                   rstd.data_ptr<float>(), optp(dres), dx.data_ptr(), dw.data_ptr(), const_cast<void*>(optp(db)),
                   workspace.data_ptr<float>(), (int)parts, rows, H, rms, cur()));
}

static void rope_qkv(torch::Tensor& qkv, const torch::Tensor& freqs, const c10::optional<torch::Tensor>& pos,
                     int64_t tokens, int64_t batch, int64_t n_groups, int64_t heads_per_group, int64_t hn,
                     int64_t pos_offset, bool inverse, int64_t token_stride) {
  c10::cuda::CUDAGuard guard(qkv.device());
  CHK(mlb_rope_qkv(dt(qkv), qkv.data_ptr(), freqs.data_ptr(), pos.has_value() ? (const long long*)pos->data_ptr() : nullptr,
                   (int)tokens, (int)batch, (int)n_groups, (int)heads_per_group, (int)hn, (int)pos_offset, inverse,
                   token_stride, cur()));
}

static void rope_heads(torch::Tensor& x, const torch::Tensor& freqs, const c10::optional<torch::Tensor>& pos,
                       int64_t tokens, int64_t batch, int64_t heads, int64_t hn, int64_t pos_offset, bool inverse,
                       int64_t token_stride, int64_t head_stride) {
  c10::cuda::CUDAGuard guard(x.device());
  CHK(mlb_rope_heads(dt(x), x.data_ptr(), freqs.data_ptr(), pos.has_value() ? (const long long*)pos->data_ptr() : nullptr,
                     (int)tokens, (int)batch, (int)heads, (int)hn, (int)pos_offset, inverse, token_stride, head_stride,
                     cur()));
}

static void glu_fwd(const torch::Tensor& x, torch::Tensor& y, int64_t kind) {
  c10::cuda::CUDAGuard guard(x.device());
  const int F = y.size(-1);
  CHK(mlb_glu_fwd(dt(x), x.data_ptr(), y.data_ptr(), y.numel() / F, F, (int)kind, cur()));
}
static void glu_bwd(const torch::Tensor& dy, const torch::Tensor& x, torch::Tensor& dx, int64_t kind) {
  c10::cuda::CUDAGuard guard(x.device());
  const int F = dy.size(-1);
  CHK(mlb_glu_bwd(dt(x), dy.data_ptr(), x.data_ptr(), dx.data_ptr(), dy.numel() / F, F, (int)kind, cur()));
}
static void gelu(const torch::Tensor& x, const c10::optional<torch::Tensor>& bias, const c10::optional<torch::Tensor>& dy,
                 torch::Tensor& out, bool approx, bool backward) {
  c10::cuda::CUDAGuard guard(x.device());
  const int F = x.size(-1);
  CHK(mlb_gelu(dt(x), x.data_ptr(), optp(bias), optp(dy), out.data_ptr(), x.numel() / F, F, approx, backward, cur()));
}
static void bias_dropout_add(const torch::Tensor& x, const c10::optional<torch::Tensor>& bias,
                             const c10::optional<torch::Tensor>& residual, torch::Tensor& out, double p, int64_t seed,
                             bool backward) {
  c10::cuda::CUDAGuard guard(x.device());
  const int F = x.size(-1);
  CHK(mlb_bias_dropout_add(dt(x), x.data_ptr(), optp(bias), optp(residual), out.data_ptr(), x.numel() / F, F, (float)p,
                           (unsigned long long)seed, backward, cur()));
}

// ids [batch, seq] int64; weight [rows_local, H]; out [seq, batch, H] (sbh) or [batch, seq, H]; ids outside
// [vocab_start, vocab_start + rows_local) give zero rows
static void embedding_fwd(const torch::Tensor& ids, const torch::Tensor& weight, torch::Tensor& out, int64_t vocab_start,
                          bool sbh) {
  c10::cuda::CUDAGuard guard(weight.device());
  TORCH_CHECK(ids.scalar_type() == torch::kInt64 && ids.dim() == 2 && ids.is_contiguous());
  TORCH_CHECK(weight.dim() == 2 && weight.is_contiguous() && out.is_contiguous() && out.scalar_type() == weight.scalar_type());
  TORCH_CHECK(out.numel() == ids.numel() * weight.size(1));
  CHK(mlb_embedding_fwd(dt(weight), (const long long*)ids.data_ptr(), weight.data_ptr(), out.data_ptr(), (int)ids.size(0),
                        (int)ids.size(1), (int)weight.size(1), vocab_start, weight.size(0), sbh, cur()));
}
// dweight(fp32) [rows_local, H] += scatter of dout rows by token id
static void embedding_bwd(const torch::Tensor& ids, const torch::Tensor& dout, torch::Tensor& dweight, int64_t vocab_start,
                          bool sbh) {
  c10::cuda::CUDAGuard guard(dout.device());
  TORCH_CHECK(ids.scalar_type() == torch::kInt64 && ids.dim() == 2 && ids.is_contiguous() && dout.is_contiguous());
  TORCH_CHECK(dweight.scalar_type() == torch::kFloat32 && dweight.dim() == 2 && dweight.is_contiguous());
  TORCH_CHECK(dout.numel() == ids.numel() * dweight.size(1));
  CHK(mlb_embedding_bwd(dt(dout), (const long long*)ids.data_ptr(), dout.data_ptr(), dweight.data_ptr<float>(),
                        (int)ids.size(0), (int)ids.size(1), (int)dweight.size(1), vocab_start, dweight.size(0), sbh, cur()));
}

static void ce_stats(const torch::Tensor& logits, const torch::Tensor& target, torch::Tensor& stats, int64_t vocab_start) {
  c10::cuda::CUDAGuard guard(logits.device());
  CHK(mlb_ce_stats(dt(logits), logits.data_ptr(), (const long long*)target.data_ptr(), stats.data_ptr<float>(),
                   (int)logits.size(0), (int)logits.size(1), (int)vocab_start, logits.stride(0), cur()));
}
static void ce_bwd(const torch::Tensor& logits, torch::Tensor& out, const torch::Tensor& target, const torch::Tensor& M,
                   const torch::Tensor& logS, const torch::Tensor& g, int64_t vocab_start, double smoothing,
                   int64_t vocab_size) {
  c10::cuda::CUDAGuard guard(logits.device());
  CHK(mlb_ce_bwd(dt(logits), logits.data_ptr(), out.data_ptr(), (const long long*)target.data_ptr(), M.data_ptr<float>(),
                 logS.data_ptr<float>(), g.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1), (int)vocab_start,
                 (float)smoothing, (int)vocab_size, logits.stride(0), cur()));
}

static void adamw_flat(torch::Tensor& p, const torch::Tensor& g, torch::Tensor& m, torch::Tensor& v,
                       const c10::optional<torch::Tensor>& p16, int64_t global_offset, const torch::Tensor& seg_start,
                       const torch::Tensor& seg_wd, const c10::optional<torch::Tensor>& seg_lr_mult, double lr,
                       double beta1, double beta2, double eps, double bc1, double bc2,
                       const c10::optional<torch::Tensor>& grad_scale, const c10::optional<torch::Tensor>& skip,
                       const std::vector<int64_t>& p16_peers) {
  c10::cuda::CUDAGuard guard(p.device());
  // ZeRO-1 fused cast + parameter all-gather: device pointers of every DP peer's 16-bit buffer at this shard's start
  TORCH_CHECK(p16_peers.empty() || (p16.has_value() && p16_peers.size() <= 8), "adamw_flat: bad peer list");
  long long peers[8] = {0};
  for (size_t i = 0; i < p16_peers.size(); ++i) peers[i] = p16_peers[i];
  CHK(mlb_adamw_flat(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(),
                     const_cast<void*>(optp(p16)), p16.has_value() ? dt(*p16) : 0, p.numel(), global_offset,
                     (const long long*)seg_start.data_ptr(), seg_wd.data_ptr<float>(),
                     seg_lr_mult.has_value() ? seg_lr_mult->data_ptr<float>() : nullptr, (int)seg_wd.numel(), (float)lr,
                     (float)beta1, (float)beta2, (float)eps, (float)bc1, (float)bc2,
                     grad_scale.has_value() ? grad_scale->data_ptr<float>() : nullptr,
                     skip.has_value() ? skip->data_ptr<int>() : nullptr, peers, (int)p16_peers.size(), cur()));
}
static void sgd_flat(torch::Tensor& p, const torch::Tensor& g, torch::Tensor& mom, const c10::optional<torch::Tensor>& p16,
                     int64_t global_offset, const torch::Tensor& seg_start, const torch::Tensor& seg_wd,
                     const torch::Tensor& seg_lr_mult, double lr, double momentum, bool first_step,
                     const c10::optional<torch::Tensor>& grad_scale,
                     const c10::optional<torch::Tensor>& skip) {
  c10::cuda::CUDAGuard guard(p.device());
  CHK(mlb_sgd_flat(p.data_ptr<float>(), g.data_ptr<float>(), mom.data_ptr<float>(), const_cast<void*>(optp(p16)),
                   p16.has_value() ? dt(*p16) : 0, p.numel(), global_offset, (const long long*)seg_start.data_ptr(),
                   seg_wd.data_ptr<float>(), seg_lr_mult.data_ptr<float>(), (int)seg_wd.numel(), (float)lr,
                   (float)momentum, first_step,
                   grad_scale.has_value() ? grad_scale->data_ptr<float>() : nullptr,
                   skip.has_value() ? skip->data_ptr<int>() : nullptr, cur()));
}
static void sqnorm_flat(const torch::Tensor& x, int64_t global_offset, const c10::optional<torch::Tensor>& seg_start,
                        const c10::optional<torch::Tensor>& seg_weight, torch::Tensor& workspace, torch::Tensor& out,
                        bool accumulate) {
  c10::cuda::CUDAGuard guard(x.device());
  CHK(mlb_sqnorm_flat(dt(x), x.data_ptr(), x.numel(), global_offset,
                      seg_start.has_value() ? (const long long*)seg_start->data_ptr() : nullptr,
                      seg_weight.has_value() ? seg_weight->data_ptr<float>() : nullptr,
                      seg_weight.has_value() ? (int)seg_weight->numel() : 0, workspace.data_ptr<float>(),
                      out.data_ptr<float>(), accumulate, cur()));
}
static void clip_coef(const torch::Tensor& total_sq, double max_norm, torch::Tensor& norm_out, torch::Tensor& coef_out,
                      const c10::optional<torch::Tensor>& found_inf, double extra_scale) {
  c10::cuda::CUDAGuard guard(total_sq.device());
  CHK(mlb_clip_coef(total_sq.data_ptr<float>(), (float)max_norm, norm_out.data_ptr<float>(), coef_out.data_ptr<float>(),
                    found_inf.has_value() ? found_inf->data_ptr<int>() : nullptr, (float)extra_scale, cur()));
}
static void scale_cast(const torch::Tensor& x, torch::Tensor& y, double scale, const c10::optional<torch::Tensor>& scale_ptr) {
  c10::cuda::CUDAGuard guard(x.device());
  CHK(mlb_scale_cast(dt(x), dt(y), x.data_ptr(), y.data_ptr(), x.numel(), (float)scale,
                     scale_ptr.has_value() ? scale_ptr->data_ptr<float>() : nullptr, cur()));
}
static void accumulate(const torch::Tensor& x, torch::Tensor& y) {
  c10::cuda::CUDAGuard guard(x.device());
  CHK(mlb_accumulate(dt(x), x.data_ptr(), y.data_ptr<float>(), x.numel(), cur()));
}
static void softmax_fwd(const torch::Tensor& x, torch::Tensor& y, const c10::optional<torch::Tensor>& mask, double scale,
                        int64_t sq, int64_t sk, int64_t np, int64_t mode) {
  c10::cuda::CUDAGuard guard(x.device());
  CHK(mlb_softmax_fwd(dt(x), x.data_ptr(), y.data_ptr(), mask.has_value() ? (const unsigned char*)mask->data_ptr() : nullptr,
                      (float)scale, x.numel() / sk, (int)sq, (int)sk, (int)np, mask.has_value() ? (int)mask->size(0) : 1,
                      (int)mode, cur()));
}
static void softmax_bwd(torch::Tensor& dy, const torch::Tensor& y, double scale, int64_t sk) {
  c10::cuda::CUDAGuard guard(y.device());
  CHK(mlb_softmax_bwd(dt(y), dy.data_ptr(), y.data_ptr(), (float)scale, y.numel() / sk, (int)sk, cur()));
}

void register_attention(pybind11::module_& m);
void register_comm(pybind11::module_& m);

static torch::Tensor gemm2_debug() {
  auto t = torch::zeros({512, 8}, torch::dtype(torch::kInt64));
  CHK(mlb_gemm2_debug_read(reinterpret_cast<unsigned long long*>(t.data_ptr<int64_t>()), 512 * 8));
  return t;
}

PYBIND11_MODULE(_C_b200, m) {
  m.def("gemm", &gemm);
  m.def("gemm2_debug", &gemm2_debug);
  m.def("norm_fwd", &norm_fwd);
  m.def("norm_bwd", &norm_bwd);
  m.def("rope_qkv", &rope_qkv);
  m.def("rope_heads", &rope_heads);
  m.def("glu_fwd", &glu_fwd);
  m.def("glu_bwd", &glu_bwd);
  m.def("gelu", &gelu);
  m.def("bias_dropout_add", &bias_dropout_add);
  m.def("embedding_fwd", &embedding_fwd);
  m.def("embedding_bwd", &embedding_bwd);
  m.def("ce_stats", &ce_stats);
  m.def("ce_bwd", &ce_bwd);
  m.def("adamw_flat", &adamw_flat);
  m.def("sgd_flat", &sgd_flat);
  m.def("sqnorm_flat", &sqnorm_flat);
  m.def("clip_coef", &clip_coef);
  m.def("scale_cast", &scale_cast);
  m.def("accumulate", &accumulate);
  m.def("softmax_fwd", &softmax_fwd);
  m.def("softmax_bwd", &softmax_bwd);
  m.def("num_sms", &num_sms);
  register_attention(m);
  register_comm(m);
}

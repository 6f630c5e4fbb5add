// Race check of the SIMT kernels on the CPU: the kernel sources (compiled for the host by host_build.py, here with
// -fsanitize=thread) run one OS thread per CUDA thread, so a missing __syncthreads / __syncwarp around shared memory is
// a data race ThreadSanitizer reports -- a CPU-side stand-in for compute-sanitizer's racecheck on the kernels it can
// execute.  This driver only launches the kernels on random data (numerics are checked in test_kernel_emulation.py);
// the process exits non-zero if ThreadSanitizer saw a race.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef void* cudaStream_t;
extern "C" {
int mlb_attn_decode(int dtype, const void* q, const void* k, const void* v, const long long* q_str, const long long* k_str,
                    const long long* v_str, int batch, int sq, int sk, int nq, int nkv, int head_dim, int window,
                    float softmax_scale, int n_splits, int keys_per_split, float* part_o, float* part_ml, void* out,
                    cudaStream_t stream);
int mlb_norm_fwd(int dtype, const void* x, const void* res_in, const void* w, const void* b, void* y, void* res_out,
                 float* mean, float* rstd, int rows, int H, float eps, int rms, cudaStream_t st);
int mlb_norm_bwd(int dtype, const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                 const void* dres, void* dx, void* dw, void* db, float* workspace, int parts, int rows, int H, int rms,
                 cudaStream_t st);
int mlb_ce_stats(int dtype, const void* logits, const long long* target, float* stats, int rows, int Vp, int vocab_start,
                 long long row_stride, cudaStream_t st);
int mlb_ce_bwd(int dtype, const void* logits, void* out, const long long* target, const float* M, const float* logS,
               const float* g, int rows, int Vp, int vocab_start, float smoothing, int vocab_size, long long row_stride,
               cudaStream_t st);
int mlb_softmax_fwd(int dtype, const void* x, void* y, const unsigned char* mask, float scale, long long rows, int sq,
                    int sk, int np, int mask_batch, int mode, cudaStream_t st);
int mlb_softmax_bwd(int dtype, void* dy, const void* y, float scale, long long rows, int sk, cudaStream_t st);
int mlb_sqnorm_flat(int dtype, const void* x, long long n, long long global_offset, const long long* seg_start,
                    const float* seg_weight, int nseg, float* workspace, float* out, int accumulate, cudaStream_t st);
int mlb_embedding_bwd(int dtype, const long long* ids, const void* dout, float* dweight, int batch, int seq, int H,
                      long long vocab_start, long long rows_local, int sbh, cudaStream_t st);
}

static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static uint16_t bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static std::vector<uint16_t> rand_bf16(size_t n) { std::vector<uint16_t> v(n); for (auto& x : v) x = bf16(frand()); return v; }
static std::vector<float> rand_f32(size_t n) { std::vector<float> v(n); for (auto& x : v) x = frand(); return v; }
#define RUN(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d\n", #call, rc_); return 2; } } while (0)

int main(int argc, char** argv) {
  const bool only_decode = argc > 1 && !strcmp(argv[1], "decode");
  {  // decode attention: 16 rows per group (two passes), 3 splits, window, both head dims
    for (int hn : {128, 64}) {
      const int b = 1, sq = 2, sk = 150, nq = 16, nkv = 2, splits = 3, kps = 64, g = nq / nkv;
      auto q = rand_bf16((size_t)b * sq * nq * hn), k = rand_bf16((size_t)b * sk * nkv * hn), v = rand_bf16((size_t)b * sk * nkv * hn);
      std::vector<float> po((size_t)b * nkv * splits * sq * g * hn), pml((size_t)b * nkv * splits * sq * g * 2);
      std::vector<uint16_t> out((size_t)b * sq * nq * hn);
      long long qs[3] = {(long long)sq * nq * hn, (long long)nq * hn, hn}, ks[3] = {(long long)sk * nkv * hn, (long long)nkv * hn, hn};
      RUN(mlb_attn_decode(0, q.data(), k.data(), v.data(), qs, ks, ks, b, sq, sk, nq, nkv, hn, 40, 0.1f, splits, kps,
                          po.data(), pml.data(), out.data(), nullptr));
    }
  }
  if (only_decode) return 0;
#ifndef RACE_DECODE_ONLY   // (the mutation check links the decode kernel only)
  {  // norms (block reductions, per-part partial sums, column sums)
    const int rows = 5, H = 1024;
    auto x = rand_bf16((size_t)rows * H), res = rand_bf16((size_t)rows * H), w = rand_bf16(H), bias = rand_bf16(H), dy = rand_bf16((size_t)rows * H);
    std::vector<uint16_t> y((size_t)rows * H), ro((size_t)rows * H), dx((size_t)rows * H), dw(H), db(H);
    std::vector<float> mean(rows), rstd(rows), ws((size_t)2 * 296 * H);
    for (int rms : {1, 0}) {
      RUN(mlb_norm_fwd(0, x.data(), res.data(), w.data(), rms ? nullptr : bias.data(), y.data(), ro.data(),
                       rms ? nullptr : mean.data(), rstd.data(), rows, H, 1e-5f, rms, nullptr));
      RUN(mlb_norm_bwd(0, dy.data(), ro.data(), w.data(), rms ? nullptr : mean.data(), rstd.data(), res.data(), dx.data(),
                       dw.data(), rms ? nullptr : db.data(), ws.data(), rows, rows, H, rms, nullptr));
    }
  }
  {  // cross entropy
    const int rows = 3, V = 2056;
    auto logits = rand_f32((size_t)rows * V);
    std::vector<long long> target = {5, 2055, 9999};
    std::vector<float> stats((size_t)rows * 4), M(rows, 1.f), logS(rows, 2.f), g(rows, 1.f), out((size_t)rows * V);
    RUN(mlb_ce_stats(2, logits.data(), target.data(), stats.data(), rows, V, 0, V, nullptr));
    RUN(mlb_ce_bwd(2, logits.data(), out.data(), target.data(), M.data(), logS.data(), g.data(), rows, V, 0, 0.1f, V, V, nullptr));
  }
  {  // softmax family (64-, 128- and 256-thread blocks)
    for (int sk : {200, 700, 1100}) {
      const int sq = 3, np = 2, b = 1;
      auto x = rand_f32((size_t)b * np * sq * sk);
      std::vector<float> y(x.size()), dy = rand_f32(x.size());
      std::vector<unsigned char> mask((size_t)b * sq * sk, 0);
      for (int mode : {0, 1, 2}) RUN(mlb_softmax_fwd(2, x.data(), y.data(), mask.data(), 0.5f, (long long)b * np * sq, sq, sk, np, b, mode, nullptr));
      RUN(mlb_softmax_bwd(2, dy.data(), y.data(), 0.5f, (long long)b * np * sq, sk, nullptr));
    }
  }
  {  // weighted squared norm (block reduce + finalize) and the scatter-add of the embedding backward (atomics)
    const long long n = 5000;
    auto x = rand_f32(n);
    std::vector<long long> seg = {0, 3000, n};
    std::vector<float> wgt = {1.f, 0.5f}, ws(148 * 8), out(1);
    RUN(mlb_sqnorm_flat(2, x.data(), n, 0, seg.data(), wgt.data(), 2, ws.data(), out.data(), 0, nullptr));
    const int b = 2, s = 40, H = 64;
    std::vector<long long> ids((size_t)b * s, 7);                     // every token hits the same row
    auto dout = rand_bf16((size_t)b * s * H);
    std::vector<float> dw((size_t)16 * H);
    RUN(mlb_embedding_bwd(0, ids.data(), dout.data(), dw.data(), b, s, H, 0, 16, 0, nullptr));
  }
#endif
  return 0;
}

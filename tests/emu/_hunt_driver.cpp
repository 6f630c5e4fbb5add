// ThreadSanitizer run of the tcgen05 attention kernels on the functional model (tcgen05_model.h): forward (single-tile
// and two-tile kernels), delta, dK/dV and dQ, with dropout on, on random data.  Tensor-memory and shared-memory accesses
// of the model are plain loads / stores ordered ONLY by what the kernels synchronise on (mbarriers, named barriers,
// __syncthreads), so a consumer that is not ordered after its producer is a data race here.  Exit code != 0 = race.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef void* cudaStream_t;
extern "C" {
int mlb_attn_fwd_ex(const void* q, const void* k, const void* v, const long long* q_str, const long long* k_str,
                    const long long* v_str, int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map,
                    int q_per_kv, int seq, int batch, int heads, int window, float softmax_scale, void* out,
                    long long out_s_stride, long long out_b_stride, float* lse, int head_dim, int fp16, float dropout_p,
                    unsigned long long seed, cudaStream_t stream);
int mlb_attn_bwd_ex(const void* q, const void* k, const void* v, const void* o, const void* dout, const long long* q_str,
                    const long long* k_str, const long long* v_str, const long long* o_str, const long long* do_str,
                    int q_map_heads, int k_map_heads, int v_map_heads, const int* head_map, int q_per_kv, int seq,
                    int batch, int heads, int window, float softmax_scale, const float* lse, float* delta, void* dq,
                    void* dk, void* dv, const long long* dq_str, const long long* dk_str, const long long* dv_str,
                    int head_dim, int fp16, float dropout_p, unsigned long long seed, cudaStream_t stream);
}
static uint32_t st = 1;
static float fr() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static uint16_t bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
struct Buf {
  std::vector<uint16_t> raw; uint16_t* p;
  explicit Buf(size_t n, bool rnd) : raw(n + 16) {
    p = (uint16_t*)(((uintptr_t)raw.data() + 15) & ~(uintptr_t)15);
    for (size_t i = 0; i < n; ++i) p[i] = rnd ? bf(fr()) : 0;
  }
};

int main(int argc, char** argv) {
  const int s = argc > 1 ? atoi(argv[1]) : 384, hn = argc > 2 ? atoi(argv[2]) : 128;
  const int b = 1, n = argc > 3 ? atoi(argv[3]) : 2, nkv = argc > 4 ? atoi(argv[4]) : 1, g = n / nkv; const int win = argc > 5 ? atoi(argv[5]) : s / 2;
  const float p = 0.2f; const unsigned long long seed = 99;
  Buf q((size_t)s * n * hn, true), k((size_t)s * nkv * hn, true), v((size_t)s * nkv * hn, true), dout((size_t)s * n * hn, true);
  Buf out((size_t)s * n * hn, false), dq((size_t)s * n * hn, false), dk((size_t)s * nkv * hn, false), dv((size_t)s * nkv * hn, false);
  std::vector<float> lse((size_t)n * s), delta((size_t)n * s);
  long long qs[3] = {hn, (long long)n * hn, (long long)s * n * hn}, ks[3] = {hn, (long long)nkv * hn, (long long)s * nkv * hn};
  int hm[6] = {g, 0, 1, 0, 1, 0};
  int rc = mlb_attn_fwd_ex(q.p, k.p, v.p, qs, ks, ks, n, nkv, nkv, hm, g, s, b, n, win, 0.1f, out.p, (long long)b * n * hn,
                           (long long)n * hn, lse.data(), hn, 0, p, seed, nullptr);
  if (rc) return 2;
  rc = mlb_attn_bwd_ex(q.p, k.p, v.p, out.p, dout.p, qs, ks, ks, qs, qs, n, nkv, nkv, hm, g, s, b, n, win, 0.1f, lse.data(),
                       delta.data(), dq.p, dk.p, dv.p, qs, ks, ks, hn, 0, p, seed, nullptr);
  return rc ? 3 : 0;
}

// ThreadSanitizer run of the 2-CTA (cta_group::2, cluster of two CTAs) tcgen05 GEMM kernel on the functional model: the
// pair runs concurrently -- TMA loads of both CTAs completing on the leader's barriers, the leader's MMAs writing both
// tensor memories, multicast commits, the remote arrive of the epilogue warps, TMA-store epilogue through swizzled
// shared memory -- several tiles per pair.  Exit code != 0 = data race.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef void* cudaStream_t;
extern "C" int mlb_gemm_bf16_2cta(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                                  int a_mn_major, int b_mn_major, int epilogue, int fp16, int num_sms, cudaStream_t stream);
static uint32_t st = 7;
static uint16_t rb() { st = st * 1664525u + 1013904223u; float f = ((st >> 8) & 0xFFFF) / 32768.0f - 1.0f; uint32_t u; memcpy(&u, &f, 4); return u >> 16; }
int main(int argc, char** argv) {
  const int epi = argc > 1 ? atoi(argv[1]) : 0;            // 0: 16-bit TMA store, 1: fp32 reduce-add, 3: 16-bit accumulate
  const int M = 768, N = 256, K = 320;                     // 3 tiles of 256 x 256 on one CTA pair, 5 k-blocks each
  std::vector<uint16_t> a((size_t)M * K + 8), b((size_t)N * K + 8);
  std::vector<float> c((size_t)M * N + 8, 0.f);
  for (auto& x : a) x = rb();
  for (auto& x : b) x = rb();
  auto al = [](auto& v) { return (void*)(((uintptr_t)v.data() + 15) & ~(uintptr_t)15); };
  return mlb_gemm_bf16_2cta(al(a), al(b), al(c), M, N, K, K, K, N, 0, 0, epi, 0, 2, nullptr);
}

// ThreadSanitizer run of the 1-CTA tcgen05 GEMM kernel on the functional model: a persistent CTA loops over several
// tiles (4-stage TMA ring, two accumulator stages in tensor memory, 4 epilogue warps).  Exit code != 0 = data race.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef void* cudaStream_t;
extern "C" int mlb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                             int a_mn_major, int b_mn_major, int epilogue, int block_n, int fp16, int num_sms,
                             cudaStream_t stream);
static uint32_t st = 7;
static uint16_t rb() { st = st * 1664525u + 1013904223u; float f = ((st >> 8) & 0xFFFF) / 32768.0f - 1.0f; uint32_t u; memcpy(&u, &f, 4); return u >> 16; }
int main() {
  const int M = 384, N = 256, K = 320;                    // 6 tiles of 128 x 128 on 2 "SMs": 3 tiles per CTA, 5 k-blocks
  std::vector<uint16_t> a((size_t)M * K + 8), b((size_t)N * K + 8), c((size_t)M * N + 8);
  for (auto& x : a) x = rb();
  for (auto& x : b) x = rb();
  auto al = [](std::vector<uint16_t>& v) { return (uint16_t*)(((uintptr_t)v.data() + 15) & ~(uintptr_t)15); };
  return mlb_gemm_bf16(al(a), al(b), al(c), M, N, K, K, K, N, 0, 0, 0, 128, 0, 2, nullptr);
}

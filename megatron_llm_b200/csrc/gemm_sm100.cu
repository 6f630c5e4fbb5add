// Host launchers for the tcgen05 GEMM (see gemm_sm100.cuh).
#include "gemm_sm100.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdio.h>

namespace mlb {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 2-D bf16 tensor map: inner dim `inner` (contiguous), outer dim `outer` with row stride `ld` elements.
// box = {box_inner (=64 elements = 128B swizzle span), box_outer}
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                  cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N>;
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int num_m = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  int grid = num_m * num_n;
  if (grid > num_sms) grid = num_sms;
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA, tmB, p);
  return (int)cudaGetLastError();
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                        cudaStream_t stream) {
  switch (epi) {
    case EPI_BF16: return launch<BLOCK_N, A_MN, B_MN, EPI_BF16>(tmA, tmB, p, num_sms, stream);
    case EPI_F32_ACCUM: return launch<BLOCK_N, A_MN, B_MN, EPI_F32_ACCUM>(tmA, tmB, p, num_sms, stream);
    case EPI_F32: return launch<BLOCK_N, A_MN, B_MN, EPI_F32>(tmA, tmB, p, num_sms, stream);
    case EPI_BF16_ACCUM: return launch<BLOCK_N, A_MN, B_MN, EPI_BF16_ACCUM>(tmA, tmB, p, num_sms, stream);
  }
  return -2;
}

template <int BLOCK_N>
static int dispatch_major(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (!a_mn && !b_mn) return dispatch_epi<BLOCK_N, false, false>(epi, tmA, tmB, p, num_sms, stream);
  if (!a_mn && b_mn) return dispatch_epi<BLOCK_N, false, true>(epi, tmA, tmB, p, num_sms, stream);
  if (a_mn && b_mn) return dispatch_epi<BLOCK_N, true, true>(epi, tmA, tmB, p, num_sms, stream);
  return dispatch_epi<BLOCK_N, true, false>(epi, tmA, tmB, p, num_sms, stream);
}

}  // namespace mlb

// D[M,N] = A * B^T with fp32 accumulation on tcgen05.
//   a_mn_major = 0: A is [M, lda] (K contiguous)   | 1: A is [K, lda] (M contiguous)
//   b_mn_major = 0: B is [N, ldb] (K contiguous)   | 1: B is [K, ldb] (N contiguous)
// `comm` may be null (plain GEMM).  Returns 0 on success.
extern "C" int mlb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                             int ldc, int a_mn_major, int b_mn_major, int epilogue, int block_n,
                             const mlb::GemmComm* comm, int num_sms, cudaStream_t stream) {
  using namespace mlb;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap tmA, tmB;
  int r;
  if (!a_mn_major) r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, GEMM_BLOCK_M);
  else r = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BLOCK_K);
  if (r) return 1000 + r;
  if (block_n != 128 && block_n != 256) {
    // heuristic: fewest waves, then the larger tile
    auto waves = [&](int bn) {
      long tiles = (long)((M + 127) / 128) * ((N + bn - 1) / bn);
      long w = (tiles + num_sms - 1) / num_sms;
      return w * bn;  // cost ~ waves * tile width
    };
    block_n = (waves(256) <= waves(128)) ? 256 : 128;
  }
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, block_n);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return 2000 + r;
  GemmParams p;
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  if (comm) p.comm = *comm;
  else {
    p.comm.a_ready_flags = nullptr; p.comm.a_chunk_rows = 0; p.comm.a_ready_epoch = 0; p.comm.m_rotate_blocks = 0;
    p.comm.out_chunk_rows = 0;
    for (int i = 0; i < GEMM_MAX_PEERS; ++i) { p.comm.out_ptrs[i] = nullptr; p.comm.tile_counters[i] = nullptr; }
  }
  if (block_n == 256) return dispatch_major<256>(a_mn_major, b_mn_major, epilogue, tmA, tmB, p, num_sms, stream);
  return dispatch_major<128>(a_mn_major, b_mn_major, epilogue, tmA, tmB, p, num_sms, stream);
}

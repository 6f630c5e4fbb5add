// Host launchers for the tcgen05 GEMM (see gemm_sm100.cuh).
#include "gemm_sm100.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdio.h>
#include <string.h>

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
  mlb_bind_context();
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

// 2-D output map (fp32 or bf16), 128-byte inner box (32 fp32 / 64 bf16 columns), 128B swizzle: epilogue TMA stores
int make_tmap_2d_out(CUtensorMap* tm, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_outer) {
  auto fn = get_encode_fn();
  mlb_bind_context();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, int MODE = MODE_PLAIN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                  cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N>;
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, EPI, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int num_m = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  int grid = num_m * num_n;
  if (MODE == MODE_AG_GEMM) {
    const int compute = num_sms - p.comm.num_comm_ctas;
    if (grid > compute) grid = compute;
    grid += p.comm.num_comm_ctas;   // trailing CTAs are the NVLink pullers; all CTAs must be co-resident
  } else if (grid > num_sms) {
    grid = num_sms;
  }
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

template <int BLOCK_N>
static int dispatch_fused(int mode, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                          int num_sms, cudaStream_t stream) {
  if (mode == MODE_AG_GEMM)
    return b_mn ? launch<BLOCK_N, false, true, EPI_BF16, MODE_AG_GEMM>(tmA, tmB, p, num_sms, stream)
                : launch<BLOCK_N, false, false, EPI_BF16, MODE_AG_GEMM>(tmA, tmB, p, num_sms, stream);
  if (mode == MODE_GEMM_RS)
    return b_mn ? launch<BLOCK_N, false, true, EPI_BF16, MODE_GEMM_RS>(tmA, tmB, p, num_sms, stream)
                : launch<BLOCK_N, false, false, EPI_BF16, MODE_GEMM_RS>(tmA, tmB, p, num_sms, stream);
  return -3;
}

static int pick_block_n(int M, int N, int num_sms) {
  // cost ~ waves x tile width / efficiency.  128x128 tiles read A+B from smem at the tensor pipe's full rate (measured
  // ~0.78 of the 128x256 kernel's throughput on B200), so they only win when they save a whole wave.
  auto cost = [&](int bn) {
    long tiles = (long)((M + 127) / 128) * ((N + bn - 1) / bn);
    long w = (tiles + num_sms - 1) / num_sms;
    return (double)w * bn / (bn == 256 ? 1.0 : 0.78);
  };
  return (cost(256) <= cost(128)) ? 256 : 128;
}

}  // namespace mlb

// D[M,N] = A * B^T with fp32 accumulation on tcgen05.
//   a_mn_major = 0: A is [M, lda] (K contiguous)   | 1: A is [K, lda] (M contiguous)
//   b_mn_major = 0: B is [N, ldb] (K contiguous)   | 1: B is [K, ldb] (N contiguous)
// `comm` may be null (plain GEMM).  Returns 0 on success.
extern "C" int mlb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                             int ldc, int a_mn_major, int b_mn_major, int epilogue, int block_n,
                             int fp16, int num_sms, cudaStream_t stream) {
  using namespace mlb;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap tmA, tmB;
  int r;
  if (!a_mn_major) r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, GEMM_BLOCK_M);
  else r = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BLOCK_K);
  if (r) return 1000 + r;
  if (block_n != 128 && block_n != 256) block_n = pick_block_n(M, N, num_sms);
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, block_n);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return 2000 + r;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.fp16 = fp16;
  if (block_n == 256) return dispatch_major<256>(a_mn_major, b_mn_major, epilogue, tmA, tmB, p, num_sms, stream);
  return dispatch_major<128>(a_mn_major, b_mn_major, epilogue, tmA, tmB, p, num_sms, stream);
}

// Fused GEMM + collective (see gemm_sm100.cuh).  A is always K-major ([M, lda]); bf16 output.
//   mode 1 (all-gather -> GEMM): A = comm->ag_dst (local gathered buffer the puller CTAs fill), C = out [M, ldc]
//   mode 2 (GEMM -> reduce-scatter): A = local activations; tiles go to comm->rs_dst[], C unused
extern "C" int mlb_gemm_bf16_fused(int mode, const void* A, const void* B, void* C, int M, int N, int K, int lda,
                                   int ldb, int ldc, int b_mn_major, const mlb::GemmComm* comm, int num_sms,
                                   cudaStream_t stream) {
  using namespace mlb;
  if (M <= 0 || N <= 0 || K <= 0 || comm == nullptr) return -1;
  CUtensorMap tmA, tmB;
  int r = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, GEMM_BLOCK_M);
  if (r) return 1000 + r;
  const int compute_sms = num_sms - (mode == MODE_AG_GEMM ? comm->num_comm_ctas : 0);
  const int block_n = pick_block_n(M, N, compute_sms);
  if (!b_mn_major) r = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, block_n);
  else r = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BLOCK_K);
  if (r) return 2000 + r;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.comm = *comm;
  if (block_n == 256) return dispatch_fused<256>(mode, b_mn_major, tmA, tmB, p, num_sms, stream);
  return dispatch_fused<128>(mode, b_mn_major, tmA, tmB, p, num_sms, stream);
}

// Bindings for the fused GEMM + collective kernels (gemm_sm100.cuh MODE_AG_GEMM / MODE_GEMM_RS) and the
// peer-memory data-parallel gradient reduction (comm.cu).
#include <algorithm>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdlib>

#include <cstdio>
#include <stdexcept>

#include <cuda_runtime.h>
#include <string.h>

#include "gemm_types.h"

extern "C" {
int mlb_gemm_bf16_2cta_ag(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                          int b_mn_major, const mlb::GemmComm* comm, int num_sms, cudaStream_t stream);
int mlb_copy2(const void* src, void* d1, void* d2, long long bytes, int num_sms, cudaStream_t stream);
int mlb_p2p_bench(int mode, const void* src, void* dst, long long bytes, int ctas, int piece_bytes, int stages,
                  int row_bytes, long long dst_stride, cudaStream_t stream);
int mlb_set_ints3(int* dst, int a, int b, int c, cudaStream_t stream);
int mlb_gemm_bf16_2cta_rs(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int b_mn_major,
                          mlb::GemmComm* comm, int prev_total, int num_sms, cudaStream_t stream);
int mlb_gemm_bf16_fused(int mode, const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb,
                        int ldc, int b_mn_major, const mlb::GemmComm* comm, int num_sms, cudaStream_t stream);
int mlb_dp_reduce_nvls(int reduce_scatter, float* local, float* mc, int* pad_local, const long long* pad_peer_ptrs,
                       long long n, int rank, int world, int epoch, float scale, int num_ctas, cudaStream_t st);
int mlb_peer_barrier(int* pad_local, const long long* pad_peer_ptrs, int rank, int world, int epoch, int slot,
                     cudaStream_t st);
int mlb_dp_reduce(int reduce_scatter, float* local, const long long* peer_ptrs, int* pad_local,
                  const long long* pad_peer_ptrs, long long n, int rank, int world, int epoch, float scale,
                  int num_ctas, cudaStream_t st);
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

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static void fill_pads(mlb::GemmComm& c, int64_t pad_local, const std::vector<int64_t>& pad_peers) {
  c.pad_local = reinterpret_cast<int*>(pad_local);
  for (size_t i = 0; i < pad_peers.size() && i < (size_t)mlb::GEMM_MAX_PEERS; ++i)
    c.pad_peer[i] = reinterpret_cast<int*>(pad_peers[i]);
}

// Tile-group height of the 2-CTA fused kernels (256-row blocks).  A group shares every B panel, so a tall group cuts
// the B traffic; but its first tiles need all of its rows, so it should not reach past the rows that are available
// first (one rank's shard) unless B is too large to be re-read from L2 for every shard.
static int pick_group_blocks(int64_t rows_per_rank, int64_t b_bytes) {
  static const char* forced = getenv("MLB200_FUSED_GROUP");      // (tuning aid)
  if (forced) return atoi(forced);
  const int per_rank = (int)(rows_per_rank / 256);
  if (b_bytes > (24LL << 20) || per_rank >= 4) return 4;
  return per_rank >= 2 ? 2 : 1;
}

// Number of puller CTAs for all-gather -> GEMM.  Measured on B200 (profiles/nvlink_n2_r2.jsonl): a bulk-copy puller CTA
// moves ~44 GB/s alone and ~29 GB/s when 16 share the link, which saturates near 470 GB/s per direction for 16 MB
// messages.  The GEMM consumes its own rows first (1/world of the time), then the remote rows in arrival order, so the
// pullers must (a) sustain the consumption rate of the remote rows and (b) have the first tile group's chunks there
// when the own rows are done; every puller costs the GEMM one SM.  The first version of this model minimised
// max(gemm, pull + one group) with an optimistic per-puller rate and left no slack: at TP=2 the producers were blocked
// ~70 us per call (exposed_tp_collective in bench.py).  Now: the rate requirement with 1.4x headroom, at least 4.
static int pick_comm_ctas(int64_t M, int64_t N, int64_t K, int64_t rows_per_rank, int world, int max_ctas, int sms) {
  static const char* fixed = getenv("MLB200_AG_CTAS_FIXED");
  if (fixed) return std::min(max_ctas, atoi(fixed));
  constexpr double PULL_GBPS = 28.0, GEMM_TFLOPS = 1350.0, HEADROOM = 1.4;
  const double gemm_us = 2.0 * M * N * K / (GEMM_TFLOPS * 1e6) + 4.0;
  const double remote_bytes = (double)(world - 1) * rows_per_rank * K * 2;
  const double remote_window_us = gemm_us * (world - 1) / world;                       // time spent on remote rows
  const double own_us = gemm_us / world;                                                // ... and before them
  const double group_bytes = 256.0 * pick_group_blocks(rows_per_rank, N * K * 2) * K * 2;
  const double need_gbps = std::max(remote_bytes / (remote_window_us * 1e3), group_bytes / (own_us * 1e3));
  int c = (int)(need_gbps * HEADROOM / PULL_GBPS + 0.999);
  c = (c + 1) & ~1;
  (void)sms;
  return std::max(4, std::min(c, max_ctas & ~1));
}

// out[M, N] = all_gather(shards)[M, K] @ W^T (b_mn=false, W [N, K]) or @ W (b_mn=true, W [K, N]).
// `gathered` is the local [M, K] buffer the puller CTAs fill from the peers' published shards `ag_src`.
static void fused_ag_gemm(torch::Tensor& gathered, const torch::Tensor& weight, torch::Tensor& out, bool b_mn,
                          const std::vector<int64_t>& ag_src, int64_t rows_per_rank, torch::Tensor& chunk_flags,
                          torch::Tensor& read_counters, int64_t pad_local, const std::vector<int64_t>& pad_peers,
                          int64_t rank, int64_t world, int64_t epoch, int64_t num_comm_ctas, int64_t sms,
                          int64_t state_ptr, int64_t stats_ptr) {
  c10::cuda::CUDAGuard guard(gathered.device());
  const int M = gathered.size(0), K = gathered.size(1);
  const int N = b_mn ? weight.size(1) : weight.size(0);
  TORCH_CHECK(gathered.is_contiguous() && out.stride(1) == 1 && weight.stride(1) == 1);
  TORCH_CHECK(rows_per_rank % mlb::GEMM_BLOCK_M == 0, "rows per rank must be a multiple of 128");
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0);
  mlb::GemmComm c;
  memset(&c, 0, sizeof(c));
  c.rank = rank; c.world = world; c.epoch = epoch;
  c.num_comm_ctas = pick_comm_ctas(M, N, K, rows_per_rank, (int)world, (int)num_comm_ctas, sms > 0 ? (int)sms : sm_count());
  num_comm_ctas = c.num_comm_ctas;
  c.m_group_blocks = pick_group_blocks(rows_per_rank, (int64_t)N * K * 2);
  c.state = reinterpret_cast<const int*>(state_ptr);
  c.stats = reinterpret_cast<unsigned long long*>(stats_ptr);
  c.m_rotate_blocks = (int)(rank * rows_per_rank / mlb::GEMM_BLOCK_M);
  for (int i = 0; i < world; ++i) c.ag_src[i] = reinterpret_cast<const void*>(ag_src[i]);
  c.ag_dst = gathered.data_ptr();
  c.ag_rows_per_rank = rows_per_rank;
  c.ag_row_bytes = K * 2;
  c.ag_chunk_flags = chunk_flags.data_ptr<int>();
  c.ag_read_counters = read_counters.data_ptr<int>();
  fill_pads(c, pad_local, pad_peers);
  // the 2-CTA (256x256-tile, TMA-store epilogue) kernel when the shard is a whole number of its row blocks
  static const bool force_1cta = getenv("MLB200_FUSED_1CTA") != nullptr;
  if (!force_1cta && rows_per_rank % 256 == 0 && num_comm_ctas % 2 == 0 && N >= 256 && (out.stride(0) * 2) % 16 == 0) {
    CHK(mlb_gemm_bf16_2cta_ag(gathered.data_ptr(), weight.data_ptr(), out.data_ptr(), M, N, K, K,
                              (int)weight.stride(0), (int)out.stride(0), b_mn, &c, sms > 0 ? (int)sms : sm_count(),
                              cur()));
    return;
  }
  CHK(mlb_gemm_bf16_fused(mlb::MODE_AG_GEMM, gathered.data_ptr(), weight.data_ptr(), out.data_ptr(), M, N, K, K,
                          (int)weight.stride(0), (int)out.stride(0), b_mn, &c, sms > 0 ? (int)sms : sm_count(), cur()));
}

// NVLS form of the all-gather -> GEMM (2-CTA kernel only): `gathered` = this rank's symmetric gather buffer of the call's
// parity viewed as [M, K]; the pusher CTAs multimem.st `x_shard` into everybody's buffer (`mc_dst`) and release the
// chunk flags `flags[d]`; the GEMM reads A from `gathered`.  Returns false if the shape needs the pull kernel.
static bool fused_ag_gemm_nvls(torch::Tensor& gathered, const torch::Tensor& x_shard, const torch::Tensor& weight,
                               torch::Tensor& out, bool b_mn, int64_t mc_dst, const std::vector<int64_t>& flags,
                               torch::Tensor& done_counter, int64_t rows_per_rank, int64_t pad_local,
                               const std::vector<int64_t>& pad_peers, int64_t rank, int64_t world, int64_t epoch,
                               int64_t num_push_ctas, int64_t sms, int64_t state_ptr, int64_t stats_ptr) {
  c10::cuda::CUDAGuard guard(gathered.device());
  const int M = gathered.size(0), K = gathered.size(1);
  const int N = b_mn ? weight.size(1) : weight.size(0);
  TORCH_CHECK(gathered.is_contiguous() && x_shard.is_contiguous() && out.stride(1) == 1 && weight.stride(1) == 1);
  TORCH_CHECK(x_shard.size(0) == rows_per_rank && x_shard.size(1) == K && M == rows_per_rank * world);
  TORCH_CHECK((uintptr_t)x_shard.data_ptr() % 16 == 0 && mc_dst % 16 == 0 && K % 8 == 0 && N % 8 == 0);
  if (rows_per_rank % 256 != 0 || N < 256 || (out.stride(0) * 2) % 16 != 0 || num_push_ctas < 2 || (num_push_ctas & 1))
    return false;
  const int per_rank = (int)(rows_per_rank / 256);
  const int G = per_rank % 4 == 0 ? 4 : (per_rank % 2 == 0 ? 2 : 1);      // whole tile groups per source rank
  mlb::GemmComm c;
  memset(&c, 0, sizeof(c));
  c.rank = rank; c.world = world; c.epoch = epoch;
  c.num_comm_ctas = (int)num_push_ctas;
  c.m_group_blocks = G;
  c.m_stripe = 1;
  c.state = reinterpret_cast<const int*>(state_ptr);
  c.stats = reinterpret_cast<unsigned long long*>(stats_ptr);
  c.ag_nvls = 1;
  c.ag_local_src = x_shard.data_ptr();
  c.ag_mc_dst = reinterpret_cast<void*>(mc_dst);
  for (int i = 0; i < world; ++i) c.ag_flag_peer[i] = reinterpret_cast<int*>(flags[i]);
  c.ag_dst = gathered.data_ptr();
  c.ag_rows_per_rank = rows_per_rank;
  c.ag_row_bytes = K * 2;
  c.ag_done_counter = done_counter.data_ptr<int>();
  fill_pads(c, pad_local, pad_peers);
  CHK(mlb_gemm_bf16_2cta_ag(gathered.data_ptr(), weight.data_ptr(), out.data_ptr(), M, N, K, K, (int)weight.stride(0),
                            (int)out.stride(0), b_mn, &c, sms > 0 ? (int)sms : sm_count(), cur()));
  return true;
}

// rs_out[M/world, N] = reduce_scatter(x[M, K] @ W^T or @ W) over the group; tiles travel through rs_dst[] (peer slots).
// ``prev_total``: cumulative arrivals every source had delivered per destination before this call; returns the new
// cumulative count (the tile granularity depends on the kernel variant that is chosen here).
static int64_t fused_gemm_rs(const torch::Tensor& x, const torch::Tensor& weight, torch::Tensor& rs_out, bool b_mn,
                             const std::vector<int64_t>& rs_dst, int64_t rs_slots, int64_t rows_per_rank,
                             int64_t prev_total, int64_t tiles_1cta, torch::Tensor& reduce_counter, int64_t pad_local,
                             const std::vector<int64_t>& pad_peers, int64_t rank, int64_t world, int64_t epoch,
                             int64_t sms, int64_t state_ptr, const std::vector<int64_t>& ar_dst, int64_t stats_ptr) {
  c10::cuda::CUDAGuard guard(x.device());
  const int M = x.size(0), K = x.size(1);
  const int N = b_mn ? weight.size(1) : weight.size(0);
  TORCH_CHECK(x.stride(1) == 1 && weight.stride(1) == 1 && rs_out.is_contiguous());
  TORCH_CHECK(rows_per_rank % mlb::GEMM_BLOCK_M == 0 && M == rows_per_rank * world);
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0);
  mlb::GemmComm c;
  memset(&c, 0, sizeof(c));
  c.rank = rank; c.world = world; c.epoch = epoch;
  c.state = reinterpret_cast<const int*>(state_ptr);
  c.stats = reinterpret_cast<unsigned long long*>(stats_ptr);
  c.m_rotate_blocks = (int)(((rank + 1) % world) * rows_per_rank / mlb::GEMM_BLOCK_M);  // remote chunks first
  c.m_group_blocks = pick_group_blocks(rows_per_rank, (int64_t)N * K * 2);
  c.m_interleave = (world == 2 && (M / 256) % (2 * c.m_group_blocks) == 0) ? 1 : 0;
  for (int i = 0; i < world; ++i) c.rs_dst[i] = reinterpret_cast<void*>(rs_dst[i]);
  c.rs_slots = reinterpret_cast<const void*>(rs_slots);
  c.rs_out = rs_out.data_ptr();
  c.rs_rows_per_rank = rows_per_rank;
  c.rs_reduce_counter = reduce_counter.data_ptr<int>();
  // GEMM -> all-reduce: every rank's [M, N] output buffer of this parity (empty list = plain reduce-scatter)
  TORCH_CHECK(ar_dst.empty() || (int64_t)ar_dst.size() == world, "ar_dst: one buffer per rank");
  for (size_t i = 0; i < ar_dst.size(); ++i) c.ar_dst[i] = reinterpret_cast<void*>(ar_dst[i]);
  fill_pads(c, pad_local, pad_peers);
  static const bool force_1cta = getenv("MLB200_FUSED_1CTA") != nullptr;
  if (!force_1cta && rows_per_rank % 256 == 0 && N >= 256) {
    const int got = mlb_gemm_bf16_2cta_rs(x.data_ptr(), weight.data_ptr(), M, N, K, (int)x.stride(0),
                                          (int)weight.stride(0), b_mn, &c, (int)prev_total,
                                          sms > 0 ? (int)sms : sm_count(), cur());
    if (got > 0) return prev_total + got;
    TORCH_CHECK(got == -3, "fused_gemm_rs (2-CTA) failed with code ", got);
  }
  c.rs_expected_total = (int)(prev_total + tiles_1cta);
  CHK(mlb_gemm_bf16_fused(mlb::MODE_GEMM_RS, x.data_ptr(), weight.data_ptr(), nullptr, M, N, K, (int)x.stride(0),
                          (int)weight.stride(0), N, b_mn, &c, sms > 0 ? (int)sms : sm_count(), cur()));
  return prev_total + tiles_1cta;
}

// Data-parallel gradient reduction over peer memory: every rank's fp32 bucket lives in symmetric memory.
//   reduce_scatter=0: all-reduce (each rank reduces its 1/world slice from all peers, scales by `scale`, and writes the
//                     result back into every peer's bucket)            reduce_scatter=1: only the own slice.
static void dp_reduce(torch::Tensor& local, const std::vector<int64_t>& peer_ptrs, int64_t pad_local,
                      const std::vector<int64_t>& pad_peers, int64_t rank, int64_t world, int64_t epoch,
                      double scale, bool reduce_scatter, int64_t num_ctas) {
  c10::cuda::CUDAGuard guard(local.device());
  TORCH_CHECK(local.scalar_type() == torch::kFloat32 && local.is_contiguous());
  long long pp[mlb::GEMM_MAX_PEERS] = {0}, pads[mlb::GEMM_MAX_PEERS] = {0};
  for (int i = 0; i < world; ++i) { pp[i] = peer_ptrs[i]; pads[i] = pad_peers[i]; }
  // pointer tables are passed by value through a tiny device-visible staging struct inside the launcher
  CHK(mlb_dp_reduce(reduce_scatter, local.data_ptr<float>(), pp, reinterpret_cast<int*>(pad_local), pads,
                    local.numel(), (int)rank, (int)world, (int)epoch, (float)scale, (int)num_ctas, cur()));
}

// NVLS form of dp_reduce: ``mc_ptr`` = multicast address of ``local`` (in-switch fp32 reduction, multicast write-back)
static void dp_reduce_nvls(torch::Tensor& local, int64_t mc_ptr, int64_t pad_local, const std::vector<int64_t>& pad_peers,
                           int64_t rank, int64_t world, int64_t epoch, double scale, bool reduce_scatter,
                           int64_t num_ctas) {
  c10::cuda::CUDAGuard guard(local.device());
  TORCH_CHECK(local.scalar_type() == torch::kFloat32 && local.is_contiguous() && mc_ptr != 0 && mc_ptr % 16 == 0);
  long long pads[mlb::GEMM_MAX_PEERS] = {0};
  for (int i = 0; i < world; ++i) pads[i] = pad_peers[i];
  CHK(mlb_dp_reduce_nvls(reduce_scatter, local.data_ptr<float>(), reinterpret_cast<float*>(mc_ptr),
                         reinterpret_cast<int*>(pad_local), pads, local.numel(), (int)rank, (int)world, (int)epoch,
                         (float)scale, (int)num_ctas, cur()));
}

// all ranks of the group meet on the current stream (flags in pad slots [slot, slot + world))
static void peer_barrier(int64_t pad_local, const std::vector<int64_t>& pad_peers, int64_t rank, int64_t world,
                         int64_t epoch, int64_t slot) {
  long long pads[mlb::GEMM_MAX_PEERS] = {0};
  TORCH_CHECK(world <= mlb::GEMM_MAX_PEERS && (int64_t)pad_peers.size() >= world);
  for (int i = 0; i < world; ++i) pads[i] = pad_peers[i];
  CHK(mlb_peer_barrier(reinterpret_cast<int*>(pad_local), pads, (int)rank, (int)world, (int)epoch, (int)slot, cur()));
}

// state[0..2] = {a, b, c} on the current stream (the offsets a replayed graph's fused kernels add to their epochs)
static void comm_set_state(torch::Tensor& state, int64_t a, int64_t b, int64_t c) {
  c10::cuda::CUDAGuard guard(state.device());
  TORCH_CHECK(state.scalar_type() == torch::kInt32 && state.numel() >= 3);
  CHK(mlb_set_ints3(state.data_ptr<int>(), (int)a, (int)b, (int)c, cur()));
}

// d1[:] = d2[:] = src[:] (contiguous, same byte size, a multiple of 16 bytes)
static void comm_copy2(const torch::Tensor& src, torch::Tensor& d1, torch::Tensor& d2) {
  c10::cuda::CUDAGuard guard(src.device());
  const long long bytes = (long long)src.numel() * src.element_size();
  TORCH_CHECK(src.is_contiguous() && d1.is_contiguous() && d2.is_contiguous());
  TORCH_CHECK((long long)d1.numel() * d1.element_size() == bytes && (long long)d2.numel() * d2.element_size() == bytes);
  TORCH_CHECK(bytes % 16 == 0 && (uintptr_t)src.data_ptr() % 16 == 0 && (uintptr_t)d1.data_ptr() % 16 == 0 &&
              (uintptr_t)d2.data_ptr() % 16 == 0);
  CHK(mlb_copy2(src.data_ptr(), d1.data_ptr(), d2.data_ptr(), bytes, sm_count(), cur()));
}

// NVLink transfer micro-benchmark (raw pointers: symmetric-memory peer / multicast addresses); see p2p_bench.cu
static void p2p_bench(int64_t mode, int64_t src_ptr, int64_t dst_ptr, int64_t bytes, int64_t ctas, int64_t piece_bytes,
                      int64_t stages, int64_t row_bytes, int64_t dst_stride) {
  TORCH_CHECK(bytes % 16 == 0 && src_ptr % 16 == 0 && dst_ptr % 16 == 0);
  TORCH_CHECK(mode == 0 || mode == 3 || (piece_bytes % 16 == 0 && stages >= 2 && stages * piece_bytes <= 200 * 1024));
  TORCH_CHECK(mode != 2 || (row_bytes % 16 == 0 && piece_bytes % row_bytes == 0 && bytes % piece_bytes == 0));
  CHK(mlb_p2p_bench((int)mode, reinterpret_cast<const void*>(src_ptr), reinterpret_cast<void*>(dst_ptr), bytes,
                    (int)ctas, (int)piece_bytes, (int)stages, (int)row_bytes, dst_stride, cur()));
}

void register_comm(pybind11::module_& m) {
  m.def("p2p_bench", &p2p_bench);
  m.def("comm_copy2", &comm_copy2);
  m.def("comm_set_state", &comm_set_state);
  m.def("fused_ag_gemm", &fused_ag_gemm);
  m.def("fused_ag_gemm_nvls", &fused_ag_gemm_nvls);
  m.def("fused_gemm_rs", &fused_gemm_rs);
  m.def("dp_reduce", &dp_reduce);
  m.def("peer_barrier", &peer_barrier);
  m.def("dp_reduce_nvls", &dp_reduce_nvls);
}

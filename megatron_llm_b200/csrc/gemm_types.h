// Plain-C++ declarations shared by the CUDA kernels and the host bindings (no device code in here).
#pragma once
#include <stdint.h>

namespace mlb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = one 128B swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_MAX_PEERS = 8;

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // C(bf16) = acc
  EPI_F32_ACCUM = 1,   // C(fp32) += acc      (wgrad into main_grad)
  EPI_F32 = 2,         // C(fp32) = acc
  EPI_BF16_ACCUM = 3,  // C(bf16) += acc
};

enum GemmMode : int { MODE_PLAIN = 0, MODE_AG_GEMM = 1, MODE_GEMM_RS = 2 };

// int32 slots of the per-rank symmetric signal pad (every rank has the same layout, peers write into it)
enum PadSlot : int {
  PAD_AG_READY = 0,   // [src]  : src published its shard for epoch e
  PAD_AG_ACK = 8,     // [rdr]  : rdr finished reading my shard of epoch e
  PAD_RS_ARRIVED = 16,  // [src]: cumulative number of tiles src delivered into my receive slots
  PAD_RS_FREE = 24,   // [dst]  : dst finished reducing epoch e (its receive slot of parity e may be reused)
  PAD_ERROR = 32,     // spin-wait timeout marker (checked once per training step)
  PAD_AR_DONE = 40,   // [src]  : (GEMM -> all-reduce) src stored its reduced row slice of epoch e into my output buffer
  PAD_INTS = 64,
};

struct GemmComm {
  int rank, world, epoch;
  int num_comm_ctas;     // AG: trailing CTAs of the grid that pull peer shards
  int m_rotate_blocks;   // first m-block processed
  int m_group_blocks;    // 2-CTA fused modes: 256-row blocks per tile group (n-major inside a group); 0 = default
  int m_interleave;      // reduce-scatter, world 2: alternate remote / own tile groups (evens out the NVLink stores)
  // ---- all-gather side
  const void* ag_src[GEMM_MAX_PEERS];  // ag_src[p]: peer p's published shard [rows_per_rank, K] (NVLink-mapped)
  void* ag_dst;                        // local gathered activations [world*rows_per_rank, K]
  int ag_rows_per_rank;
  int ag_row_bytes;                    // K * 2
  int* ag_chunk_flags;                 // local, one per 128-row chunk of the gathered buffer: set to epoch
  int* ag_read_counters;               // local, [world]: puller CTAs done with peer p (for the ack)
  // ---- NVLS all-gather (2-CTA kernel): the owner's pusher CTAs ``multimem.st`` their shard into the symmetric gather
  //      buffer of every rank at once (egress 1x instead of (world-1)x pulled, posted writes, 4-8 SIMT CTAs instead of
  //      16-32 bulk-copy pullers) and release one flag per 128-row chunk at every destination; A is read from that buffer
  int ag_nvls;
  int m_stripe;                        // tile groups visit the sources round-robin (chunks of ALL sources arrive together)
  const void* ag_local_src;            // this rank's shard [rows_per_rank, K] (any local tensor)
  void* ag_mc_dst;                     // multicast address of the gather buffer of this epoch parity
  int* ag_flag_peer[GEMM_MAX_PEERS];   // rank d's chunk flags of this parity ([rank] = local): set to the epoch
  int* ag_done_counter;                // local: CTAs of this launch done with the buffer (last one acks to the pushers)
  // ---- reduce-scatter side
  void* rs_dst[GEMM_MAX_PEERS];        // rs_dst[d]: my receive slot on rank d: [rows_per_rank, N] bf16
  const void* rs_slots;                // local receive buffer of this epoch parity: [world][rows_per_rank, N]
  void* rs_out;                        // local reduced output [rows_per_rank, N]
  int rs_rows_per_rank;
  int rs_expected_total;               // cumulative tiles every source will have delivered after this call
  int* rs_reduce_counter;              // local: CTAs that finished the reduction (last one frees the slot)
  // ---- GEMM -> all-reduce (non-sequence-parallel Row forward / Column dgrad): reduce-scatter as above, then every rank
  //      stores its reduced slice into ALL ranks' symmetric output buffers [M, N] (all-gather by posted stores)
  void* ar_dst[GEMM_MAX_PEERS];        // rank d's output buffer of this epoch parity (nullptr = plain reduce-scatter)
  // ---- device-resident offsets added to `epoch` / `rs_expected_total` (nullptr = 0).  A kernel node of a replayed
  //      CUDA graph keeps the arguments of its capture; the host writes {ag epoch, rs epoch, rs arrivals} deltas here
  //      before every replay so the captured calls continue the live sequence.
  const int* state;
  // ---- exposed-communication accounting (nullptr = off), %globaltimer nanoseconds averaged over the CTAs of a launch:
  //      [0] all-gather: time the TMA producers spent blocked on chunk flags (data not there yet)
  //      [1] reduce-scatter / all-reduce: time from a CTA's last tile to the end of the slot reduction + handshakes
  unsigned long long* stats;
  // ---- signal pads
  int* pad_local;
  int* pad_peer[GEMM_MAX_PEERS];
};

struct GemmParams {
  void* C;
  int M, N, K;
  int ldc;  // elements
  int fp16; // 1: A/B (and a 16-bit C) are IEEE fp16 instead of bf16 (same tiles; kind::f16 operand-format bits differ)
  GemmComm comm;
};

}  // namespace mlb

// cuTensorMapEncodeTiled (driver API) needs a CUDA context current on the calling thread.  PyTorch's autograd worker
// threads only get one once a runtime call that needs it runs there (a CUDAGuard for device 0 is a no-op), so every
// tensor-map builder binds the primary context of the thread's current device first (one runtime call per thread).
#if defined(__CUDACC__) || defined(__CUDA_RUNTIME_H__) || defined(CUDART_VERSION)
static inline void mlb_bind_context() {
  static thread_local bool bound = false;
  if (!bound) { cudaFree(nullptr); bound = true; }
}
#endif

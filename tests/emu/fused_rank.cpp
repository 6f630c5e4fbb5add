// One emulated tensor-parallel rank of the fused GEMM -> reduce-scatter / all-reduce kernels (csrc/gemm_sm100.cuh
// MODE_GEMM_RS with one CTA, csrc/gemm2_sm100.cu with one CTA pair), on the functional tcgen05 model.  Ranks are
// processes; symmetric memory (receive slots, signal pads, all-reduce outputs) is a set of files every rank maps:
//
//     fused_rank <dir> <rank> <world> <rows_per_rank> <N> <K> <mode: rs1 | rs2 | ar1> <calls> [CTAs (pairs) per rank]
// (more than one CTA / pair per rank needs MLB_EMU_CONCURRENT_BLOCKS=1: the slot reduction waits for all of them)
//
// Row-parallel forward: rank r holds X_r [M, K] and W_r [N, K] (its K-shard); Y = sum_r X_r W_r^T; rank d ends up with
// rows [d m, d m + m) of Y (reduce-scatter) or all of Y (all-reduce).  Tiles travel from the epilogue of every rank into
// the destination's receive slot (plain stores / TMA stores into the mapped file), arrival counters and free flags go
// through the pads, the slot reduction runs when all sources delivered.  Every rank recomputes the reference from the
// seeds and exits 0 only if its result matches and no spin-wait timed out.  `calls` > 1 alternates the two slot
// parities and exercises the "slot is free again" handshake (PAD_RS_FREE).
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gemm_types.h"

typedef void* cudaStream_t;
extern "C" {
int mlb_gemm_bf16_fused(int mode, const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                        int b_mn_major, const mlb::GemmComm* comm, int num_sms, cudaStream_t stream);
int mlb_gemm_bf16_2cta_rs(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int b_mn_major,
                          mlb::GemmComm* comm, int prev_total, int num_sms, cudaStream_t stream);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return uint16_t(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
static void fill(std::vector<uint16_t>& v, uint32_t seed) {
  uint32_t s = seed * 2654435761u + 12345u;
  for (auto& x : v) { s = s * 1664525u + 1013904223u; x = f2bf((((s >> 8) & 0xFFFF) / 32768.0f - 1.0f) * 0.5f); }
}
static void* map_file(const std::string& path, size_t bytes) {
  const int fd = open(path.c_str(), O_RDWR);
  if (fd < 0) { perror(path.c_str()); exit(9); }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { perror("mmap"); exit(9); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 9) return 8;
  const std::string dir = argv[1], mode = argv[7];
  const int rank = atoi(argv[2]), world = atoi(argv[3]), m = atoi(argv[4]), N = atoi(argv[5]), K = atoi(argv[6]);
  const int calls = atoi(argv[8]), M = m * world, ctas = argc > 9 ? atoi(argv[9]) : 1;
  const bool all_reduce = mode == "ar1", two_cta = mode == "rs2";
  const size_t slot_elems = (size_t)world * m * N;                       // one parity: [world][m][N] bf16
  std::vector<uint16_t*> slots(world), arout(world);
  std::vector<int*> pads(world);
  for (int r = 0; r < world; ++r) {
    slots[r] = (uint16_t*)map_file(dir + "/slots" + std::to_string(r) + ".bin", 2 * slot_elems * 2);
    arout[r] = (uint16_t*)map_file(dir + "/arout" + std::to_string(r) + ".bin", 2 * (size_t)M * N * 2);
    pads[r] = (int*)map_file(dir + "/pad" + std::to_string(r) + ".bin", 64 * 4);
  }
  // rendezvous (process start-up must not count against the kernels' bounded spins)
  { FILE* f = fopen((dir + "/ready" + std::to_string(rank)).c_str(), "w"); fclose(f); }
  for (int r = 0; r < world; ++r)
    while (access((dir + "/ready" + std::to_string(r)).c_str(), F_OK) != 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));

  // FUSED_RANK_GRAPH_STATE=1: like a replayed CUDA graph -- the launch arguments stay those of the "capture" (epoch of
  // the parity's first call, arrivals of one call) and the live epoch / arrival total come from GemmComm::state
  const bool graph_state = getenv("FUSED_RANK_GRAPH_STATE") != nullptr;
  int state[3] = {0, 0, 0};
  int reduce_counter = 0, total = 0, rc = 0;
  for (int call = 1; call <= calls && rc == 0; ++call) {
    const int parity = call & 1;
    std::vector<std::vector<uint16_t>> X(world, std::vector<uint16_t>((size_t)M * K)), W(world, std::vector<uint16_t>((size_t)N * K));
    for (int r = 0; r < world; ++r) { fill(X[r], 1000 * call + r); fill(W[r], 2000 * call + r); }
    std::vector<uint16_t> out((size_t)m * N, 0x7fc0);
    mlb::GemmComm c;
    memset(&c, 0, sizeof(c));
    c.rank = rank; c.world = world; c.epoch = call;
    const int prev_total = total;
    if (graph_state) {
      c.epoch = 2 - parity;                                   // the epoch this parity's kernel node was captured with
      state[1] = call - c.epoch;                              // STATE_RS_EPOCH
      state[2] = prev_total;                                  // STATE_RS_TOTAL (the captured node expects one call's arrivals)
      c.state = state;
    }
    c.m_rotate_blocks = ((rank + 1) % world) * m / mlb::GEMM_BLOCK_M;            // remote chunks first
    c.m_group_blocks = 1;
    for (int d = 0; d < world; ++d) {
      c.rs_dst[d] = slots[d] + parity * slot_elems + (size_t)rank * m * N;        // my receive slot on rank d
      c.pad_peer[d] = pads[d];
      if (all_reduce) c.ar_dst[d] = arout[d] + (size_t)parity * M * N;
    }
    c.rs_slots = slots[rank] + parity * slot_elems;
    c.rs_out = out.data();
    c.rs_rows_per_rank = m;
    c.rs_reduce_counter = &reduce_counter;
    c.pad_local = pads[rank];
    if (two_cta) {
      const int got = mlb_gemm_bf16_2cta_rs(X[rank].data(), W[rank].data(), M, N, K, K, K, 0, &c, graph_state ? 0 : total, 2 * ctas, nullptr);
      if (got <= 0) { fprintf(stderr, "2cta rs -> %d\n", got); return 3; }
      total += got;
    } else {
      const int tiles = (m / mlb::GEMM_BLOCK_M) * ((N + 127) / 128);                // N <= 128 here: one column of tiles
      c.rs_expected_total = (graph_state ? 0 : total) + tiles;
      total += tiles;
      const int e = mlb_gemm_bf16_fused(mlb::MODE_GEMM_RS, X[rank].data(), W[rank].data(), nullptr, M, N, K, K, K, N, 0, &c, ctas, nullptr);
      if (e) { fprintf(stderr, "fused -> %d\n", e); return 3; }
    }
    if (pads[rank][mlb::PAD_ERROR]) { fprintf(stderr, "rank %d: a spin-wait timed out\n", rank); return 4; }
    // reference: the rows this rank must hold, summed over the sources in bf16 slot precision like the kernel
    const int row0 = all_reduce ? 0 : rank * m, nrows = all_reduce ? M : m;
    const uint16_t* got = all_reduce ? arout[rank] + (size_t)parity * M * N : out.data();
    double worst = 0, scale = 0;
    for (int i = 0; i < nrows; ++i)
      for (int n = 0; n < N; ++n) {
        float sum = 0.f;
        for (int r = 0; r < world; ++r) {
          float acc = 0.f;
          for (int k = 0; k < K; ++k) acc += bf2f(X[r][(size_t)(row0 + i) * K + k]) * bf2f(W[r][(size_t)n * K + k]);
          sum += bf2f(f2bf(acc));                                                  // tiles travel as bf16
        }
        const double d = std::fabs(bf2f(got[(size_t)i * N + n]) - sum);
        worst = d > worst ? d : worst;
        scale = std::fabs(sum) > scale ? std::fabs(sum) : scale;
      }
    if (!(worst <= 0.02 * (scale > 1 ? scale : 1))) { fprintf(stderr, "rank %d call %d: max err %g (scale %g)\n", rank, call, worst, scale); rc = 5; }
  }
  return rc;
}

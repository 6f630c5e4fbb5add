// One emulated tensor-parallel rank of the fused all-gather -> GEMM kernel (csrc/gemm_sm100.cuh MODE_AG_GEMM, 1-CTA
// kernel) on the functional model: the puller CTAs and the compute CTAs of the launch run concurrently
// (MLB_EMU_CONCURRENT_BLOCKS=1), ranks are processes, the published shards and the signal pads are files all ranks map.
//
//     fused_ag_rank <dir> <rank> <world> <rows_per_rank> <N> <K> <pullers> <calls> [2cta | nvls]
// "2cta": csrc/gemm2_sm100.cu -- puller CLUSTERS next to compute clusters, rows per rank a multiple of 256.
// "nvls": the push transport of the 2-CTA kernel -- pusher CTAs store the own shard into EVERY rank's gather buffer
//         through the multicast mapping (`multimem.st`, emulated over the mapped copies) and release one flag per chunk
//         at every destination; the own rows arrive that way too; two buffer parities, PAD_AG_ACK two calls back.
//
// Column-parallel forward under sequence parallelism: rank r owns the activation shard X_r [m, K] and the weight shard
// W_r [N, K]; out_r = concat_p(X_p) W_r^T.  The own rows are placed by the host before the launch (as the real caller
// does), the pullers copy the peers' shards piece by piece through shared memory (bulk copies), release one flag per
// 128-row chunk, and the TMA producers of the compute CTAs wait for the flag of the chunk they are about to read; read
// acknowledgements let the next call overwrite a published shard.  Exit 0 = result matches the reference recomputed
// from the seeds and no bounded spin timed out.
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
int mlb_gemm_bf16_2cta_ag(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                          int b_mn_major, const mlb::GemmComm* comm, int num_sms, cudaStream_t stream);
void emu_register_multicast(const float* mc_base, const long long* copies, int world, long long n);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return uint16_t(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
static void fill(uint16_t* v, size_t n, uint32_t seed) {
  uint32_t s = seed * 2654435761u + 12345u;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; v[i] = f2bf((((s >> 8) & 0xFFFF) / 32768.0f - 1.0f) * 0.5f); }
}
static void* map_file(const std::string& path, size_t bytes) {
  const int fd = open(path.c_str(), O_RDWR);
  if (fd < 0) { perror(path.c_str()); exit(9); }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { perror("mmap"); exit(9); }
  return p;
}
static void rendezvous(const std::string& dir, const char* tag, int call, int rank, int world) {
  { FILE* f = fopen((dir + "/" + tag + std::to_string(call) + "_" + std::to_string(rank)).c_str(), "w"); fclose(f); }
  for (int r = 0; r < world; ++r)
    while (access((dir + "/" + tag + std::to_string(call) + "_" + std::to_string(r)).c_str(), F_OK) != 0)
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
}

int main(int argc, char** argv) {
  if (argc < 9) return 8;
  const std::string dir = argv[1];
  const int rank = atoi(argv[2]), world = atoi(argv[3]), m = atoi(argv[4]), N = atoi(argv[5]), K = atoi(argv[6]);
  const int pullers = atoi(argv[7]), calls = atoi(argv[8]), M = m * world;
  const bool nvls = argc > 9 && !strcmp(argv[9], "nvls");
  const bool two_cta = nvls || (argc > 9 && !strcmp(argv[9], "2cta"));
  std::vector<uint16_t*> shard(world);
  std::vector<int*> pads(world);
  for (int r = 0; r < world; ++r) {
    shard[r] = (uint16_t*)map_file(dir + "/shard" + std::to_string(r) + ".bin", (size_t)m * K * 2);
    pads[r] = (int*)map_file(dir + "/pad" + std::to_string(r) + ".bin", 64 * 4);
  }
  // NVLS: symmetric gather buffers (two parities) and chunk flags of every rank
  std::vector<uint16_t*> gbuf(world);
  std::vector<int*> gflags(world);
  if (nvls)
    for (int r = 0; r < world; ++r) {
      gbuf[r] = (uint16_t*)map_file(dir + "/gather" + std::to_string(r) + ".bin", 2 * (size_t)M * K * 2);
      gflags[r] = (int*)map_file(dir + "/gflags" + std::to_string(r) + ".bin", 2 * (size_t)(M / 128) * 4);
    }
  std::vector<float> mc_va((size_t)M * K / 2 + 16);          // only its address range is used: the "multicast VA"
  int done_counter = 0;
  std::vector<uint16_t> gathered((size_t)M * K + 64), W((size_t)N * K), out((size_t)M * N);
  uint16_t* G = (uint16_t*)(((uintptr_t)gathered.data() + 127) & ~(uintptr_t)127);
  std::vector<int> chunk_flags(M / 128, 0), read_counters(world, 0);
  int rc = 0;
  for (int call = 1; call <= calls && rc == 0; ++call) {
    // stream order of the real caller: write + publish the own shard, place the own rows, then launch
    fill(shard[rank], (size_t)m * K, 1000 * call + rank);
    if (!nvls) memcpy(G + (size_t)rank * m * K, shard[rank], (size_t)m * K * 2);
    fill(W.data(), W.size(), 2000 * call + rank);
    for (auto& x : out) x = 0x7fc0;
    rendezvous(dir, "ready", call, rank, world);       // (keeps interpreter / process start-up out of the bounded spins)
    mlb::GemmComm c;
    memset(&c, 0, sizeof(c));
    c.rank = rank; c.world = world; c.epoch = call;
    c.num_comm_ctas = pullers;
    c.m_rotate_blocks = rank * m / mlb::GEMM_BLOCK_M;                   // own rows first
    for (int p = 0; p < world; ++p) { c.ag_src[p] = shard[p]; c.pad_peer[p] = pads[p]; }
    c.ag_dst = G;
    c.ag_rows_per_rank = m;
    c.ag_row_bytes = K * 2;
    c.ag_chunk_flags = chunk_flags.data();
    c.ag_read_counters = read_counters.data();
    c.pad_local = pads[rank];
    c.m_group_blocks = 1;
    const uint16_t* A = G;
    if (nvls) {
      const int parity = call & 1;
      std::vector<long long> copies(world);
      for (int p = 0; p < world; ++p) copies[p] = (long long)(gbuf[p] + (size_t)parity * M * K);
      emu_register_multicast(mc_va.data(), copies.data(), world, (long long)M * K / 2);
      c.ag_nvls = 1;
      c.m_stripe = 1;
      c.ag_local_src = shard[rank];
      c.ag_mc_dst = mc_va.data();
      for (int p = 0; p < world; ++p) c.ag_flag_peer[p] = gflags[p] + (size_t)parity * (M / 128);
      c.ag_done_counter = &done_counter;
      c.ag_dst = gbuf[rank] + (size_t)parity * M * K;
      A = gbuf[rank] + (size_t)parity * M * K;
    }
    const int e = two_cta ? mlb_gemm_bf16_2cta_ag(A, W.data(), out.data(), M, N, K, K, K, N, 0, &c, 4 + pullers, nullptr)
                          : mlb_gemm_bf16_fused(mlb::MODE_AG_GEMM, G, W.data(), out.data(), M, N, K, K, K, N, 0, &c, 2 + pullers, nullptr);
    if (e) { fprintf(stderr, "fused ag -> %d\n", e); return 3; }
    if (pads[rank][mlb::PAD_ERROR]) { fprintf(stderr, "rank %d: a spin-wait timed out\n", rank); return 4; }
    double worst = 0;
    std::vector<uint16_t> xs((size_t)m * K);
    for (int p = 0; p < world; ++p) {
      fill(xs.data(), xs.size(), 1000 * call + p);
      for (int i = 0; i < m; ++i)
        for (int n = 0; n < N; ++n) {
          float acc = 0.f;
          for (int k = 0; k < K; ++k) acc += bf2f(xs[(size_t)i * K + k]) * bf2f(W[(size_t)n * K + k]);
          const double d = std::fabs(bf2f(out[((size_t)p * m + i) * N + n]) - acc);
          worst = d > worst ? d : worst;
        }
    }
    if (!(worst <= 0.05)) { fprintf(stderr, "rank %d call %d: max err %g\n", rank, call, worst); rc = 5; }
    rendezvous(dir, "done", call, rank, world);        // nobody rewrites its shard while a slow peer still verifies
  }
  return rc;
}

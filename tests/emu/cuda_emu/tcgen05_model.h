// A FUNCTIONAL model of the Blackwell pieces the attention kernels are written against, for running their source on
// CPU threads (tests/test_attention_kernel_model.py).  It models data movement and ordering, not timing:
//   * shared memory  : one 1024-byte aligned buffer per block; smem_u32() = offset in it (so the 128-byte swizzle, which
//                      is a function of address bits 4-9, behaves as on the device)
//   * mbarrier       : phase bit / pending arrivals / transaction bytes; wait(parity) returns once the phase with that
//                      parity has completed (so waiting on the "previous" parity of a fresh barrier returns at once)
//   * TMA            : cuTensorMapEncodeTiled records (base, dims, byte strides, box); a 4-D tile load copies the box row
//                      by row into shared memory through the 128B swizzle, zero-fills out-of-bounds rows and completes
//                      the transaction bytes on the barrier
//   * tensor memory  : 128 lanes x 512 columns of 32 bits per block; tcgen05.ld / st move 32 lanes x N columns per warp
//   * tcgen05.mma    : executed at issue (kind::f16, M = 128, fp32 accumulate): operands are read through the shared-
//                      memory descriptors (start address, LBO, SBO, SWIZZLE_128B; K-major or MN-major per the instruction
//                      descriptor) or from tensor memory (16-bit pairs per column); tcgen05.commit is an immediate arrive
// Everything the kernels synchronise on is therefore still required to be correct: a consumer that does not wait for the
// right barrier phase reads stale data here as well.  What the model cannot show is asynchrony-only bugs (e.g. a missing
// tcgen05.fence) and performance.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <map>
#include <mutex>

#include "cuda_emu.h"

// ---- driver API types used to build tensor maps --------------------------------------------------------------------
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
struct CUtensorMap_st {
  const uint8_t* base;
  uint64_t dims[5], strides[5];      // strides in bytes; strides[0] = element size
  uint32_t box[5];
  uint32_t rank, elem_bytes, swizzle;
};
typedef CUtensorMap_st CUtensorMap;
typedef CUresult (*PFN_cuTensorMapEncodeTiled_v12000)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                      CUtensorMapFloatOOBfill);
inline CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                                           const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle sw,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  tm->base = static_cast<const uint8_t*>(base);
  tm->rank = rank;
  tm->elem_bytes = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : (dt == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1 : 2);
  tm->swizzle = sw;
  tm->strides[0] = tm->elem_bytes;
  for (cuuint32_t i = 0; i < rank; ++i) {
    tm->dims[i] = dims[i];
    tm->box[i] = box[i];
    if (i > 0) tm->strides[i] = strides[i - 1];
    if (i > 0 && strides[i - 1] % 16 != 0) return 1;        // the driver rejects strides that are not 16-byte multiples
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || box[0] * tm->elem_bytes > 128) return 1;
  return CUDA_SUCCESS;
}
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
constexpr int cudaEnableDefault = 0;
inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, int, cudaDriverEntryPointQueryResult* q) {
  *fn = reinterpret_cast<void*>(&emu_cuTensorMapEncodeTiled);
  *q = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
#define __grid_constant__

namespace cuda_emu {

// ---- per-block state (created by launch_dyn for every block) --------------------------------------------------------
struct BlockModel {
  uint8_t* smem = nullptr;             // 1024-byte aligned dynamic shared memory
  size_t smem_bytes = 0;
  std::vector<uint32_t> tmem;          // [128][512]
  std::map<int, std::unique_ptr<std::barrier<>>> named;
  // thread-block cluster (cta_group::2 kernels): the CTAs of the cluster, this CTA's rank, a barrier over all threads.
  // Shared-memory addresses carry the CTA rank in bit 24 (what the kernels' "leader's copy" address arithmetic assumes).
  BlockModel* cluster[2] = {this, nullptr};
  uint32_t cluster_rank = 0;
  std::barrier<>* cluster_bar = nullptr;
  BlockModel() : tmem(128 * 512, 0xDEADBEEFu) {}
};
inline thread_local BlockModel* bm = nullptr;
inline std::mutex& mbar_mutex() { static std::mutex m; return m; }       // every mbarrier operation (any CTA)

// Schedule fuzzing (MLB_EMU_CHAOS=<seed>): random short stalls in front of barrier operations, TMA loads, MMAs and tensor
// memory accesses, different per thread and per run, so that warps drift apart by whole tiles -- the situations in which
// a barrier protocol that only works "because the other side is always faster" gives wrong results or hangs.
inline int chaos_seed() {
  static const int seed = [] { const char* e = std::getenv("MLB_EMU_CHAOS"); return e ? std::atoi(e) : 0; }();
  return seed;
}
inline void chaos() {
  if (chaos_seed() == 0) return;
  static thread_local uint32_t rng = 0;
  if (rng == 0) rng = uint32_t(chaos_seed()) * 2654435761u ^ (uint32_t(std::hash<std::thread::id>()(std::this_thread::get_id())) | 1u);
  rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5;
  if ((rng & 7u) == 0) std::this_thread::sleep_for(std::chrono::microseconds((rng >> 8) % 300));
  else if ((rng & 3u) == 1) std::this_thread::yield();
}

inline uint32_t& tmem_at(uint32_t lane, uint32_t col) { return bm->tmem[(lane & 127) * 512 + (col & 511)]; }

// the 128-byte swizzle: 16-byte chunk index (address bits 4-6) XOR row-in-atom (address bits 7-9)
inline uint32_t swizzle128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }
inline uint8_t* smem_ptr(uint32_t addr) { return bm->cluster[(addr >> 24) & 1]->smem + (addr & 0xFFFFFFu); }

struct MBar { uint8_t expected, pending, phase, pad; int32_t tx; };
static_assert(sizeof(MBar) == 8, "an mbarrier occupies 8 bytes of shared memory");
inline void mbar_check(MBar* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->expected; }
}

inline void named_barrier(int id, int count) {
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> g(mbar_mutex());
    auto& slot = bm->named[id];
    if (!slot) slot.reset(new std::barrier<>(count));
    b = slot.get();
  }
  b->arrive_and_wait();
}

// All blocks of a 1-D grid at once (MLB_EMU_CONCURRENT_BLOCKS=1): for launches whose CTAs wait for each other -- the
// puller CTAs and the compute CTAs of the fused all-gather -> GEMM kernel.  (Kernels with `static` __shared__ variables
// must not be run this way; the fused all-gather path has none.)
template <class F>
void launch_dyn_concurrent(dim3 grid, unsigned threads, size_t smem_bytes, F kernel) {
  const unsigned nb = grid.x;
  std::vector<std::unique_ptr<Block>> blocks;
  std::vector<std::unique_ptr<BlockModel>> models;
  for (unsigned b = 0; b < nb; ++b) {
    blocks.emplace_back(new Block(threads));
    models.emplace_back(new BlockModel());
    models[b]->smem_bytes = smem_bytes;
    models[b]->smem = static_cast<uint8_t*>(std::aligned_alloc(1024, (smem_bytes + 1023) / 1024 * 1024 + 1024));
    std::memset(models[b]->smem, 0xCD, smem_bytes);
  }
  std::vector<std::thread> pool;
  pool.reserve(size_t(nb) * threads);
  for (unsigned b = 0; b < nb; ++b)
    for (unsigned t = 0; t < threads; ++t)
      pool.emplace_back([&, b, t] {
        blk = blocks[b].get();
        bm = models[b].get();
        threadIdx = uint3{t, 0, 0};
        blockIdx = uint3{b, 0, 0};
        blockDim = dim3(threads);
        gridDim = grid;
        kernel();
        blocks[b]->warp_bar[t >> 5]->arrive_and_drop();
        blocks[b]->block_bar.arrive_and_drop();
      });
  for (auto& th : pool) th.join();
  for (auto& m : models) std::free(m->smem);
}

// launch with dynamic shared memory and the block model
template <class F>
void launch_dyn(dim3 grid, unsigned threads, size_t smem_bytes, F kernel) {
  static const bool concurrent = std::getenv("MLB_EMU_CONCURRENT_BLOCKS") != nullptr;
  if (concurrent && grid.y == 1 && grid.z == 1) { launch_dyn_concurrent(grid, threads, smem_bytes, kernel); return; }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block block(threads);
        BlockModel model;
        model.smem_bytes = smem_bytes;
        model.smem = static_cast<uint8_t*>(std::aligned_alloc(1024, (smem_bytes + 1023) / 1024 * 1024 + 1024));
        std::memset(model.smem, 0xCD, smem_bytes);
        std::vector<std::thread> pool;
        pool.reserve(threads);
        for (unsigned t = 0; t < threads; ++t)
          pool.emplace_back([&, t] {
            blk = &block;
            bm = &model;
            threadIdx = uint3{t, 0, 0};
            blockIdx = uint3{bx, by, bz};
            blockDim = dim3(threads);
            gridDim = grid;
            kernel();
            block.warp_bar[t >> 5]->arrive_and_drop();
            block.block_bar.arrive_and_drop();
          });
        for (auto& th : pool) th.join();
        std::free(model.smem);
      }
}

// all CTA pairs of the grid at once (MLB_EMU_CONCURRENT_BLOCKS=1): puller clusters next to compute clusters
template <class F>
void launch_cluster2_concurrent(dim3 grid, unsigned threads, size_t smem_bytes, F kernel) {
  const unsigned pairs = grid.x / 2;
  std::vector<std::unique_ptr<Block>> blocks;
  std::vector<std::unique_ptr<BlockModel>> models;
  std::vector<std::unique_ptr<std::barrier<>>> cbars;
  for (unsigned b = 0; b < 2 * pairs; ++b) {
    blocks.emplace_back(new Block(threads));
    models.emplace_back(new BlockModel());
    models[b]->smem_bytes = smem_bytes;
    models[b]->smem = static_cast<uint8_t*>(std::aligned_alloc(1024, (smem_bytes + 1023) / 1024 * 1024 + 1024));
    std::memset(models[b]->smem, 0xCD, smem_bytes);
  }
  for (unsigned pr = 0; pr < pairs; ++pr) {
    cbars.emplace_back(new std::barrier<>(2 * threads));
    for (unsigned r = 0; r < 2; ++r) {
      BlockModel* m = models[2 * pr + r].get();
      m->cluster[0] = models[2 * pr].get(); m->cluster[1] = models[2 * pr + 1].get();
      m->cluster_rank = r;
      m->cluster_bar = cbars[pr].get();
    }
  }
  std::vector<std::thread> pool;
  pool.reserve(size_t(2 * pairs) * threads);
  for (unsigned b = 0; b < 2 * pairs; ++b)
    for (unsigned t = 0; t < threads; ++t)
      pool.emplace_back([&, b, t] {
        blk = blocks[b].get();
        bm = models[b].get();
        threadIdx = uint3{t, 0, 0};
        blockIdx = uint3{b, 0, 0};
        blockDim = dim3(threads);
        gridDim = grid;
        kernel();
        blocks[b]->warp_bar[t >> 5]->arrive_and_drop();
        blocks[b]->block_bar.arrive_and_drop();
        cbars[b / 2]->arrive_and_drop();
      });
  for (auto& th : pool) th.join();
  for (auto& m : models) std::free(m->smem);
}

// clusters of two CTAs (__cluster_dims__(2, 1, 1)): the pair runs concurrently, pairs one after the other
template <class F>
void launch_cluster2(dim3 grid, unsigned threads, size_t smem_bytes, F kernel) {
  static const bool concurrent = std::getenv("MLB_EMU_CONCURRENT_BLOCKS") != nullptr;
  if (concurrent) { launch_cluster2_concurrent(grid, threads, smem_bytes, kernel); return; }
  for (unsigned pair = 0; pair < grid.x / 2; ++pair) {
    Block block0(threads), block1(threads);
    Block* blocks[2] = {&block0, &block1};
    BlockModel model[2];
    std::barrier<> cbar(2 * threads);
    for (int r = 0; r < 2; ++r) {
      model[r].smem_bytes = smem_bytes;
      model[r].smem = static_cast<uint8_t*>(std::aligned_alloc(1024, (smem_bytes + 1023) / 1024 * 1024 + 1024));
      std::memset(model[r].smem, 0xCD, smem_bytes);
      model[r].cluster[0] = &model[0]; model[r].cluster[1] = &model[1];
      model[r].cluster_rank = r;
      model[r].cluster_bar = &cbar;
    }
    std::vector<std::thread> pool;
    pool.reserve(2 * threads);
    for (unsigned r = 0; r < 2; ++r)
      for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, r, t] {
          blk = blocks[r];
          bm = &model[r];
          threadIdx = uint3{t, 0, 0};
          blockIdx = uint3{2 * pair + r, 0, 0};
          blockDim = dim3(threads);
          gridDim = grid;
          kernel();
          blocks[r]->warp_bar[t >> 5]->arrive_and_drop();
          blocks[r]->block_bar.arrive_and_drop();
          cbar.arrive_and_drop();
        });
    for (auto& th : pool) th.join();
    for (int r = 0; r < 2; ++r) std::free(model[r].smem);
  }
}
#define __cluster_dims__(...)

inline float to_float16bits(uint16_t h, bool is_bf16) {
  if (is_bf16) { uint32_t u = uint32_t(h) << 16; float f; std::memcpy(&f, &u, 4); return f; }
  _Float16 v; std::memcpy(&v, &h, 2); return float(v);
}

}  // namespace cuda_emu

inline bool __any_sync(unsigned, bool pred) {          // warp-wide OR
  int v = pred ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v != 0;
}

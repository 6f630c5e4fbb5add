// Data-parallel gradient reduction over NVLink peer memory (sm_100a), fused with the 1/DP scale.
//
// Replaces the reference's single blocking NCCL all-reduce of the whole fp32 grad buffer after backward
// (megatron/model/distributed.py:202-209) and the distributed optimizer's reduce_scatter_tensor
// (optimizer/distrib_optimizer.py:553-567).  Every rank's fp32 bucket lives in symmetric memory mapped into all
// peers.  Two-shot algorithm, one kernel per bucket, launched from the backward hooks on a side stream:
//   1. handshake: publish "my bucket is complete" to all peers / wait for theirs        (st.release.sys / ld.acquire.sys)
//   2. reduce-scatter: rank r sums slice r of every peer's bucket with 16-byte loads over NVLink, scales by 1/DP,
//      writes its own slice;
//   3. (all-reduce only) all-gather: the reduced slice is stored into every peer's bucket over NVLink;
//   4. handshake out: a rank's kernel retires only when every peer has finished reading from / writing into its bucket.
#include <string.h>

#include "gemm_types.h"
#include "ptx.cuh"

namespace mlb {

enum DpPadSlot : int { DP_READY = 0, DP_DONE = 8, DP_ERROR = 32 /* == PAD_ERROR */, DP_CTA_COUNTER = 40 };

// bounded spin: a lost peer must not hang the box; the timeout is recorded in the pad (polled once per training step)
__device__ __forceinline__ void dp_spin_until_ge(const int* flag, int value, int* pad_local) {
  long long polls = 0;
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(100);
    if (++polls > (1LL << 25)) {
      st_release_sys(pad_local + DP_ERROR, 1);
      break;
    }
  }
}

struct DpArgs {
  float* peer[GEMM_MAX_PEERS];
  int* pad_peer[GEMM_MAX_PEERS];
  int* pad_local;
  long long n;
  int rank, world, epoch;
  float scale;
  int reduce_scatter;
};

__global__ void __launch_bounds__(512) dp_reduce_kernel(const DpArgs a) {
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        if (p != a.rank) st_release_sys(a.pad_peer[p] + DP_READY + a.rank, a.epoch);
    }
    for (int p = 0; p < a.world; ++p)
      if (p != a.rank) dp_spin_until_ge(a.pad_local + DP_READY + p, a.epoch, a.pad_local);
  }
  __syncthreads();

  const long long slice = a.n / a.world;          // n is padded to a multiple of world*4 by the caller
  const long long begin = slice * a.rank;
  const long long nvec = slice / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e = begin + i * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < GEMM_MAX_PEERS; ++k) {
      if (k < a.world) {
        const int p = (a.rank + k) % a.world;      // start with the local copy, spread NVLink reads across peers
        const uint4 u = ld_v4_relaxed_sys(a.peer[p] + e);
        acc.x += __uint_as_float(u.x); acc.y += __uint_as_float(u.y);
        acc.z += __uint_as_float(u.z); acc.w += __uint_as_float(u.w);
      }
    }
    acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
    const uint4 o = make_uint4(__float_as_uint(acc.x), __float_as_uint(acc.y), __float_as_uint(acc.z),
                               __float_as_uint(acc.w));
    if (a.reduce_scatter) {
      st_v4(a.peer[a.rank] + e, o);
    } else {
#pragma unroll
      for (int k = 0; k < GEMM_MAX_PEERS; ++k)
        if (k < a.world) st_v4(a.peer[(a.rank + k) % a.world] + e, o);
    }
  }

  // handshake out (last CTA of this rank)
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(a.pad_local + DP_CTA_COUNTER, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    a.pad_local[DP_CTA_COUNTER] = 0;
    __threadfence_system();
    for (int p = 0; p < a.world; ++p)
      if (p != a.rank) st_release_sys(a.pad_peer[p] + DP_DONE + a.rank, a.epoch);
    for (int p = 0; p < a.world; ++p)
      if (p != a.rank) dp_spin_until_ge(a.pad_local + DP_DONE + p, a.epoch, a.pad_local);
  }
}

// NVLS variant: the bucket is mapped through an NVSwitch multicast object.  ``multimem.ld_reduce`` returns the SUM of the
// element over every rank's copy -- the addition happens inside the switch, so a rank pulls 1 x its slice instead of
// (world - 1) x -- and ``multimem.st`` writes the scaled result into every rank's copy at once (all-reduce form).
struct DpNvlsArgs {
  float* mc;                 // multicast address of the bucket start
  float* local;              // this rank's unicast address of the bucket start
  int* pad_peer[GEMM_MAX_PEERS];
  int* pad_local;
  long long n;
  int rank, world, epoch;
  float scale;
  int reduce_scatter;
};

__global__ void __launch_bounds__(512) dp_reduce_nvls_kernel(const DpNvlsArgs a) {
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
#pragma unroll
      for (int p = 0; p < GEMM_MAX_PEERS; ++p)
        if (p < a.world && p != a.rank) st_release_sys(a.pad_peer[p] + DP_READY + a.rank, a.epoch);
    }
    for (int p = 0; p < a.world; ++p)
      if (p != a.rank) dp_spin_until_ge(a.pad_local + DP_READY + p, a.epoch, a.pad_local);
  }
  __syncthreads();
  const long long slice = a.n / a.world;
  const long long begin = slice * a.rank;
  const long long nvec = slice / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 4;       // independent 16-byte in-switch reductions in flight per thread
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(a.mc + begin + i * 4)
                     : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      const float4 o = make_float4(v[u].x * a.scale, v[u].y * a.scale, v[u].z * a.scale, v[u].w * a.scale);
      if (a.reduce_scatter) {
        *reinterpret_cast<float4*>(a.local + begin + i * 4) = o;
      } else {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + begin + i * 4),
                     "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                     : "memory");
      }
    }
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(a.pad_local + DP_CTA_COUNTER, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    a.pad_local[DP_CTA_COUNTER] = 0;
    __threadfence_system();
#pragma unroll
    for (int p = 0; p < GEMM_MAX_PEERS; ++p)
      if (p < a.world && p != a.rank) st_release_sys(a.pad_peer[p] + DP_DONE + a.rank, a.epoch);
    for (int p = 0; p < a.world; ++p)
      if (p != a.rank) dp_spin_until_ge(a.pad_local + DP_DONE + p, a.epoch, a.pad_local);
  }
}

}  // namespace mlb

// dst1[i] = dst2[i] = src[i] (16-byte vectors): publishes a shard to the symmetric buffer and places it into the local
// gathered buffer with one read
__global__ void __launch_bounds__(256) copy2_kernel(const uint4* __restrict__ src, uint4* __restrict__ d1,
                                                    uint4* __restrict__ d2, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) v[u] = __ldcs(src + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) { d1[i] = v[u]; d2[i] = v[u]; }
    }
  }
}

extern "C" int mlb_copy2(const void* src, void* d1, void* d2, long long bytes, int num_sms, cudaStream_t stream) {
  const long long n = bytes / 16;
  long long want = (n + 256 * 4 - 1) / (256 * 4);
  int grid = (int)(want < 1 ? 1 : (want > 4LL * num_sms ? 4LL * num_sms : want));
  copy2_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(d1),
                                         reinterpret_cast<uint4*>(d2), n);
  return (int)cudaGetLastError();
}

__global__ void set_ints3_kernel(int* dst, int a, int b, int c) {
  dst[0] = a; dst[1] = b; dst[2] = c;
}

extern "C" int mlb_set_ints3(int* dst, int a, int b, int c, cudaStream_t stream) {
  set_ints3_kernel<<<1, 1, 0, stream>>>(dst, a, b, c);
  return (int)cudaGetLastError();
}

// all ranks of a group meet: used after the ZeRO-1 optimizer kernel has stored its updated 16-bit shard into every
// peer's parameter buffer (stream order: the stores are complete when this kernel starts)
__global__ void peer_barrier_kernel(int* pad_local, mlb::DpArgs a, int slot) {
  __threadfence_system();
  for (int p = 0; p < a.world; ++p)
    if (p != a.rank) mlb::st_release_sys(a.pad_peer[p] + slot + a.rank, a.epoch);
  for (int p = 0; p < a.world; ++p)
    if (p != a.rank) mlb::dp_spin_until_ge(pad_local + slot + p, a.epoch, pad_local);
}

extern "C" int mlb_peer_barrier(int* pad_local, const long long* pad_peer_ptrs, int rank, int world, int epoch, int slot,
                                cudaStream_t st) {
  if (world > mlb::GEMM_MAX_PEERS) return -2;
  mlb::DpArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < world; ++i) a.pad_peer[i] = reinterpret_cast<int*>(pad_peer_ptrs[i]);
  a.pad_local = pad_local; a.rank = rank; a.world = world; a.epoch = epoch;
  peer_barrier_kernel<<<1, 1, 0, st>>>(pad_local, a, slot);
  return (int)cudaGetLastError();
}

extern "C" int mlb_dp_reduce_nvls(int reduce_scatter, float* local, float* mc, int* pad_local,
                                  const long long* pad_peer_ptrs, long long n, int rank, int world, int epoch,
                                  float scale, int num_ctas, cudaStream_t st) {
  using namespace mlb;
  if (world > GEMM_MAX_PEERS || n % (world * 4) != 0) return -2;
  DpNvlsArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < world; ++i) a.pad_peer[i] = reinterpret_cast<int*>(pad_peer_ptrs[i]);
  a.mc = mc; a.local = local; a.pad_local = pad_local;
  a.n = n; a.rank = rank; a.world = world; a.epoch = epoch; a.scale = scale; a.reduce_scatter = reduce_scatter;
  dp_reduce_nvls_kernel<<<num_ctas > 0 ? num_ctas : 16, 512, 0, st>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int mlb_dp_reduce(int reduce_scatter, float* local, const long long* peer_ptrs, int* pad_local,
                             const long long* pad_peer_ptrs, long long n, int rank, int world, int epoch,
                             float scale, int num_ctas, cudaStream_t st) {
  using namespace mlb;
  if (world > GEMM_MAX_PEERS || n % (world * 4) != 0) return -2;
  DpArgs a;
  for (int i = 0; i < GEMM_MAX_PEERS; ++i) {
    a.peer[i] = i < world ? reinterpret_cast<float*>(peer_ptrs[i]) : nullptr;
    a.pad_peer[i] = i < world ? reinterpret_cast<int*>(pad_peer_ptrs[i]) : nullptr;
  }
  a.peer[rank] = local;
  a.pad_local = pad_local;
  a.n = n; a.rank = rank; a.world = world; a.epoch = epoch; a.scale = scale; a.reduce_scatter = reduce_scatter;
  dp_reduce_kernel<<<num_ctas > 0 ? num_ctas : 32, 512, 0, st>>>(a);
  return (int)cudaGetLastError();
}

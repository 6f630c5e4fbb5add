// Attention-probability dropout mask of the sm_100a attention kernels: a counter-based hash of
// (seed, batch * heads + head, query position, key position), so forward, dK/dV and dQ regenerate the same mask from
// their own tile layouts and nothing is stored (the reference gets this from the FA-2 library's Philox stream,
// megatron/model/transformer.py:538-553 `dropout_p`).
//
//   head_key  = mix(seed_hi ^ bh * C1)                              (one per CTA / head)
//   row_key   = mix(seed_lo ^ row * C2) + head_key                  (one per query row)
//   bytes     = mix(row_key ^ (key >> 2) * C3)                      (4 decisions: keys 4q .. 4q+3, one byte each)
//   dropped   = byte(key & 3) < threshold,  threshold = round(p * 256)  -> the effective rate is threshold / 256 and
//               kept probabilities are scaled by 256 / (256 - threshold), which is exact for that rate.
// Plain C++ (also compiled for the host by tests/emu to pin the Python replica of the mask used by the tests).
#pragma once
#include <stdint.h>

#ifndef MLB_HD
#if defined(__CUDACC__)
#define MLB_HD __host__ __device__ __forceinline__
#else
#define MLB_HD inline
#endif
#endif

namespace mlb {

struct DropoutParams {
  uint32_t seed_lo, seed_hi;
  uint32_t threshold;      // 0 = no dropout; dropped iff byte < threshold
  float inv_keep;          // 256 / (256 - threshold)
};

MLB_HD uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
MLB_HD uint32_t drop_head_key(uint32_t seed_hi, uint32_t bh) { return drop_mix(seed_hi ^ (bh * 0x85EBCA77u)); }
MLB_HD uint32_t drop_row_key(uint32_t seed_lo, uint32_t head_key, uint32_t row) {
  return drop_mix(seed_lo ^ (row * 0x9E3779B1u)) + head_key;
}
// the four decision bytes of keys 4 * key_quad .. 4 * key_quad + 3
MLB_HD uint32_t drop_bytes(uint32_t row_key, uint32_t key_quad) { return drop_mix(row_key ^ (key_quad * 0xC2B2AE3Du)); }
MLB_HD bool drop_is_dropped(uint32_t bytes, uint32_t key, uint32_t threshold) {
  return ((bytes >> ((key & 3u) * 8u)) & 0xFFu) < threshold;
}

static inline DropoutParams make_dropout_params(float p, unsigned long long seed) {
  DropoutParams d;
  d.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull);
  d.seed_hi = (uint32_t)(seed >> 32);
  int t = (int)(p * 256.0f + 0.5f);
  t = t < 0 ? 0 : (t > 255 ? 255 : t);
  d.threshold = (uint32_t)t;
  d.inv_keep = 256.0f / (256.0f - (float)t);
  return d;
}

}  // namespace mlb

// Host build of csrc/attention_dropout.cuh: the keep / drop decisions exactly as the attention kernels compute them.
#include <cstddef>

#include "../../megatron_llm_b200/csrc/attention_dropout.cuh"

// keep[bh][row][key] (1 = kept) for bh in [0, n_bh), row in [0, rows), key in [0, keys); returns the threshold
extern "C" int emu_dropout_keep(float p, unsigned long long seed, int n_bh, int rows, int keys, unsigned char* keep,
                                float* inv_keep) {
  using namespace mlb;
  const DropoutParams d = make_dropout_params(p, seed);
  *inv_keep = d.inv_keep;
  for (int bh = 0; bh < n_bh; ++bh) {
    const uint32_t hk = drop_head_key(d.seed_hi, (uint32_t)bh);
    for (int r = 0; r < rows; ++r) {
      const uint32_t rk = drop_row_key(d.seed_lo, hk, (uint32_t)r);
      for (int k = 0; k < keys; ++k)
        keep[((size_t)bh * rows + r) * keys + k] = drop_is_dropped(drop_bytes(rk, (uint32_t)k >> 2), (uint32_t)k, d.threshold) ? 0 : 1;
    }
  }
  return (int)d.threshold;
}

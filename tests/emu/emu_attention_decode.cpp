// Runs csrc/attention_decode.cu (the real kernel source, compiled for the host against tests/emu/cuda_emu) on CPU threads.
#define MLB_HOST_EMULATION 1
#include "cuda_emu.h"

#include "../../megatron_llm_b200/csrc/attention_decode.cu"

template <int D, typename T>
static void run(const mlb::DecodeParams& p) {
  cuda_emu::launch(dim3(p.n_splits, p.nkv, p.batch), mlb::DEC_THREADS, [&] { mlb::attn_decode_split_kernel<D, T>(p); });
  const long long rows = (long long)p.batch * p.nkv * p.sq * p.g;
  cuda_emu::launch(dim3((unsigned)((rows + 3) / 4)), 128, [&] { mlb::attn_decode_merge_kernel<D, T>(p); });
}

extern "C" int emu_attn_decode(int dtype, const void* q, const void* k, const void* v, const long long* q_str,
                               const long long* k_str, const long long* v_str, int batch, int sq, int sk, int nq,
                               int nkv, int head_dim, int window, float softmax_scale, int n_splits,
                               int keys_per_split, float* part_o, float* part_ml, void* out) {
  using namespace mlb;
  DecodeParams p;
  const int r = fill_decode_params(p, q, k, v, q_str, k_str, v_str, batch, sq, sk, nq, nkv, head_dim, window,
                                   softmax_scale, n_splits, keys_per_split, part_o, part_ml, out);
  if (r) return r;
  if (dtype == DT_BF16) { if (head_dim == 128) run<128, __nv_bfloat16>(p); else run<64, __nv_bfloat16>(p); return 0; }
  if (dtype == DT_F16) { if (head_dim == 128) run<128, __half>(p); else run<64, __half>(p); return 0; }
  return -100;
}

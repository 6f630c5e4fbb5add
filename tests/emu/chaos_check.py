"""Run the tcgen05 kernels on the functional model with schedule fuzzing (MLB_EMU_CHAOS=<seed> in the environment: random
stalls in front of barrier operations, TMA loads, MMAs and tensor-memory accesses) and compare with the oracle.

    MLB_EMU_CHAOS=3 python chaos_check.py <attention.so> <gemm.so>

Prints one JSON line {case: relative error}.  A barrier protocol that only holds while the producer stays ahead of the
consumer shows up as a wrong result or a hang (the caller's timeout)."""
import ctypes
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from megatron_llm_b200.ops.attention import attention_reference, dropout_keep_mask  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr())
S3 = lambda t: (ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0))
rel = lambda a, r: ((a.float() - r).norm() / r.norm()).item()
SEED = 0x0BAD_5EED_0000_0001


def attention(lib, s, n, nkv, hn, window, p):
    g = torch.Generator().manual_seed(s)
    q, k, v, do = (torch.randn(1, s, h, hn, generator=g).bfloat16() for h in (n, nkv, nkv, n))
    b, gq, w = 1, n // nkv, -1 if window is None else window
    hm = (ctypes.c_int * 6)(gq, 0, 1, 0, 1, 0)
    sc, seed = ctypes.c_float(1 / math.sqrt(hn)), ctypes.c_ulonglong(SEED if p else 0)
    out = torch.zeros(s, b, n, hn).bfloat16().permute(1, 0, 2, 3)
    lse, delta = torch.zeros(b, n, s), torch.zeros(b, n, s)
    assert lib.mlb_attn_fwd_ex(P(q), P(k), P(v), S3(q), S3(k), S3(v), n, nkv, nkv, hm, gq, s, b, n, w, sc, P(out),
                               ctypes.c_longlong(b * n * hn), ctypes.c_longlong(n * hn), P(lse), hn, 0, ctypes.c_float(p), seed, None) == 0
    dq, dk, dv = (torch.zeros(s, b, h, hn).bfloat16().permute(1, 0, 2, 3) for h in (n, nkv, nkv))
    assert lib.mlb_attn_bwd_ex(P(q), P(k), P(v), P(out), P(do), S3(q), S3(k), S3(v), S3(out), S3(do), n, nkv, nkv, hm, gq, s, b,
                               n, w, sc, P(lse), P(delta), P(dq), P(dk), P(dv), S3(dq), S3(dk), S3(dv), hn, 0, ctypes.c_float(p),
                               seed, None) == 0
    qf, kf, vf = (x.float().requires_grad_() for x in (q, k, v))
    keep = dropout_keep_mask(SEED, p, b, n, s, s) if p else None
    ref = attention_reference(qf, kf, vf, True, window, None, p, keep)
    ref.backward(do.float())
    return max(rel(out, ref), rel(dq, qf.grad), rel(dk, kf.grad), rel(dv, vf.grad))


def gemm(lib, two_cta, M, N, K):
    torch.manual_seed(M)
    A, B = torch.randn(M, K).bfloat16(), torch.randn(N, K).bfloat16()
    c = torch.zeros(M, N).bfloat16()
    if two_cta:
        assert lib.mlb_gemm_bf16_2cta(P(A), P(B), P(c), M, N, K, K, K, N, 0, 0, 0, 0, 2, None) == 0
    else:
        assert lib.mlb_gemm_bf16(P(A), P(B), P(c), M, N, K, K, K, N, 0, 0, 0, 128, 0, 2, None) == 0
    return rel(c, A.float() @ B.float().t())


if __name__ == "__main__":
    attn, gm = ctypes.CDLL(sys.argv[1]), ctypes.CDLL(sys.argv[2])
    res = {"attn_two_tile_s512": attention(attn, 512, 2, 1, 128, None, 0.0),
           "attn_one_tile_s640_window_dropout": attention(attn, 640, 2, 2, 64, 200, 0.2),
           "gemm_1cta_5_tiles": gemm(gm, False, 640, 128, 320),
           "gemm_2cta_3_tiles": gemm(gm, True, 768, 256, 320)}
    print("RESULT " + json.dumps(res))

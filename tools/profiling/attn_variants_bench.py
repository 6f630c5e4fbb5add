"""Timing of the attention kernel variants that have no hardware numbers yet (decode, dropout, fp16) against the
``flash_attn`` library; CUDA events, median of 20 after 5 warm-up calls, a 256 MB buffer written between calls so every
call starts from a cold L2.  One JSON line per case.

  python tools/profiling/attn_variants_bench.py decode|dropout|fp16 > profiles/attn_variants_<what>.jsonl

decode reports the achieved fraction of MEASURED_PEAKS.json's copy bandwidth (K and V are each read once per KV group)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from megatron_llm_b200.ops import _ext  # noqa: E402

DEV = "cuda"
_flush = None


def timed(fn, n=20, warm=5):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def peak_bw():
    try:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        return float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6485.0


def decode():
    from flash_attn import flash_attn_func
    mod = _ext.load()
    for b, sk, nq, nkv, hn in ((1, 4096, 32, 32, 128), (8, 4096, 32, 32, 128), (1, 32768, 64, 8, 128), (16, 2048, 128, 8, 64),
                               (4, 8192, 71, 1, 64)):
        kmem = torch.randn(sk, b, nkv, hn, device=DEV).bfloat16()
        vmem = torch.randn(sk, b, nkv, hn, device=DEV).bfloat16()
        q = torch.randn(b, 1, nq, hn, device=DEV).bfloat16()
        k, v = kmem.transpose(0, 1), vmem.transpose(0, 1)
        ours = timed(lambda: mod.attn_decode(q, k, v, -1, 1.0 / math.sqrt(hn), 0))
        lib = timed(lambda: flash_attn_func(q, k, v, causal=True))
        gb = 2 * sk * b * nkv * hn * 2 / 1e9
        print(json.dumps(dict(case="decode", b=b, sk=sk, nq=nq, nkv=nkv, hn=hn, ours_us=ours * 1e3, flash_attn_us=lib * 1e3,
                              ours_gbps=gb / (ours * 1e-3), frac_of_copy_bw=gb / (ours * 1e-3) / peak_bw())), flush=True)


def training(dtype, p):
    from flash_attn import flash_attn_func
    mod = _ext.load()
    for b, s, nq, nkv, hn in ((1, 4096, 32, 32, 128), (2, 2048, 16, 16, 64), (1, 4096, 64, 8, 128)):
        q, k, v, do = (torch.randn(b, s, n, hn, device=DEV).to(dtype) for n in (nq, nkv, nkv, nq))
        sc = 1.0 / math.sqrt(hn)
        out, lse = mod.attn_fwd(q, k, v, True, -1, sc, p, 1234)
        fwd = timed(lambda: mod.attn_fwd(q, k, v, True, -1, sc, p, 1234))
        bwd = timed(lambda: mod.attn_bwd(do, q, k, v, out, lse, True, -1, sc, p, 1234))
        base_f = timed(lambda: mod.attn_fwd(q.bfloat16(), k.bfloat16(), v.bfloat16(), True, -1, sc, 0.0, 0)) if p or dtype != torch.bfloat16 else fwd
        ql, kl, vl = (t.clone().requires_grad_() for t in (q, k, v))
        lib_f = timed(lambda: flash_attn_func(ql, kl, vl, dropout_p=p, causal=True))
        flops = 4 * b * nq * s * s * hn / 2
        print(json.dumps(dict(case="training", dtype=str(dtype)[6:], dropout=p, b=b, s=s, nq=nq, nkv=nkv, hn=hn, fwd_us=fwd * 1e3,
                              bwd_us=bwd * 1e3, fwd_tflops=flops / (fwd * 1e-3) / 1e12, bf16_nodrop_fwd_us=base_f * 1e3,
                              flash_attn_fwd_us=lib_f * 1e3)), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "decode"
    if what == "decode":
        decode()
    elif what == "dropout":
        training(torch.bfloat16, 0.1)
    else:
        training(torch.float16, 0.0)

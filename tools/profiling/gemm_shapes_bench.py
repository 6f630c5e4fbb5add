"""Time the 1-CTA and 2-CTA tcgen05 GEMM kernels against cuBLAS on the Llama-2-7B shapes (TP=1 and TP=8)."""
import json, os, sys, torch
sys.path.insert(0, ".")
from megatron_llm_b200 import ops

def bench(fn, iters=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

shapes = [  # (kind, M, N, K)
    ("nt", 4096, 12288, 4096), ("nt", 4096, 4096, 4096), ("nt", 4096, 22016, 4096), ("nt", 4096, 4096, 11008),
    ("nt", 4096, 32000, 4096), ("nn", 4096, 4096, 12288), ("nn", 4096, 4096, 22016), ("nn", 4096, 11008, 4096),
    ("tn", 12288, 4096, 4096), ("tn", 22016, 4096, 4096), ("tn", 4096, 11008, 4096),
    ("nt", 4096, 1536, 4096), ("nt", 4096, 4096, 512), ("nt", 4096, 2752, 4096), ("nt", 4096, 4096, 1376),
]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
res = []
for kind, M, N, K in shapes:
    torch.manual_seed(0)
    if kind == "nt":
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        f = lambda out=None: ops.gemm_nt(a, b, out=out); lib = lambda: torch.matmul(a, b.t()); ref = a.float() @ b.float().t()
    elif kind == "nn":
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); b = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
        f = lambda out=None: ops.gemm_nn(a, b, out=out); lib = lambda: torch.matmul(a, b); ref = a.float() @ b.float()
    else:
        a = torch.randn(K, M, device="cuda", dtype=torch.bfloat16); b = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
        acc = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        f = lambda out=None: ops.gemm_tn(a, b, out=acc, accumulate=True); lib = lambda: torch.matmul(a.t(), b); ref = None
    row = {"kind": kind, "M": M, "N": N, "K": K}
    flops = 2.0 * M * N * K
    for mode in ("0", "1"):
        os.environ["MLB200_GEMM_2CTA"] = mode
        try:
            out = f()
            torch.cuda.synchronize()
            if ref is not None:
                err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
            else:
                acc.zero_(); f(); torch.cuda.synchronize()
                err = ((acc - a.float().t() @ b.float()).abs().max() / acc.abs().max()).item()
            t = bench(f)
            row[f"cta{int(mode)+1}_tflops"] = round(flops / t / 1e9, 1); row[f"cta{int(mode)+1}_relerr"] = round(err, 5)
        except Exception as e:
            row[f"cta{int(mode)+1}_error"] = repr(e)[:200]
    tl = bench(lib)
    row["cublas_tflops"] = round(flops / tl / 1e9, 1)
    print(json.dumps(row), flush=True)
    res.append(row)

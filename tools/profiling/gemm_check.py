"""GPU sanity + timing for the tcgen05 GEMM variants (run under gpurun, each case in its own timeout)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from megatron_llm_b200 import ops

def bench(fn, iters=20):
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

def main():
    which = sys.argv[1]
    M, N, K = [int(x) for x in sys.argv[2:5]]
    torch.manual_seed(0)
    dev = "cuda"
    res = {"case": which, "M": M, "N": N, "K": K}
    if which == "nt":
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        out = ops.gemm_nt(a, b); ref = a.float() @ b.float().t()
        fn = lambda: ops.gemm_nt(a, b, out=out); lib = lambda: torch.matmul(a, b.t())
    elif which == "nn":
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        out = ops.gemm_nn(a, b); ref = a.float() @ b.float()
        fn = lambda: ops.gemm_nn(a, b, out=out); lib = lambda: torch.matmul(a, b)
    elif which == "tn":
        a = torch.randn(K, M, device=dev, dtype=torch.bfloat16); b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        out = ops.gemm_tn(a, b); ref = a.float().t() @ b.float()
        fn = lambda: ops.gemm_tn(a, b, out=out); lib = lambda: torch.matmul(a.t(), b)
    elif which == "tn_acc":
        a = torch.randn(K, M, device=dev, dtype=torch.bfloat16); b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        base = torch.randn(M, N, device=dev, dtype=torch.float32)
        out = base.clone(); ops.gemm_tn(a, b, out=out, accumulate=True); ref = base + a.float().t() @ b.float()
        fn = lambda: ops.gemm_tn(a, b, out=out, accumulate=True); lib = lambda: torch.matmul(a.t(), b)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    res.update(max_abs_err=err, rel_err=rel, ok=bool(rel < 2e-2))
    if res["ok"] and len(sys.argv) > 5:
        t = bench(fn); tl = bench(lib)
        res.update(ms=t, tflops=2 * M * N * K / t / 1e9, cublas_ms=tl, cublas_tflops=2 * M * N * K / tl / 1e9)
    print(json.dumps(res))

main()

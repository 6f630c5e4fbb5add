"""Runs every memory-bound hot kernel at its Llama-2-7B shape (4096 tokens) a few times, for ncu captures and CUDA-event
timings of the per-kernel roofline table (profiles/README.md).   python tools/profiling/ew_drive.py [--time]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from megatron_llm_b200 import ops  # noqa: E402
from megatron_llm_b200.ops import _ext  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16
T, H, F, V = 4096, 4096, 11008, 32000
torch.manual_seed(0)
mod = _ext.load()


def timeit(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for _ in range(n):
        flush.zero_()                      # > L2: every timed launch starts cold
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


cases = {}
x = torch.randn(T, H, device=DEV, dtype=bf, requires_grad=True)
res = torch.randn(T, H, device=DEV, dtype=bf)
w = torch.ones(H, device=DEV, dtype=bf, requires_grad=True)
y, hid = ops.rmsnorm(x, w, 1e-5, residual=res)
gy = torch.randn_like(y)
cases["rmsnorm_fwd(+residual)"] = (lambda: ops.rmsnorm(x, w, 1e-5, residual=res), 4 * T * H * 2)
cases["rmsnorm_bwd"] = (lambda: torch.autograd.grad(y, x, gy, retain_graph=True), 3 * T * H * 2)
h2 = torch.randn(T, 2 * F, device=DEV, dtype=bf, requires_grad=True)
g_out = ops.glu(h2, "swiglu")
gg = torch.randn_like(g_out)
cases["swiglu_fwd"] = (lambda: ops.glu(h2, "swiglu"), 3 * T * F * 2)
cases["swiglu_bwd"] = (lambda: torch.autograd.grad(g_out, h2, gg, retain_graph=True), 5 * T * F * 2)
qkv = torch.randn(T, 1, 3 * H, device=DEV, dtype=bf)
tab = ops.rope_table(128, T, device=DEV)
cases["rope_qkv"] = (lambda: ops.rope_qkv_(qkv, 32, 3, 128, tab), 2 * 2 * T * H * 2)
logits = torch.randn(T, V, device=DEV, dtype=bf)
tgt = torch.randint(0, V, (T,), device=DEV)
cases["ce_stats"] = (lambda: ops.ce_local_stats(logits, tgt, 0), T * V * 2)
n = 202_000_000
p = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
p16 = torch.empty(n, device=DEV, dtype=bf)
seg_start = torch.tensor([0, n], device=DEV, dtype=torch.int64); seg_wd = torch.zeros(1, device=DEV)
coef = torch.ones(1, device=DEV); inf = torch.zeros(1, device=DEV, dtype=torch.int32)
cases["adamw_flat(1 layer)"] = (lambda: mod.adamw_flat(p, g, m, v, p16, 0, seg_start, seg_wd, None, 1e-4, 0.9, 0.95, 1e-8,
                                                       0.1, 0.05, coef, inf, []), n * (4 * 4 + 3 * 4 + 2))
ids = torch.randint(0, V, (1, T), device=DEV)
emb = torch.randn(V, H, device=DEV, dtype=bf)
emb.main_grad = torch.zeros(V, H, device=DEV)
cases["embedding_fwd"] = (lambda: ops.embedding_lookup(ids, emb, 0, sbh=True), 2 * T * H * 2)
dout = torch.randn(T, 1, H, device=DEV, dtype=bf)
cases["embedding_bwd(main_grad)"] = (lambda: mod.embedding_bwd(ids, dout, emb.main_grad, 0, True), T * H * (2 + 8))
xb = torch.randn(T, H, device=DEV, dtype=bf); bias = torch.randn(H, device=DEV, dtype=bf)
cases["bias_dropout_add(p=0.1)"] = (lambda: ops.bias_dropout_add(xb, bias, res, 0.1, True), 3 * T * H * 2)

if "--time" in sys.argv:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
    for name, (fn, nbytes) in cases.items():
        us = timeit(fn)
        print(json.dumps({"kernel": name, "us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1),
                          "of_measured_copy_bw": round(nbytes / us / 1e3 / peak, 3), "algorithmic_MB": round(nbytes / 2 ** 20, 1)}), flush=True)
else:
    for name, (fn, _) in cases.items():
        for _ in range(3):
            fn()
    torch.cuda.synchronize()

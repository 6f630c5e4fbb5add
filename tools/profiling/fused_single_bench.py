"""Cost of the fused kernels' tile order / flag checks / epilogue WITHOUT any transfer: one virtual rank (loopback,
world = 1) vs the plain 2-CTA GEMM on the same shapes, CUDA-graph replay timing.
  python tools/profiling/fused_single_bench.py [M]        (MLB200_FUSED_GROUP=2|4|8 overrides the tile-group height)"""
import json, os, sys
import torch
sys.path.insert(0, ".")
from megatron_llm_b200 import ops
from megatron_llm_b200.parallel.symm import TPCommunicator

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
h, ffn = 4096, 11008
bf = torch.bfloat16
comm = TPCommunicator.loopback_group(1, max_rows_per_rank=M, max_k=2 * ffn, max_n=2 * ffn, num_comm_ctas=2, sms=148)[0]


def graph_time(fn, n=10, reps=5):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        before = comm.counters()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            for _ in range(n):
                fn()
        adv = comm.end_capture(before)
        comm.begin_replay(before, adv); g.replay(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            comm.begin_replay(before, adv); g.replay()
        e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (n * reps) * 1e3


shapes = [("ag", h, 3 * h, False, "qkv"), ("ag", h, 2 * ffn, False, "mlp up+gate"), ("rs", h, h, False, "attn dense"),
          ("rs", ffn, h, False, "mlp down"), ("ag", h, ffn, True, "dgrad mlp down"), ("ag", h, h, True, "dgrad attn dense"),
          ("rs", 2 * ffn, h, True, "dgrad up+gate"), ("rs", 3 * h, h, True, "dgrad qkv")]
for kind, K, N, tw, what in shapes:
    w = (torch.randn(K, N, device="cuda", dtype=bf) if tw else torch.randn(N, K, device="cuda", dtype=bf)) * 0.02
    x = torch.randn(M, K, device="cuda", dtype=bf)
    out = torch.empty(M, N, device="cuda", dtype=bf)
    gemm = (lambda: ops.gemm_nn(x, w, out=out)) if tw else (lambda: ops.gemm_nt(x, w, out=out))
    t_gemm = graph_time(gemm)
    if kind == "ag":
        t_f = graph_time(lambda: comm.ag_gemm(x, w, tw, out=out, keep=False))
    else:
        t_f = graph_time(lambda: comm.gemm_rs(x, w, tw))
    print(json.dumps({"kind": kind, "what": what, "M": M, "K": K, "N": N, "gemm_us": round(t_gemm, 1),
                      "fused_world1_us": round(t_f, 1), "group": os.environ.get("MLB200_FUSED_GROUP", "auto")}), flush=True)
assert comm.error_flag() == 0

"""Per-shape timing of the fused TP kernels against (NCCL collective + plain GEMM) and the plain GEMM alone.

torchrun --nproc-per-node N tools/profiling/fused_bench.py [hidden ffn seq]   -> one JSON line per shape on rank 0
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timeit_graph(fn, comm, n=10, reps=5):
    """Device time per call without the host: ``n`` calls captured in one CUDA graph, replayed ``reps`` times
    (the fused kernels' epochs continue through TPCommunicator.begin_replay)."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        before = comm.counters()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            for _ in range(n):
                fn()
        adv = comm.end_capture(before)
        comm.begin_replay(before, adv)
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            comm.begin_replay(before, adv)
            g.replay()
        e.record()
        torch.cuda.synchronize()
    torch.cuda.current_stream().wait_stream(st)
    t = torch.tensor([s.elapsed_time(e) / (n * reps)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from megatron_llm_b200 import ops
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.parallel.symm import TPCommunicator
    ps.initialize_model_parallel(world, 1)
    group = ps.get_tensor_model_parallel_group()
    h, ffn, seq = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 11008, 4096)))
    in_graph = os.environ.get("MLB200_BENCH_IN_GRAPH", "1") == "1"     # time graph replays (no host overhead)
    t = (lambda fn: timeit_graph(fn, comm)) if in_graph else timeit
    ctas = int(os.environ.get("MLB200_COMM_CTAS", "32"))
    m = seq // world
    comm = TPCommunicator(group, max_rows_per_rank=m, max_k=max(h, ffn), max_n=max(h, 2 * ffn // world, 3 * h // world),
                          num_comm_ctas=ctas)
    dev = torch.device("cuda", rank)
    bf = torch.bfloat16
    shapes = [  # (kind, K, N, transposed_weight, what)
        ("ag", h, 3 * h // world, False, "fwd qkv"),
        ("ag", h, 2 * ffn // world, False, "fwd mlp up+gate"),
        ("rs", h // world, h, False, "fwd attn dense"),
        ("rs", ffn // world, h, False, "fwd mlp down"),
        ("ag", h, ffn // world, True, "bwd dgrad mlp down"),
        ("ag", h, h // world, True, "bwd dgrad attn dense"),
        ("rs", 2 * ffn // world, h, True, "bwd dgrad mlp up+gate"),
        ("rs", 3 * h // world, h, True, "bwd dgrad qkv"),
    ]
    for kind, K, N, tw, what in shapes:
        w = (torch.randn(K, N, device=dev, dtype=bf) if tw else torch.randn(N, K, device=dev, dtype=bf)) * 0.02
        res = {"kind": kind, "what": what, "M": seq, "K": K, "N": N, "world": world, "comm_ctas": ctas,
               "timing": "cuda-graph replay" if in_graph else "eager launches"}
        if kind == "ag":
            x = torch.randn(m, K, device=dev, dtype=bf)
            full = torch.empty(seq, K, device=dev, dtype=bf)
            out = torch.empty(seq, N, device=dev, dtype=bf)
            gemm = (lambda: ops.gemm_nn(full, w, out=out)) if tw else (lambda: ops.gemm_nt(full, w, out=out))
            res["gemm_ms"] = t(gemm)
            res["nccl_ms"] = t(lambda: dist.all_gather_into_tensor(full, x, group=group))
            res["nccl_plus_gemm_ms"] = t(lambda: (dist.all_gather_into_tensor(full, x, group=group), gemm()))
            res["fused_ms"] = t(lambda: comm.ag_gemm(x, w, tw, out=out, keep=not tw))   # bwd consumes the gather at once
        else:
            a = torch.randn(seq, K, device=dev, dtype=bf)
            part = torch.empty(seq, N, device=dev, dtype=bf)
            red = torch.empty(m, N, device=dev, dtype=bf)
            gemm = (lambda: ops.gemm_nn(a, w, out=part)) if tw else (lambda: ops.gemm_nt(a, w, out=part))
            res["gemm_ms"] = t(gemm)
            res["nccl_ms"] = t(lambda: dist.reduce_scatter_tensor(red, part, group=group))
            res["nccl_plus_gemm_ms"] = t(lambda: (gemm(), dist.reduce_scatter_tensor(red, part, group=group)))
            res["fused_ms"] = t(lambda: comm.gemm_rs(a, w, tw))
        res["tflops_fused"] = 2.0 * seq * K * N / res["fused_ms"] / 1e9
        res["tflops_gemm"] = 2.0 * seq * K * N / res["gemm_ms"] / 1e9
        if rank == 0:
            print(json.dumps(res), flush=True)
    assert comm.error_flag() == 0
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


main()

"""Data-parallel gradient reduction: peer-pointer kernel vs NVLS kernel vs NCCL on one big fp32 bucket.
  torchrun --nproc-per-node N tools/profiling/dp_bench.py [MB]     -> one JSON line per variant on rank 0"""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from megatron_llm_b200.parallel import state as ps
from megatron_llm_b200.parallel.symm import DPCommunicator
ps.initialize_model_parallel(1, 1)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = (mb << 20) // 4
n -= n % (world * 64)
group = ps.get_data_parallel_group()


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for rs in (True, False):
    res = {"world": world, "MB": mb, "op": "reduce_scatter" if rs else "all_reduce"}
    for variant in ("peer", "nvls"):
        os.environ["MLB200_DP_NVLS"] = "1" if variant == "nvls" else "0"
        for ctas in (16, 32, 64):
            comm = DPCommunicator(group, n, num_ctas=ctas)
            if variant == "nvls" and not comm.use_nvls:
                res[f"{variant}_{ctas}ctas_ms"] = None
                continue
            comm.buffer.normal_()
            def run():
                comm.reduce_bucket(comm.buffer, 0, n, reduce_scatter=rs).wait()
            res[f"{variant}_{ctas}ctas_ms"] = round(timeit(run), 3)
            del comm
            torch.cuda.empty_cache()
    buf = torch.randn(n, device="cuda")
    if rs:
        out = torch.empty(n // world, device="cuda")
        res["nccl_ms"] = round(timeit(lambda: dist.reduce_scatter_tensor(out, buf, op=dist.ReduceOp.AVG, group=group)), 3)
    else:
        res["nccl_ms"] = round(timeit(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)), 3)
    if rank == 0:
        print(json.dumps(res), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)

"""Hermetic multi-process harness: spawn ``world_size`` processes on CPU/Gloo (or CUDA/NCCL), run ``fn(rank, world,
*args)`` in each, re-raise the first failure.  The reference's distributed tests need 8 real GPUs under torchrun
(tests/test_utilities.py:6-29); these run anywhere."""
import os
import socket
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, backend, fn, args, errq):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
        if backend == "nccl":
            os.environ["MLB200_FORCE_CPU"] = "0"      # (never inherit a CPU-only switch from the parent process)
            torch.cuda.set_device(rank)
        else:
            os.environ["MLB200_FORCE_CPU"] = "1"
        dist.init_process_group(backend, rank=rank, world_size=world)
        fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world_size, *args, backend="gloo"):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, backend, fn, args, errq)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    failed = [p for p in procs if p.exitcode != 0]
    for p in procs:
        if p.is_alive():
            p.terminate()
    if failed or not errq.empty():
        msgs = []
        while not errq.empty():
            r, tb = errq.get()
            msgs.append(f"--- rank {r} ---\n{tb}")
        raise RuntimeError("distributed test failed:\n" + "\n".join(msgs))

"""Process start-up (parity: megatron/initialize.py:26-275).

Differences: runs on CUDA/NCCL or CPU/Gloo (the reference asserts CUDA, :36); no JIT kernel compilation at
start-up (the extension is built ahead of time); no nvFuser flags (gone from torch); after the process groups
exist, symmetric-memory communicators are bound to the TP and DP groups for the fused collective kernels.
"""
from __future__ import annotations

import os
import random
import time
from datetime import timedelta

import numpy as np
import torch
import torch.distributed as dist

from . import global_vars
from .arguments import parse_args, validate_args
from .global_vars import get_adlr_autoresume, get_args, get_tensorboard_writer, set_global_variables
from .parallel import state as ps
from .parallel.random import model_parallel_cuda_manual_seed
from .utils.device import use_cuda


def initialize_megatron(extra_args_provider=None, args_defaults={}, ignore_unknown_args=False,
                        allow_no_cuda=True, args_list=None):
    """Parse/validate args, set globals, init torch.distributed + model-parallel groups, seed RNGs."""
    args = parse_args(extra_args_provider, ignore_unknown_args, args_list=args_list)
    if args.use_checkpoint_args or args_defaults.get("use_checkpoint_args", False):
        assert args.load is not None, "--use_checkpoint_args requires --load argument"
        from .checkpointing import load_args_from_checkpoint
        load_args_from_checkpoint(args)
    validate_args(args, args_defaults)
    set_global_variables(args)

    def finish_mpu_init():
        args = get_args()
        _initialize_distributed()
        _use_side_stream_for_graphs()
        if args.rank == 0:
            print("> setting random seeds to {} ...".format(args.seed))
        _set_random_seed(args.seed, args.data_parallel_random_init)

    finish_mpu_init()
    _init_autoresume()
    _bind_symmetric_communicators()
    return None


def _use_side_stream_for_graphs():
    """CUDA-graph capture cannot touch the legacy default stream, and autograd's AccumulateGrad nodes run on the stream
    that was current when they were created (model construction): with --cuda_graph_microbatch do everything on one
    ordinary side stream from the very beginning and capture on that same stream."""
    args = get_args()
    if use_cuda() and getattr(args, "cuda_graph_microbatch", False) \
            and torch.cuda.current_stream() == torch.cuda.default_stream():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(side)


def _destroy_gloo_group():
    """Orderly exit of a CPU (gloo) run.  gloo sends complete locally, so a rank that finishes first (the first
    pipeline stage of an evaluation loop, say) would close its sockets while its peers still have receives to post:
    wait for everybody -- bounded, and not at all after an uncaught exception -- before tearing the group down."""
    import sys
    if not dist.is_initialized():
        return
    try:
        if getattr(sys, "last_value", None) is None and getattr(sys, "last_exc", None) is None:
            dist.monitored_barrier(timeout=timedelta(seconds=60))
    except Exception:
        pass
    try:
        dist.destroy_process_group()
    except Exception:
        pass


def _initialize_distributed():
    args = get_args()
    device_count = torch.cuda.device_count() if use_cuda() else 0
    if dist.is_initialized():
        if args.rank == 0:
            print("torch distributed is already initialized, skipping initialization ...", flush=True)
        args.rank = dist.get_rank()
        args.world_size = dist.get_world_size()
    else:
        if args.rank == 0:
            print("> initializing torch distributed ...", flush=True)
        if device_count > 0:
            device = args.rank % device_count
            if args.local_rank is not None:
                assert args.local_rank == device, "expected local-rank to be the same as rank % device-count."
            else:
                args.local_rank = device
            torch.cuda.set_device(device)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if device_count > 0 and args.distributed_backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=args.distributed_backend, world_size=args.world_size, rank=args.rank,
                                timeout=timedelta(minutes=10), **kwargs)
        if args.distributed_backend == "gloo":
            # a gloo group that is still alive when the interpreter finalises aborts the process ("terminate called
            # without an active exception") whenever the ranks do not exit in lock step, turning a finished run into
            # a non-zero exit code; the group this function created is torn down at exit.  (NCCL runs keep their
            # validated behaviour: bench.py shuts down explicitly, with a watchdog.)
            import atexit
            atexit.register(_destroy_gloo_group)
    if ps.model_parallel_is_initialized():
        print("model parallel is already initialized")
    else:
        ps.initialize_model_parallel(args.tensor_model_parallel_size, args.pipeline_model_parallel_size,
                                     args.virtual_pipeline_model_parallel_size,
                                     args.pipeline_model_parallel_split_rank)
        if args.rank == 0:
            print(f"> initialized tensor model parallel with size {ps.get_tensor_model_parallel_world_size()}")
            print(f"> initialized pipeline model parallel with size {ps.get_pipeline_model_parallel_world_size()}")


def _bind_symmetric_communicators():
    """Fused GEMM+collective (TP) and peer-memory grad reduction (DP) need a symmetric heap per group."""
    args = get_args()
    if not use_cuda() or args.distributed_backend != "nccl":
        return
    try:
        from .parallel import symm
        tp = ps.get_tensor_model_parallel_world_size()
        flag = getattr(args, "fused_tp_comm", None)
        if flag is None:      # automatic: a single TP group over the whole job
            flag = tp > 1 and tp == dist.get_world_size()
        want = os.environ.get("MLB200_FUSED_TP", "1" if flag else "0") == "1"
        if want and tp > 1:
            symm.bind_tp_communicator(args)
    except Exception as e:  # never fatal: the NCCL path is the checked fallback
        if args.rank == 0:
            print(f"WARNING: symmetric-memory communicators unavailable ({e!r}); using NCCL collectives", flush=True)


def _init_autoresume():
    autoresume = get_adlr_autoresume()
    if autoresume:
        dist.barrier()
        autoresume.init()
        dist.barrier()


def _set_random_seed(seed_, data_parallel_random_init=False):
    """Different seed per pipeline stage (and per DP rank if requested); TP ranks share the default generator."""
    if seed_ is not None and seed_ > 0:
        seed = seed_ + (100 * ps.get_pipeline_model_parallel_rank())
        if data_parallel_random_init:
            seed = seed + (10 * ps.get_data_parallel_rank())
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        model_parallel_cuda_manual_seed(seed)
    else:
        raise ValueError("Seed ({}) should be a positive integer.".format(seed_))


def write_args_to_tensorboard():
    args = get_args()
    writer = get_tensorboard_writer()
    if writer:
        for arg in vars(args):
            writer.add_text(arg, str(getattr(args, arg)), global_step=args.iteration)


def set_jit_fusion_options(args=None):
    """nvFuser is gone from torch; the fused ops are hand-written kernels, so there is nothing to configure."""
    return None


def _compile_dependencies(args=None):
    """The sm_100a extension is built ahead of time (``python -m megatron_llm_b200.ops.build``)."""
    return None

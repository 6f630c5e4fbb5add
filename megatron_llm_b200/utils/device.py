"""Device abstraction.

The reference hard-asserts CUDA (megatron/initialize.py:36) and allocates on
``torch.cuda.current_device()`` everywhere.  This framework runs the same code on a
B200 (NCCL) or on CPU (Gloo) -- BASELINE config #1 is a CPU/Gloo plumbing run -- so
every allocation goes through :func:`current_device`.
"""
from __future__ import annotations

import os

import torch

_FORCE_CPU = False


def force_cpu(flag: bool = True) -> None:
    global _FORCE_CPU
    _FORCE_CPU = flag


def use_cuda() -> bool:
    if _FORCE_CPU or os.environ.get("MLB200_FORCE_CPU", "0") == "1":
        return False
    return torch.cuda.is_available()


def current_device() -> torch.device:
    if use_cuda():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def synchronize() -> None:
    if use_cuda():
        torch.cuda.synchronize()


def default_backend() -> str:
    return "nccl" if use_cuda() else "gloo"

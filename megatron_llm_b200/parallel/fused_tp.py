"""Fused GEMM+collective dispatch for the tensor-parallel group (see parallel/symm.py and csrc/comm.cu).

``active(x)`` is True when a symmetric-memory communicator is bound to the TP group and ``x`` is a bf16 CUDA
tensor; the layers then call :func:`ag_gemm` / :func:`gemm_rs` instead of NCCL + GEMM.
"""
from __future__ import annotations

import torch

_COMM = None  # bound by parallel.symm.bind_tp_communicator()


def bind(comm) -> None:
    global _COMM
    _COMM = comm


def communicator():
    return _COMM


def active(x: torch.Tensor) -> bool:
    return _COMM is not None and x.is_cuda and x.dtype == torch.bfloat16 and _COMM.enabled


def ag_gemm(x_shard: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False, out=None):
    """all-gather(x_shard along dim 0) then GEMM.  Returns (out2d [s*b, N], gathered input)."""
    return _COMM.ag_gemm(x_shard, weight, transposed_weight, out=out)


def gemm_rs(x2d: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False):
    """GEMM then reduce-scatter along dim 0.  Returns out2d [rows/tp, N]."""
    return _COMM.gemm_rs(x2d, weight, transposed_weight)

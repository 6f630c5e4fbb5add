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


def active_column(x_shard: torch.Tensor, weight: torch.Tensor) -> bool:
    """Column-parallel linear under sequence parallelism: forward = all-gather(x) -> GEMM, backward = GEMM ->
    reduce-scatter.  True when the bound communicator's buffers fit this call (else the layer uses NCCL)."""
    if not active(x_shard):
        return False
    k = x_shard.size(-1)
    rows = x_shard.numel() // k
    return (rows % 128 == 0 and rows <= _COMM.max_rows and k % 8 == 0 and k <= min(_COMM.max_k, _COMM.max_n)
            and weight.size(0) % 8 == 0 and weight.dtype == torch.bfloat16)


def active_row(x_full: torch.Tensor, weight: torch.Tensor) -> bool:
    """Row-parallel linear under sequence parallelism: forward = GEMM -> reduce-scatter, backward = all-gather(dy)
    -> GEMM."""
    if not active(x_full):
        return False
    k = x_full.size(-1)
    total = x_full.numel() // k
    n = weight.size(0)
    if total % _COMM.world != 0:
        return False
    rows = total // _COMM.world
    return (rows % 128 == 0 and rows <= _COMM.max_rows and k % 8 == 0 and n % 8 == 0
            and n <= min(_COMM.max_n, _COMM.max_k) and weight.dtype == torch.bfloat16)


def ag_gemm(x_shard: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False, out=None,
            keep: bool = True):
    """all-gather(x_shard along dim 0) then GEMM.  Returns (out2d [s*b, N], gathered input); ``keep=False`` if the
    gathered input is consumed before the next fused call (lets the push variant skip a copy)."""
    return _COMM.ag_gemm(x_shard, weight, transposed_weight, out=out, keep=keep)


def gemm_rs(x2d: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False):
    """GEMM then reduce-scatter along dim 0.  Returns out2d [rows/tp, N]."""
    return _COMM.gemm_rs(x2d, weight, transposed_weight)


def active_all_reduce(x2d: torch.Tensor, n_out: int) -> bool:
    """GEMM -> all-reduce (no sequence parallelism: Row-parallel forward, Column-parallel dgrad): ``x2d`` [M, K] is
    this rank's full-length operand, the result is [M, n_out] on every rank."""
    if not active(x2d) or getattr(_COMM, "ar_n", 0) <= 0 or x2d.dim() != 2:
        return False
    M, k = x2d.shape
    if M % _COMM.world != 0:
        return False
    rows = M // _COMM.world
    return (rows % 128 == 0 and rows <= _COMM.max_rows and k % 8 == 0 and n_out % 8 == 0
            and n_out <= min(_COMM.ar_n, _COMM.max_n))


def gemm_ar(x2d: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False, keep: bool = True):
    """GEMM then all-reduce over the TP group.  Returns out2d [M, N]."""
    return _COMM.gemm_ar(x2d, weight, transposed_weight, keep=keep)

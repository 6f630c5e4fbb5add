"""Fused GEMM+collective dispatch for the tensor-parallel group (see parallel/symm.py and csrc/comm.cu).

``active(x)`` is True when a symmetric-memory communicator is bound to the TP group and ``x`` is a bf16 CUDA
tensor; the layers then call :func:`ag_gemm` / :func:`gemm_rs` instead of NCCL + GEMM.
"""
from __future__ import annotations

import torch

_COMM = None  # bound by parallel.symm.bind_tp_communicator()
_PREFER = {}  # (kind, M, N, K) -> take the fused kernel?  (see prefer_fused)


def bind(comm) -> None:
    global _COMM
    _COMM = comm
    _PREFER.clear()          # the fused-vs-library decisions depend on the communicator (world size, transports)


def communicator():
    return _COMM


def active(x: torch.Tensor) -> bool:
    return _COMM is not None and x.is_cuda and x.dtype == torch.bfloat16 and _COMM.enabled


# ------------------------------------------------------------------------------------------------------------------
# Per-call choice between the fused kernel and (library collective + plain GEMM), from measured rates.
#   fused   ~ max(bytes / R_fused(world), 1.15 x GEMM)      the transfer runs inside the GEMM; 15 % = puller SMs / tile order
#   library ~ GEMM + bytes / R_nccl(world)                   serialised on one stream
# Rates in GB/s per GPU (profiles/README.md round 2: fused per-pair tables at TP=2 and TP=8, graph-replay timing):
#   in-kernel transports (bulk-copy pullers / TMA-store pushes)  2 GPUs ~450, 8 GPUs ~350;  NVLS gather 8 GPUs ~520
#   NCCL 2.28 (NVLS)                                             2 GPUs ~320, 8 GPUs ~560
# At TP=2/4 the fused kernels win on every layer shape; at TP=8 (7 peers, link-bound) they win only behind the large
# GEMMs (h->4h forward, 4h->h / qkv dgrad), and the library call wins behind the small ones.  MLB200_FUSED_TP_FORCE=1
# takes the fused kernel whenever it is applicable.
import os as _os

_R_FUSED = {2: 450.0, 4: 400.0, 8: 350.0}
_R_FUSED_NVLS_AG = {2: 450.0, 4: 480.0, 8: 520.0}
_R_NCCL = {2: 320.0, 4: 450.0, 8: 560.0}
_GEMM_TFLOPS = 1700.0


def prefer_fused(kind: str, M: int, N: int, K: int) -> bool:
    """kind 'ag' (all-gather -> GEMM, M = gathered rows) or 'rs' (GEMM -> reduce-scatter, M = full rows)."""
    if _os.environ.get("MLB200_FUSED_TP_FORCE", "0") == "1" or _COMM is None:
        return True
    key = (kind, M, N, K)
    hit = _PREFER.get(key)
    if hit is not None:
        return hit
    w = _COMM.world
    near = min(_R_NCCL, key=lambda k: abs(k - w))
    gemm_us = 2.0 * M * N * K / (_GEMM_TFLOPS * 1e6)
    moved = (w - 1) / w * M * (K if kind == "ag" else N) * 2          # bytes in (gather) or out (scatter) per GPU
    r_f = (_R_FUSED_NVLS_AG if (kind == "ag" and getattr(_COMM, "nvls_ag", False)) else _R_FUSED)[near]
    fused_us = max(moved / (r_f * 1e3), 1.15 * gemm_us)
    lib_us = gemm_us + moved / (_R_NCCL[near] * 1e3)
    _PREFER[key] = fused_us <= lib_us
    return _PREFER[key]


def active_column(x_shard: torch.Tensor, weight: torch.Tensor) -> bool:
    """Column-parallel linear under sequence parallelism: forward = all-gather(x) -> GEMM, backward = GEMM ->
    reduce-scatter.  True when the bound communicator's buffers fit this call (else the layer uses NCCL)."""
    if not active(x_shard):
        return False
    k = x_shard.size(-1)
    rows = x_shard.numel() // k
    return (rows % 128 == 0 and rows <= _COMM.max_rows and k % 8 == 0 and k <= min(_COMM.max_k, _COMM.max_n)
            and weight.size(0) % 8 == 0 and weight.dtype == torch.bfloat16
            and prefer_fused("ag", rows * _COMM.world, weight.size(0), k))


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
            and n <= min(_COMM.max_n, _COMM.max_k) and weight.dtype == torch.bfloat16
            and prefer_fused("rs", total, n, k))


def column_backward_fused() -> bool:
    """Column-parallel backward: dX = reduce_scatter(dY @ W).  The unfused form launches the reduce-scatter
    asynchronously and hides it behind the independent wgrad GEMM (reference layers.py:285-296); the fused
    GEMM -> reduce-scatter kernel overlaps the transfer only with its own tiles and leaves its tail exposed, which
    measured 50-250 us slower per call at TP=2 and TP=8.  Default: unfused + overlapped; MLB200_COLUMN_BWD_FUSED=1 (or
    MLB200_FUSED_TP_FORCE=1) takes the fused kernel."""
    return _os.environ.get("MLB200_COLUMN_BWD_FUSED", "0") == "1" or _os.environ.get("MLB200_FUSED_TP_FORCE", "0") == "1"


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

"""Tensor/sequence-parallel autograd mappings.

Parity target: megatron/core/tensor_parallel/mappings.py (:13-124 primitives, :127-246 the
seven autograd pairs, :253-278 wrappers).  All seven conjugate pairs are generated from one
table of (forward-primitive, backward-primitive) instead of seven hand-written classes; the
class names of the reference are kept as aliases because its tests call
``Cls.forward/backward/symbolic`` directly.

These are the *unfused* NCCL/Gloo paths -- the oracle and fallback for the fused sm_100a
GEMM+collective kernels in :mod:`megatron_llm_b200.parallel.fused_tp`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import state as ps
from .tp_utils import split_tensor_along_last_dim


def _tp():
    return ps.get_tensor_model_parallel_world_size()


def _group():
    return ps.get_tensor_model_parallel_group()


# ---- primitives --------------------------------------------------------------------------

def _identity(x):
    return x


def _reduce(x):
    if _tp() == 1:
        return x
    if not x.is_contiguous():
        x = x.contiguous()   # e.g. the stride-0 expanded grad of ``sum()``; never reduce into aliased storage
    dist.all_reduce(x, group=_group())
    return x


def _split_along_last_dim(x):
    if _tp() == 1:
        return x
    parts = split_tensor_along_last_dim(x, _tp())
    return parts[ps.get_tensor_model_parallel_rank()].contiguous()


def _split_along_first_dim(x):
    if _tp() == 1:
        return x
    n = x.size(0)
    assert n % _tp() == 0, "First dimension of the tensor should be divisible by tensor parallel size"
    loc = n // _tp()
    r = ps.get_tensor_model_parallel_rank()
    return x[r * loc:(r + 1) * loc].contiguous()


def _gather_along_last_dim(x):
    world = _tp()
    if world == 1:
        return x
    x = x.contiguous()
    flat = torch.empty((world * x.size(0),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=_group())
    # [w*d0, ..., c] -> [d0, ..., w*c]
    return torch.cat(list(flat.chunk(world, dim=0)), dim=-1).contiguous()


def _gather_along_first_dim(x):
    world = _tp()
    if world == 1:
        return x
    shape = list(x.shape)
    shape[0] *= world
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=_group())
    return out


def _reduce_scatter_along_first_dim(x):
    world = _tp()
    if world == 1:
        return x
    shape = list(x.shape)
    assert shape[0] % world == 0, "First dimension of the tensor should be divisible by tensor parallel size"
    shape[0] //= world
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x.contiguous(), group=_group())
    return out


# ---- autograd pairs ----------------------------------------------------------------------

def _make_pair(name, fwd, bwd, doc):
    def symbolic(graph, input_):
        return fwd(input_)

    def forward(ctx, input_):
        return fwd(input_)

    def backward(ctx, grad_output):
        return bwd(grad_output)

    return type(name, (torch.autograd.Function,), {
        "__doc__": doc,
        "symbolic": staticmethod(symbolic),
        "forward": staticmethod(forward),
        "backward": staticmethod(backward),
    })


_CopyToModelParallelRegion = _make_pair(
    "_CopyToModelParallelRegion", _identity, _reduce,
    "identity forward / all-reduce backward")
_ReduceFromModelParallelRegion = _make_pair(
    "_ReduceFromModelParallelRegion", _reduce, _identity,
    "all-reduce forward / identity backward")
_ScatterToModelParallelRegion = _make_pair(
    "_ScatterToModelParallelRegion", _split_along_last_dim, _gather_along_last_dim,
    "split-last forward / gather-last backward")
_GatherFromModelParallelRegion = _make_pair(
    "_GatherFromModelParallelRegion", _gather_along_last_dim, _split_along_last_dim,
    "gather-last forward / split-last backward")
_ScatterToSequenceParallelRegion = _make_pair(
    "_ScatterToSequenceParallelRegion", _split_along_first_dim, _gather_along_first_dim,
    "split-first forward / gather-first backward")
_ReduceScatterToSequenceParallelRegion = _make_pair(
    "_ReduceScatterToSequenceParallelRegion", _reduce_scatter_along_first_dim,
    _gather_along_first_dim, "reduce-scatter forward / gather-first backward")


class _GatherFromSequenceParallelRegion(torch.autograd.Function):
    """gather-first forward; backward is a reduce-scatter when the consumer computed on the
    full sequence in tensor-parallel fashion, else a plain split."""

    @staticmethod
    def symbolic(graph, input_, tensor_parallel_output_grad=True):
        return _gather_along_first_dim(input_)

    @staticmethod
    def forward(ctx, input_, tensor_parallel_output_grad=True):
        ctx.tensor_parallel_output_grad = tensor_parallel_output_grad
        return _gather_along_first_dim(input_)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.tensor_parallel_output_grad:
            return _reduce_scatter_along_first_dim(grad_output), None
        return _split_along_first_dim(grad_output), None


# ---- public wrappers ---------------------------------------------------------------------

def copy_to_tensor_model_parallel_region(input_):
    return _CopyToModelParallelRegion.apply(input_)


def reduce_from_tensor_model_parallel_region(input_):
    return _ReduceFromModelParallelRegion.apply(input_)


def scatter_to_tensor_model_parallel_region(input_):
    return _ScatterToModelParallelRegion.apply(input_)


def gather_from_tensor_model_parallel_region(input_):
    return _GatherFromModelParallelRegion.apply(input_)


def scatter_to_sequence_parallel_region(input_):
    return _ScatterToSequenceParallelRegion.apply(input_)


def gather_from_sequence_parallel_region(input_, tensor_parallel_output_grad=True):
    return _GatherFromSequenceParallelRegion.apply(input_, tensor_parallel_output_grad)


def reduce_scatter_to_sequence_parallel_region(input_):
    return _ReduceScatterToSequenceParallelRegion.apply(input_)

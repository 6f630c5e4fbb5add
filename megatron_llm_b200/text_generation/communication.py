"""Small collectives used by generation across pipeline stages and ranks (parity: text_generation/communication.py).
All helpers are device-agnostic (CUDA/NCCL or CPU/Gloo)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..utils.device import current_device


def recv_from_prev_pipeline_rank_(recv_buffer=None):
    if not ps.is_pipeline_first_stage():
        assert recv_buffer is not None
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, recv_buffer, ps.get_pipeline_model_parallel_prev_rank())])
        for r in reqs:
            r.wait()


def send_to_next_pipeline_rank(tensor=None):
    if not ps.is_pipeline_last_stage():
        assert tensor is not None
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, tensor, ps.get_pipeline_model_parallel_next_rank())])
        for r in reqs:
            r.wait()


def _is_cuda(tensor):
    assert tensor is not None


def _is_cuda_contiguous(tensor):
    _is_cuda(tensor)
    assert tensor.is_contiguous()


def _alloc(size, dtype):
    return torch.empty(size, dtype=dtype, device=current_device())


def broadcast_from_last_pipeline_stage(size, dtype, tensor=None):
    """Last stage -> every stage of the pipeline group."""
    is_last = ps.is_pipeline_last_stage()
    if ps.is_pipeline_first_stage() and is_last:
        return tensor
    if is_last:
        _is_cuda_contiguous(tensor)
    else:
        tensor = _alloc(size, dtype)
    dist.broadcast(tensor, ps.get_pipeline_model_parallel_last_rank(), group=ps.get_pipeline_model_parallel_group())
    return tensor


def broadcast_from_last_to_first_pipeline_stage(size, dtype, tensor=None):
    """Last stage -> first stage only (through the embedding group, which is exactly {first, last})."""
    is_last, is_first = ps.is_pipeline_last_stage(), ps.is_pipeline_first_stage()
    if is_first and is_last:
        return tensor
    if is_last or is_first:
        if is_last:
            _is_cuda_contiguous(tensor)
        else:
            tensor = _alloc(size, dtype)
        dist.broadcast(tensor, ps.get_pipeline_model_parallel_last_rank(), group=ps.get_embedding_group())
    else:
        tensor = None
    return tensor


def copy_from_last_to_first_pipeline_stage(size, dtype, tensor=None):
    """In-place version of the above: the first stage's ``tensor`` receives the last stage's values."""
    is_last, is_first = ps.is_pipeline_last_stage(), ps.is_pipeline_first_stage()
    if is_first and is_last:
        return
    if is_last or is_first:
        _is_cuda(tensor)
        is_contiguous = tensor.is_contiguous()
        if is_contiguous:
            buf = tensor
        else:
            buf = tensor.contiguous() if is_last else _alloc(size, dtype)
        dist.broadcast(buf, ps.get_pipeline_model_parallel_last_rank(), group=ps.get_embedding_group())
        if is_first and not is_contiguous:
            tensor[...] = buf


def broadcast_tensor(size, dtype, tensor=None, rank=0):
    if dist.get_rank() == rank:
        _is_cuda_contiguous(tensor)
    else:
        tensor = _alloc(size, dtype)
    dist.broadcast(tensor, rank)
    return tensor


def broadcast_list(size, dtype, list_values=None, rank=0):
    tensor = None
    if dist.get_rank() == rank:
        tensor = torch.tensor(list_values, dtype=dtype, device=current_device())
    return broadcast_tensor(size, dtype, tensor=tensor, rank=rank)


def broadcast_int_list(size, int_list=None, rank=0):
    return broadcast_list(size, torch.int64, list_values=int_list, rank=rank)


def broadcast_float_list(size, float_list=None, rank=0):
    return broadcast_list(size, torch.float32, list_values=float_list, rank=rank)

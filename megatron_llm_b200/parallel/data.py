"""Broadcast a dict of same-dtype tensors from TP-rank 0 to its TP group.

Parity target: megatron/core/tensor_parallel/data.py:65-105 (two broadcasts: sizes, then one
flattened payload).  Here sizes travel as one small int64 tensor and the payload is a single
flat buffer that receivers slice zero-copy.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..utils.device import current_device
from . import state as ps

_MAX_DATA_DIM = 5


def _check_data_types(keys, data, target_dtype):
    for key in keys:
        assert data[key].dtype == target_dtype, (
            f"{key} has data type {data[key].dtype} which is different than {target_dtype}")


def _build_key_size_numel_dictionaries(keys, data):
    max_dim = _MAX_DATA_DIM
    sizes = [0] * (max_dim * len(keys))
    if ps.get_tensor_model_parallel_rank() == 0:
        for i, key in enumerate(keys):
            assert data[key].dim() < max_dim, "you should increase MAX_DATA_DIM"
            for j, s in enumerate(data[key].size()):
                sizes[i * max_dim + j] = s
    sizes_t = torch.tensor(sizes, dtype=torch.int64, device=current_device())
    if ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(sizes_t, ps.get_tensor_model_parallel_src_rank(),
                       group=ps.get_tensor_model_parallel_group())
    sizes_cpu = sizes_t.cpu().tolist()
    key_size, key_numel, total = {}, {}, 0
    for i, key in enumerate(keys):
        shape = []
        for j in range(max_dim):
            s = sizes_cpu[i * max_dim + j]
            if s <= 0:
                break
            shape.append(s)
        n = 1
        for s in shape:
            n *= s
        key_size[key], key_numel[key] = shape, n
        total += n
    return key_size, key_numel, total


def broadcast_data(keys, data, datatype):
    """data is only read on TP-rank 0; every rank returns {key: tensor on device}."""
    key_size, key_numel, total = _build_key_size_numel_dictionaries(keys, data)
    dev = current_device()
    if ps.get_tensor_model_parallel_rank() == 0:
        _check_data_types(keys, data, datatype)
        flat = torch.cat([data[k].contiguous().view(-1) for k in keys], dim=0).to(dev, non_blocking=True)
    else:
        flat = torch.empty(total, device=dev, dtype=datatype)
    if ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(flat, ps.get_tensor_model_parallel_src_rank(),
                       group=ps.get_tensor_model_parallel_group())
    out, off = {}, 0
    for k in keys:
        out[k] = flat[off:off + key_numel[k]].view(key_size[k])
        off += key_numel[k]
    return out

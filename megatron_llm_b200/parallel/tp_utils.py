"""Tensor-parallel slicing helpers (parity: megatron/core/tensor_parallel/utils.py)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

from ..utils.core_utils import divide
from ..utils.device import current_device
from . import state as ps


def split_tensor_along_last_dim(tensor: torch.Tensor, num_partitions: int,
                                contiguous_split_chunks: bool = False) -> List[torch.Tensor]:
    last = tensor.dim() - 1
    chunk = divide(tensor.size(last), num_partitions)
    parts = torch.split(tensor, chunk, dim=last)
    if contiguous_split_chunks:
        return tuple(p.contiguous() for p in parts)
    return parts


def split_tensor_into_1d_equal_chunks(tensor: torch.Tensor, new_buffer: bool = False) -> torch.Tensor:
    """This TP rank's 1/tp slice of the flattened tensor."""
    n = tensor.numel() // ps.get_tensor_model_parallel_world_size()
    start = n * ps.get_tensor_model_parallel_rank()
    flat = tensor.reshape(-1)[start:start + n]
    if new_buffer:
        out = torch.empty(n, dtype=tensor.dtype, device=tensor.device, requires_grad=False)
        out.copy_(flat)
        return out
    return flat


def gather_split_1d_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`split_tensor_into_1d_equal_chunks`."""
    world = ps.get_tensor_model_parallel_world_size()
    out = torch.empty(world * tensor.numel(), dtype=tensor.dtype, device=tensor.device,
                      requires_grad=False)
    with torch.no_grad():      # pure data movement (the received chunk may carry requires_grad)
        dist.all_gather_into_tensor(out, tensor.detach().contiguous(), group=ps.get_tensor_model_parallel_group())
    return out


class VocabUtility:
    """Vocabulary is split into ``world_size`` contiguous ranges ``[first, last)``."""

    @staticmethod
    def vocab_range_from_per_partition_vocab_size(per_partition_vocab_size: int, rank: int,
                                                  world_size: int) -> Sequence[int]:
        first = rank * per_partition_vocab_size
        return first, first + per_partition_vocab_size

    @staticmethod
    def vocab_range_from_global_vocab_size(global_vocab_size: int, rank: int,
                                           world_size: int) -> Sequence[int]:
        per = divide(global_vocab_size, world_size)
        return VocabUtility.vocab_range_from_per_partition_vocab_size(per, rank, world_size)

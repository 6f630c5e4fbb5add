"""Model/data-parallel process-group state ("mpu").

Parity target: megatron/core/parallel_state.py (reference :51-205 group construction,
:259-341 getters, :424-471 rank helpers, :484-494 global memory buffer, :497-524 teardown).

Design: instead of nested loops that rebuild rank lists, the world is a 3-D grid
``[pp, dp, tp]`` (tp fastest).  Every group family is a slice of that grid, so the
rank algebra is a handful of integer formulas and works identically under NCCL (B200) and
Gloo (CPU).  Groups are created in a fixed global order so ``new_group`` (collective over
the world) stays consistent on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist

from ..utils.device import current_device


@dataclass
class _State:
    initialized: bool = False
    world_size: int = 1
    rank: int = 0
    tp: int = 1
    pp: int = 1
    dp: int = 1
    vpp: Optional[int] = None
    vpp_rank: Optional[int] = None
    split_rank: Optional[int] = None
    # groups
    tp_group: object = None
    pp_group: object = None
    dp_group: object = None
    mp_group: object = None
    embedding_group: object = None
    position_embedding_group: object = None
    # global ranks
    tp_ranks: List[int] = field(default_factory=list)
    pp_ranks: List[int] = field(default_factory=list)
    dp_ranks: List[int] = field(default_factory=list)
    mp_ranks: List[int] = field(default_factory=list)
    embedding_ranks: List[int] = field(default_factory=list)
    position_embedding_ranks: List[int] = field(default_factory=list)
    # overrides used by the offline tools (checkpoint resharder) to fake a world
    tp_world_override: Optional[int] = None
    pp_world_override: Optional[int] = None
    tp_rank_override: Optional[int] = None
    pp_rank_override: Optional[int] = None
    memory_buffer: object = None


_S = _State()


# ----------------------------------------------------------------------------------------
# rank algebra (pure functions; unit-tested without any process group)
# ----------------------------------------------------------------------------------------

def grid_coords(rank: int, tp: int, pp: int, world: int):
    """rank -> (pp_rank, dp_rank, tp_rank) for the [pp, dp, tp] grid."""
    dp = world // (tp * pp)
    return rank // (dp * tp), (rank // tp) % dp, rank % tp


def tensor_group_ranks(tp: int, pp: int, world: int) -> List[List[int]]:
    return [list(range(i * tp, (i + 1) * tp)) for i in range(world // tp)]


def pipeline_group_ranks(tp: int, pp: int, world: int) -> List[List[int]]:
    stride = world // pp
    return [list(range(i, world, stride)) for i in range(stride)]


def data_group_ranks(tp: int, pp: int, world: int) -> List[List[int]]:
    per_stage = world // pp
    out = []
    for p in range(pp):
        for t in range(tp):
            out.append(list(range(p * per_stage + t, (p + 1) * per_stage, tp)))
    return out


def model_group_ranks(tp: int, pp: int, world: int) -> List[List[int]]:
    dp = world // (tp * pp)
    dgroups = data_group_ranks(tp, pp, world)
    return [[g[d] for g in dgroups] for d in range(dp)]


def embedding_ranks_of(pipe_ranks: List[int], split_rank: Optional[int]):
    """first+last stage (and the split stage for encoder-decoder) share the tied embedding;
    position embeddings live on the first stage (+ split stage)."""
    if len(pipe_ranks) == 1:
        return list(pipe_ranks), list(pipe_ranks)
    emb = [pipe_ranks[0], pipe_ranks[-1]]
    pos = [pipe_ranks[0]]
    if split_rank is not None:
        if pipe_ranks[split_rank] not in emb:
            emb = [pipe_ranks[0], pipe_ranks[split_rank], pipe_ranks[-1]]
        if pipe_ranks[split_rank] not in pos:
            pos = [pipe_ranks[0], pipe_ranks[split_rank]]
    return emb, pos


# ----------------------------------------------------------------------------------------
# construction / teardown
# ----------------------------------------------------------------------------------------

def initialize_model_parallel(tensor_model_parallel_size: int = 1,
                              pipeline_model_parallel_size: int = 1,
                              virtual_pipeline_model_parallel_size: Optional[int] = None,
                              pipeline_model_parallel_split_rank: Optional[int] = None) -> None:
    assert dist.is_initialized(), "torch.distributed is not initialized"
    world = dist.get_world_size()
    rank = dist.get_rank()
    tp, pp = tensor_model_parallel_size, pipeline_model_parallel_size
    if world % (tp * pp) != 0:
        raise RuntimeError(
            f"world_size ({world}) is not divisible by tensor_model_parallel_size ({tp}) x "
            f"pipeline_model_parallel_size ({pp})")
    if virtual_pipeline_model_parallel_size is not None:
        if not pp > 2:
            raise RuntimeError("pipeline-model-parallel size should be greater than 2 with "
                               "interleaved schedule")
    assert not _S.initialized, "model parallel already initialized"
    _S.world_size, _S.rank, _S.tp, _S.pp, _S.dp = world, rank, tp, pp, world // (tp * pp)
    _S.vpp = virtual_pipeline_model_parallel_size
    _S.vpp_rank = 0 if virtual_pipeline_model_parallel_size is not None else None
    _S.split_rank = pipeline_model_parallel_split_rank

    def build(groups):
        mine_g, mine_r = None, None
        for ranks in groups:
            g = dist.new_group(ranks)
            if rank in ranks:
                mine_g, mine_r = g, ranks
        return mine_g, mine_r

    _S.dp_group, _S.dp_ranks = build(data_group_ranks(tp, pp, world))
    _S.mp_group, _S.mp_ranks = build(model_group_ranks(tp, pp, world))
    _S.tp_group, _S.tp_ranks = build(tensor_group_ranks(tp, pp, world))
    pgroups = pipeline_group_ranks(tp, pp, world)
    _S.pp_group, _S.pp_ranks = build(pgroups)
    emb_lists, pos_lists = zip(*[embedding_ranks_of(r, pipeline_model_parallel_split_rank)
                                 for r in pgroups])
    g, r = build(list(emb_lists))
    _S.embedding_group, _S.embedding_ranks = g, (r or [])
    g, r = build(list(pos_lists))
    _S.position_embedding_group, _S.position_embedding_ranks = g, (r or [])
    _S.memory_buffer = GlobalMemoryBuffer()
    _S.initialized = True


def model_parallel_is_initialized() -> bool:
    return _S.initialized


def destroy_model_parallel() -> None:
    """Forget all groups (the communicators themselves are released with the default process group)."""
    global _S
    _S = _State()


# ----------------------------------------------------------------------------------------
# getters
# ----------------------------------------------------------------------------------------

def _need(g, name):
    assert g is not None, f"{name} is not initialized"
    return g


def get_model_parallel_group():
    return _need(_S.mp_group, "model parallel group")


def get_tensor_model_parallel_group():
    return _need(_S.tp_group, "intra_layer_model parallel group")


def get_pipeline_model_parallel_group():
    return _need(_S.pp_group, "pipeline_model parallel group")


def get_data_parallel_group():
    return _need(_S.dp_group, "data parallel group")


def get_embedding_group():
    return _need(_S.embedding_group, "embedding group")


def get_position_embedding_group():
    return _need(_S.position_embedding_group, "position embedding group")


def set_tensor_model_parallel_world_size(n):
    _S.tp_world_override = n


def set_pipeline_model_parallel_world_size(n):
    _S.pp_world_override = n


def set_tensor_model_parallel_rank(r):
    _S.tp_rank_override = r


def set_pipeline_model_parallel_rank(r):
    _S.pp_rank_override = r


def set_pipeline_model_parallel_split_rank(r):
    _S.split_rank = r


def get_tensor_model_parallel_world_size() -> int:
    if _S.tp_world_override is not None:
        return _S.tp_world_override
    return _S.tp if _S.initialized else 1


def get_pipeline_model_parallel_world_size() -> int:
    if _S.pp_world_override is not None:
        return _S.pp_world_override
    return _S.pp if _S.initialized else 1


def get_tensor_model_parallel_rank() -> int:
    if _S.tp_rank_override is not None:
        return _S.tp_rank_override
    return _S.rank % _S.tp if _S.initialized else 0


def get_pipeline_model_parallel_rank() -> int:
    if _S.pp_rank_override is not None:
        return _S.pp_rank_override
    return _S.rank // (_S.dp * _S.tp) if _S.initialized else 0


def get_data_parallel_world_size() -> int:
    return _S.dp if _S.initialized else 1


def get_data_parallel_rank() -> int:
    return (_S.rank // _S.tp) % _S.dp if _S.initialized else 0


def get_pipeline_model_parallel_split_rank():
    return _S.split_rank


def get_virtual_pipeline_model_parallel_rank():
    return _S.vpp_rank


def set_virtual_pipeline_model_parallel_rank(r):
    _S.vpp_rank = r


def get_virtual_pipeline_model_parallel_world_size():
    return _S.vpp


def set_virtual_pipeline_model_parallel_world_size(n):
    _S.vpp = n


def is_pipeline_first_stage(ignore_virtual: bool = False) -> bool:
    if not ignore_virtual and _S.vpp is not None and _S.vpp_rank != 0:
        return False
    return get_pipeline_model_parallel_rank() == 0


def is_pipeline_last_stage(ignore_virtual: bool = False) -> bool:
    if not ignore_virtual and _S.vpp is not None and _S.vpp_rank != _S.vpp - 1:
        return False
    return get_pipeline_model_parallel_rank() == get_pipeline_model_parallel_world_size() - 1


def is_rank_in_embedding_group(ignore_virtual: bool = False) -> bool:
    if not _S.initialized:
        return True
    rank = _S.rank
    if ignore_virtual:
        return rank in _S.embedding_ranks
    if rank in _S.embedding_ranks:
        if rank == _S.embedding_ranks[0]:
            return is_pipeline_first_stage(ignore_virtual=False)
        if rank == _S.embedding_ranks[-1]:
            return is_pipeline_last_stage(ignore_virtual=False)
        return True
    return False


def is_rank_in_position_embedding_group() -> bool:
    return (not _S.initialized) or _S.rank in _S.position_embedding_ranks


def is_pipeline_stage_before_split(rank=None) -> bool:
    if get_pipeline_model_parallel_world_size() == 1:
        return True
    rank = get_pipeline_model_parallel_rank() if rank is None else rank
    return _S.split_rank is None or rank < _S.split_rank


def is_pipeline_stage_after_split(rank=None) -> bool:
    if get_pipeline_model_parallel_world_size() == 1:
        return True
    rank = get_pipeline_model_parallel_rank() if rank is None else rank
    return _S.split_rank is None or rank >= _S.split_rank


def is_pipeline_stage_at_split() -> bool:
    r = get_pipeline_model_parallel_rank()
    return is_pipeline_stage_before_split(r) and is_pipeline_stage_after_split(r + 1)


def get_tensor_model_parallel_src_rank() -> int:
    """global rank of tp-rank 0 of the caller's TP group"""
    if not _S.initialized:
        return 0
    return (_S.rank // _S.tp) * _S.tp


def get_data_parallel_src_rank() -> int:
    return _S.dp_ranks[0] if _S.initialized else 0


def get_pipeline_model_parallel_first_rank() -> int:
    return _S.pp_ranks[0] if _S.initialized else 0


def get_pipeline_model_parallel_last_rank() -> int:
    return _S.pp_ranks[-1] if _S.initialized else 0


def get_pipeline_model_parallel_next_rank() -> int:
    r = get_pipeline_model_parallel_rank()
    return _S.pp_ranks[(r + 1) % get_pipeline_model_parallel_world_size()]


def get_pipeline_model_parallel_prev_rank() -> int:
    r = get_pipeline_model_parallel_rank()
    return _S.pp_ranks[(r - 1) % get_pipeline_model_parallel_world_size()]


def get_tensor_model_parallel_ranks() -> List[int]:
    return list(_S.tp_ranks) if _S.initialized else [0]


def get_data_parallel_ranks() -> List[int]:
    return list(_S.dp_ranks) if _S.initialized else [0]


# ----------------------------------------------------------------------------------------
# grow-only named scratch (reference core/utils.py:24-42)
# ----------------------------------------------------------------------------------------

class GlobalMemoryBuffer:
    """Named, grow-only scratch tensors.  Callers must not use one name concurrently."""

    def __init__(self):
        self.buffer = {}

    def get_tensor(self, tensor_shape, dtype, name):
        n = 1
        for d in tensor_shape:
            n *= int(d)
        key = (name, dtype)
        buf = self.buffer.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=dtype, device=current_device(), requires_grad=False)
            self.buffer[key] = buf
        return buf[:n].view(*tensor_shape)


def get_global_memory_buffer() -> GlobalMemoryBuffer:
    if _S.memory_buffer is None:
        _S.memory_buffer = GlobalMemoryBuffer()
    return _S.memory_buffer

"""Data-parallel wrapper with contiguous parameter + gradient buffers, buckets and overlapped reduction.

Parity target: megatron/model/distributed.py (MemoryBuffer :15-39, DistributedDataParallel :75-232:
``main_grad`` views allocated in reverse parameter order, grad-accumulator hooks, ``zero_grad_buffer``,
``allreduce_gradients``, ``broadcast_params``).

B200-first differences:
  * the reference all-reduces the WHOLE fp32 buffer once after the full fwd/bwd (distributed.py:202-209: not
    bucketed, not overlapped).  Here the buffer is cut into buckets in backward order; as soon as every param
    of a bucket has its final gradient (last micro-batch) the bucket's reduction is launched asynchronously --
    by the hand-written peer-memory reduce-scatter/all-reduce kernel (parallel/symm.py, fused with the 1/DP
    scale) when a symmetric communicator is bound to the DP group, else by NCCL/Gloo -- and overlaps the
    rest of backward.
  * model weights are re-homed into one contiguous buffer per dtype with the same offsets as the grad
    buffer, so the optimizer is a single flat AdamW kernel and the ZeRO-1 param all-gather runs in place.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..models.module import MegatronModule
from ..utils.device import current_device
from . import state as ps

_ALIGN = 64  # elements; keeps every param 256B-aligned in the fp32 buffers (16B vector kernels need 8)


class MemoryBuffer:
    """Zero-initialised contiguous buffer; ``get(shape, start)`` returns a view."""

    def __init__(self, numel: int, numel_padded: int, dtype: torch.dtype, device=None):
        self.numel, self.numel_padded, self.dtype = numel, numel_padded, dtype
        self.data = torch.zeros(numel_padded, dtype=dtype, device=device if device is not None else current_device(),
                                requires_grad=False)

    def zero(self):
        self.data.zero_()

    def get(self, shape, start_index):
        end = start_index + int(torch.Size(shape).numel())
        assert end <= self.numel, "requested tensor is out of the buffer range."
        return self.data[start_index:end].view(shape)


class Bucket:
    def __init__(self, index: int, start: int, end: int, params: List[torch.nn.Parameter], deferred: bool = False):
        self.index, self.start, self.end, self.params = index, start, end, params
        self.pending = set()
        self.handle = None
        self.launched = False
        # deferred buckets are never reduced from the backward hooks: their params get further gradient
        # contributions / reductions after backward (tied embeddings, SP norm grads) -- see _deferred_params
        self.deferred = deferred


class DistributedDataParallelBase(MegatronModule):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def allreduce_gradients(self):
        raise NotImplementedError

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def state_dict(self, prefix="", keep_vars=False):
        return self.module.state_dict(prefix=prefix, keep_vars=keep_vars)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def load_state_dict(self, state_dict, strict=True):
        self.module.load_state_dict(state_dict, strict=strict)

    def set_input_tensor(self, input_tensor):
        return self.module.set_input_tensor(input_tensor)


class DistributedDataParallel(DistributedDataParallelBase):
    def __init__(self, module, accumulate_allreduce_grads_in_fp32: bool, use_contiguous_buffers: bool = True,
                 bucket_size_mb: int = 256, overlap_grad_reduce: bool = True, flatten_params: bool = True,
                 use_distributed_optimizer: bool = False):
        super().__init__(module)
        self.accumulate_allreduce_grads_in_fp32 = accumulate_allreduce_grads_in_fp32
        self.use_contiguous_buffers = True  # the flat optimizer requires it; flag kept for CLI parity
        self.overlap_grad_reduce = overlap_grad_reduce
        self.use_distributed_optimizer = use_distributed_optimizer
        self._grad_sync_enabled = False
        self._dp_world = ps.get_data_parallel_world_size()
        self._dp_group = ps.get_data_parallel_group() if ps.model_parallel_is_initialized() else None
        self._symm = None  # bound by parallel.symm.bind_dp_communicator

        def grad_dtype(p):
            return torch.float if accumulate_allreduce_grads_in_fp32 else p.dtype

        # ---- layout: params in reverse order (== order in which backward produces their grads) ----
        params = [p for p in self.module.parameters() if p.requires_grad]
        by_dtype: Dict[torch.dtype, List[torch.nn.Parameter]] = {}
        for p in params:
            by_dtype.setdefault(grad_dtype(p), []).append(p)
        self._grad_buffers: Dict[torch.dtype, MemoryBuffer] = {}
        self._param_buffers: Dict[torch.dtype, Dict[torch.dtype, torch.Tensor]] = {}
        self._grad_buffer_param_index_map: Dict[torch.dtype, Dict[torch.nn.Parameter, tuple]] = {}
        self._buckets: Dict[torch.dtype, List[Bucket]] = {}
        self._param_to_bucket = {}
        bucket_elems = max(1, bucket_size_mb) * 1024 * 1024 // 4
        pad_to = _ALIGN * max(1, self._dp_world)   # bucket boundaries are DP-shardable
        deferred = self._deferred_params(module)
        for gdt, plist in by_dtype.items():
            offset, cur_start = 0, 0
            index_map, buckets, cur = {}, [], []
            # backward order first; the deferred params (tiny norm weights, the embeddings whose gradient arrives
            # last anyway) live behind them in buckets of their own that are reduced after backward
            ordered = [p for p in reversed(plist) if p not in deferred]
            late = [p for p in reversed(plist) if p in deferred]
            for group, is_deferred in ((ordered, False), (late, True)):
                if is_deferred and cur:
                    offset = int(math.ceil(offset / pad_to) * pad_to)
                    buckets.append(Bucket(len(buckets), cur_start, offset, cur))
                    cur, cur_start = [], offset
                for p in group:
                    n = p.data.nelement()
                    index_map[p] = (offset, offset + n)
                    offset += int(math.ceil(n / _ALIGN) * _ALIGN)
                    cur.append(p)
                    if offset - cur_start >= bucket_elems:
                        offset = int(math.ceil(offset / pad_to) * pad_to)
                        buckets.append(Bucket(len(buckets), cur_start, offset, cur, deferred=is_deferred))
                        cur, cur_start = [], offset
            numel = offset
            numel_padded = int(math.ceil(numel / pad_to) * pad_to)
            if cur or not buckets:
                buckets.append(Bucket(len(buckets), cur_start, numel_padded, cur, deferred=bool(late)))
            buf = MemoryBuffer(numel, numel_padded, gdt)
            self._grad_buffers[gdt] = buf
            self._grad_buffer_param_index_map[gdt] = index_map
            self._buckets[gdt] = buckets
            for b in buckets:
                for p in b.params:
                    self._param_to_bucket[p] = (gdt, b)
            # main_grad views + flat parameter storage with identical offsets
            pbufs = {}
            for p in plist:
                s, e = index_map[p]
                p.main_grad = buf.get(p.data.shape, s)
                if flatten_params:
                    pb = pbufs.get(p.dtype)
                    if pb is None:
                        pb = torch.zeros(numel_padded, dtype=p.dtype, device=p.device)
                        pbufs[p.dtype] = pb
                    view = pb[s:e].view(p.data.shape)
                    view.copy_(p.data)
                    p.data = view
            self._param_buffers[gdt] = pbufs

        # ---- grad hooks ----
        self.grad_accs = []
        for p in params:
            p_tmp = p.expand_as(p)
            acc = p_tmp.grad_fn.next_functions[0][0]
            acc.register_hook(self._make_param_hook(p))
            self.grad_accs.append(acc)
            p._grad_ready_callback = self._make_ready_callback(p)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _deferred_params(module) -> set:
        """Params whose gradient is NOT final when their own backward contribution has been accumulated:

        * sequence-parallel params (norm weights, Row biases): TP all-reduced after backward
          (``allreduce_layernorm_grads``) -- a DP reduction in flight would race with it;
        * embedding tables: a tied word embedding receives the LM-head wgrad (fused epilogue, reported through
          ``_grad_ready_callback``) AND the lookup gradient (autograd hook), and with PP>1 / T5 the ``shared`` copies
          are all-reduced over the embedding groups after backward."""
        out = set()
        for p in module.parameters():
            if getattr(p, "sequence_parallel", False) or getattr(p, "shared", False):
                out.add(p)
        from .layers import VocabParallelEmbedding
        for m in module.modules():
            if isinstance(m, (VocabParallelEmbedding, torch.nn.Embedding)):
                out.update(m.parameters(recurse=False))
        return out

    def bind_symmetric_communicator(self, comm):
        self._symm = comm

    def rehome_grad_buffer(self, gdt, new_storage: torch.Tensor):
        """Move the grad buffer of dtype ``gdt`` into ``new_storage`` (symmetric memory) and re-point every
        ``param.main_grad`` view at it."""
        buf = self._grad_buffers[gdt]
        assert new_storage.numel() >= buf.numel_padded and new_storage.dtype == gdt
        new_storage[:buf.numel_padded].copy_(buf.data)
        buf.data = new_storage[:buf.numel_padded]
        for p, (s, e) in self._grad_buffer_param_index_map[gdt].items():
            p.main_grad = buf.get(p.data.shape, s)

    def rehome_param_buffer(self, gdt, pdtype, new_storage: torch.Tensor):
        """Move the flat parameter buffer of dtype ``pdtype`` (grad dtype ``gdt``) into ``new_storage`` (symmetric
        memory: the ZeRO-1 optimizer kernel stores updated weights straight into every DP peer's copy) and re-point
        every ``param.data`` view at it."""
        old = self._param_buffers[gdt][pdtype]
        assert new_storage.numel() >= old.numel() and new_storage.dtype == pdtype
        new = new_storage[:old.numel()]
        new.copy_(old)
        self._param_buffers[gdt][pdtype] = new
        for p, (s, e) in self._grad_buffer_param_index_map[gdt].items():
            if p.dtype == pdtype:
                p.data = new[s:e].view(p.data.shape)

    def _make_param_hook(self, param):
        def hook(*unused):
            if param.grad is not None:
                ops.accumulate_(param.main_grad.view(-1), param.grad.data.contiguous().view(-1)) \
                    if param.main_grad.dtype == torch.float32 else param.main_grad.add_(param.grad.data)
                param.grad = None
            self._on_param_ready(param)
        return hook

    def _make_ready_callback(self, param):
        def cb():
            self._on_param_ready(param)
        return cb

    def enable_grad_sync(self, flag: bool = True):
        """Schedules call this with True right before the LAST micro-batch's backward: from then on a bucket is
        reduced as soon as all of its params have reported their final gradient."""
        self._grad_sync_enabled = flag and self.overlap_grad_reduce and self._dp_world > 1
        if self._grad_sync_enabled:
            for buckets in self._buckets.values():
                for b in buckets:
                    b.pending = set(b.params)
                    b.launched = False
                    b.handle = None

    def _on_param_ready(self, param):
        if not self._grad_sync_enabled:
            return
        gdt, bucket = self._param_to_bucket[param]
        bucket.pending.discard(param)
        if not bucket.pending and not bucket.launched and not bucket.deferred:
            self._launch_bucket(gdt, bucket, async_op=True)

    def _launch_bucket(self, gdt, bucket, async_op):
        buf = self._grad_buffers[gdt]
        view = buf.data[bucket.start:bucket.end]
        bucket.launched = True
        if self._dp_world == 1:
            return
        if self._symm is not None and self._symm.enabled and view.is_cuda and view.dtype == torch.float32:
            bucket.handle = self._symm.reduce_bucket(view, bucket.start, buf.numel_padded,
                                                     reduce_scatter=self.use_distributed_optimizer)
            return
        if view.is_cuda and self.use_distributed_optimizer:
            # ZeRO-1: rank r keeps the r-th slice of every bucket (in-place reduce-scatter)
            n = (bucket.end - bucket.start) // self._dp_world
            r = ps.get_data_parallel_rank()
            bucket.handle = dist.reduce_scatter_tensor(view[r * n:(r + 1) * n], view, op=dist.ReduceOp.AVG,
                                                       group=self._dp_group, async_op=async_op)
        elif view.is_cuda:
            bucket.handle = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self._dp_group, async_op=async_op)
        else:
            view.div_(self._dp_world)
            bucket.handle = dist.all_reduce(view, group=self._dp_group, async_op=async_op)

    # ------------------------------------------------------------------------------------------
    def zero_grad_buffer(self):
        """Set the grad buffer data to zero. Needs to be called at the beginning of each iteration."""
        for buf in self._grad_buffers.values():
            buf.zero()
        self._grad_sync_enabled = False
        for buckets in self._buckets.values():
            for b in buckets:
                b.launched, b.handle = False, None

    def broadcast_params(self):
        for param in self.module.parameters():
            dist.broadcast(param.data, src=ps.get_data_parallel_src_rank(), group=ps.get_data_parallel_group())

    def exposed_reduce_ms(self, reset: bool = True) -> float:
        """Device time the compute stream spent in :meth:`allreduce_gradients` since the last reset (launching the
        deferred buckets and waiting for every bucket): the part of the DP reduction that backward did not hide."""
        pairs = getattr(self, "_exposed_pairs", [])
        if not pairs:
            return 0.0
        torch.cuda.synchronize()
        total = sum(a.elapsed_time(b) for a, b in pairs)
        if reset:
            self._exposed_pairs = []
        return total

    def allreduce_gradients(self):
        """Finish the data-parallel reduction: launch whatever was not overlapped, then wait."""
        account = getattr(self, "account_exposed", False) and self._dp_world > 1 and torch.cuda.is_available()
        if account:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        self._finish_reduction()
        if account:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self._exposed_pairs = getattr(self, "_exposed_pairs", []) + [(ev0, ev1)]

    def _finish_reduction(self):
        if self._dp_world > 1:
            for gdt, buckets in self._buckets.items():
                for b in buckets:
                    if not b.launched:
                        self._launch_bucket(gdt, b, async_op=True)
            for buckets in self._buckets.values():
                for b in buckets:
                    if b.handle is not None:
                        b.handle.wait()
                        b.handle = None
        self._grad_sync_enabled = False

    # accessors used by the flat optimizers
    def grad_buffers(self):
        return self._grad_buffers

    def param_buffers(self):
        return self._param_buffers

    def param_index_maps(self):
        return self._grad_buffer_param_index_map

    def buckets(self):
        return self._buckets

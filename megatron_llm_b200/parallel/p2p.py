"""Pipeline-parallel point-to-point communication.

Parity target: megatron/p2p_communication.py (``_communicate`` :101-251, shape exchange for
``--variable_seq_lengths`` :11-98, the nine send/recv wrappers :254-405, scatter/gather of the activation over
the TP group :148-165,233-249).

Stays on NCCL p2p (north star) but without the reference's ``torch.cuda.synchronize()`` after every transfer
(:231): receives are waited on the *stream* right before use, sends are left in flight and retired lazily, so
a stage's send overlaps its next compute step.
"""
from __future__ import annotations

import operator
from functools import reduce
from typing import List, Optional, Tuple, Union

import torch
import torch.distributed as dist

from ..utils.device import current_device
from . import state as ps
from .tp_utils import gather_split_1d_tensor, split_tensor_into_1d_equal_chunks
from ..utils.core_utils import make_viewless_tensor

Shape = Union[List[int], torch.Size]

_PENDING_SENDS = []  # (requests, tensors kept alive)

# Exposed pipeline-p2p accounting (bench.py: ``exposed_pp_p2p_ms_per_step``): the compute stream waits for every
# receive right where it is posted, so the device time between "posted" and "received" is time this stage did nothing
# else -- pipeline bubble + transfer.  CUDA-event pairs, summed on request (one sync), off by default.
_ACCOUNT = {"on": False, "pairs": []}


def enable_accounting(flag: bool = True) -> None:
    _ACCOUNT["on"] = bool(flag) and torch.cuda.is_available()
    _ACCOUNT["pairs"] = []


def exposed_recv_ms(reset: bool = True) -> float:
    """Sum over the receives since the last reset of (receive complete - receive posted) on the compute stream, ms."""
    if not _ACCOUNT["pairs"]:
        return 0.0
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in _ACCOUNT["pairs"])
    if reset:
        _ACCOUNT["pairs"] = []
    return total


def _args():
    from ..global_vars import get_args
    return get_args()


def drain_pending_sends():
    """Retire outstanding isends (called at the end of every schedule and before buffers are reused)."""
    global _PENDING_SENDS
    for reqs, _keep in _PENDING_SENDS:
        for r in reqs:
            r.wait()
    _PENDING_SENDS = []


def _communicate_shapes(tensor_send_next, tensor_send_prev, recv_prev, recv_next):
    """Exchange the 3 dims of the tensors about to be sent/received (variable sequence lengths)."""
    dev = current_device()
    recv_prev_shape_tensor = torch.empty((3,), device=dev, dtype=torch.int64) if recv_prev else None
    recv_next_shape_tensor = torch.empty((3,), device=dev, dtype=torch.int64) if recv_next else None
    send_prev_shape_tensor = torch.tensor(tensor_send_prev.size(), device=dev, dtype=torch.int64) \
        if tensor_send_prev is not None else None
    send_next_shape_tensor = torch.tensor(tensor_send_next.size(), device=dev, dtype=torch.int64) \
        if tensor_send_next is not None else None
    ops = []
    if send_prev_shape_tensor is not None:
        ops.append(dist.P2POp(dist.isend, send_prev_shape_tensor, ps.get_pipeline_model_parallel_prev_rank()))
    if recv_prev_shape_tensor is not None:
        ops.append(dist.P2POp(dist.irecv, recv_prev_shape_tensor, ps.get_pipeline_model_parallel_prev_rank()))
    if send_next_shape_tensor is not None:
        ops.append(dist.P2POp(dist.isend, send_next_shape_tensor, ps.get_pipeline_model_parallel_next_rank()))
    if recv_next_shape_tensor is not None:
        ops.append(dist.P2POp(dist.irecv, recv_next_shape_tensor, ps.get_pipeline_model_parallel_next_rank()))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    recv_prev_shape = recv_prev_shape_tensor.tolist() if recv_prev_shape_tensor is not None else [0, 0, 0]
    recv_next_shape = recv_next_shape_tensor.tolist() if recv_next_shape_tensor is not None else [0, 0, 0]
    return recv_prev_shape, recv_next_shape


def _communicate(tensor_send_next: Optional[torch.Tensor], tensor_send_prev: Optional[torch.Tensor],
                 recv_prev: bool, recv_next: bool, tensor_shape: Shape, dtype_: Optional[torch.dtype] = None,
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    args = _args()
    tensor_recv_prev = tensor_recv_next = None
    override_scatter_gather = False
    if not args.variable_seq_lengths:
        if tensor_shape is None:
            # inference / downstream-task callers give no shape (reference p2p_communication.py:132-137)
            tensor_shape = (args.seq_length, args.micro_batch_size, args.hidden_size)
        recv_prev_shape = recv_next_shape = tensor_shape
    else:
        recv_prev_shape, recv_next_shape = _communicate_shapes(tensor_send_next, tensor_send_prev, recv_prev,
                                                               recv_next)
    tp = ps.get_tensor_model_parallel_world_size()
    scatter_gather = args.scatter_gather_tensors_in_pipeline and not args.sequence_parallel and tp > 1

    def chunk_shape(shape):
        if scatter_gather and not override_scatter_gather:
            n = reduce(operator.mul, shape, 1)
            if n % tp == 0:
                return [n // tp]
        return list(shape)

    if scatter_gather:
        n0 = reduce(operator.mul, tensor_shape, 1) if tensor_shape is not None else 0
        if n0 % tp != 0:
            override_scatter_gather = True
    dtype = args.params_dtype
    if args.fp32_residual_connection:
        dtype = torch.float
    requires_grad = True
    if dtype_ is not None:
        dtype = dtype_
        requires_grad = False
    dev = current_device()
    if recv_prev:
        tensor_recv_prev = torch.empty(chunk_shape(recv_prev_shape), requires_grad=requires_grad, device=dev,
                                       dtype=dtype)
    if recv_next:
        tensor_recv_next = torch.empty(chunk_shape(recv_next_shape), requires_grad=requires_grad, device=dev,
                                       dtype=dtype)
    if scatter_gather and not override_scatter_gather:
        if tensor_send_next is not None:
            tensor_send_next = split_tensor_into_1d_equal_chunks(tensor_send_next)
        if tensor_send_prev is not None:
            tensor_send_prev = split_tensor_into_1d_equal_chunks(tensor_send_prev)

    def send_alias(t):
        """A separate tensor object over the same storage: the schedule shrinks ``out.data`` of a sent activation
        right after the send is *posted* (deallocate_output_tensor); the in-flight isend must keep the real storage."""
        c = t.contiguous()
        return c.detach() if c is t else c

    ops, kinds = [], []
    if tensor_send_prev is not None:
        ops.append(dist.P2POp(dist.isend, send_alias(tensor_send_prev), ps.get_pipeline_model_parallel_prev_rank()))
        kinds.append("s")
    if tensor_recv_prev is not None:
        ops.append(dist.P2POp(dist.irecv, tensor_recv_prev, ps.get_pipeline_model_parallel_prev_rank()))
        kinds.append("r")
    if tensor_send_next is not None:
        ops.append(dist.P2POp(dist.isend, send_alias(tensor_send_next), ps.get_pipeline_model_parallel_next_rank()))
        kinds.append("s")
    if tensor_recv_next is not None:
        ops.append(dist.P2POp(dist.irecv, tensor_recv_next, ps.get_pipeline_model_parallel_next_rank()))
        kinds.append("r")
    if ops:
        ev0 = None
        if _ACCOUNT["on"] and "r" in kinds and dev.type == "cuda":
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        reqs = dist.batch_isend_irecv(ops)
        if len(reqs) == len(ops) and "r" in kinds and "s" in kinds:
            # per-op requests: wait receives now, retire sends lazily
            sends = [r for r, k in zip(reqs, kinds) if k == "s"]
            for r, k in zip(reqs, kinds):
                if k == "r":
                    r.wait()
            _PENDING_SENDS.append((sends, [o.tensor for o, k in zip(ops, kinds) if k == "s"]))
        elif "r" in kinds:
            for r in reqs:
                r.wait()
        else:
            _PENDING_SENDS.append((reqs, [o.tensor for o in ops]))
        if ev0 is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()                       # (after the receives' stream-waits)
            _ACCOUNT["pairs"].append((ev0, ev1))
        if len(_PENDING_SENDS) > 8:
            drain_pending_sends()

    if scatter_gather and not override_scatter_gather:
        if recv_prev:
            tensor_recv_prev = gather_split_1d_tensor(tensor_recv_prev).view(list(recv_prev_shape)).requires_grad_()
            tensor_recv_prev = make_viewless_tensor(tensor_recv_prev, requires_grad=True, keep_graph=False)
        if recv_next:
            tensor_recv_next = gather_split_1d_tensor(tensor_recv_next).view(list(recv_next_shape)).requires_grad_()
            tensor_recv_next = make_viewless_tensor(tensor_recv_next, requires_grad=True, keep_graph=False)
    return tensor_recv_prev, tensor_recv_next


def _timed(name, timers, fn):
    if timers is not None:
        timers(name, log_level=2).start()
    out = fn()
    if timers is not None:
        timers(name).stop()
    return out


def recv_forward(tensor_shape=None, dtype_=None, timers=None):
    """Receive activations from the previous pipeline stage."""
    if ps.is_pipeline_first_stage():
        return None
    return _timed("forward-recv", timers, lambda: _communicate(None, None, True, False, tensor_shape, dtype_)[0])


def recv_backward(tensor_shape=None, timers=None):
    """Receive output gradients from the next pipeline stage."""
    if ps.is_pipeline_last_stage():
        return None
    return _timed("backward-recv", timers, lambda: _communicate(None, None, False, True, tensor_shape)[1])


def send_forward(output_tensor, tensor_shape=None, dtype_=None, timers=None):
    if not ps.is_pipeline_last_stage():
        _timed("forward-send", timers, lambda: _communicate(output_tensor, None, False, False, tensor_shape, dtype_))


def send_backward(input_tensor_grad, tensor_shape=None, timers=None):
    if not ps.is_pipeline_first_stage():
        _timed("backward-send", timers, lambda: _communicate(None, input_tensor_grad, False, False, tensor_shape))


def send_forward_recv_backward(output_tensor, tensor_shape=None, timers=None):
    if ps.is_pipeline_last_stage():
        return None
    return _timed("forward-send-backward-recv", timers,
                  lambda: _communicate(output_tensor, None, False, True, tensor_shape)[1])


def send_backward_recv_forward(input_tensor_grad, tensor_shape=None, timers=None):
    if ps.is_pipeline_first_stage():
        return None
    return _timed("backward-send-forward-recv", timers,
                  lambda: _communicate(None, input_tensor_grad, True, False, tensor_shape)[0])


def send_forward_recv_forward(output_tensor, recv_prev, tensor_shape=None, timers=None):
    return _timed("forward-send-forward-recv", timers,
                  lambda: _communicate(output_tensor, None, recv_prev, False, tensor_shape)[0])


def send_backward_recv_backward(input_tensor_grad, recv_next, tensor_shape=None, timers=None):
    return _timed("backward-send-backward-recv", timers,
                  lambda: _communicate(None, input_tensor_grad, False, recv_next, tensor_shape)[1])


def send_forward_backward_recv_forward_backward(output_tensor, input_tensor_grad, recv_prev, recv_next,
                                                tensor_shape=None, timers=None):
    return _timed("forward-backward-send-forward-backward-recv", timers,
                  lambda: _communicate(output_tensor, input_tensor_grad, recv_prev, recv_next, tensor_shape))

"""Gradient clipping / zero counting over lists of tensors (parity: megatron/optimizer/clip_grads.py:16-136).

The training path uses the flat-buffer kernels in optimizer.py; these list-based versions keep the reference's
public functions (tools, tests, torch-DDP mode) and run the same sm_100a reduction kernel per tensor on CUDA.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from .. import ops
from ..models.module import param_is_not_shared
from ..parallel.layers import param_is_not_tensor_parallel_duplicate
from ..utils.device import current_device


def _sq_norm_list(grads):
    dev = grads[0].device if grads else current_device()
    total = torch.zeros(1, dtype=torch.float32, device=dev)
    if grads and ops.cuda_ops_available(grads[0]):
        ws = torch.zeros(148 * 8, dtype=torch.float32, device=dev)
        for g in grads:
            ops._C().sqnorm_flat(g.contiguous().view(-1), 0, None, None, ws, total, True)
            ops._count(2)
    else:
        for g in grads:
            total += g.float().pow(2).sum()
    return total


def clip_grad_norm_fp32(parameters, grads_for_norm, max_norm, norm_type=2, model_parallel_group=None):
    """Clip in place the ``main_grad``/``grad`` of ``parameters`` to a global norm of ``max_norm``; the norm is
    computed over ``grads_for_norm`` (unique grads only) and summed over ``model_parallel_group``."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    if isinstance(grads_for_norm, torch.Tensor):
        grads_for_norm = [grads_for_norm]
    grads = []
    for p in parameters:
        g = getattr(p, "main_grad", None)
        if g is None:
            g = p.grad
        if g is not None:
            grads.append(g.detach())
    max_norm, norm_type = float(max_norm), float(norm_type)
    if norm_type == math.inf:
        total = torch.tensor([max((g.abs().max().item() for g in grads_for_norm), default=0.0)],
                             dtype=torch.float32, device=current_device())
        if model_parallel_group is not None and dist.is_initialized():
            dist.all_reduce(total, op=dist.ReduceOp.MAX, group=model_parallel_group)
        total_norm = total[0].item()
    else:
        if norm_type == 2.0:
            total = _sq_norm_list(grads_for_norm)
        else:
            total = torch.zeros(1, dtype=torch.float32, device=current_device())
            for g in grads_for_norm:
                total += torch.norm(g.float(), norm_type) ** norm_type
        if model_parallel_group is not None and dist.is_initialized():
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=model_parallel_group)
        total_norm = total.item() ** (1.0 / norm_type)
    clip_coeff = max_norm / (total_norm + 1.0e-6)
    if clip_coeff < 1.0:
        for g in grads:
            g.mul_(clip_coeff)
    return total_norm


def count_zeros_fp32(parameters, model_parallel_group=None):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    total = torch.zeros(1, dtype=torch.float32, device=current_device())
    for p in parameters:
        g = getattr(p, "main_grad", None)
        if g is None:
            g = p.grad
        if g is not None and param_is_not_shared(p) and param_is_not_tensor_parallel_duplicate(p):
            total += g.numel() - torch.count_nonzero(g.detach())
    if model_parallel_group is not None and dist.is_initialized():
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=model_parallel_group)
    return total.item()

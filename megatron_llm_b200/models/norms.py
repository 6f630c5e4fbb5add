"""LayerNorm / RMSNorm modules backed by the sm_100a kernels in csrc/norm.cu.

Parity: megatron/model/fused_layer_norm.py (MixedFusedLayerNorm :56-122 -> apex kernels; RMSNorm :125-139 ->
five unfused torch ops).  Both take ``sequence_parallel`` and flag their weights so the optimizer all-reduces
their grads over the TP group (optimizer.py:257-277).  ``forward(x, residual=r)`` fuses the preceding residual
add and returns ``(norm(x+r), x+r)``.
"""
from __future__ import annotations

import numbers

import torch
from torch.nn import init
from torch.nn.parameter import Parameter

from .. import ops
from ..utils.device import current_device


class MixedFusedLayerNorm(torch.nn.Module):
    def __init__(self, normalized_shape, eps=1e-5, no_persist_layer_norm=True, sequence_parallel=False):
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        self.eps = eps
        self.weight = Parameter(torch.empty(*normalized_shape, device=current_device()))
        self.bias = Parameter(torch.empty(*normalized_shape, device=current_device()))
        self.reset_parameters()
        self.no_persist_layer_norm = no_persist_layer_norm
        self.sequence_parallel = sequence_parallel
        setattr(self.weight, "sequence_parallel", sequence_parallel)
        setattr(self.bias, "sequence_parallel", sequence_parallel)

    def reset_parameters(self):
        init.ones_(self.weight)
        init.zeros_(self.bias)

    def forward(self, input, residual=None):
        return ops.layernorm(input, self.weight, self.bias, self.eps, residual=residual)


LayerNorm = MixedFusedLayerNorm


class FusedLayerNormAffineFunction:
    """apex-style functional entry (reference fused_layer_norm.py:33-53): ``apply(input, weight, bias, shape, eps)``;
    differentiable through the same autograd function the modules use."""

    @staticmethod
    def apply(input, weight, bias, normalized_shape, eps):
        assert tuple(input.shape[-len(tuple(normalized_shape)):]) == tuple(normalized_shape)
        return ops.layernorm(input, weight, bias, eps)


class RMSNorm(torch.nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6, sequence_parallel: bool = False):
        super().__init__()
        self.eps = eps
        self.weight = Parameter(torch.ones(dim, device=current_device()))
        self.sequence_parallel = sequence_parallel
        setattr(self.weight, "sequence_parallel", sequence_parallel)

    def forward(self, x, residual=None):
        return ops.rmsnorm(x, self.weight, self.eps, residual=residual)

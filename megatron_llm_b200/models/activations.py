"""Activations (parity: megatron/model/glu_activations.py, fused_bias_gelu.py, utils.py init helpers)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops


class _GLU(nn.Module):
    """x -> x1 * act(x2) with (x1, x2) = chunk(x, 2, -1): first half = up projection, second = gate."""
    kind = "liglu"

    def forward(self, x):
        return ops.glu(x, self.kind)


class LiGLU(_GLU):
    kind = "liglu"


class GEGLU(_GLU):
    kind = "geglu"


class ReGLU(_GLU):
    kind = "reglu"


class SwiGLU(_GLU):
    kind = "swiglu"


liglu, geglu, reglu, swiglu = LiGLU(), GEGLU(), ReGLU(), SwiGLU()

GLU_ACTIVATIONS = {"geglu": geglu, "liglu": liglu, "reglu": reglu, "swiglu": swiglu}


def bias_gelu_impl(x, bias):
    """tanh-approximate gelu(x + bias) (reference fused_bias_gelu.py:14-43)."""
    return ops.gelu(x, bias, approximate=True)


def init_method_normal(sigma):
    def init_(tensor):
        return nn.init.normal_(tensor, mean=0.0, std=sigma)
    return init_


def scaled_init_method_normal(sigma, num_layers):
    std = sigma / math.sqrt(2.0 * num_layers)

    def init_(tensor):
        return nn.init.normal_(tensor, mean=0.0, std=std)
    return init_


def attention_mask_func(attention_scores, attention_mask):
    attention_scores.masked_fill_(attention_mask, -10000.0)
    return attention_scores


def get_linear_layer(rows, columns, init_method, perform_initialization=True):
    layer = nn.Linear(rows, columns)
    if perform_initialization:
        init_method(layer.weight)
    with torch.no_grad():
        layer.bias.zero_()
    return layer


def gelu_impl(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def openai_gelu(x):
    return gelu_impl(x)


def erf_gelu(x):
    return x * 0.5 * (torch.erf(x / 1.41421).to(dtype=x.dtype) + torch.ones_like(x).to(dtype=x.dtype))

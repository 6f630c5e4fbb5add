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


def bias_gelu(bias, y):
    """Reference argument order (fused_bias_gelu.py:14): tanh-approximate gelu(y + bias)."""
    return ops.gelu(y, bias, approximate=True)


def bias_gelu_back(g, bias, y):
    """d/dy of ``bias_gelu`` times ``g`` (fused_bias_gelu.py:22-30); the training path gets it from the backward of
    ``ops.gelu`` -- this closed form is for callers that apply it by hand."""
    x = (bias + y).float()
    t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
    ff = 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)
    return (ff * g.float()).to(g.dtype)


class GeLUFunction(torch.autograd.Function):
    """autograd wrapper returning (dinput, dbias) like the reference's (fused_bias_gelu.py:32-43)."""

    @staticmethod
    def forward(ctx, input, bias):
        ctx.save_for_backward(input, bias)
        return bias_gelu(bias, input)

    @staticmethod
    def backward(ctx, grad_output):
        input, bias = ctx.saved_tensors
        tmp = bias_gelu_back(grad_output, bias, input)
        dbias = tmp.reshape(-1, tmp.size(-1)).sum(0).view_as(bias) if bias.dim() == 1 else tmp
        return tmp, dbias


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

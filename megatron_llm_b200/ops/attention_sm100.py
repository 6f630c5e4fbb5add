"""Python side of the hand-written sm_100a attention kernels (csrc/attention_sm100.cu)."""
from __future__ import annotations

import os

import torch

from . import _ext


def _head_dim_ok(hn: int) -> bool:
    """128 (Llama / Mistral) and 64 (Falcon-40B: 128 heads x 64, GPT-2 style models); both instantiations are
    validated against the fp32 reference in tests/test_ops_gpu.py (``MLB200_ATTN_HD64=0`` routes 64 to the library)."""
    return hn == 128 or (hn == 64 and os.environ.get("MLB200_ATTN_HD64", "1") == "1")


def supported(q, k, v, causal, window, dropout_p) -> bool:
    if os.environ.get("MLB200_ATTN", "1") == "0" or os.environ.get("MLB200_DISABLE_KERNELS", "0") == "1":
        return False
    _ext.load()      # a CUDA tensor without the built extension is an error, never a silent library fallback
    hn = q.size(-1)
    return (q.dtype == torch.bfloat16 and _head_dim_ok(hn) and dropout_p == 0.0 and causal
            and q.size(1) == k.size(1) and q.size(1) % 128 == 0 and q.size(2) % k.size(2) == 0)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, window, scale):
        mod = _ext.load()
        out, lse = mod.attn_fwd(q, k, v, causal, -1 if window is None else int(window), float(scale))
        _ext.count()
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.causal, ctx.window, ctx.scale = causal, window, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        mod = _ext.load()
        dq, dk, dv = mod.attn_bwd(dout if dout.stride(-1) == 1 else dout.contiguous(), q, k, v, out, lse, ctx.causal,
                                  -1 if ctx.window is None else int(ctx.window), float(ctx.scale))
        _ext.count(3)
        return dq, dk, dv, None, None, None


def attention(q, k, v, causal, window, scale):
    import math
    scale = scale if scale is not None else 1.0 / math.sqrt(q.size(-1))
    return _AttnFn.apply(q, k, v, causal, window, scale)


# ------------------------------------------------------------------------------------------------ packed QKV path
def packed_supported(mixed, nkv, g, hn, dropout_p) -> bool:
    """``mixed`` = QKV projection output [s, b, nkv * (g + 2) * hn] (per KV group: g query heads, then k, then v)."""
    if os.environ.get("MLB200_ATTN", "1") == "0" or os.environ.get("MLB200_ATTN_PACKED", "1") == "0" \
            or os.environ.get("MLB200_DISABLE_KERNELS", "0") == "1":
        return False
    if not (mixed.is_cuda and mixed.dtype == torch.bfloat16 and _head_dim_ok(hn) and dropout_p == 0.0
            and mixed.dim() == 3 and mixed.stride(2) == 1 and mixed.size(0) % 128 == 0):
        return False
    _ext.load()      # (fails loudly on a GPU box without the extension)
    return True


class _PackedAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mixed, nkv, g, window, scale, hn):
        mod = _ext.load()
        out, lse = mod.attn_fwd_packed(mixed, nkv, g, -1 if window is None else int(window), float(scale), hn)
        _ext.count()
        ctx.save_for_backward(mixed, out, lse)
        ctx.cfg = (nkv, g, window, scale, hn)
        return out

    @staticmethod
    def backward(ctx, dout):
        mixed, out, lse = ctx.saved_tensors
        nkv, g, window, scale, hn = ctx.cfg
        mod = _ext.load()
        d = dout if dout.stride(-1) == 1 else dout.contiguous()
        dmixed = mod.attn_bwd_packed(d, mixed, out, lse, nkv, g, -1 if window is None else int(window), float(scale),
                                     hn)
        _ext.count(3)
        return dmixed, None, None, None, None, None


def packed_attention(mixed, nkv, g, window=None, scale=None, hn=None):
    """Causal attention straight from the packed (already rotated) QKV buffer -> context [s, b, nkv * g * hn]."""
    import math
    hn = hn if hn is not None else mixed.size(-1) // (nkv * (g + 2))
    scale = scale if scale is not None else 1.0 / math.sqrt(hn)
    return _PackedAttnFn.apply(mixed, nkv, g, window, scale, hn)

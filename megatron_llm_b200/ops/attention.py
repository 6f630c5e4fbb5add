"""Causal self-attention with native GQA/MQA and optional sliding window.

Parity target: ``flash_attn.flash_attn_func`` as called from megatron/model/transformer.py:538-553 (causal,
window_size=(w,w), dropout_p), except that K/V are NOT broadcast to the query head count first
(transformer.py:458-465): the kernels consume ``n_kv`` heads directly.  ``sk > sq`` (KV-cache decode) aligns the causal
mask bottom-right like flash-attention: query i sits at position sk - sq + i.

q: [b, sq, nq, hn]   k, v: [b, sk, nkv, hn]   (any strides with a contiguous last dim)  ->  [b, sq, nq, hn]
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import _ext


_M32 = 0xFFFFFFFF


def _drop_mix(x):
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def dropout_threshold(p: float):
    """(threshold, 1 / keep) as the kernels derive them: the rate is quantised to threshold / 256."""
    t = min(255, max(0, int(p * 256.0 + 0.5)))
    return t, 256.0 / (256.0 - t)


def dropout_keep_mask(seed: int, p: float, batch: int, heads: int, sq: int, sk: int, device=None):
    """Replica of the attention kernels' dropout mask (csrc/attention_dropout.cuh): bool [batch, heads, sq, sk], True =
    kept, for the 64-bit ``seed`` the kernels were launched with.  Queries / keys are absolute positions 0 .. s-1."""
    t, _ = dropout_threshold(p)
    i64 = dict(dtype=torch.int64, device=device)
    lo, hi = seed & _M32, (seed >> 32) & _M32
    bh = torch.arange(batch * heads, **i64)
    head_key = _drop_mix(hi ^ ((bh * 0x85EBCA77) & _M32))                                     # [bh]
    rows = torch.arange(sq, **i64)
    row_key = (_drop_mix(lo ^ ((rows * 0x9E3779B1) & _M32))[None, :] + head_key[:, None]) & _M32   # [bh, sq]
    keys = torch.arange(sk, **i64)
    quad = ((keys >> 2) * 0xC2B2AE3D) & _M32
    bytes4 = _drop_mix(row_key[:, :, None] ^ quad[None, None, :])                             # [bh, sq, sk]
    byte = (bytes4 >> ((keys & 3) * 8)[None, None, :]) & 0xFF
    return (byte >= t).view(batch, heads, sq, sk)


def attention_reference(q, k, v, causal=True, window: Optional[int] = None, scale: Optional[float] = None,
                        dropout_p: float = 0.0, keep_mask=None):
    """fp32 math oracle (also the CPU path).  ``keep_mask`` (bool [b, nq, sq, sk], with ``dropout_p``) applies a given
    dropout mask with the kernels' quantised rate instead of drawing one."""
    b, sq, nq, hn = q.shape
    sk, nkv = k.size(1), k.size(2)
    g = nq // nkv
    scale = scale if scale is not None else 1.0 / math.sqrt(hn)
    qf = q.float().permute(0, 2, 1, 3).reshape(b, nkv, g, sq, hn)
    kf = k.float().permute(0, 2, 1, 3).unsqueeze(2)  # [b,nkv,1,sk,hn]
    vf = v.float().permute(0, 2, 1, 3).unsqueeze(2)
    scores = torch.matmul(qf, kf.transpose(-1, -2)) * scale  # [b,nkv,g,sq,sk]
    if causal or window is not None:
        qi = torch.arange(sq, device=q.device).view(sq, 1) + (sk - sq)
        ki = torch.arange(sk, device=q.device).view(1, sk)
        allowed = torch.ones(sq, sk, dtype=torch.bool, device=q.device)
        if causal:
            allowed &= ki <= qi
        if window is not None:
            allowed &= ki >= qi - window
        scores = scores.masked_fill(~allowed, float("-inf"))
    p = torch.softmax(scores, dim=-1)
    if keep_mask is not None and dropout_p > 0:
        p = p * keep_mask.view(b, nkv, g, sq, sk).to(p.dtype) * dropout_threshold(dropout_p)[1]
    elif dropout_p > 0:
        p = F.dropout(p, dropout_p)
    out = torch.matmul(p, vf)  # [b,nkv,g,sq,hn]
    return out.reshape(b, nq, sq, hn).permute(0, 2, 1, 3).to(q.dtype)


def _library_flash(q, k, v, causal, window, scale, dropout_p):
    from flash_attn import flash_attn_func
    ws = (-1, -1) if window is None else (window, window)
    return flash_attn_func(q, k, v, dropout_p=dropout_p, softmax_scale=scale, causal=causal, window_size=ws)


def flash_attention(q, k, v, causal: bool = True, window: Optional[int] = None, scale: Optional[float] = None,
                    dropout_p: float = 0.0):
    if q.is_cuda and q.dtype in (torch.bfloat16, torch.float16):
        from . import attention_sm100
        if attention_sm100.supported(q, k, v, causal, window, dropout_p):
            return attention_sm100.attention(q, k, v, causal, window, scale, dropout_p)
        if attention_sm100.decode_supported(q, k, v, causal, dropout_p):
            return attention_sm100.decode_attention(q, k, v, window, scale)
        # outside the hand-written kernels' envelope (non-causal, sq != sk with gradients or many query positions, a
        # kernel variant that failed its self-test): the FA-2 library
        return _library_flash(q, k, v, causal, window, scale, dropout_p)
    return attention_reference(q, k, v, causal, window, scale, dropout_p)

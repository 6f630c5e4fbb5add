"""Causal self-attention with native GQA/MQA and optional sliding window.

Parity target: ``flash_attn.flash_attn_func`` as called from megatron/model/transformer.py:538-553 (causal,
window_size=(w,w)), except that K/V are NOT broadcast to the query head count first
(transformer.py:458-465): the kernels consume ``n_kv`` heads directly.

q: [b, sq, nq, hn]   k, v: [b, sk, nkv, hn]   (any strides with a contiguous last dim)  ->  [b, sq, nq, hn]
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import _ext


def attention_reference(q, k, v, causal=True, window: Optional[int] = None, scale: Optional[float] = None,
                        dropout_p: float = 0.0):
    """fp32 math oracle (also the CPU path)."""
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
    if dropout_p > 0:
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
            return attention_sm100.attention(q, k, v, causal, window, scale)
        # TODO(perf): shapes outside the tcgen05 kernel's envelope fall back to the FA-2 library
        return _library_flash(q, k, v, causal, window, scale, dropout_p)
    return attention_reference(q, k, v, causal, window, scale, dropout_p)

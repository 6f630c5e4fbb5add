"""Rotary position embeddings (parity: megatron/model/positional_embeddings.py:7-51).

``precompute_freqs_cis`` / ``apply_rotary_emb`` keep the reference's complex API for tools and tests; the
training path uses the real-valued (cos, sin) table + in-place kernel ``ops.rope_qkv_`` instead (no complex
math, no casts, no per-call host->device copy of the table).
"""
from __future__ import annotations

import torch


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, scaling_factor: float = 1.0) -> torch.Tensor:
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    t = torch.arange(end, device=freqs.device) / scaling_factor
    return torch.polar(torch.ones(end, dim // 2), torch.outer(t, freqs).float())  # complex64


def reshape_for_broadcast(freqs_cis: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    ndim = x.ndim
    assert 0 <= 1 < ndim
    assert freqs_cis.shape == (x.shape[0], x.shape[-1])
    shape = [d if i == 0 or i == ndim - 1 else 1 for i, d in enumerate(x.shape)]
    return freqs_cis.view(*shape)


def apply_rotary_emb(xq: torch.Tensor, xk: torch.Tensor, freqs_cis: torch.Tensor, position_ids=None):
    """xq/xk: [s, b, n, hn] (interleaved-pair convention).  position_ids: [b, s] or None."""
    freqs_cis = freqs_cis.to(xq.device)
    if position_ids is None:
        freqs_q = reshape_for_broadcast(freqs_cis, torch.view_as_complex(xq.float().reshape(*xq.shape[:-1], -1, 2)))
        freqs_k = freqs_q
    else:
        f = freqs_cis[position_ids].transpose(0, 1)[:, :, None, :]  # [s, b, 1, hn/2]
        freqs_q = freqs_k = f
    xq_ = torch.view_as_complex(xq.float().reshape(*xq.shape[:-1], -1, 2))
    xk_ = torch.view_as_complex(xk.float().reshape(*xk.shape[:-1], -1, 2))
    xq_out = torch.view_as_real(xq_ * freqs_q).flatten(3)
    xk_out = torch.view_as_real(xk_ * freqs_k).flatten(3)
    return xq_out.type_as(xq), xk_out.type_as(xk)

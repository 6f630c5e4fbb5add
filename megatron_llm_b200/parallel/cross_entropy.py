"""Vocab-parallel cross entropy with ONE packed cross-rank reduction.

Parity target: megatron/core/tensor_parallel/cross_entropy.py (:14-127 loss with label
smoothing, :146-175 distributed argmax).  The reference issues three TP all-reduces
(MAX, SUM, SUM) and materialises the fp32 ``[s,b,V/t]`` softmax twice.  Here every rank
reduces its vocab shard to four per-token statistics in a single pass
``(local max, local sum-exp, target logit if owned, sum of logits)``, one all-gather of
``[tokens, 4]`` fp32 combines them, and backward recomputes the softmax from the saved
bf16 logits + ``(max, log-sum-exp)`` and overwrites the logits buffer with the gradient.
On a B200 both passes are hand-written sm_100a kernels (csrc/ce.cu); the torch code below
is the CPU path and the test oracle.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops
from . import state as ps
from .tp_utils import VocabUtility


def _local_stats_torch(logits2d: torch.Tensor, target: torch.Tensor, vocab_start: int):
    """logits2d [T, Vp] (any float dtype), target [T] global ids -> stats [T,4] fp32."""
    x = logits2d.float()
    vp = x.size(1)
    m = x.max(dim=1).values
    s = torch.exp(x - m[:, None]).sum(dim=1)
    local_t = target - vocab_start
    owned = (local_t >= 0) & (local_t < vp)
    idx = local_t.clamp(0, vp - 1)
    tl = torch.where(owned, x.gather(1, idx[:, None]).squeeze(1), torch.zeros_like(m))
    return torch.stack([m, s, tl, x.sum(dim=1)], dim=1)


def _combine(stats_all: torch.Tensor):
    """stats_all [W, T, 4] -> (global max M, log-sum-exp relative to M, target logit, sum logits)."""
    m_r, s_r, t_r, x_r = stats_all.unbind(-1)
    M = m_r.max(dim=0).values
    S = (s_r * torch.exp(m_r - M[None])).sum(dim=0)
    return M, torch.log(S), t_r.sum(dim=0), x_r.sum(dim=0)


class _VocabParallelCrossEntropy(torch.autograd.Function):

    @staticmethod
    def forward(ctx, vocab_parallel_logits, target, label_smoothing=0.0):
        world = ps.get_tensor_model_parallel_world_size()
        rank = ps.get_tensor_model_parallel_rank()
        vp = vocab_parallel_logits.size(-1)
        vocab_start, _ = VocabUtility.vocab_range_from_per_partition_vocab_size(vp, rank, world)
        lead = vocab_parallel_logits.shape[:-1]
        logits2d = vocab_parallel_logits.reshape(-1, vp)
        tgt = target.reshape(-1)

        if ops.cuda_ops_available(logits2d):
            stats = ops.ce_local_stats(logits2d, tgt, vocab_start)
        else:
            stats = _local_stats_torch(logits2d, tgt, vocab_start)
        if world > 1:
            gathered = torch.empty((world * stats.size(0), stats.size(1)), dtype=stats.dtype, device=stats.device)
            dist.all_gather_into_tensor(gathered, stats.contiguous(),
                                        group=ps.get_tensor_model_parallel_group())
            gathered = gathered.view(world, stats.size(0), stats.size(1))
        else:
            gathered = stats[None]
        M, logS, tlogit, xsum = _combine(gathered)
        loss = logS + M - tlogit
        vocab_size = vp * world
        smoothing = 0.0
        if label_smoothing > 0:
            assert 1.0 > label_smoothing > 0.0
            smoothing = label_smoothing * vocab_size / (vocab_size - 1)
            mean_log_probs = xsum / vocab_size - (logS + M)
            loss = (1.0 - smoothing) * loss - smoothing * mean_log_probs
        ctx.smoothing, ctx.vocab_size, ctx.vocab_start = smoothing, vocab_size, vocab_start
        ctx.save_for_backward(logits2d, tgt, M, logS)
        ctx.lead = lead
        return loss.view(lead)

    @staticmethod
    def backward(ctx, grad_output):
        logits2d, tgt, M, logS = ctx.saved_tensors
        g = grad_output.reshape(-1).float().contiguous()
        if ops.cuda_ops_available(logits2d):
            grad = ops.ce_backward_(logits2d, tgt, M, logS, g, ctx.vocab_start, ctx.smoothing,
                                    ctx.vocab_size)
        else:
            x = logits2d.float()
            vp = x.size(1)
            p = torch.exp(x - (M + logS)[:, None])
            local_t = tgt - ctx.vocab_start
            owned = (local_t >= 0) & (local_t < vp)
            onehot = torch.zeros_like(p)
            onehot.scatter_(1, local_t.clamp(0, vp - 1)[:, None], owned.float()[:, None])
            if ctx.smoothing > 0:
                grad = p - (1.0 - ctx.smoothing) * onehot - ctx.smoothing / ctx.vocab_size
            else:
                grad = p - onehot
            grad = (grad * g[:, None]).to(logits2d.dtype)
        return grad.view(*ctx.lead, -1), None, None


def vocab_parallel_cross_entropy(vocab_parallel_logits, target, label_smoothing=0.0):
    """Per-token loss when logits are split along the vocab dim across the TP group.

    vocab_parallel_logits: [..., vocab/tp]; target: [...] global token ids."""
    return _VocabParallelCrossEntropy.apply(vocab_parallel_logits, target, label_smoothing)


def vocab_parallel_max_indices(logits: torch.Tensor) -> torch.Tensor:
    """Distributed argmax over the vocab dim: returns global ids, shape logits.shape[:-1]."""
    world = ps.get_tensor_model_parallel_world_size()
    rank = ps.get_tensor_model_parallel_rank()
    vp = logits.size(-1)
    vals, idx = logits.float().max(dim=-1)
    idx = idx + rank * vp
    if world == 1:
        return idx
    group = ps.get_tensor_model_parallel_group()
    packed = torch.stack([vals, idx.to(vals.dtype)], dim=0).contiguous()
    allp = torch.empty((world * packed.size(0),) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(allp, packed, group=group)
    allp = allp.view((world,) + tuple(packed.shape))
    best = allp[:, 0].argmax(dim=0, keepdim=True)
    return allp[:, 1].gather(0, best).squeeze(0).long()

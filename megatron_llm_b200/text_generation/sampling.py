"""Token sampling: greedy / top-k / top-p with temperature (parity: text_generation/sampling.py:12-93)."""
from __future__ import annotations

import torch


def modify_logits_for_top_k_filtering(logits, top_k):
    """Everything below the k-th largest logit of each row -> -inf (in place)."""
    kth = torch.topk(logits, top_k)[0][..., -1, None]
    logits.masked_fill_(logits < kth, float("-Inf"))


def modify_logits_for_top_p_filtering(logits, top_p):
    """Nucleus filtering (in place): keep the smallest prefix of the sorted distribution whose mass exceeds top_p
    (the token that crosses the threshold is kept, and at least one token always survives)."""
    sorted_logits, sorted_indices = torch.sort(logits, descending=True)
    cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    drop = cumulative > top_p
    drop[:, 1:] = drop[:, :-1].clone()
    drop[..., 0] = False
    drop = drop.scatter(1, sorted_indices, drop)
    logits.masked_fill_(drop, float("-Inf"))


def sample(logits, top_k=0, top_p=0.0, temperature=1.0, vocab_size=None):
    """logits [b, v] fp32 -> sampled ids [b]; ``vocab_size`` clamps away padded-vocab ids."""
    assert logits.ndim == 2, "expected the logits to be of [b, v] shape."
    assert logits.dtype == torch.float32, "input logits should be floats."
    if top_k == 1:
        assert top_p == 0.0, "cannot set both greedy and top-p samplings."
        samples = torch.argmax(logits, dim=-1)
    else:
        logits = logits.clone()
        if temperature != 1.0:
            logits.div_(temperature)
        if top_k > 1:
            assert top_p == 0.0, "cannot set both top-k and top-p samplings."
            assert top_k <= logits.size(1), "top-k is larger than logit size."
            if vocab_size:
                assert top_k < vocab_size, "top-k is larger than vocab size."
            modify_logits_for_top_k_filtering(logits, top_k)
        elif top_p > 0.0:
            assert top_p <= 1.0, "top-p should be in (0, 1]."
            modify_logits_for_top_p_filtering(logits, top_p)
        samples = torch.multinomial(logits.softmax(dim=-1), num_samples=1).view(-1)
    if vocab_size:
        samples = torch.clamp(samples, min=0, max=(vocab_size - 1))
    return samples

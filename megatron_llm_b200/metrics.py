"""Validation-metric registry selected with ``--metrics`` (parity: megatron/metrics.py:100-110)."""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch


class MetricInput:
    """Bundle handed to every metric: the batch tuple, the vocab-parallel logits and the scalar loss."""

    def __init__(self, batch: tuple, output: torch.Tensor, loss: torch.Tensor):
        self.tokens, self.labels, self.loss_mask, self.attention_mask, self.position_ids = batch
        self.output, self.loss = output, loss
        self._max_indices: Optional[torch.Tensor] = None
        self._instruct_mask: Optional[torch.Tensor] = None
        self._instruct_mask_done = False

    @property
    def max_indices(self) -> torch.Tensor:
        if self._max_indices is None:
            from .parallel.cross_entropy import vocab_parallel_max_indices
            self._max_indices = vocab_parallel_max_indices(self.output)
        return self._max_indices

    @property
    def instruct_mask(self) -> Optional[torch.Tensor]:
        """loss_mask minus the ChatML scaffolding: ``<|im_start|>role\\n`` and ``<|im_end|>\\n`` + next token."""
        if self._instruct_mask_done:
            return self._instruct_mask
        self._instruct_mask_done = True
        from .global_vars import get_tokenizer
        tok = get_tokenizer()
        (start_id,) = tok.tokenize("<|im_start|>")
        (end_id,) = tok.tokenize("<|im_end|>")
        keep = torch.ones_like(self.loss_mask)
        width = keep.size(1)
        for marker in (start_id, end_id):
            rows, cols = torch.nonzero(self.labels == marker, as_tuple=True)
            if torch.any(cols + 2 >= width):
                print("Error calculating instruct mask")
                return None
            for d in range(3):
                keep[rows, cols + d] = 0.0
        self._instruct_mask = self.loss_mask * keep
        return self._instruct_mask


def _dp_average(vals):
    from .utils import average_losses_across_data_parallel_group
    return average_losses_across_data_parallel_group(vals)


def _masked_accuracy(inputs: MetricInput, mask: torch.Tensor) -> torch.Tensor:
    hit = (inputs.labels == inputs.max_indices) & (mask != 0)
    return torch.count_nonzero(hit) / torch.count_nonzero(mask)


def perplexity(inputs: MetricInput) -> Dict[str, float]:
    return {"ppl": math.exp(min(20, inputs.loss.item()))}


def accuracy(inputs: MetricInput) -> Dict[str, float]:
    return {"lm accuracy": _dp_average([_masked_accuracy(inputs, inputs.loss_mask)])[0]}


def instruct_accuracy(inputs: MetricInput) -> Dict[str, float]:
    if inputs.instruct_mask is None:
        acc = torch.tensor(float("nan"), device=inputs.labels.device)
    else:
        acc = _masked_accuracy(inputs, inputs.instruct_mask)
    return {"instruct accuracy": _dp_average([acc])[0]}


def count_loss_mask(inputs: MetricInput) -> Dict[str, float]:
    return {"count loss mask": torch.count_nonzero(inputs.loss_mask) / inputs.loss_mask.size(0)}


def count_instruct_mask(inputs: MetricInput) -> Dict[str, float]:
    if inputs.instruct_mask is None:
        return {}
    return {"count instruct mask": torch.count_nonzero(inputs.instruct_mask) / inputs.instruct_mask.size(0)}


METRICS: Dict[str, Callable[[MetricInput], Dict[str, float]]] = {
    "perplexity": perplexity,
    "accuracy": accuracy,
    "instruct_accuracy": instruct_accuracy,
    "count_loss_mask": count_loss_mask,
    "count_instruct_mask": count_instruct_mask,
}


def get_metric(name: str) -> Callable[[MetricInput], Dict[str, float]]:
    return METRICS[name]

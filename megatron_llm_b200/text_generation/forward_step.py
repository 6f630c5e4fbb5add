"""Inference forward pass with KV cache and pipelining (parity: text_generation/forward_step.py:18-204).

``InferenceParams`` holds the per-layer key/value cache.  Unlike the reference (which caches K/V already broadcast to
the query-head count, transformer.py:412-419), the cache stores ``n_kv/tp`` heads and keys enter it already rotated at
their absolute positions, so RoPE/GQA models (Llama, Falcon, Mistral) decode incrementally."""
from __future__ import annotations

from collections.abc import Iterable

import torch

from ..parallel import state as ps
from ..utils.device import current_device
from .communication import recv_from_prev_pipeline_rank_, send_to_next_pipeline_rank


def _args():
    from ..global_vars import get_args
    return get_args()


class InferenceParams:
    def __init__(self, max_batch_size, max_sequence_len):
        self.max_sequence_len = max_sequence_len
        self.max_batch_size = max_batch_size
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict = {}

    def swap_key_value_dict(self, batch_idx):
        """Reorder the cached batch entries (beam search keeps the surviving beams)."""
        if len(self.key_value_memory_dict) == 0:
            raise ValueError("should not swap when dict in empty")
        for layer_number, (k, v) in self.key_value_memory_dict.items():
            assert len(batch_idx) == k.shape[1]
            self.key_value_memory_dict[layer_number] = (k[:, batch_idx].contiguous(), v[:, batch_idx].contiguous())


class ForwardStep:
    """Forward step that pipelines micro-batches when ``batch * seq`` exceeds the configured threshold."""

    def __init__(self, model, max_batch_size, max_sequence_len):
        assert not isinstance(model, Iterable), "interleaving schedule is not supported for inference"
        model.eval()
        self.model = model
        self.inference_params = InferenceParams(max_batch_size, max_sequence_len)
        args = _args()
        self.pipeline_size_larger_than_one = args.pipeline_model_parallel_size > 1
        self.pipelining_batch_x_seqlen = args.inference_batch_times_seqlen_threshold

    def __call__(self, tokens, position_ids, attention_mask):
        if self.pipeline_size_larger_than_one:
            current_batch_x_seqlen = tokens.size(0) * tokens.size(1)
            if current_batch_x_seqlen >= self.pipelining_batch_x_seqlen:
                micro_batch_size = max(1, self.pipelining_batch_x_seqlen // tokens.size(1))
                return _with_pipelining_forward_step(self.model, tokens, position_ids, attention_mask,
                                                     self.inference_params, micro_batch_size)
        return _no_pipelining_forward_step(self.model, tokens, position_ids, attention_mask, self.inference_params)


def _get_recv_buffer_dtype(args):
    return torch.float if args.fp32_residual_connection else args.params_dtype


def _allocate_recv_buffer(batch_size, sequence_length):
    if ps.is_pipeline_first_stage():
        return None
    args = _args()
    return torch.empty((sequence_length, batch_size, args.hidden_size), dtype=_get_recv_buffer_dtype(args),
                       device=current_device())


def _forward_step_helper(model, tokens, position_ids, attention_mask, inference_params, recv_buffer=None):
    batch_size, sequence_length = tokens.size(0), tokens.size(1)
    if recv_buffer is None:
        recv_buffer = _allocate_recv_buffer(batch_size, sequence_length)
    recv_from_prev_pipeline_rank_(recv_buffer)
    from ..utils import unwrap_model
    unwrap_model(model).set_input_tensor(recv_buffer)
    output_tensor = model(tokens, position_ids, attention_mask, inference_params=inference_params)
    send_to_next_pipeline_rank(output_tensor)
    return output_tensor


def _no_pipelining_forward_step(model, tokens, position_ids, attention_mask, inference_params, recv_buffer=None):
    output_tensor = _forward_step_helper(model, tokens, position_ids, attention_mask, inference_params,
                                         recv_buffer=recv_buffer)
    inference_params.sequence_len_offset += tokens.size(1)
    return output_tensor if ps.is_pipeline_last_stage() else None


def _with_pipelining_forward_step(model, tokens, position_ids, attention_mask, inference_params, micro_batch_size):
    sequence_length, batch_size = tokens.size(1), tokens.size(0)
    num_micro_batches, last_chunk = divmod(batch_size, micro_batch_size)
    if last_chunk > 0:
        num_micro_batches += 1
    logits = None
    if ps.is_pipeline_last_stage():
        args = _args()
        logits = torch.empty((batch_size, sequence_length, args.padded_vocab_size), dtype=torch.float32,
                             device=current_device())
    recv_buffer = _allocate_recv_buffer(micro_batch_size, sequence_length)
    for mb in range(num_micro_batches):
        start = mb * micro_batch_size
        end = min(start + micro_batch_size, batch_size)
        this_size = end - start
        if this_size != micro_batch_size:
            recv_buffer = None
        output = _forward_step_helper(model, tokens[start:end], position_ids[start:end], attention_mask,
                                      inference_params, recv_buffer=recv_buffer)
        inference_params.batch_size_offset += this_size
        if ps.is_pipeline_last_stage():
            logits[start:end] = output
    inference_params.sequence_len_offset += sequence_length
    inference_params.batch_size_offset = 0
    return logits

"""Inference API (parity: text_generation/api.py:20-201): ``generate_and_post_process``, ``generate``,
``beam_search_and_post_process``, ``beam_search``.  Rank 0 owns the request; its parameters are broadcast to every rank
as one float list so all ranks enter the same generation loop."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..parallel import state as ps
from .communication import broadcast_float_list
from .generation import (beam_search_and_return_on_first_stage, generate_tokens_probs_and_return_on_first_stage,
                         score_and_return_on_first_stage)
from .tokenization import detokenize_generations, tokenize_prompts


def generate_and_post_process(model, prompts=None, tokens_to_generate=0, return_output_log_probs=False,
                              top_k_sampling=0, top_p_sampling=0.0, top_p_decay=0.0, top_p_bound=0.0,
                              temperature=1.0, add_BOS=False, use_eod_token_for_early_termination=True,
                              stop_on_double_eol=False, stop_on_eol=False, prevent_newline_after_colon=False,
                              random_seed=-1):
    """Returns (texts, segments, log-probs, tokens) on the first pipeline stage, None elsewhere."""
    tokens, lengths, output_log_probs = generate(
        model, prompts=prompts, tokens_to_generate=tokens_to_generate,
        return_output_log_probs=return_output_log_probs, top_k_sampling=top_k_sampling,
        top_p_sampling=top_p_sampling, top_p_decay=top_p_decay, top_p_bound=top_p_bound, temperature=temperature,
        add_BOS=add_BOS, use_eod_token_for_early_termination=use_eod_token_for_early_termination,
        stop_on_double_eol=stop_on_double_eol, stop_on_eol=stop_on_eol,
        prevent_newline_after_colon=prevent_newline_after_colon, random_seed=random_seed)
    if ps.is_pipeline_first_stage():
        tokens, prompts_plus_generations, segments = detokenize_generations(tokens, lengths, True)
        if return_output_log_probs:
            output_log_probs = output_log_probs.cpu().numpy().tolist()
            for i, (prob, seg) in enumerate(zip(output_log_probs, segments)):
                output_log_probs[i] = prob[:len(seg) - 1]
        return prompts_plus_generations, segments, output_log_probs, tokens
    return None


def generate(model, prompts=None, tokens_to_generate=0, return_output_log_probs=False, top_k_sampling=0,
             top_p_sampling=0.0, top_p_decay=0.0, top_p_bound=0.0, temperature=1.0, add_BOS=False,
             use_eod_token_for_early_termination=True, stop_on_double_eol=False, stop_on_eol=False,
             prevent_newline_after_colon=False, random_seed=-1):
    """tokens [b, len], lengths [b], output_log_probs [b, len-1] (or None)."""
    values = [tokens_to_generate, return_output_log_probs, top_k_sampling, top_p_sampling, top_p_decay, top_p_bound,
              temperature, add_BOS, use_eod_token_for_early_termination, stop_on_double_eol, stop_on_eol,
              prevent_newline_after_colon, random_seed]
    if dist.is_initialized() and dist.get_world_size() > 1:
        values = broadcast_float_list(len(values), float_list=values).tolist()
    tokens_to_generate = int(values[0])
    return_output_log_probs = bool(values[1])
    top_k_sampling = int(values[2])
    top_p_sampling, top_p_decay, top_p_bound, temperature = values[3], values[4], values[5], values[6]
    add_BOS = bool(values[7])
    use_eod_token_for_early_termination = bool(values[8])
    stop_on_double_eol, stop_on_eol = bool(values[9]), bool(values[10])
    prevent_newline_after_colon = bool(values[11])
    random_seed = int(values[12])
    if random_seed != -1:
        torch.random.manual_seed(random_seed)
    if (not dist.is_initialized()) or dist.get_rank() == 0:
        assert prompts is not None
    context_tokens, context_lengths = tokenize_prompts(prompts=prompts, tokens_to_generate=tokens_to_generate,
                                                       add_BOS=add_BOS)
    if tokens_to_generate == 0:
        return score_and_return_on_first_stage(model, context_tokens, context_lengths)
    return generate_tokens_probs_and_return_on_first_stage(
        model, context_tokens, context_lengths, return_output_log_probs=return_output_log_probs,
        top_k=top_k_sampling, top_p=top_p_sampling, top_p_decay=top_p_decay, top_p_bound=top_p_bound,
        temperature=temperature, use_eod_token_for_early_termination=use_eod_token_for_early_termination,
        stop_on_double_eol=stop_on_double_eol, stop_on_eol=stop_on_eol,
        prevent_newline_after_colon=prevent_newline_after_colon)


def beam_search_and_post_process(model, prompts=None, tokens_to_generate=0, beam_size=0, add_BOS=False,
                                 stop_token=50256, num_return_gen=1, length_penalty=1,
                                 prevent_newline_after_colon=False):
    tokens, scores = beam_search(model, prompts=prompts, tokens_to_generate=tokens_to_generate, beam_size=beam_size,
                                 add_BOS=add_BOS, stop_token=stop_token, num_return_gen=num_return_gen,
                                 length_penalty=length_penalty,
                                 prevent_newline_after_colon=prevent_newline_after_colon)
    if ps.is_pipeline_first_stage():
        lengths = tokens.size(1) * torch.ones(beam_size, dtype=torch.int64, device=tokens.device)
        tokens, prompts_plus_generations, segments = detokenize_generations(tokens, lengths, True)
        return prompts_plus_generations, segments, scores.cpu().numpy().tolist()
    return None


def beam_search(model, prompts=None, tokens_to_generate=0, beam_size=0, add_BOS=False, stop_token=50256,
                num_return_gen=1, length_penalty=1, prevent_newline_after_colon=False):
    values = [tokens_to_generate, beam_size, add_BOS, stop_token, num_return_gen, length_penalty,
              prevent_newline_after_colon]
    if dist.is_initialized() and dist.get_world_size() > 1:
        values = broadcast_float_list(len(values), float_list=values).tolist()
    tokens_to_generate, beam_size = int(values[0]), int(values[1])
    add_BOS, stop_token, num_return_gen = bool(values[2]), int(values[3]), int(values[4])
    length_penalty, prevent_newline_after_colon = values[5], bool(values[6])
    context_tokens, context_lengths = tokenize_prompts(prompts=prompts, tokens_to_generate=tokens_to_generate,
                                                       add_BOS=add_BOS)
    return beam_search_and_return_on_first_stage(model, context_tokens, context_lengths, beam_size,
                                                 stop_token=stop_token, num_return_gen=num_return_gen,
                                                 length_penalty=length_penalty,
                                                 prevent_newline_after_colon=prevent_newline_after_colon)

"""Token generation loops: scoring, sampling-based generation and beam search.

Parity target: text_generation/generation.py (score :20-86, generate :89-286, beam search :288-415,
``_build_attention_mask_and_position_ids`` :418-429).  The model is expected to return FULL-vocabulary logits
(``parallel_output=False``) like the reference's inference model."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..parallel import state as ps
from ..utils import get_ltor_masks_and_position_ids
from ..utils.device import current_device
from .beam_utils import BeamHypotheses
from .communication import (broadcast_from_last_pipeline_stage, broadcast_from_last_to_first_pipeline_stage,
                            copy_from_last_to_first_pipeline_stage)
from .forward_step import ForwardStep
from .sampling import sample


def _glob():
    from ..global_vars import get_args, get_tokenizer
    return get_args(), get_tokenizer()


def score_and_return_on_first_stage(model, tokens, lengths):
    """Log-probabilities of the given tokens (no generation): output_log_probs [b, s-1] on the first stage."""
    args, _ = _glob()
    batch_size, max_sequence_length = tokens.size(0), tokens.size(1)
    assert max_sequence_length == lengths.max().item()
    if max_sequence_length > args.max_position_embeddings:
        raise ValueError("Length of prompt + tokens_to_generate longer than allowed")
    if max_sequence_length * batch_size > args.max_tokens_to_oom:
        raise ValueError("Too many tokens.  " + str(max_sequence_length * batch_size) + " is greater than " +
                         str(args.max_tokens_to_oom))
    forward_step = ForwardStep(model, batch_size, max_sequence_length)
    output_log_probs = None
    size = (batch_size, max_sequence_length - 1)
    if ps.is_pipeline_last_stage():
        output_log_probs = torch.empty(size, dtype=torch.float32, device=current_device())
    with torch.no_grad():
        attention_mask, position_ids = _build_attention_mask_and_position_ids(tokens)
        logits = forward_step(tokens, position_ids, attention_mask)
        if ps.is_pipeline_last_stage():
            assert logits is not None
            log_probs = F.log_softmax(logits.float(), dim=2)
            idx = torch.unsqueeze(tokens[:, 1:], 2)
            output_log_probs = torch.gather(log_probs, 2, idx).squeeze(2)
    output_log_probs = broadcast_from_last_to_first_pipeline_stage(size, torch.float32, output_log_probs)
    return tokens, lengths, output_log_probs


def generate_tokens_probs_and_return_on_first_stage(model, tokens, lengths, return_output_log_probs=False, top_k=0,
                                                    top_p=0.0, top_p_decay=0.0, top_p_bound=0.0, temperature=1.0,
                                                    use_eod_token_for_early_termination=True,
                                                    stop_on_double_eol=False, stop_on_eol=False,
                                                    prevent_newline_after_colon=True):
    """Incremental decoding: prefill up to the shortest prompt, then one token per step with the KV cache.
    Returns (tokens [b, <=max], generated lengths [b], output_log_probs [b, len-1] | None) on the first stage."""
    args, tokenizer = _glob()
    batch_size = tokens.size(0)
    min_prompt_length = lengths.min().item()
    max_sequence_length = tokens.size(1)
    if max_sequence_length > args.max_position_embeddings:
        raise ValueError("Length of prompt + tokens_to_generate longer than allowed")
    if max_sequence_length * batch_size > args.max_tokens_to_oom:
        raise ValueError("Too many tokens.  " + str(max_sequence_length * batch_size) + " is greater than " +
                         str(args.max_tokens_to_oom))
    forward_step = ForwardStep(model, batch_size, max_sequence_length)
    termination_id = args.eos_id if hasattr(args, "eos_id") else tokenizer.eod
    dev = current_device()
    output_log_probs = None
    log_probs_size = (batch_size, max_sequence_length - 1)
    generated_sequence_lengths = None
    if ps.is_pipeline_last_stage():
        if return_output_log_probs:
            output_log_probs = torch.empty(log_probs_size, dtype=torch.float32, device=dev)
        generated_sequence_lengths = torch.ones(batch_size, dtype=torch.int64, device=dev) * max_sequence_length
    is_generation_done = torch.zeros(batch_size, dtype=torch.uint8, device=dev)
    colon_id = newline_id = None
    if prevent_newline_after_colon:
        try:
            colon_id, newline_id = tokenizer.tokenize(":")[0], tokenizer.tokenize("\n")[0]
        except Exception:
            prevent_newline_after_colon = False
    context_length = min_prompt_length
    with torch.no_grad():
        attention_mask, position_ids = _build_attention_mask_and_position_ids(tokens)
        prev = 0
        for context_length in range(min_prompt_length, max_sequence_length):
            tokens2use = tokens[:, prev:context_length]
            positions2use = position_ids[:, prev:context_length]
            mask2use = attention_mask[..., prev:context_length, :context_length] if attention_mask is not None else None
            logits = forward_step(tokens2use, positions2use, mask2use)
            if ps.is_pipeline_last_stage():
                assert logits is not None
                logits = logits.float()
                if prevent_newline_after_colon:
                    logits[tokens2use[:, -1] == colon_id, -1, newline_id] = -1e10
                new_sample = sample(logits[:, -1, :].contiguous(), top_k=top_k, top_p=top_p, temperature=temperature,
                                    vocab_size=tokenizer.vocab_size)
                if top_p > 0.0 and top_p_decay > 0.0:
                    top_p = top_p * top_p_decay
                    if top_p_bound > 0.0:
                        top_p = max(top_p, top_p_bound)
                started = lengths <= context_length
                tokens[started, context_length] = new_sample[started]
                if return_output_log_probs:
                    log_probs = F.log_softmax(logits, dim=2)
                    idx = torch.unsqueeze(tokens[:, (prev + 1):(context_length + 1)], 2)
                    output_log_probs[:, prev:context_length] = torch.gather(log_probs, 2, idx).squeeze(2)
            copy_from_last_to_first_pipeline_stage(batch_size, torch.int64, tokens[:, context_length])
            prev = context_length
            done = None
            if ps.is_pipeline_last_stage():
                started_b = started.byte()
                if stop_on_double_eol:
                    done_token = ((new_sample == 628).byte() & started_b) | \
                        ((new_sample == 198).byte() & (tokens[:, context_length - 1] == 198).byte() & started_b)
                elif stop_on_eol:
                    done_token = ((new_sample == 628).byte() & started_b) | ((new_sample == 198).byte() & started_b)
                else:
                    done_token = (new_sample == termination_id).byte() & started_b
                just_finished = (done_token & ~is_generation_done).bool()
                generated_sequence_lengths[just_finished.view(-1)] = context_length + 1
                is_generation_done = is_generation_done | done_token
                done = torch.all(is_generation_done)
            done = broadcast_from_last_pipeline_stage(1, torch.uint8, tensor=done)
            if use_eod_token_for_early_termination and done:
                break
    tokens = tokens[:, :(context_length + 1)]
    if ps.is_pipeline_last_stage() and return_output_log_probs:
        output_log_probs = output_log_probs[:, :context_length]
    generated_sequence_lengths = broadcast_from_last_to_first_pipeline_stage(batch_size, torch.int64,
                                                                             generated_sequence_lengths)
    if return_output_log_probs:
        output_log_probs = broadcast_from_last_to_first_pipeline_stage((batch_size, context_length), torch.float32,
                                                                       output_log_probs)
    return tokens, generated_sequence_lengths, output_log_probs


def beam_search_and_return_on_first_stage(model, tokens, lengths, beam_size, stop_token, num_return_gen,
                                          length_penalty, prevent_newline_after_colon=True):
    """Batch-size-1 beam search keeping 2*beam candidates per step; the KV cache is re-ordered to follow the beams."""
    args, tokenizer = _glob()
    batch_size = tokens.size(0)
    assert batch_size == 1
    prompt_length = lengths.item()
    final_sequence_length = tokens.size(1)
    final_sequence_length = min(final_sequence_length, args.max_position_embeddings)
    if prompt_length >= final_sequence_length:
        raise ValueError("context length + tokens_to_generate too large")
    forward_step = ForwardStep(model, beam_size, final_sequence_length)
    dev = current_device()
    beam_hyp = BeamHypotheses(beam_size, length_penalty)
    best_batches = None
    done = torch.zeros(1, dtype=torch.uint8, device=dev)
    scores = torch.zeros(beam_size, dtype=torch.float32, device=dev).unsqueeze(1)
    scores_size_tensor, tokens_size_tensor = None, None
    with torch.no_grad():
        tokens = tokens.repeat(beam_size, 1)
        attention_mask, position_ids = _build_attention_mask_and_position_ids(tokens)
        prev = 0
        for context_length in range(prompt_length, final_sequence_length):
            tokens2use = tokens[:, prev:context_length]
            positions2use = position_ids[:, prev:context_length]
            mask2use = attention_mask[..., prev:context_length, :context_length] if attention_mask is not None else None
            logits = forward_step(tokens2use, positions2use, mask2use)
            if ps.is_pipeline_last_stage():
                logits = logits.float()
                if prevent_newline_after_colon:
                    try:
                        logits[tokens2use[:, -1] == tokenizer.tokenize(":")[0], -1, tokenizer.tokenize("\n")[0]] = -1e10
                    except Exception:
                        pass
                vocab_size = logits.size(2)
                log_probs = F.log_softmax(logits, dim=2)
                new_scores = log_probs[:, -1, :] + scores
                if context_length == prompt_length:  # all beams are identical at the first step
                    sorted_scores, indices = torch.sort(new_scores[0, :], descending=True)
                else:
                    sorted_scores, indices = torch.sort(new_scores.view(-1), descending=True)
                best_beam_ids = torch.div(indices[: 2 * beam_size], vocab_size).trunc().long()
                best_words = indices[: 2 * beam_size] % vocab_size
                best_scores = sorted_scores[: 2 * beam_size]
                next_beams = []
                for rank_, (token_id, beam_score, beam_id) in enumerate(zip(best_words, best_scores, best_beam_ids)):
                    if token_id.item() == stop_token:
                        if rank_ >= beam_size:   # a finished hypothesis outside the top beam_size is dropped
                            continue
                        beam_hyp.add(tokens[beam_id].clone(), beam_score, context_length + 1 - prompt_length)
                    else:
                        next_beams.append((token_id, beam_score, beam_id))
                    if len(next_beams) == beam_size:
                        break
                if beam_hyp.is_done(best_scores.max().item(), context_length + 1 - prompt_length):
                    done = torch.ones(1, dtype=torch.uint8, device=dev)
                best_batches = tokens.new([item[2] for item in next_beams])
                tokens = tokens[best_batches, :]
                tokens[:, context_length] = tokens.new([item[0] for item in next_beams])
                scores = scores.new([item[1] for item in next_beams]).unsqueeze(1)
            done = broadcast_from_last_pipeline_stage(1, torch.uint8, done)
            if done:
                break
            copy_from_last_to_first_pipeline_stage(tokens.size(), torch.int64, tokens)
            best_batches = broadcast_from_last_pipeline_stage(beam_size, torch.int64, best_batches)
            forward_step.inference_params.swap_key_value_dict(best_batches)
            prev = context_length
        if ps.is_pipeline_last_stage():
            if not done:  # ran out of length: every live beam becomes a hypothesis
                for beam_id in range(beam_size):
                    beam_hyp.add(tokens[beam_id].clone(), scores[beam_id].squeeze(), context_length + 1 - prompt_length)
            sorted_hyps = sorted(beam_hyp.beams, key=lambda x: x[0], reverse=True)
            num_return_gen = min(num_return_gen, len(sorted_hyps))
            scores = torch.stack([sorted_hyps[i][0] for i in range(num_return_gen)], dim=0)
            tokens = torch.stack([sorted_hyps[i][1] for i in range(num_return_gen)], dim=0)
            scores_size_tensor = torch.tensor(scores.shape, dtype=torch.int64, device=dev)
            tokens_size_tensor = torch.tensor(tokens.shape, dtype=torch.int64, device=dev)
        scores_size_tensor = broadcast_from_last_pipeline_stage(1, torch.int64, scores_size_tensor)
        tokens_size_tensor = broadcast_from_last_pipeline_stage(2, torch.int64, tokens_size_tensor)
        scores = broadcast_from_last_to_first_pipeline_stage(tuple(scores_size_tensor), torch.float32, scores)
        tokens = broadcast_from_last_to_first_pipeline_stage(tuple(tokens_size_tensor), torch.int64, tokens)
    return tokens, scores


def _build_attention_mask_and_position_ids(tokens):
    """Causal mask (only materialised for the unfused attention path) and position ids for a prompt batch."""
    args, tokenizer = _glob()
    attention_mask, _, position_ids = get_ltor_masks_and_position_ids(
        data=tokens, eod_token=None, reset_position_ids=False, reset_attention_mask=False, eod_mask_loss=False,
        build_attention_mask=not args.use_flash_attn)
    return attention_mask, position_ids

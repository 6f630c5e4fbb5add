"""Prompt tokenisation / detokenisation for generation (parity: text_generation/tokenization.py:14-118): rank 0
tokenises, pads every prompt to ``max_prompt_len + tokens_to_generate`` with EOD and broadcasts sizes, tokens and
lengths."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..utils.device import current_device
from .communication import broadcast_int_list, broadcast_tensor


def _tok():
    from ..global_vars import get_args, get_tokenizer
    return get_args(), get_tokenizer()


def detokenize_generations(tokens_gpu_tensor, lengths_gpu_tensor, return_segments):
    args, tokenizer = _tok()
    prompts_plus_generations, segments_all = [], []
    tokens = tokens_gpu_tensor.cpu().numpy().tolist()
    lengths = lengths_gpu_tensor.cpu().numpy().tolist()
    for seq, length in zip(tokens, lengths):
        seq = seq[:length]
        prompts_plus_generations.append(tokenizer.detokenize(seq))
        if return_segments:
            words = []
            for token in seq:
                if args.tokenizer_type in ("SentencePieceTokenizer", "FalconTokenizer", "NullTokenizer"):
                    words.append(tokenizer.detokenize([token]))
                else:  # GPT-2 byte-level BPE: map the token string back through the byte decoder
                    piece = tokenizer.tokenizer.decoder[token]
                    words.append(bytearray(tokenizer.tokenizer.byte_decoder[c] for c in piece).decode(
                        "utf-8", errors="replace"))
            segments_all.append(words)
    if return_segments:
        return tokens, prompts_plus_generations, segments_all
    return tokens, prompts_plus_generations


def tokenize_prompts(prompts=None, tokens_to_generate=None, add_BOS=None, rank=0):
    """Returns (tokens [b, max_len] padded with EOD, lengths [b]) on every rank."""
    sizes_list = prompts_tokens = prompts_length = None
    me = dist.get_rank() if dist.is_initialized() else 0
    if me == rank:
        assert prompts is not None and tokens_to_generate is not None
        prompts_tokens, prompts_length = _tokenize_prompts_and_batch(prompts, tokens_to_generate, add_BOS)
        sizes_list = [prompts_tokens.size(0), prompts_tokens.size(1)]
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return prompts_tokens, prompts_length
    sizes = broadcast_int_list(2, int_list=sizes_list, rank=rank).tolist()
    tokens = broadcast_tensor(sizes, torch.int64, tensor=prompts_tokens, rank=rank)
    lengths = broadcast_tensor(sizes[0], torch.int64, tensor=prompts_length, rank=rank)
    return tokens, lengths


def _tokenize_prompts_and_batch(prompts, tokens_to_generate, add_BOS):
    args, tokenizer = _tok()
    if add_BOS:
        bos = getattr(tokenizer, "bos_token_id", None)
        bos = bos if bos is not None else tokenizer.eod
        toks = [[bos] + tokenizer.tokenize(p) for p in prompts]
    else:
        toks = [tokenizer.tokenize(p) for p in prompts]
    lengths = [len(t) for t in toks]
    total = max(lengths) + tokens_to_generate
    for t, n in zip(toks, lengths):
        t.extend([tokenizer.eod] * (total - n))
    dev = current_device()
    return torch.tensor(toks, dtype=torch.long, device=dev), torch.tensor(lengths, dtype=torch.long, device=dev)

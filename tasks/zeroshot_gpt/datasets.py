"""Zero-shot evaluation datasets: WikiText-103 perplexity (overlapping windows) and LAMBADA last-word accuracy
(parity: tasks/zeroshot_gpt/datasets.py)."""
import json
import math

import numpy as np
import torch

from megatron_llm_b200 import get_args, get_tokenizer, print_rank_0

from .detokenizer import get_detokenizer


def build_dataset(task):
    if task == "LAMBADA":
        return _build_lambada_dataset()
    if task == "WIKITEXT103":
        return _build_wikitext103_dataset()
    raise NotImplementedError("dataset for {} task is not implemented.".format(task))


def _pad(tokens, pad_mask, length, pad_idx):
    short = length - len(tokens)
    if short > 0:
        tokens = tokens + [pad_idx] * short
        pad_mask = pad_mask + [0] * short
    return tokens, pad_mask


class _LMDataset(torch.utils.data.Dataset):
    """Windows of ``seq_len + 1`` tokens every ``overlapping_eval`` tokens; only the new tokens of a window score."""

    def __init__(self, tokens, seq_len, pad_idx, num_original_tokens, num_tokenized_tokens, overalapping_eval=None):
        self.tokens, self.seq_len, self.pad_idx = tokens, seq_len, pad_idx
        self.overalapping_eval = max(1, overalapping_eval if overalapping_eval is not None else seq_len)
        self.num_original_tokens, self.num_tokenized_tokens = num_original_tokens, num_tokenized_tokens
        self.total_targets = len(tokens) - 1
        rest = max(self.total_targets - self.overalapping_eval, 0)
        self.total_sequences = max(math.ceil(rest / self.overalapping_eval) + 1, 1)

    def __len__(self):
        return self.total_sequences

    def __getitem__(self, idx):
        start = idx * self.overalapping_eval
        tokens = list(self.tokens[start:start + self.seq_len + 1])
        tokens, pad_mask = _pad(tokens, [1] * len(tokens), self.seq_len + 1, self.pad_idx)
        pad_mask = np.array(pad_mask[1:])
        if self.overalapping_eval != self.seq_len and idx != 0:
            pad_mask[:-self.overalapping_eval] *= 0
        return {"text": np.array(tokens), "pad_mask": pad_mask}


class _LambadaDataset(torch.utils.data.Dataset):
    def __init__(self, path, pad_idx, tokenizer, seq_len, strict=False):
        print_rank_0("> building lambada dataset from {} ...".format(path))
        self.seq_len, self.pad_idx, self.tokenizer, self.strict = seq_len, pad_idx, tokenizer, strict
        self.tokens, self.labels = [], []
        with open(path, "r") as f:
            for line in f:
                t, l = self.get_tokens(json.loads(line)["text"])
                self.tokens.append(t)
                self.labels.append(l)

    def get_tokens(self, text):
        """non-strict: the last BPE token is the target; strict: all tokens of the last whitespace word."""
        if not self.strict:
            tokens = self.tokenizer.tokenize(text)
            return tokens[:-1], [tokens[-1]]
        last = text.split()[-1]
        start = text.rfind(last)
        return self.tokenizer.tokenize(text[:start].strip()), self.tokenizer.tokenize(" " + last)

    def __len__(self):
        return len(self.tokens)

    def __getitem__(self, idx):
        ctx, labels = self.tokens[idx], self.labels[idx]
        tokens, pad_mask = _pad(ctx + labels, [0] * len(ctx) + [1] * len(labels), self.seq_len + 1, self.pad_idx)
        return {"text": np.array(tokens), "pad_mask": np.array(pad_mask[1:])}


def _build_lambada_dataset():
    args, tok = get_args(), get_tokenizer()
    assert len(args.valid_data) == 1
    ds = _LambadaDataset(args.valid_data[0], tok.eod, tok, args.seq_length, args.strict_lambada)
    print_rank_0(" > found {} samples.".format(len(ds)))
    return ds


def _build_wikitext103_dataset():
    args, tok = get_args(), get_tokenizer()
    assert len(args.valid_data) == 1
    with open(args.valid_data[0], "rb") as f:
        data = f.read().decode("utf-8")
    num_original_tokens = len(data.strip().split(" "))
    detok = get_detokenizer(args.valid_data[0])
    tokenized = tok.tokenize(detok(data) if detok else data)
    ds = _LMDataset(tokenized, args.seq_length, tok.eod, num_original_tokens, len(tokenized), args.overlapping_eval)
    print_rank_0(" > number of original tokens: {}, number of detokenized tokens: {}".format(num_original_tokens,
                                                                                               len(tokenized)))
    return ds

"""Instruction-tuning dataset: parallel ``<prefix>-text`` / ``<prefix>-role`` indexed datasets + padding collator.

Parity: megatron/data/instruction_dataset.py (Role :20-24, InstructionDataset :26-52, builders :54-315,
instruction_collator :321-355).  Samples are whole conversations (never packed): each epoch is a fresh permutation of
the documents of a split; the collator right-pads to ``seq_length + 1`` (or, with ``--variable_seq_lengths``, to the
longest sample rounded up to 16) and emits the attention / assistant / pad masks the loss uses."""
from __future__ import annotations

import time
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from .blendable_dataset import BlendableDataset
from .dataset_utils import get_datasets_weights_and_num_samples, get_train_valid_test_split_
from .indexed_dataset import make_dataset


def _print0(msg):
    from ..utils import print_rank_0
    print_rank_0(msg)


class Role(IntEnum):
    system = 0
    prompter = 1
    assistant = 2


class InstructionDataset(Dataset):
    def __init__(self, name: str, sample_indices: np.ndarray, indexed_datasets: dict, seq_length: int):
        self.indexed_text, self.indexed_role = indexed_datasets["text"], indexed_datasets["role"]
        assert len(self.indexed_text) == len(self.indexed_role)
        assert sample_indices.min() >= 0 and sample_indices.max() < len(self.indexed_text)
        self.name, self.sample_indices, self.seq_length = name, sample_indices, seq_length

    def __len__(self) -> int:
        return int(self.sample_indices.shape[0])

    def __getitem__(self, idx) -> dict:
        doc = self.sample_indices[idx]
        text, role = self.indexed_text.get(doc), self.indexed_role.get(doc)
        assert text is not None and role is not None and text.shape == role.shape
        return {"text": text.astype(np.int64), "role": role.astype(np.int64)}


def get_indexed_datasets_(data_prefix: str, data_impl: str, skip_warmup: bool) -> dict:
    _print0(" > building dataset index ...")
    t0 = time.time()
    text = make_dataset(f"{data_prefix}-text", data_impl, skip_warmup)
    role = make_dataset(f"{data_prefix}-role", data_impl, skip_warmup)
    assert text is not None
    _print0(f" > finished creating indexed dataset in {time.time() - t0:4f} seconds")
    _print0(f"    number of documents: {len(text)}")
    _print0(f"    number of tokens: {int(np.sum(text.sizes))}")
    return {"text": text, "role": role}


def _sample_dataset(np_rng, document_indices, indexed_datasets, name, num_samples, seq_length):
    """``num_samples`` draws: full permutations of the split's documents, the last one truncated."""
    assert num_samples > 0
    epochs = []
    remaining = num_samples
    while remaining > 0:
        take = min(remaining, len(document_indices))
        epochs.append(np_rng.permutation(document_indices)[:take])
        remaining -= take
    return InstructionDataset(name, np.concatenate(epochs), indexed_datasets, seq_length)


def _whole_prefix_dataset(name, prefix, data_impl, num_samples, seq_length, seed, skip_warmup):
    """Dataset over every document of one prefix (used when train/valid/test paths are given separately)."""
    ds = get_indexed_datasets_(prefix, data_impl, skip_warmup)
    n = len(ds["text"])
    _print0(f"    {name}:\n     document indices in [0, {n}) total of {n} documents")
    return _sample_dataset(np.random.RandomState(seed=seed), np.arange(n, dtype=np.int32), ds, name, num_samples,
                           seq_length)


def _build_dataset(name, data_prefix: Sequence[str], data_impl, num_samples, seq_length, seed, skip_warmup):
    if len(data_prefix) == 1:
        return _whole_prefix_dataset(name, data_prefix[0], data_impl, num_samples, seq_length, seed, skip_warmup)
    prefixes, weights, per_ds = get_datasets_weights_and_num_samples(data_prefix, num_samples)
    parts = [_whole_prefix_dataset(name, p, data_impl, n, seq_length, seed, skip_warmup)
             for p, n in zip(prefixes, per_ds)]
    parts = [p for p in parts if p]
    return BlendableDataset(parts, weights) if parts else None


def _build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples, seq_length,
                                     seed, skip_warmup):
    ds = get_indexed_datasets_(data_prefix, data_impl, skip_warmup)
    n = len(ds["text"])
    splits = get_train_valid_test_split_(splits_string, n)
    _print0(" > dataset split:")
    names = ("train", "validation", "test")
    for i, name in enumerate(names):
        _print0(f"    {name}\n    document indices in [{splits[i]}, {splits[i + 1]}) total of "
                f"{splits[i + 1] - splits[i]}")
    rng = np.random.RandomState(seed=seed)
    order = rng.permutation(n)
    out = []
    for i, name in enumerate(names):
        lo, hi = splits[i], splits[i + 1]
        out.append(None if hi <= lo else _sample_dataset(rng, order[lo:hi], ds, name,
                                                         train_valid_test_num_samples[i], seq_length))
    return tuple(out)


def build_train_valid_test_datasets(data_prefix: Optional[Sequence[str]], data_impl: str, splits_string: str,
                                    train_valid_test_num_samples, seq_length: int, seed: int, skip_warmup: bool,
                                    train_data_prefix=None, valid_data_prefix=None, test_data_prefix=None):
    if data_prefix:
        _print0("Single data path provided for train, valid & test")
        if len(data_prefix) == 1:
            return _build_train_valid_test_datasets(data_prefix[0], data_impl, splits_string,
                                                    train_valid_test_num_samples, seq_length, seed, skip_warmup)
        prefixes, weights, per_ds = get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples)
        columns = ([], [], [])
        for p, n in zip(prefixes, per_ds):
            for col, d in zip(columns, _build_train_valid_test_datasets(p, data_impl, splits_string, n, seq_length,
                                                                        seed, skip_warmup)):
                if d:
                    col.append(d)
        return tuple(BlendableDataset(c, weights) if c else None for c in columns)
    _print0("Separate data paths provided for train, valid & test. Split string will be ignored.")
    out = []
    for name, prefix, n, warm in (("train", train_data_prefix, train_valid_test_num_samples[0], skip_warmup),
                                  ("valid", valid_data_prefix, train_valid_test_num_samples[1], False),
                                  ("test", test_data_prefix, train_valid_test_num_samples[2], False)):
        out.append(None if prefix is None else _build_dataset(name, prefix, data_impl, n, seq_length, seed, warm))
    return tuple(out)


def round_to_multiple_of(x: int, y: int) -> int:
    return ((x + y - 1) // y) * y


def collate(data, seq_length: int, pad_id: int, variable_seq_lengths: bool = False) -> dict:
    """Pad/truncate to ``seq_len + 1`` tokens (one extra so labels = tokens shifted by one)."""
    seq_len = seq_length
    if variable_seq_lengths:
        seq_len = min(seq_length, round_to_multiple_of(max(len(x["text"]) for x in data), 16))
    seq_len += 1
    n = len(data)
    tokens = torch.full((n, seq_len), pad_id, dtype=torch.long)
    role = torch.full((n, seq_len), -1, dtype=torch.long)
    attention_mask = torch.zeros((n, seq_len), dtype=torch.long)
    for i, x in enumerate(data):
        k = min(len(x["text"]), seq_len)
        tokens[i, :k] = torch.from_numpy(x["text"][:k])
        role[i, :k] = torch.from_numpy(x["role"][:k])
        attention_mask[i, :k] = 1
    return {"text": tokens, "attention_mask": attention_mask,
            "assistant_mask": (role == Role.assistant.value).long(), "pad_mask": (tokens == pad_id).long()}


def instruction_collator(data):
    from ..global_vars import get_args, get_tokenizer
    args = get_args()
    return collate(data, args.seq_length, get_tokenizer().pad, args.variable_seq_lengths)

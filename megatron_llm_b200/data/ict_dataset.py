"""Inverse-cloze-task dataset: (pseudo-query sentence, evidence block) pairs.  Parity: megatron/data/ict_dataset.py."""
from __future__ import annotations

import itertools
import random

import numpy as np
from torch.utils.data import Dataset

from .biencoder_dataset_utils import get_block_samples_mapping, make_attention_mask
from .dataset_utils import get_indexed_dataset_


def get_ict_dataset(use_titles=True, query_in_block_prob=1):
    """Single-epoch dataset used for block indexing (``get_block``) rather than training."""
    from ..global_vars import get_args
    args = get_args()
    blocks = get_indexed_dataset_(args.data_path, "mmap", True)
    titles = get_indexed_dataset_(args.titles_data_path, "mmap", True)
    return ICTDataset(name="full", block_dataset=blocks, title_dataset=titles, data_prefix=args.data_path,
                      num_epochs=1, max_num_samples=None, max_seq_length=args.seq_length, seed=1,
                      query_in_block_prob=query_in_block_prob, use_titles=use_titles,
                      use_one_sent_docs=args.use_one_sent_docs)


class ICTDataset(Dataset):
    def __init__(self, name, block_dataset, title_dataset, data_prefix, num_epochs, max_num_samples, max_seq_length,
                 query_in_block_prob, seed, use_titles=True, use_one_sent_docs=False, binary_head=False,
                 tokenizer=None):
        self.name, self.seed, self.max_seq_length = name, seed, max_seq_length
        self.query_in_block_prob = query_in_block_prob
        self.block_dataset, self.title_dataset = block_dataset, title_dataset
        self.rng = random.Random(seed)
        self.use_titles, self.use_one_sent_docs = use_titles, use_one_sent_docs
        self.samples_mapping = get_block_samples_mapping(block_dataset, title_dataset, data_prefix, num_epochs,
                                                         max_num_samples, max_seq_length, seed, name,
                                                         use_one_sent_docs)
        if tokenizer is None:
            from ..global_vars import get_tokenizer
            tokenizer = get_tokenizer()
        self.tokenizer = tokenizer
        self.cls_id, self.sep_id, self.mask_id, self.pad_id = tokenizer.cls, tokenizer.sep, tokenizer.mask, tokenizer.pad

    def __len__(self):
        return len(self.samples_mapping)

    def __getitem__(self, idx):
        sample = self.samples_mapping[idx]
        start, end, doc, _ = sample.as_tuple()
        title = self.title_dataset[int(doc)] if self.use_titles else None
        reserve = 3 + len(title) if self.use_titles else 2
        block = [self.block_dataset[i] for i in range(start, end)]
        assert len(block) > 1 or self.use_one_sent_docs or self.query_in_block_prob == 1
        pick = self.rng.randint(0, len(block) - 1)
        # the query sentence stays in its block with probability query_in_block_prob
        query = block[pick].copy() if self.rng.random() < self.query_in_block_prob else block.pop(pick)
        query = query[:self.max_seq_length - 2]
        block = list(itertools.chain(*block))[:self.max_seq_length - reserve]
        q_tokens, q_pad = self.concat_and_pad_tokens(query)
        c_tokens, c_pad = self.concat_and_pad_tokens(block, title)
        return {"query_tokens": q_tokens, "query_mask": make_attention_mask(q_tokens, q_tokens),
                "query_pad_mask": q_pad, "context_tokens": c_tokens,
                "context_mask": make_attention_mask(c_tokens, c_tokens), "context_pad_mask": c_pad,
                "block_data": sample.as_array()}

    def get_block(self, start_idx, end_idx, doc_idx):
        block = [self.block_dataset[i] for i in range(start_idx, end_idx)]
        title = self.title_dataset[int(doc_idx)]
        block = list(itertools.chain(*block))[:self.max_seq_length - (3 + len(title))]
        return self.concat_and_pad_tokens(block, title)

    def get_null_block(self):
        return self.concat_and_pad_tokens([], [])

    def concat_and_pad_tokens(self, tokens, title=None):
        tokens = list(tokens)
        if title is None:
            tokens = [self.cls_id] + tokens + [self.sep_id]
        else:
            tokens = [self.cls_id] + list(title) + [self.sep_id] + tokens + [self.sep_id]
        assert len(tokens) <= self.max_seq_length
        pad = self.max_seq_length - len(tokens)
        return np.array(tokens + [self.pad_id] * pad), np.array([1] * len(tokens) + [0] * pad)

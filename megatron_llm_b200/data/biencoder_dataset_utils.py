"""Shared utilities of the ICT / REALM / ORQA retrieval datasets.

Parity: megatron/data/biencoder_dataset_utils.py and realm_dataset_utils.py (the reference keeps two near-identical
copies; ``realm_dataset_utils`` re-exports this module here)."""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..parallel.data import broadcast_data
from ..utils import print_rank_0
from .data_samplers import MegatronPretrainingSampler


def make_attention_mask(source_block, target_block):
    """[len(source), len(target)] keep-mask over real (id >= 1) tokens."""
    return ((target_block[None, :] >= 1) * (source_block[:, None] >= 1)).astype(np.int64)


def get_one_epoch_dataloader(dataset, micro_batch_size=None):
    """Exactly one pass, no dropped tail -- used by the indexing jobs."""
    from ..global_vars import get_args
    args = get_args()
    mbs = micro_batch_size if micro_batch_size is not None else args.micro_batch_size
    sampler = MegatronPretrainingSampler(total_samples=len(dataset), consumed_samples=0, micro_batch_size=mbs,
                                         data_parallel_rank=ps.get_data_parallel_rank(),
                                         data_parallel_size=ps.get_data_parallel_world_size(), drop_last=False)
    return torch.utils.data.DataLoader(dataset, batch_sampler=sampler, num_workers=args.num_workers,
                                       pin_memory=torch.cuda.is_available())


def get_ict_batch(data_iterator):
    keys = ["query_tokens", "query_mask", "context_tokens", "context_mask", "block_data"]
    data = None if data_iterator is None else next(data_iterator)
    d = broadcast_data(keys, data, torch.int64)
    return (d["query_tokens"].long(), d["query_mask"] < 0.5, d["context_tokens"].long(), d["context_mask"] < 0.5,
            d["block_data"].long())


def join_str_list(str_list):
    """WordPiece pieces -> text ('##' continuation pieces are glued to the previous piece)."""
    out = ""
    for s in str_list:
        out += s[2:] if s.startswith("##") else " " + s
    return out


class BlockSampleData:
    """(first sentence, last sentence + 1, document, block id) of one evidence block."""

    def __init__(self, start_idx, end_idx, doc_idx, block_idx):
        self.start_idx, self.end_idx, self.doc_idx, self.block_idx = start_idx, end_idx, doc_idx, block_idx

    def as_array(self):
        return np.array([self.start_idx, self.end_idx, self.doc_idx, self.block_idx]).astype(np.int64)

    def as_tuple(self):
        return self.start_idx, self.end_idx, self.doc_idx, self.block_idx


class BlockSamplesMapping:
    def __init__(self, mapping_array):
        assert mapping_array.shape[1] == 4
        self.mapping_array = mapping_array

    def __len__(self):
        return self.mapping_array.shape[0]

    def __getitem__(self, idx):
        return BlockSampleData(*self.mapping_array[idx])


def get_block_samples_mapping(block_dataset, title_dataset, data_prefix, num_epochs, max_num_samples, max_seq_length,
                              seed, name, use_one_sent_docs=False):
    """Fixed-size evidence blocks (title lengths are accounted for), cached as ``.npy`` next to the data."""
    if not num_epochs:
        if not max_num_samples:
            raise ValueError("Need to specify either max_num_samples or num_epochs")
        num_epochs = np.iinfo(np.int32).max - 1
    if not max_num_samples:
        max_num_samples = np.iinfo(np.int64).max - 1
    fname = f"{data_prefix}_{name}_indexmap"
    if num_epochs != np.iinfo(np.int32).max - 1:
        fname += f"_{num_epochs}ep"
    if max_num_samples != np.iinfo(np.int64).max - 1:
        fname += f"_{max_num_samples}mns"
    fname += f"_{max_seq_length}msl_{seed}s" + ("_1sentok" if use_one_sent_docs else "") + ".npy"
    rank0 = (not dist.is_initialized()) or ps.get_data_parallel_rank() == 0
    if rank0 and not os.path.isfile(fname):
        print(f" > WARNING: could not find index map file {fname}, building the indices on rank 0 ...")
        assert block_dataset.doc_idx.dtype == np.int64 and block_dataset.sizes.dtype == np.int32
        from . import helpers
        t0 = time.time()
        mapping = helpers.build_blocks_mapping(block_dataset.doc_idx, block_dataset.sizes, title_dataset.sizes,
                                               num_epochs, max_num_samples, max_seq_length - 3, seed, True,
                                               use_one_sent_docs)
        np.save(fname, mapping, allow_pickle=True)
        print_rank_0(f" > saved the index mapping in {fname} ({time.time() - t0:4f} s)")
    from .dataset_utils import _sync_after_index_build
    _sync_after_index_build()          # (not a world barrier: only TP-rank-0 ranks build datasets)
    mapping = np.load(fname, allow_pickle=True, mmap_mode="r")
    print_rank_0(f"    total number of samples: {mapping.shape[0]}")
    return BlockSamplesMapping(mapping)

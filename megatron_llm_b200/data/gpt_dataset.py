"""GPT-style dataset: documents laid end to end, samples of ``seq_length + 1`` tokens.

Parity target: megatron/data/gpt_dataset.py (builders :20-238, GPTDataset :241-269, index mappings :272-406 with the
same ``*_indexmap_{ns}ns_{sl}sl_{s}s_{doc,sample,shuffle}_idx.npy`` cache files next to the data).  The cache is
built on global rank 0 and the other ranks wait on a barrier-free counter all-reduce (works on CUDA or CPU)."""
from __future__ import annotations

import os
import time
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..utils import print_rank_0
from ..utils.device import current_device
from . import indexed_dataset as idx_ds
from .blendable_dataset import BlendableDataset
from .dataset_utils import get_datasets_weights_and_num_samples, get_train_valid_test_split_


def build_train_valid_test_datasets(data_prefix: Optional[List[str]], data_impl: str, splits_string: str,
                                    train_valid_test_num_samples: List[int], seq_length: int, seed: int,
                                    skip_warmup: bool, train_data_prefix=None, valid_data_prefix=None,
                                    test_data_prefix=None):
    if data_prefix:
        print_rank_0("Single data path provided for train, valid & test")
        if len(data_prefix) == 1:
            return _build_train_valid_test_datasets(data_prefix[0], data_impl, splits_string,
                                                    train_valid_test_num_samples, seq_length, seed, skip_warmup)
        prefixes, weights, per_ds_samples = get_datasets_weights_and_num_samples(data_prefix,
                                                                                 train_valid_test_num_samples)
        groups = ([], [], [])
        for prefix, nums in zip(prefixes, per_ds_samples):
            for g, ds in zip(groups, _build_train_valid_test_datasets(prefix, data_impl, splits_string, nums,
                                                                      seq_length, seed, skip_warmup)):
                if ds:
                    g.append(ds)
        return tuple(BlendableDataset(g, weights) if g else None for g in groups)
    print_rank_0("Separate data paths provided for train, valid & test. Split string will be ignored.")
    out = []
    for name, prefix, n, warm in (("train", train_data_prefix, train_valid_test_num_samples[0], skip_warmup),
                                  ("valid", valid_data_prefix, train_valid_test_num_samples[1], False),
                                  ("test", test_data_prefix, train_valid_test_num_samples[2], False)):
        out.append(_build_dataset(name, prefix, data_impl, n, seq_length, seed, warm) if prefix is not None else None)
    return tuple(out)


def _build_dataset(dataset_name, data_prefix, data_impl, num_samples, seq_length, seed, skip_warmup):
    if len(data_prefix) == 1:
        return _build_dataset_kernel(dataset_name, data_prefix[0], data_impl, num_samples, seq_length, seed,
                                     skip_warmup)
    prefixes, weights, per_ds = get_datasets_weights_and_num_samples(data_prefix, num_samples)
    datasets = [d for d in (_build_dataset_kernel(dataset_name, p, data_impl, n, seq_length, seed, skip_warmup)
                            for p, n in zip(prefixes, per_ds)) if d]
    return BlendableDataset(datasets, weights) if datasets else None


def _build_dataset_kernel(dataset_name, data_prefix, data_impl, num_samples, seq_length, seed, skip_warmup):
    indexed = get_indexed_dataset_(data_prefix, data_impl, skip_warmup)
    total = indexed.sizes.shape[0]
    print_rank_0("    {}:".format(dataset_name))
    print_rank_0("     document indices in [0, {}) total of {} documents".format(total, total))
    documents = np.arange(0, total, 1, dtype=np.int32)
    return GPTDataset(dataset_name, data_prefix, documents, indexed, num_samples, seq_length, seed)


def _build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples, seq_length,
                                     seed, skip_warmup):
    indexed = get_indexed_dataset_(data_prefix, data_impl, skip_warmup)
    total = indexed.sizes.shape[0]
    splits = get_train_valid_test_split_(splits_string, total)
    print_rank_0(" > dataset split:")
    for i, name in enumerate(("train", "validation", "test")):
        print_rank_0("    {}:".format(name))
        print_rank_0("     document indices in [{}, {}) total of {} documents".format(
            splits[i], splits[i + 1], splits[i + 1] - splits[i]))

    def make(i, name):
        if splits[i + 1] <= splits[i]:
            return None
        documents = np.arange(splits[i], splits[i + 1], 1, dtype=np.int32)
        return GPTDataset(name, data_prefix, documents, indexed, train_valid_test_num_samples[i], seq_length, seed)

    return make(0, "train"), make(1, "valid"), make(2, "test")


def get_indexed_dataset_(data_prefix, data_impl, skip_warmup):
    print_rank_0(" > building dataset index ...")
    t0 = time.time()
    indexed = idx_ds.make_dataset(data_prefix, data_impl, skip_warmup)
    assert indexed is not None
    print_rank_0(" > finished creating indexed dataset in {:4f} seconds".format(time.time() - t0))
    print_rank_0("    number of documents: {}".format(indexed.sizes.shape[0]))
    print_rank_0("    number of tokens: {}".format(int(np.sum(indexed.sizes))))
    return indexed


class GPTDataset(torch.utils.data.Dataset):
    def __init__(self, name, data_prefix, documents, indexed_dataset, num_samples, seq_length, seed):
        self.name = name
        self.indexed_dataset = indexed_dataset
        assert np.min(documents) >= 0
        assert np.max(documents) < indexed_dataset.sizes.shape[0]
        self.doc_idx, self.sample_idx, self.shuffle_idx = _build_index_mappings(
            self.name, data_prefix, documents, self.indexed_dataset.sizes, num_samples, seq_length, seed)

    def __len__(self):
        return self.sample_idx.shape[0] - 1

    def __getitem__(self, idx):
        idx = self.shuffle_idx[idx]
        d0, o0 = self.sample_idx[idx]
        d1, o1 = self.sample_idx[idx + 1]
        get = self.indexed_dataset.get
        if d0 == d1:
            sample = get(self.doc_idx[d0], offset=o0, length=o1 - o0 + 1)
        else:
            pieces = [get(self.doc_idx[d0], offset=o0)]
            pieces.extend(get(self.doc_idx[i]) for i in range(d0 + 1, d1))
            pieces.append(get(self.doc_idx[d1], length=o1 + 1))
            sample = np.concatenate(pieces)
        return {"text": np.array(sample, dtype=np.int64)}


def _build_index_mappings(name, data_prefix, documents, sizes, num_samples, seq_length, seed):
    """doc_idx: shuffled documents for all epochs; sample_idx: (doc_idx position, offset) of every sample start;
    shuffle_idx: random permutation of the samples."""
    tokens_per_epoch = _num_tokens(documents, sizes)
    num_epochs = _num_epochs(tokens_per_epoch, seq_length, num_samples)
    np_rng = np.random.RandomState(seed=seed)
    base = f"{data_prefix}_{name}_indexmap_{num_samples}ns_{seq_length}sl_{seed}s"
    files = {k: f"{base}_{k}_idx.npy" for k in ("doc", "sample", "shuffle")}
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0 and not all(os.path.isfile(f) for f in files.values()):
        print_rank_0(" > WARNING: could not find index map files, building the indices on rank 0 ...")
        if num_epochs == 1:
            separate_last_epoch = False
            print(" > only one epoch required, setting separate_last_epoch to False", flush=True)
        else:
            before_last = ((num_epochs - 1) * tokens_per_epoch - 1) // seq_length
            last_epoch_samples = num_samples - before_last
            assert last_epoch_samples >= 0, "last epoch number of samples should be non-negative."
            per_epoch = (tokens_per_epoch - 1) // seq_length
            assert last_epoch_samples < per_epoch + 1, "last epoch number of samples exceeded max value."
            separate_last_epoch = last_epoch_samples < int(0.80 * per_epoch)
            print(" > last epoch number of samples ({}) is {} than 80% of number of samples per epoch ({}), setting "
                  "separate_last_epoch to {}".format(last_epoch_samples, "smaller" if separate_last_epoch else "larger",
                                                     per_epoch, separate_last_epoch), flush=True)
        t0 = time.time()
        doc_idx = _build_doc_idx(documents, num_epochs, np_rng, separate_last_epoch)
        np.save(files["doc"], doc_idx, allow_pickle=True)
        print_rank_0(" > elasped time to build and save doc-idx mapping (seconds): {:4f}".format(time.time() - t0))
        t0 = time.time()
        from . import helpers
        assert doc_idx.dtype == np.int32 and sizes.dtype == np.int32
        sample_idx = helpers.build_sample_idx(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch)
        np.save(files["sample"], sample_idx, allow_pickle=True)
        print_rank_0(" > elasped time to build and save sample-idx mapping (seconds): {:4f}".format(time.time() - t0))
        t0 = time.time()
        first = ((num_epochs - 1) * tokens_per_epoch - 1) // seq_length if separate_last_epoch \
            else sample_idx.shape[0] - 1
        shuffle_idx = _build_shuffle_idx(first, sample_idx.shape[0] - 1, np_rng)
        np.save(files["shuffle"], shuffle_idx, allow_pickle=True)
        print_rank_0(" > elasped time to build and save shuffle-idx mapping (seconds): {:4f}".format(time.time() - t0))
    # other ranks wait for rank 0 (same cross-group counter trick as the reference, device-agnostic)
    if dist.is_initialized() and ps.model_parallel_is_initialized():
        counts = torch.ones(1, dtype=torch.long, device=current_device())
        dist.all_reduce(counts, group=ps.get_data_parallel_group())
        dist.all_reduce(counts, group=ps.get_pipeline_model_parallel_group())
        assert counts[0].item() == dist.get_world_size() // dist.get_world_size(group=ps.get_tensor_model_parallel_group())
    t0 = time.time()
    doc_idx = np.load(files["doc"], allow_pickle=True, mmap_mode="r")
    sample_idx = np.load(files["sample"], allow_pickle=True, mmap_mode="r")
    shuffle_idx = np.load(files["shuffle"], allow_pickle=True, mmap_mode="r")
    print_rank_0("    loaded indexed file in {:3.3f} seconds".format(time.time() - t0))
    print_rank_0("    total number of samples: {}".format(sample_idx.shape[0]))
    print_rank_0("    total number of epochs: {}".format(num_epochs))
    return doc_idx, sample_idx, shuffle_idx


def _num_tokens(documents, sizes):
    return int(np.sum(sizes[documents]))


def _num_epochs(tokens_per_epoch, seq_length, num_samples):
    """Smallest e with (e * tokens_per_epoch - 1) // seq_length >= num_samples."""
    num_epochs, total = 0, 0
    while True:
        num_epochs += 1
        total += tokens_per_epoch
        if (total - 1) // seq_length >= num_samples:
            return num_epochs


def _build_doc_idx(documents, num_epochs, np_rng, separate_last_epoch):
    if not separate_last_epoch or num_epochs == 1:
        doc_idx = np.tile(np.asarray(documents, dtype=np.int32), num_epochs)
        np_rng.shuffle(doc_idx)
        return doc_idx
    first = _build_doc_idx(documents, num_epochs - 1, np_rng, False)
    last = _build_doc_idx(documents, 1, np_rng, False)
    return np.concatenate((first, last))


def _build_sample_idx(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch):
    """Pure-python twin of the C++ builder (tests compare the two)."""
    num_samples = (num_epochs * tokens_per_epoch - 1) // seq_length
    out = np.zeros([num_samples + 1, 2], dtype=np.int32)
    cursor, offset = 0, 0
    for s in range(1, num_samples + 1):
        need = seq_length + 1
        while need > 0:
            avail = sizes[doc_idx[cursor]] - offset
            if avail >= need:
                offset += need - 1
                need = 0
            else:
                need -= avail
                cursor += 1
                offset = 0
        out[s] = (cursor, offset)
    return out


def _build_shuffle_idx(num_samples, total_size, np_rng):
    print(" > building shuffle index with split [0, {}) and [{}, {}) ...".format(num_samples, num_samples,
                                                                                total_size), flush=True)
    dtype_ = np.uint32 if total_size < (np.iinfo(np.uint32).max - 1) else np.int64
    first = np.arange(0, num_samples, 1, dtype=dtype_)
    np_rng.shuffle(first)
    if num_samples == total_size:
        return first
    last = np.arange(num_samples, total_size, 1, dtype=dtype_)
    np_rng.shuffle(last)
    return np.concatenate((first, last))

"""Batch samplers + loader factory (parity: megatron/data/data_samplers.py:14-187).

``MegatronPretrainingSampler``: every DP rank walks the same global index stream and keeps its own
micro-batch slice.  ``MegatronPretrainingRandomSampler``: epoch-seeded permutation, optionally sharded per
DP rank.  Loaders use pinned memory so the H2D copy of the next micro-batch is asynchronous.
"""
from __future__ import annotations

import random

import numpy as np
import torch
from torch.utils.data import Dataset

from ..parallel import state as ps


def _args():
    from ..global_vars import get_args
    return get_args()


def build_pretraining_data_loader(dataset, consumed_samples, collate_fn=None):
    if dataset is None:
        return None
    args = _args()
    if args.dataloader_type == "single":
        batch_sampler = MegatronPretrainingSampler(
            total_samples=len(dataset), consumed_samples=consumed_samples, micro_batch_size=args.micro_batch_size,
            data_parallel_rank=ps.get_data_parallel_rank(), data_parallel_size=ps.get_data_parallel_world_size())
    elif args.dataloader_type == "cyclic":
        batch_sampler = MegatronPretrainingRandomSampler(
            dataset, total_samples=len(dataset), consumed_samples=consumed_samples,
            micro_batch_size=args.micro_batch_size, data_parallel_rank=ps.get_data_parallel_rank(),
            data_parallel_size=ps.get_data_parallel_world_size(), data_sharding=args.data_sharding)
    else:
        raise Exception("{} dataloader type is not supported.".format(args.dataloader_type))
    return torch.utils.data.DataLoader(dataset, batch_sampler=batch_sampler, num_workers=args.num_workers,
                                       pin_memory=torch.cuda.is_available(), collate_fn=collate_fn)


class MegatronPretrainingSampler:
    def __init__(self, total_samples, consumed_samples, micro_batch_size, data_parallel_rank, data_parallel_size,
                 drop_last=True):
        self.total_samples, self.consumed_samples = total_samples, consumed_samples
        self.micro_batch_size, self.data_parallel_rank = micro_batch_size, data_parallel_rank
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        self.drop_last = drop_last
        assert self.total_samples > 0, "no sample to consume: {}".format(self.total_samples)
        assert self.consumed_samples < self.total_samples, \
            "no samples left to consume: {}, {}".format(self.consumed_samples, self.total_samples)
        assert self.micro_batch_size > 0
        assert data_parallel_size > 0
        assert self.data_parallel_rank < data_parallel_size, \
            "data_parallel_rank should be smaller than data size: {}, {}".format(self.data_parallel_rank,
                                                                                 data_parallel_size)

    def __len__(self):
        return self.total_samples

    def get_start_end_idx(self):
        start = self.data_parallel_rank * self.micro_batch_size
        return start, start + self.micro_batch_size

    def __iter__(self):
        step = self.micro_batch_times_data_parallel_size
        lo, hi = self.get_start_end_idx()
        pos = self.consumed_samples
        while pos + step <= self.total_samples:
            yield list(range(pos + lo, pos + hi))
            pos += step
        if pos < self.total_samples and not self.drop_last:
            tail = list(range(pos, self.total_samples))
            yield tail[lo:hi]


class RandomSeedDataset(Dataset):
    """Re-seeds python/numpy/torch per item so augmentation is reproducible across epochs and workers."""

    def __init__(self, dataset):
        args = _args()
        self.base_seed = self.curr_seed = args.seed
        self.dataset = dataset

    def __len__(self):
        return len(self.dataset)

    def set_epoch(self, epoch):
        self.curr_seed = self.base_seed + epoch

    def __getitem__(self, idx):
        seed = idx + self.curr_seed
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        return self.dataset[idx]


class MegatronPretrainingRandomSampler:
    def __init__(self, dataset, total_samples, consumed_samples, micro_batch_size, data_parallel_rank,
                 data_parallel_size, data_sharding):
        self.dataset, self.total_samples, self.consumed_samples = dataset, total_samples, consumed_samples
        self.micro_batch_size = micro_batch_size
        self.data_parallel_rank, self.data_parallel_size = data_parallel_rank, data_parallel_size
        self.data_sharding = data_sharding
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        self.last_batch_size = self.total_samples % self.micro_batch_times_data_parallel_size
        assert self.total_samples > 0, "no sample to consume: {}".format(self.total_samples)
        assert self.micro_batch_size > 0
        assert data_parallel_size > 0
        assert self.data_parallel_rank < data_parallel_size

    def __len__(self):
        return self.total_samples

    def __iter__(self):
        active = self.total_samples - self.last_batch_size
        self.epoch = self.consumed_samples // active
        in_epoch = self.consumed_samples % active
        assert in_epoch % self.micro_batch_times_data_parallel_size == 0
        if isinstance(self.dataset, RandomSeedDataset):
            self.dataset.set_epoch(self.epoch)
        g = torch.Generator()
        g.manual_seed(self.epoch)
        if self.data_sharding:
            bucket = (self.total_samples // self.micro_batch_times_data_parallel_size) * self.micro_batch_size
            offset = in_epoch // self.data_parallel_size
            start = self.data_parallel_rank * bucket
            idx_range = [start + x for x in torch.randperm(bucket, generator=g).tolist()[offset:]]
        else:
            full = (self.total_samples // self.micro_batch_size) * self.micro_batch_size
            perm = torch.randperm(full, generator=g).tolist()[in_epoch:]
            idx_range = perm[self.data_parallel_rank::self.data_parallel_size]
        batch = []
        for idx in idx_range:
            batch.append(idx)
            if len(batch) == self.micro_batch_size:
                self.consumed_samples += self.micro_batch_times_data_parallel_size
                yield batch
                batch = []

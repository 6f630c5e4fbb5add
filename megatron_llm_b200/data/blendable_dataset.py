"""Weighted interleave of several datasets (parity: megatron/data/blendable_dataset.py:14-53): sample i comes from
the dataset whose running share lags its weight the most (C++ ``build_blending_indices``)."""
from __future__ import annotations

import time

import numpy as np
import torch

from ..utils import print_rank_0


class BlendableDataset(torch.utils.data.Dataset):
    def __init__(self, datasets, weights):
        self.datasets = datasets
        num_datasets = len(datasets)
        assert num_datasets == len(weights)
        self.size = sum(len(d) for d in datasets)
        weights = np.array(weights, dtype=np.float64)
        assert np.sum(weights) > 0.0
        weights /= np.sum(weights)
        start = time.time()
        assert num_datasets < 255
        self.dataset_index = np.zeros(self.size, dtype=np.uint8)
        self.dataset_sample_index = np.zeros(self.size, dtype=np.int64)
        from . import helpers
        rank0 = (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0
        helpers.build_blending_indices(self.dataset_index, self.dataset_sample_index, weights, num_datasets,
                                       self.size, rank0)
        print_rank_0("> elapsed time for building blendable dataset indices: {:.2f} (sec)".format(time.time() - start))

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.datasets[self.dataset_index[idx]][self.dataset_sample_index[idx]]

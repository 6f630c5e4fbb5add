"""Dataset utilities shared by the GPT / BERT / T5 / ICT datasets (parity: megatron/data/dataset_utils.py)."""
from __future__ import annotations

import collections
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ..utils import print_rank_0

DSET_TYPE_BERT = "standard_bert"
DSET_TYPE_ICT = "ict"
DSET_TYPE_T5 = "t5"
DSET_TYPES = [DSET_TYPE_BERT, DSET_TYPE_ICT, DSET_TYPE_T5]


def get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples):
    """``data_prefix = [w1, prefix1, w2, prefix2, ...]`` -> (prefixes, normalised weights, per-dataset sample
    counts inflated by 0.5% so an uneven blend never runs dry)."""
    assert len(data_prefix) % 2 == 0
    weights = [float(w) for w in data_prefix[0::2]]
    prefixes = [p.strip() for p in data_prefix[1::2]]
    total = sum(weights)
    assert total > 0.0
    weights = [w / total for w in weights]
    if isinstance(train_valid_test_num_samples, list):
        per_ds = [[int(math.ceil(v * w * 1.005)) for v in train_valid_test_num_samples] for w in weights]
    else:
        per_ds = [int(math.ceil(train_valid_test_num_samples * w * 1.005)) for w in weights]
    return prefixes, weights, per_ds


def get_train_valid_test_split_(splits_string, size):
    """'969,30,1' or '90/5/5' -> four document boundaries [0, a, b, size]."""
    sep = "," if "," in splits_string else ("/" if "/" in splits_string else None)
    parts = [float(s) for s in splits_string.split(sep)] if sep else [float(splits_string)]
    parts = (parts + [0.0, 0.0, 0.0])[:3]
    total = sum(parts)
    assert total > 0.0
    bounds = [0]
    for frac in parts:
        bounds.append(bounds[-1] + int(round(frac / total * float(size))))
    diff = bounds[-1] - size
    bounds = [bounds[0]] + [b - diff for b in bounds[1:]]
    assert len(bounds) == 4 and bounds[-1] == size
    return bounds

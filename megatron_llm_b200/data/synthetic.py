"""Synthetic token datasets (no network / no files): used by bench.py, the CPU plumbing config and tests."""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset


class SyntheticGPTDataset(Dataset):
    """Deterministic pseudo-random token streams of ``seq_length + 1`` tokens, like GPTDataset.__getitem__.

    ``structured=True`` makes token t+1 a fixed function of token t (learnable), so loss must decrease."""

    def __init__(self, num_samples: int, seq_length: int, vocab_size: int, seed: int = 1234, structured: bool = True):
        self.num_samples, self.seq_length, self.vocab_size = num_samples, seq_length, vocab_size
        self.seed, self.structured = seed, structured

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        rng = np.random.RandomState((self.seed + idx) % (2 ** 31 - 1))
        n = self.seq_length + 1
        if self.structured:
            v = min(self.vocab_size, 4096)
            start = rng.randint(0, v)
            step = 1 + (idx % 7)
            toks = (start + step * np.arange(n)) % v
            noise = rng.rand(n) < 0.05
            toks = np.where(noise, rng.randint(0, v, size=n), toks)
        else:
            toks = rng.randint(0, self.vocab_size, size=n)
        return {"text": torch.from_numpy(toks.astype(np.int64))}

"""Cooperative SIGTERM handling: any rank that received the signal makes ALL ranks checkpoint and exit
(parity: megatron/dist_signal_handler.py:50-81, used from training.py:731-737)."""
from __future__ import annotations

import signal

import torch
import torch.distributed as dist

from .utils.device import current_device


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_gather_item(item, dtype, group=None, async_op=False, local_rank=None):
    if not dist.is_available() or not dist.is_initialized():
        return [item]
    device = current_device()
    tensor = torch.tensor([item], device=device, dtype=dtype)
    out = [torch.zeros(1, dtype=dtype, device=device) for _ in range(get_world_size())]
    dist.all_gather(out, tensor, group, async_op)
    return [t.item() for t in out]


class DistributedSignalHandler:
    def __init__(self, sig=signal.SIGTERM):
        self.sig = sig
        self._signal_received = False
        self.released = False
        self.original_handler = None

    def signals_received(self):
        return all_gather_item(self._signal_received, dtype=torch.int32)

    def __enter__(self):
        self._signal_received = False
        self.released = False
        self.original_handler = signal.getsignal(self.sig)

        def handler(signum, frame):
            self._signal_received = True

        signal.signal(self.sig, handler)
        return self

    def __exit__(self, type, value, tb):
        self.release()

    def release(self):
        if self.released:
            return False
        signal.signal(self.sig, self.original_handler)
        self.released = True
        return True

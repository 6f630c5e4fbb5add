"""Pre-allocated buffers (parity: megatron/memory.py:9-132): named ``MemoryBuffer`` arena with bump allocation
and a ``RingMemBuffer`` rotating over several of them.  Legacy API (unused on the main training path, which
uses the grow-only GlobalMemoryBuffer and the DDP flat buffers) kept for user scripts."""
from __future__ import annotations

import torch

from .utils.device import current_device

_MEM_BUFFS = dict()


def allocate_mem_buff(name, numel, dtype, track_usage):
    assert name not in _MEM_BUFFS, "memory buffer {} already allocated.".format(name)
    _MEM_BUFFS[name] = MemoryBuffer(name, numel, dtype, track_usage)
    return _MEM_BUFFS[name]


def get_mem_buff(name):
    return _MEM_BUFFS[name]


class MemoryBuffer:
    """Contiguous arena; ``add(tensor)`` copies the tensor in and returns the view, ``reset()`` rewinds."""

    def __init__(self, name, numel, dtype, track_usage):
        if not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0:
            element_size = torch.tensor([], dtype=dtype).element_size()
            print("> building the {} memory buffer with {} num elements and {} dtype ({:.1f} MB)...".format(
                name, numel, dtype, numel * element_size / 1024 / 1024), flush=True)
        self.name, self.numel, self.dtype = name, numel, dtype
        self.data = torch.empty(self.numel, dtype=self.dtype, device=current_device(), requires_grad=False)
        self._start = 0
        self.track_usage = track_usage
        if self.track_usage:
            self.in_use_value = 0.0
            self.total_value = 0.0

    def reset(self):
        self._start = 0

    def is_in_use(self):
        return self._start > 0

    def numel_in_use(self):
        return self._start

    def add(self, tensor):
        assert tensor.dtype == self.dtype, "Input tensor type {} different from buffer type {}".format(
            tensor.dtype, self.dtype)
        n = torch.numel(tensor)
        new_start = self._start + n
        assert new_start <= self.numel, "Not enough memory left in the buffer ({} > {})".format(
            n, self.numel - self._start)
        view = self.data[self._start:new_start].view(tensor.shape)
        self._start = new_start
        view.copy_(tensor)
        return view

    def get_data(self):
        if self.track_usage:
            self.in_use_value += float(self._start)
            self.total_value += float(self.numel)
        return self.data[:self._start]

    def print_average_usage(self):
        assert self.track_usage, "You need to enable track usage."
        if not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0:
            print(" > usage of {} memory buffer: {:.2f} %".format(
                self.name, self.in_use_value * 100.0 / max(self.total_value, 1.0)), flush=True)


class RingMemBuffer:
    """Round-robin over ``num_buffers`` arenas."""

    def __init__(self, name, num_buffers, numel, dtype, track_usage):
        self.num_buffers = num_buffers
        self.buffers = [allocate_mem_buff(name + " {}".format(i), numel, dtype, track_usage)
                        for i in range(num_buffers)]
        self._index = -1

    def get_next_buffer(self):
        self._index = (self._index + 1) % self.num_buffers
        buff = self.buffers[self._index]
        assert not buff.is_in_use(), "buffer is already in use."
        return buff

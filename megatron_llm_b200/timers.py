"""Named timers.

Parity target: megatron/timers.py (:123-304 ``Timers`` with log levels 0-2, min/max/all aggregation across
ranks, TensorBoard export).  The reference synchronises the device on every start/stop and uses host
``time.time()`` (timers.py:79-96), which perturbs compute/communication overlap at log level 2.  Here a timer
records a pair of CUDA events on the current stream (no sync) plus an NVTX range; elapsed time is resolved
lazily when the timer is read (log interval), and falls back to ``perf_counter`` on CPU.
"""
from __future__ import annotations

import time
from abc import ABC, abstractmethod

import torch
import torch.distributed as dist

from .utils.device import current_device, use_cuda


class TimerBase(ABC):
    def __init__(self, name):
        self.name = name

    @abstractmethod
    def start(self, barrier=False): ...

    @abstractmethod
    def stop(self, barrier=False): ...

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def elapsed(self, reset=True, barrier=False): ...


class DummyTimer(TimerBase):
    def __init__(self):
        super().__init__("dummy timer")

    def start(self, barrier=False):
        return

    def stop(self, barrier=False):
        return

    def reset(self):
        return

    def elapsed(self, reset=True, barrier=False):
        raise Exception("dummy timer should not be used to calculate elapsed time")


class Timer(TimerBase):
    def __init__(self, name):
        super().__init__(name)
        self._elapsed = 0.0
        self._started = False
        self._pending = []  # (start_event, stop_event) pairs not yet resolved
        self._start_event = None
        self._start_time = 0.0
        self._barrier_group = None
        self._nvtx = False

    def set_barrier_group(self, group):
        self._barrier_group = group

    def start(self, barrier=False):
        assert not self._started, "timer has already been started"
        if barrier and dist.is_initialized():
            dist.barrier(group=self._barrier_group)
        if use_cuda():
            self._start_event = torch.cuda.Event(enable_timing=True)
            self._start_event.record()
            try:
                torch.cuda.nvtx.range_push(self.name)
                self._nvtx = True
            except Exception:
                self._nvtx = False
        else:
            self._start_time = time.perf_counter()
        self._started = True

    def stop(self, barrier=False):
        assert self._started, "timer is not started"
        if barrier and dist.is_initialized():
            dist.barrier(group=self._barrier_group)
        if use_cuda():
            if self._nvtx:
                torch.cuda.nvtx.range_pop()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._pending.append((self._start_event, ev))
        else:
            self._elapsed += time.perf_counter() - self._start_time
        self._started = False

    def _resolve(self):
        if self._pending:
            self._pending[-1][1].synchronize()
            for s, e in self._pending:
                self._elapsed += s.elapsed_time(e) / 1000.0
            self._pending = []

    def reset(self):
        self._elapsed = 0.0
        self._pending = []
        self._started = False

    def elapsed(self, reset=True, barrier=False):
        started = self._started
        if started:
            self.stop(barrier=barrier)
        self._resolve()
        val = self._elapsed
        if reset:
            self.reset()
        if started:
            self.start(barrier=barrier)
        return val


class Timers:
    """Group of timers; ``timers(name, log_level)`` returns a no-op timer above the configured level."""

    def __init__(self, log_level, log_option):
        self._log_level = log_level
        self._log_option = log_option
        self._timers = {}
        self._log_levels = {}
        self._dummy_timer = DummyTimer()
        self._max_log_level = 2

    def __call__(self, name, log_level=None):
        if name in self._timers:
            if log_level is not None:
                assert log_level == self._log_levels[name], \
                    "input log level {} does not match already existing log level {} for {} timer".format(
                        log_level, self._log_levels[name], name)
            return self._timers[name]
        if log_level is None:
            log_level = self._max_log_level
        assert log_level <= self._max_log_level, \
            "log level {} is larger than max supported log level {}".format(log_level, self._max_log_level)
        if log_level > self._log_level:
            return self._dummy_timer
        self._timers[name] = Timer(name)
        self._log_levels[name] = log_level
        return self._timers[name]

    def _get_elapsed_time_all_ranks(self, names, reset, barrier):
        if barrier and dist.is_initialized():
            dist.barrier()
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        mine = torch.zeros(len(names), dtype=torch.float, device=current_device())
        for i, name in enumerate(names):
            if name in self._timers:
                mine[i] = self._timers[name].elapsed(reset=reset)
        if world == 1:
            return mine.view(1, -1)
        allt = torch.zeros((world, len(names)), dtype=torch.float, device=current_device())
        dist.all_gather_into_tensor(allt.view(-1), mine)
        return allt

    def _get_global_min_max_time(self, names, reset, barrier, normalizer):
        t = self._get_elapsed_time_all_ranks(names, reset, barrier)
        out = {}
        for i, name in enumerate(names):
            col = t[:, i]
            col = col[col > 0.0]
            if col.numel() > 0:
                out[name] = (col.min().item() / normalizer, col.max().item() / normalizer)
        return out

    def _get_global_min_max_time_string(self, names, reset, barrier, normalizer, max_only):
        mm = self._get_global_min_max_time(names, reset, barrier, normalizer)
        if not mm:
            return None
        s = "(min, max) time across ranks (ms):" if not max_only else "max time across ranks (ms):"
        for name, (mn, mx) in mm.items():
            label = (name + " ").ljust(48, ".")
            s += "\n    {}: {:.2f}".format(label, mx * 1000.0) if max_only else \
                "\n    {}: ({:.2f}, {:.2f})".format(label, mn * 1000.0, mx * 1000.0)
        return s

    def _get_all_ranks_time_string(self, names, reset, barrier, normalizer):
        t = self._get_elapsed_time_all_ranks(names, reset, barrier)
        s = "times across ranks (ms):"
        any_ = False
        for i, name in enumerate(names):
            header = False
            for r in range(t.size(0)):
                if t[r, i] > 0:
                    any_ = True
                    if not header:
                        header = True
                        s += "\n  {}:".format(name)
                    s += "\n     rank {:2d}: {:.2f}".format(r, t[r, i].item() / normalizer * 1000.0)
        return s if any_ else None

    def log(self, names, rank=None, normalizer=1.0, reset=True, barrier=False):
        assert normalizer > 0.0
        if self._log_option in ("max", "minmax"):
            s = self._get_global_min_max_time_string(names, reset, barrier, normalizer, self._log_option == "max")
        elif self._log_option == "all":
            s = self._get_all_ranks_time_string(names, reset, barrier, normalizer)
        else:
            raise Exception("unknown timing log option {}".format(self._log_option))
        if rank is None:
            rank = (dist.get_world_size() - 1) if dist.is_initialized() else 0
        me = dist.get_rank() if dist.is_initialized() else 0
        if rank == me and s is not None:
            print(s, flush=True)

    def write(self, names, writer, iteration, normalizer=1.0, reset=False, barrier=False):
        assert normalizer > 0.0
        mm = self._get_global_min_max_time(names, reset, barrier, normalizer)
        if writer is not None:
            for name, (_, mx) in mm.items():
                writer.add_scalar(name + "-time", mx, iteration)

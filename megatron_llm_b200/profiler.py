"""Step-range profiling for the training loop (the reference has only host-timed ``Timers``).

``--profile`` wraps iterations ``[--profile_step_start, --profile_step_end)`` of ``--profile_ranks`` in

* a ``torch.profiler`` trace (CPU + CUDA activities, chrome trace + a kernel table written to ``--profile_dir``), and
* ``cudaProfilerStart/Stop`` plus one NVTX range per iteration, so ``ncu --profile-from-start off`` / ``nsys
  --capture-range=cudaProfilerApi`` capture exactly those iterations.

Numbers taken while this is active are for analysis only; the benchmark (bench.py) never runs under it.
"""
from __future__ import annotations

import os
from typing import Optional

import torch


class StepProfiler:
    def __init__(self, args, rank: int):
        self.enabled = bool(getattr(args, "profile", False)) and rank in set(getattr(args, "profile_ranks", [0]))
        self.start = int(getattr(args, "profile_step_start", 10))
        self.end = int(getattr(args, "profile_step_end", 12))
        self.dir = getattr(args, "profile_dir", None) or "profiles/trace"
        self.rank = rank
        self.cuda = torch.cuda.is_available() and os.environ.get("MLB200_FORCE_CPU", "0") != "1"
        self._prof = None
        self._in_step = False
        assert self.end > self.start, "--profile_step_end must be greater than --profile_step_start"

    # ------------------------------------------------------------------ hooks called by the training loop
    def step_begin(self, iteration: int) -> None:
        if not self.enabled:
            return
        if iteration == self.start and self._prof is None:
            acts = [torch.profiler.ProfilerActivity.CPU]
            if self.cuda:
                acts.append(torch.profiler.ProfilerActivity.CUDA)
                torch.cuda.synchronize()
                torch.cuda.cudart().cudaProfilerStart()
            self._prof = torch.profiler.profile(activities=acts, record_shapes=False)
            self._prof.__enter__()
        if self._prof is not None and self.cuda:
            torch.cuda.nvtx.range_push(f"iteration {iteration}")
            self._in_step = True

    def step_end(self, iteration: int) -> Optional[str]:
        """Call after iteration ``iteration`` finished; returns the trace path when the range just closed."""
        if not self.enabled or self._prof is None:
            return None
        if self._in_step:
            torch.cuda.nvtx.range_pop()
            self._in_step = False
        if iteration + 1 < self.end:
            return None
        if self.cuda:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        self._prof.__exit__(None, None, None)
        prof, self._prof = self._prof, None
        self.enabled = False
        os.makedirs(self.dir, exist_ok=True)
        trace = os.path.join(self.dir, f"trace_rank{self.rank}_it{self.start}-{self.end}.json")
        prof.export_chrome_trace(trace)
        sort_key = "self_cuda_time_total" if self.cuda else "self_cpu_time_total"
        with open(os.path.join(self.dir, f"kernels_rank{self.rank}_it{self.start}-{self.end}.txt"), "w") as f:
            f.write(prof.key_averages().table(sort_by=sort_key, row_limit=60))
        return trace

"""ZeRO-1 distributed optimizer (parity: megatron/optimizer/distrib_optimizer.py).

Reference behaviour kept: the grad buffer is padded to a DP multiple and each DP rank owns a contiguous slice
that ignores parameter boundaries (:119-164); grads are reduce-scattered (:553-567), the fp32 shard is stepped,
and the updated weights are all-gathered (:592-600).  Differences: slices are per *bucket* so the
reduce-scatter of a bucket can overlap backward; the bf16 weights live in a flat buffer with the same offsets,
so the AdamW kernel writes the updated bf16 slice in place and the all-gather runs in place on that buffer (the
reference aliases the param buffer onto the grad buffer storage and copies back per tensor, :380-389,603-608).
Optimizer state is saved per DP rank as flat shards.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..parallel import state as ps
from .optimizer import MixedPrecisionOptimizer


class Range:
    """[start, end) of a flat buffer (reference distrib_optimizer.py:15-29)."""

    def __init__(self, start, end):
        self.start, self.end, self.size = start, end, end - start

    def normalize(self, start=0):
        return Range(start, start + self.size)

    def __str__(self):
        return "%d,%d [%d]" % (self.start, self.end, self.size)


class DistributedOptimizer(MixedPrecisionOptimizer):
    def __init__(self, optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                 use_contiguous_buffers_in_local_ddp, fp16, bf16, params_dtype, grad_scaler, models):
        assert use_contiguous_buffers_in_local_ddp
        super().__init__(optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                         use_contiguous_buffers_in_local_ddp, fp16, bf16, params_dtype, grad_scaler, models,
                         shard_over_dp=True)
        # fused cast + all-gather: when the DDP wrapper's 16-bit parameter buffer lives in symmetric memory, the AdamW
        # kernel stores the updated shard into every DP peer's buffer and gather_model_params() is a barrier
        self._fused_gather = []
        for g in self.groups:
            comm = getattr(g.ddp, "_symm", None)
            if (comm is not None and getattr(comm, "pbuf", None) is not None and g.model_param_buffer is not None
                    and g.model_param_buffer.data_ptr() == comm.pbuf.data_ptr() and not g.is_fp32_model
                    and self.config["name"] == "adam"):
                g.p16_peer_ptrs = comm.param_peer_ptrs(g.shard[0])
                if comm not in self._fused_gather:
                    self._fused_gather.append(comm)

    def shard_ranges(self):
        """This DP rank's slice of each flat group's gradient / parameter buffer."""
        return [Range(*g.shard) for g in self.groups]

    def _norm_reduce_group(self):
        """Every (tp, pp, dp) rank holds a distinct slice of the gradients: reduce over the whole world."""
        return dist.group.WORLD if dist.is_initialized() else None

    def get_model_parallel_group(self):
        return None

    def reduce_model_grads(self, args, timers):
        """SP norm grads (TP all-reduce) -> DP reduce-scatter (per bucket, possibly already overlapped with
        backward) -> tied-embedding all-reduce."""
        timers("layernorm-grads-all-reduce", log_level=1).start(barrier=args.barrier_with_L1_time)
        self.allreduce_layernorm_grads(args)
        timers("layernorm-grads-all-reduce").stop()
        timers("embedding-grads-all-reduce", log_level=1).start(barrier=args.barrier_with_L1_time)
        self.allreduce_embedding_grads(args)
        timers("embedding-grads-all-reduce").stop()
        timers("grads-reduce-scatter", log_level=1).start(barrier=args.barrier_with_L1_time)
        for model in self.models:
            model.allreduce_gradients()
        timers("grads-reduce-scatter").stop()

    def gather_model_params(self, args, timers):
        """All-gather the updated weight slices in place (bf16 buffer; fp32 buffer for fp32 models)."""
        timers("params-all-gather", log_level=1).start(barrier=args.barrier_with_L1_time)
        w = ps.get_data_parallel_world_size()
        for comm in self._fused_gather:
            comm.params_barrier()
        if w > 1:
            group = ps.get_data_parallel_group()
            for g in self.groups:
                if g.p16_peer_ptrs:
                    continue              # already stored into every peer's buffer by the optimizer kernel
                buf = g.model_param_buffer
                s, e = g.shard
                n = e - s
                r = ps.get_data_parallel_rank()
                region = buf[s - r * n: s - r * n + w * n]
                if buf.is_cuda:
                    dist.all_gather_into_tensor(region, buf[s:e], group=group)
                else:
                    parts = [torch.empty(n, dtype=buf.dtype) for _ in range(w)]
                    dist.all_gather(parts, buf[s:e].clone(), group=group)
                    region.copy_(torch.cat(parts))
        timers("params-all-gather").stop()

    # ---- per-DP-rank flat state -------------------------------------------------------------------
    def state_dict(self):
        sd = {"step_count": self.step_count,
              "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
              "shards": [{"range": g.shard, "main_param": g.main_param.clone(), "exp_avg": g.exp_avg.clone(),
                          "exp_avg_sq": g.exp_avg_sq.clone()} for g in self.groups]}
        if self.grad_scaler:
            sd["grad_scaler"] = self.grad_scaler.state_dict()
        return sd

    def load_state_dict(self, state_dict):
        self.step_count = state_dict["step_count"]
        for grp, saved in zip(self.param_groups, state_dict["param_groups"]):
            grp.update(saved)
        for g, sh in zip(self.groups, state_dict["shards"]):
            assert tuple(sh["range"]) == tuple(g.shard), "optimizer shard layout changed (different DP size?)"
            g.main_param.copy_(sh["main_param"])
            g.exp_avg.copy_(sh["exp_avg"])
            g.exp_avg_sq.copy_(sh["exp_avg_sq"])
        if self.grad_scaler and "grad_scaler" in state_dict:
            self.grad_scaler.load_state_dict(state_dict["grad_scaler"])

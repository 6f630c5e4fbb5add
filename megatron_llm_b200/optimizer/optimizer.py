"""Mixed-precision optimizers over flat buffers.

Parity target: megatron/optimizer/optimizer.py -- MegatronOptimizer :58-301 (reduce_model_grads incl. SP
layernorm-grad all-reduce :257-277 and tied-embedding grad all-reduce :203-254), MixedPrecisionOptimizer.step
:407-466, Float16OptimizerWithFloat16Params :469-695, FP32Optimizer :698-783.

Design: the DDP wrapper already keeps (bf16 weights | fp32 main_grads) in contiguous buffers with identical
offsets.  The optimizer adds three more flat fp32 buffers (master weights, exp_avg, exp_avg_sq) per grad
buffer, so one optimizer step is:
    sqnorm kernel per buffer -> (all-reduce of ONE scalar over the model-parallel group) -> clip_coef kernel
    -> ONE AdamW kernel per buffer that reads the clip coefficient from device memory and writes the master
       weights, both moments and the bf16 model weights.
No host synchronisation happens on the bf16 path (the reference syncs for the norm and, in fp16, the inf check).
On CPU the same algorithm runs with torch ops (oracle for the kernels).
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from typing import Dict, List

import torch
import torch.distributed as dist

from .. import ops
from ..models.module import param_is_not_shared
from ..parallel import state as ps
from ..parallel.layers import param_is_not_tensor_parallel_duplicate
from ..utils.device import current_device
from ..utils import unwrap_model
from .clip_grads import clip_grad_norm_fp32, count_zeros_fp32


def _get_args():
    from ..global_vars import get_args
    return get_args()


class _FlatGroup:
    """One (model chunk, grad dtype) buffer: segment table + fp32 state, for the full buffer or a DP shard."""

    def __init__(self, ddp, gdt, param_group_of, shard=None):
        self.ddp, self.gdt = ddp, gdt
        self.grad_buffer = ddp.grad_buffers()[gdt]
        self.index_map = ddp.param_index_maps()[gdt]
        pbufs = ddp.param_buffers()[gdt]
        # params sorted by offset
        self.params = sorted(self.index_map.keys(), key=lambda p: self.index_map[p][0])
        assert len(pbufs) <= 1, "mixed parameter dtypes inside one grad buffer are not supported"
        self.param_dtype = next(iter(pbufs.keys())) if pbufs else None
        self.numel = self.grad_buffer.numel_padded
        self.shard = shard if shard is not None else (0, self.numel)   # [start, end) owned by this rank
        dev = self.grad_buffer.data.device
        starts = [self.index_map[p][0] for p in self.params] + [self.numel]
        # segment k covers [start_k, start_{k+1}) (alignment padding belongs to the preceding param)
        self.seg_start = torch.tensor(starts, dtype=torch.int64, device=dev)
        self.param_group_of = param_group_of
        self.seg_wd_mult = torch.tensor([param_group_of(p)["wd_mult"] for p in self.params], dtype=torch.float32,
                                        device=dev)
        self.seg_lr_mult = torch.tensor([param_group_of(p)["lr_mult"] for p in self.params], dtype=torch.float32,
                                        device=dev)
        tp_rank0 = ps.get_tensor_model_parallel_rank() == 0
        w = []
        for p in self.params:
            keep = param_is_not_shared(p) and (getattr(p, "tensor_model_parallel", False) or tp_rank0)
            w.append(1.0 if keep else 0.0)
        self.seg_norm_weight = torch.tensor(w, dtype=torch.float32, device=dev)
        s, e = self.shard
        n = e - s
        self.is_fp32_model = self.model_param_buffer is not None and self.model_param_buffer.dtype == torch.float32
        self.p16_peer_ptrs = []      # ZeRO-1 fused param gather: every DP peer's 16-bit buffer at this shard (optional)
        if self.is_fp32_model:
            self.main_param = self.model_param_buffer[s:e]            # optimise the weights in place
        else:
            self.main_param = self.model_param_buffer[s:e].float() if self.model_param_buffer is not None else \
                torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.workspace = torch.zeros(148 * 8, dtype=torch.float32, device=dev)

    @property
    def model_param_buffer(self):
        """The DDP wrapper's flat parameter buffer (looked up on use: it may be re-homed into symmetric memory)."""
        return self.ddp.param_buffers()[self.gdt].get(self.param_dtype) if self.param_dtype is not None else None

    def main_grad(self):
        s, e = self.shard
        return self.grad_buffer.data[s:e]

    def model_shard(self):
        s, e = self.shard
        return None if (self.model_param_buffer is None or self.is_fp32_model) else self.model_param_buffer[s:e]

    def reload_main_params_from_model(self):
        if not self.is_fp32_model and self.model_param_buffer is not None:
            s, e = self.shard
            self.main_param.copy_(self.model_param_buffer[s:e])


class MegatronOptimizer(ABC):
    def __init__(self, optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                 use_contiguous_buffers_in_local_ddp, models):
        self.config = optimizer_config          # dict: optimizer name + hyper-parameters
        self.clip_grad = clip_grad
        self.log_num_zeros_in_grad = log_num_zeros_in_grad
        self.params_have_main_grad = params_have_main_grad
        self.use_contiguous_buffers_in_local_ddp = use_contiguous_buffers_in_local_ddp
        self.models = models
        self.param_groups: List[dict] = optimizer_config["param_groups"]

    # ---- helpers shared with the reference API ----------------------------------------------
    def get_parameters(self):
        return [p for g in self.param_groups for p in g["params"]]

    def get_main_grads_for_grad_norm(self):
        grads = []
        for p in self.get_parameters():
            g = getattr(p, "main_grad", None)
            if g is None:
                g = p.grad
            if g is not None and param_is_not_shared(p) and param_is_not_tensor_parallel_duplicate(p):
                grads.append(g)
        return grads

    def get_model_parallel_group(self):
        return ps.get_model_parallel_group()

    def clip_grad_norm(self, clip_grad):
        return clip_grad_norm_fp32(self.get_parameters(), self.get_main_grads_for_grad_norm(), clip_grad,
                                   model_parallel_group=self.get_model_parallel_group())

    def count_zeros(self):
        return count_zeros_fp32(self.get_parameters(), model_parallel_group=self.get_model_parallel_group())

    @abstractmethod
    def zero_grad(self, set_to_none=True): ...

    @abstractmethod
    def get_loss_scale(self): ...

    def scale_loss(self, loss):
        return self.get_loss_scale() * loss

    @abstractmethod
    def reload_model_params(self): ...

    @abstractmethod
    def state_dict(self): ...

    @abstractmethod
    def load_state_dict(self, state_dict): ...

    @abstractmethod
    def step(self, args, timers): ...

    def gather_model_params(self, args, timers):
        """Only the distributed optimizer has work to do here."""
        pass

    # ---- gradient reductions that are not the DP reduction ----------------------------------
    def allreduce_word_embedding_grads(self, args):
        """Tied embeddings with PP>1: sum the first-stage and last-stage copies' grads."""
        if ps.is_rank_in_embedding_group(ignore_virtual=True) and ps.get_pipeline_model_parallel_world_size() > 1:
            if ps.is_pipeline_first_stage(ignore_virtual=True):
                unwrapped = self.models[0]
            elif ps.is_pipeline_last_stage(ignore_virtual=True):
                unwrapped = self.models[-1]
            else:
                unwrapped = self.models[0]
            unwrapped = unwrap_model(unwrapped)
            if getattr(unwrapped, "share_word_embeddings", False):
                w = unwrapped.word_embeddings_weight()
                grad = w.main_grad if hasattr(w, "main_grad") else w.grad
                dist.all_reduce(grad, group=ps.get_embedding_group())

    def allreduce_position_embedding_grads(self, args):
        if ps.is_rank_in_position_embedding_group() and ps.get_pipeline_model_parallel_world_size() > 1 and \
                args.pipeline_model_parallel_split_rank is not None:
            unwrapped = unwrap_model(self.models[0])
            assert args.DDP_impl == "local", "T5 model is only supported with local DDP mode"
            grad = unwrapped.language_model.embedding.position_embeddings.weight.main_grad
            dist.all_reduce(grad, group=ps.get_position_embedding_group())

    def allreduce_embedding_grads(self, args):
        self.allreduce_word_embedding_grads(args)
        self.allreduce_position_embedding_grads(args)

    def allreduce_layernorm_grads(self, args):
        """Sequence parallelism: norm weights (and Row biases) see only s/tp tokens per rank -> sum over TP."""
        if ps.get_tensor_model_parallel_world_size() > 1 and args.sequence_parallel:
            grads = []
            for m in self.models:
                for p in unwrap_model(m).parameters():
                    if getattr(p, "sequence_parallel", False):
                        g = p.main_grad if hasattr(p, "main_grad") else p.grad
                        if g is not None:
                            grads.append(g.data)
            if grads:
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat, group=ps.get_tensor_model_parallel_group())
                off = 0
                for g in grads:
                    n = g.numel()
                    g.copy_(flat[off:off + n].view_as(g))
                    off += n

    def reduce_model_grads(self, args, timers):
        """All-reduce all grads, and all-reduce embeddings."""
        timers("layernorm-grads-all-reduce", log_level=1).start(barrier=args.barrier_with_L1_time)
        self.allreduce_layernorm_grads(args)
        timers("layernorm-grads-all-reduce").stop()
        if args.DDP_impl == "local":
            timers("grads-all-reduce", log_level=1).start(barrier=args.barrier_with_L1_time)
            for model in self.models:
                model.allreduce_gradients()
            timers("grads-all-reduce").stop()
        timers("embedding-grads-all-reduce", log_level=1).start(barrier=args.barrier_with_L1_time)
        self.allreduce_embedding_grads(args)
        timers("embedding-grads-all-reduce").stop()


class FlatOptimizer(MegatronOptimizer):
    """Shared implementation of Float16OptimizerWithFloat16Params / FP32Optimizer / DistributedOptimizer."""

    def __init__(self, optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                 use_contiguous_buffers_in_local_ddp, fp16, bf16, params_dtype, grad_scaler, models,
                 shard_over_dp: bool = False):
        super().__init__(optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                         use_contiguous_buffers_in_local_ddp, models)
        self.fp16, self.bf16, self.params_dtype = fp16, bf16, params_dtype
        self.grad_scaler = grad_scaler
        if self.grad_scaler is None:
            assert not self.fp16, "fp16 expects a grad scaler."
        self.shard_over_dp = shard_over_dp
        dev = current_device()
        self.found_inf = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scale_one = torch.ones(1, dtype=torch.float32, device=dev)
        self._total_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._clip_coef = torch.ones(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                group_of[p] = g
        self._group_of = group_of
        self.groups: List[_FlatGroup] = []
        for m in models:
            ddp = m
            assert hasattr(ddp, "grad_buffers"), \
                "the flat optimizers need the local DDP wrapper (contiguous param/grad buffers)"
            for gdt in ddp.grad_buffers().keys():
                if shard_over_dp:
                    # ZeRO-1: this rank owns the r-th 1/dp slice of every bucket
                    w, r = ps.get_data_parallel_world_size(), ps.get_data_parallel_rank()
                    for b in ddp.buckets()[gdt]:
                        n = (b.end - b.start) // w
                        self.groups.append(_FlatGroup(ddp, gdt, lambda p: self._group_of[p],
                                                      (b.start + r * n, b.start + (r + 1) * n)))
                else:
                    self.groups.append(_FlatGroup(ddp, gdt, lambda p: self._group_of[p], None))

    # ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """Gradients live in the DDP buffers (zeroed by ``zero_grad_buffer``); drop stray .grad tensors."""
        for p in self.get_parameters():
            if p.grad is not None:
                p.grad = None

    def get_loss_scale(self):
        if self.grad_scaler is None:
            return self._scale_one
        return self.grad_scaler.scale

    def reload_model_params(self):
        for g in self.groups:
            g.reload_main_params_from_model()

    def _hyper(self):
        """Base (un-multiplied) lr / weight decay recovered from the param groups the scheduler writes into."""
        lr, wd = 0.0, 0.0
        for grp in self.param_groups:
            lm, wm = grp.get("lr_mult", 1.0), grp.get("wd_mult", 1.0)
            if lm != 0.0:
                lr = grp["lr"] / lm
            if wm != 0.0:
                wd = grp["weight_decay"] / wm
        return {"lr": lr, "weight_decay": wd}

    # ------------------------------------------------------------------------------------------
    def _grad_sq_norm(self):
        """sum of squares of the (unique) grads of this rank's buffers -> self._total_sq (device)."""
        first = True
        for g in self.groups:
            grad = g.main_grad()
            off = g.shard[0]
            if ops.cuda_ops_available(grad):
                ops._C().sqnorm_flat(grad, off, g.seg_start, g.seg_norm_weight, g.workspace, self._total_sq,
                                     not first)
                ops._count(2)
            else:
                w = self._expand_segments(g, g.seg_norm_weight)
                val = (w * grad.float().pow(2)).sum()
                if first:
                    self._total_sq.copy_(val.view(1))
                else:
                    self._total_sq.add_(val)
            first = False
        if first:
            self._total_sq.zero_()

    @staticmethod
    def _expand_segments(g: _FlatGroup, seg_values: torch.Tensor) -> torch.Tensor:
        """per-element copy of a per-segment value for this group's shard (CPU / oracle path)."""
        s, e = g.shard
        lens = (g.seg_start[1:] - g.seg_start[:-1])
        full = torch.repeat_interleave(seg_values, lens)
        return full[s:e]

    def step(self, args, timers):
        timers("optimizer-copy-to-main-grad", log_level=1).start(barrier=args.barrier_with_L1_time)
        timers("optimizer-copy-to-main-grad").stop()   # main_grads ARE the fp32 grads: nothing to copy

        # ---- global grad norm (+ inf check) and clip coefficient, all on device ----
        timers("optimizer-clip-main-grad", log_level=1).start(barrier=args.barrier_with_L1_time)
        self._grad_sq_norm()
        group = self._norm_reduce_group()
        if group is not None and dist.get_world_size(group=group) > 1:
            dist.all_reduce(self._total_sq, op=dist.ReduceOp.SUM, group=group)
        inv_scale = 1.0
        if self.grad_scaler is not None:
            inv_scale = 1.0 / float(self.grad_scaler.scale.item())
        if ops.cuda_ops_available(self._total_sq):
            ops._C().clip_coef(self._total_sq, float(self.clip_grad or 0.0), self._grad_norm, self._clip_coef,
                               self.found_inf, inv_scale)
            ops._count()
        else:
            nrm = self._total_sq.sqrt() * inv_scale
            self._grad_norm.copy_(nrm)
            coef = torch.clamp(self.clip_grad / (nrm + 1.0e-6), max=1.0) if (self.clip_grad or 0.0) > 0 \
                else torch.ones_like(nrm)
            self._clip_coef.copy_(coef * inv_scale)
            self.found_inf.copy_((~torch.isfinite(nrm)).to(torch.int32))
        timers("optimizer-clip-main-grad").stop()

        # ---- fp16: dynamic loss scale needs the flag on the host (bf16 path has no sync) ----
        if self.grad_scaler is not None:
            timers("optimizer-unscale-and-check-inf", log_level=1).start(barrier=args.barrier_with_L1_time)
            found = bool(self.found_inf.item())
            self.grad_scaler.update(found)
            timers("optimizer-unscale-and-check-inf").stop()
            if found:
                return False, None, None

        num_zeros_in_grad = None
        if self.log_num_zeros_in_grad:
            timers("optimizer-count-zeros", log_level=1).start(barrier=args.barrier_with_L1_time)
            num_zeros_in_grad = self.count_zeros()
            timers("optimizer-count-zeros").stop()

        # ---- the update ----
        timers("optimizer-inner-step", log_level=1).start(barrier=args.barrier_with_L1_time)
        self.step_count += 1
        self._inner_step()
        timers("optimizer-inner-step").stop()

        timers("optimizer-copy-main-to-model-params", log_level=1).start(barrier=args.barrier_with_L1_time)
        timers("optimizer-copy-main-to-model-params").stop()   # fused into the AdamW kernel's bf16 write-back
        return True, self._grad_norm, num_zeros_in_grad

    def _norm_reduce_group(self):
        return ps.get_model_parallel_group() if ps.model_parallel_is_initialized() else None

    def _inner_step(self):
        h = self._hyper()
        name = self.config["name"]
        lr, wd = h["lr"], h["weight_decay"]
        for g in self.groups:
            # per-segment weight decay = group wd * wd_mult ; lr multiplier per segment
            seg_wd = g.seg_wd_mult * wd
            grad, p16 = g.main_grad(), g.model_shard()
            off = g.shard[0]
            if name == "adam":
                b1, b2 = self.config["betas"]
                eps = self.config["eps"]
                bc1 = 1.0 - b1 ** self.step_count
                bc2 = 1.0 - b2 ** self.step_count
                if ops.cuda_ops_available(grad):
                    ops._C().adamw_flat(g.main_param, grad, g.exp_avg, g.exp_avg_sq, p16, off, g.seg_start, seg_wd,
                                        g.seg_lr_mult, lr, b1, b2, eps, bc1, bc2, self._clip_coef, self.found_inf,
                                        g.p16_peer_ptrs if p16 is not None else [])
                    ops._count()
                else:
                    skip = bool(self.found_inf.item())
                    if not skip:
                        gg = grad.float() * self._clip_coef
                        wd_e = self._expand_segments(g, seg_wd)
                        lr_e = self._expand_segments(g, g.seg_lr_mult) * lr
                        g.exp_avg.mul_(b1).add_(gg, alpha=1 - b1)
                        g.exp_avg_sq.mul_(b2).addcmul_(gg, gg, value=1 - b2)
                        upd = (g.exp_avg / bc1) / ((g.exp_avg_sq / bc2).sqrt() + eps) + wd_e * g.main_param
                        g.main_param.sub_(lr_e * upd)
                        if p16 is not None:
                            p16.copy_(g.main_param)
            elif name == "sgd":
                mom = self.config["momentum"]
                if ops.cuda_ops_available(grad):
                    ops._C().sgd_flat(g.main_param, grad, g.exp_avg, p16, off, g.seg_start, seg_wd, g.seg_lr_mult, lr, mom,
                                      self.step_count == 1, self._clip_coef, self.found_inf)
                    ops._count()
                else:
                    if not bool(self.found_inf.item()):
                        gg = grad.float() * self._clip_coef + self._expand_segments(g, seg_wd) * g.main_param
                        if mom != 0.0:
                            if self.step_count == 1:
                                g.exp_avg.copy_(gg)
                            else:
                                g.exp_avg.mul_(mom).add_(gg)
                            gg = g.exp_avg
                        g.main_param.sub_(self._expand_segments(g, g.seg_lr_mult) * lr * gg)
                        if p16 is not None:
                            p16.copy_(g.main_param)
            else:
                raise Exception("{} optimizer is not supported.".format(name))

    # ------------------------------------------------------------------------------------------
    def _per_param_views(self, g: _FlatGroup, flat: torch.Tensor):
        """{param: view of ``flat``} for params fully inside this group's shard (full-buffer optimizers)."""
        s0 = g.shard[0]
        out = {}
        for p in g.params:
            s, e = g.index_map[p]
            if s >= g.shard[0] and e <= g.shard[1]:
                out[p] = flat[s - s0:e - s0].view(p.shape)
        return out

    def state_dict(self):
        """torch.optim-style layout so offline tools can reshard it: ``optimizer.state[i] = {exp_avg, exp_avg_sq}``
        indexed in param-group order, plus the fp32 master weights (``fp32_from_fp16_params``)."""
        order = self.get_parameters()
        idx = {p: i for i, p in enumerate(order)}
        state, masters = {}, {}
        for g in self.groups:
            ea, es, mp = (self._per_param_views(g, t) for t in (g.exp_avg, g.exp_avg_sq, g.main_param))
            for p in ea:
                state[idx[p]] = {"exp_avg": ea[p].clone(), "exp_avg_sq": es[p].clone(), "step": self.step_count}
                masters[idx[p]] = mp[p].clone()
        pgs, start = [], 0
        for grp in self.param_groups:
            d = {k: v for k, v in grp.items() if k != "params"}
            d["params"] = list(range(start, start + len(grp["params"])))
            start += len(grp["params"])
            pgs.append(d)
        sd = {"optimizer": {"state": state, "param_groups": pgs}, "step_count": self.step_count}
        if self.grad_scaler:
            sd["grad_scaler"] = self.grad_scaler.state_dict()
        if not all(g.is_fp32_model for g in self.groups):
            sd["fp32_from_fp16_params"] = [[masters[i] for i in pg["params"] if i in masters] for pg in pgs]
        return sd

    def load_state_dict(self, state_dict):
        opt = state_dict.get("optimizer", state_dict.get("optimizer_state_dict"))
        order = self.get_parameters()
        idx = {p: i for i, p in enumerate(order)}
        st = opt["state"]
        self.step_count = state_dict.get("step_count", 0)
        masters = None
        if "fp32_from_fp16_params" in state_dict:
            masters = [t for grp in state_dict["fp32_from_fp16_params"] for t in grp]
        for g in self.groups:
            ea, es, mp = (self._per_param_views(g, t) for t in (g.exp_avg, g.exp_avg_sq, g.main_param))
            for p in ea:
                i = idx[p]
                if i in st:
                    ea[p].copy_(st[i]["exp_avg"])
                    es[p].copy_(st[i]["exp_avg_sq"])
                    if not self.step_count:
                        self.step_count = int(st[i].get("step", 0))
                if masters is not None and not g.is_fp32_model and i < len(masters):
                    mp[p].copy_(masters[i])
        for grp, saved in zip(self.param_groups, opt["param_groups"]):
            for k, v in saved.items():
                if k != "params":
                    grp[k] = v
        if self.grad_scaler:
            if "grad_scaler" not in state_dict:
                if self.fp16:
                    print("***WARNING*** found an old checkpoint, will not load grad scaler ...")
            else:
                self.grad_scaler.load_state_dict(state_dict["grad_scaler"])
        if masters is None:
            self.reload_model_params()


class MixedPrecisionOptimizer(FlatOptimizer):
    """Optimizers with 16-bit model weights and fp32 master state (reference optimizer.py:281-457): the loss-scaling /
    unscale / found-inf logic lives in ``FlatOptimizer`` (it degenerates to a no-op for fp32 models), this class is
    the common base of the two mixed-precision optimizers so ``isinstance`` checks keep working."""


class Float16OptimizerWithFloat16Params(MixedPrecisionOptimizer):
    """fp16/bf16 model weights + fp32 master weights / moments (full copy on every DP rank)."""

    def __init__(self, optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                 use_contiguous_buffers_in_local_ddp, fp16, bf16, params_dtype, grad_scaler, models):
        super().__init__(optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                         use_contiguous_buffers_in_local_ddp, fp16, bf16, params_dtype, grad_scaler, models,
                         shard_over_dp=False)


class FP32Optimizer(FlatOptimizer):
    """fp32 model: the weights themselves are the master weights."""

    def __init__(self, optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                 use_contiguous_buffers_in_local_ddp, models):
        super().__init__(optimizer_config, clip_grad, log_num_zeros_in_grad, params_have_main_grad,
                         use_contiguous_buffers_in_local_ddp, False, False, torch.float32, None, models,
                         shard_over_dp=False)

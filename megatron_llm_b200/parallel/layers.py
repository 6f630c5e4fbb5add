"""Tensor-parallel layers: VocabParallelEmbedding, ColumnParallelLinear, RowParallelLinear.

Parity target: megatron/core/tensor_parallel/layers.py (:41-76 param attributes, :79-125 init,
:128-210 embedding, :213-407 the fused linear autograd function, :410-563 Column, :566-701 Row).

B200-first differences:
  * every GEMM (fwd ``X W^T``, dgrad ``dY W``, wgrad ``dY^T X``) is the hand-written tcgen05 kernel
    (``ops.gemm_*``); wgrad accumulates straight into the fp32 ``main_grad`` through the GEMM epilogue, which
    replaces both the apex ``fused_weight_gradient_mlp_cuda`` call (layers.py:298-307) and the DDP hook add;
  * under sequence parallelism the all-gather -> GEMM and GEMM -> reduce-scatter pairs are dispatched to the
    fused peer-memory kernels in :mod:`fused_tp` when a symmetric-memory communicator is bound to the TP group,
    else to the unfused NCCL/Gloo path below (also the oracle in tests);
  * the gathered activation is kept for backward (180 GB HBM) instead of being re-gathered (collective C2 of
    SURVEY.md 2.3 disappears), controlled by ``MLB200_KEEP_GATHERED``.
"""
from __future__ import annotations

import os
import warnings
from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F
import torch.nn.init as init
from torch.nn.parameter import Parameter

from .. import ops
from ..utils.core_utils import divide
from ..utils.device import current_device
from . import state as ps
from .mappings import (copy_to_tensor_model_parallel_region, gather_from_tensor_model_parallel_region,
                       reduce_from_tensor_model_parallel_region, reduce_scatter_to_sequence_parallel_region,
                       scatter_to_tensor_model_parallel_region)
from .random import get_cuda_rng_tracker
from .tp_utils import VocabUtility

_MODEL_PARALLEL_ATTRIBUTE_DEFAULTS = {"tensor_model_parallel": False, "partition_dim": -1,
                                      "partition_stride": 1}


def param_is_not_tensor_parallel_duplicate(param) -> bool:
    return (getattr(param, "tensor_model_parallel", False)
            or ps.get_tensor_model_parallel_rank() == 0)


def set_tensor_model_parallel_attributes(tensor, is_parallel, dim, stride):
    for attribute in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        assert not hasattr(tensor, attribute)
    setattr(tensor, "tensor_model_parallel", is_parallel)
    setattr(tensor, "partition_dim", dim)
    setattr(tensor, "partition_stride", stride)


def set_defaults_if_not_set_tensor_model_parallel_attributes(tensor):
    for attribute, default in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS.items():
        if not hasattr(tensor, attribute):
            setattr(tensor, attribute, default)


def copy_tensor_model_parallel_attributes(destination_tensor, source_tensor):
    for attribute in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        if hasattr(source_tensor, attribute):
            setattr(destination_tensor, attribute, getattr(source_tensor, attribute))


def _initialize_affine_weight_gpu(weight, init_method, partition_dim, stride=1):
    """Initialise this rank's shard directly on the device under the TP-forked RNG."""
    set_tensor_model_parallel_attributes(weight, True, partition_dim, stride)
    with get_cuda_rng_tracker().fork():
        init_method(weight)


def _initialize_affine_weight_cpu(weight, output_size, input_size, per_partition_size, partition_dim, init_method,
                                  stride=1, return_master_weight=False, *, params_dtype=torch.float32):
    """Build the full master weight on every rank and keep this rank's strided shard."""
    set_tensor_model_parallel_attributes(weight, True, partition_dim, stride)
    master = torch.empty(output_size, input_size, dtype=torch.float, requires_grad=False)
    init_method(master)
    master = master.to(dtype=params_dtype)
    per_stride = divide(per_partition_size, stride)
    pieces = torch.split(master, per_stride, dim=partition_dim)
    rank, world = ps.get_tensor_model_parallel_rank(), ps.get_tensor_model_parallel_world_size()
    with torch.no_grad():
        weight.copy_(torch.cat(pieces[rank::world], dim=partition_dim))
    return master if return_master_weight else None


class VocabParallelEmbedding(torch.nn.Module):
    """Embedding table split along the vocabulary dimension.

    Out-of-range ids are looked up at row 0, zeroed, and the partial results all-reduced over TP."""

    def __init__(self, num_embeddings: int, embedding_dim: int, *, init_method=init.xavier_normal_,
                 params_dtype: torch.dtype = torch.float32, use_cpu_initialization: bool = False,
                 perform_initialization: bool = True, gradient_accumulation_fusion: bool = False):
        super().__init__()
        self.gradient_accumulation_fusion = gradient_accumulation_fusion
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.padding_idx = None
        self.max_norm, self.norm_type, self.scale_grad_by_freq, self.sparse = None, 2.0, False, False
        self._embedding_weight = None
        self.tensor_model_parallel_size = ps.get_tensor_model_parallel_world_size()
        self.vocab_start_index, self.vocab_end_index = VocabUtility.vocab_range_from_global_vocab_size(
            num_embeddings, ps.get_tensor_model_parallel_rank(), self.tensor_model_parallel_size)
        self.num_embeddings_per_partition = self.vocab_end_index - self.vocab_start_index
        if use_cpu_initialization:
            self.weight = Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim, dtype=params_dtype))
            if perform_initialization:
                _initialize_affine_weight_cpu(self.weight, num_embeddings, embedding_dim,
                                              self.num_embeddings_per_partition, 0, init_method,
                                              params_dtype=params_dtype)
        else:
            self.weight = Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim,
                                                device=current_device(), dtype=params_dtype))
            if perform_initialization:
                _initialize_affine_weight_gpu(self.weight, init_method, partition_dim=0, stride=1)

    def forward(self, input_):
        """[b, s] ids -> [b, s, h], summed over the TP group (reference contract)."""
        if input_.dim() == 2:
            out = ops.embedding_lookup(input_, self.weight, self.vocab_start_index, sbh=False,
                                       accumulate_into_main_grad=self.gradient_accumulation_fusion)
            return reduce_from_tensor_model_parallel_region(out)
        if self.tensor_model_parallel_size > 1:
            mask = (input_ < self.vocab_start_index) | (input_ >= self.vocab_end_index)
            local = input_ - self.vocab_start_index
            local = local.masked_fill(mask, 0)
        else:
            local = input_
        out = F.embedding(local, self.weight, self.padding_idx, self.max_norm, self.norm_type,
                          self.scale_grad_by_freq, self.sparse)
        if self.tensor_model_parallel_size > 1:
            out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        return reduce_from_tensor_model_parallel_region(out)

    def forward_sbh(self, input_, sequence_parallel: bool):
        """[b, s] ids -> [s, b, h] (transpose folded into the gather kernel).  Under sequence parallelism the partial
        lookups are reduce-scattered along s (-> [s/tp, b, h]) instead of all-reduced and then split: 1/tp of the
        traffic and no scatter copy (reference: all-reduce at layers.py:208, scatter in language_model.py)."""
        out = ops.embedding_lookup(input_, self.weight, self.vocab_start_index, sbh=True,
                                   accumulate_into_main_grad=self.gradient_accumulation_fusion)
        if self.tensor_model_parallel_size == 1:
            return out
        if sequence_parallel:
            return reduce_scatter_to_sequence_parallel_region(out)
        return reduce_from_tensor_model_parallel_region(out)


# ------------------------------------------------------------------------------------------------
# the fused linear autograd function
# ------------------------------------------------------------------------------------------------

def _keep_gathered() -> bool:
    return os.environ.get("MLB200_KEEP_GATHERED", "1") == "1"


def _tp_group():
    return ps.get_tensor_model_parallel_group()


def _all_gather_first(x, async_op=False):
    world = ps.get_tensor_model_parallel_world_size()
    shape = list(x.shape)
    shape[0] *= world
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    h = dist.all_gather_into_tensor(out, x.contiguous(), group=_tp_group(), async_op=async_op)
    return out, h


def _reduce_scatter_first(x, async_op=False):
    world = ps.get_tensor_model_parallel_world_size()
    shape = list(x.shape)
    shape[0] //= world
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    h = dist.reduce_scatter_tensor(out, x.contiguous(), group=_tp_group(), async_op=async_op)
    return out, h


def _wgrad(grad_output2d, total_input2d, weight, gradient_accumulation_fusion):
    """dW = dY^T X.  With accumulation fusion the GEMM epilogue adds into fp32/bf16 ``weight.main_grad``."""
    if gradient_accumulation_fusion and getattr(weight, "main_grad", None) is not None:
        ops.gemm_tn(grad_output2d, total_input2d, out=weight.main_grad, accumulate=True)
        cb = getattr(weight, "_grad_ready_callback", None)
        if cb is not None:
            cb()  # tells the DDP bucket tracker this param's gradient is final (no autograd hook will fire)
        return None
    return ops.gemm_tn(grad_output2d, total_input2d).to(weight.dtype)


def _linear_fwd(x, weight):
    """x[..., K] @ weight[N, K]^T -> [..., N], written straight into a tensor of the final shape so the autograd
    Function returns a base tensor (not a view): downstream kernels (RoPE) may then update it in place."""
    out = torch.empty(*x.shape[:-1], weight.size(0), dtype=x.dtype, device=x.device)
    ops.gemm_nt(x.reshape(-1, x.size(-1)), weight, out=out.view(-1, weight.size(0)))
    return out


class LinearWithGradAccumulationAndAsyncCommunication(torch.autograd.Function):
    """y = x W^T (+b) with (a) SP all-gather of x in fwd, (b) dgrad overlapped with the SP re-gather / the
    TP all-reduce of dX, (c) dX reduce-scatter overlapped with wgrad, (d) wgrad accumulated into main_grad."""

    @staticmethod
    def forward(ctx, input, weight, bias, gradient_accumulation_fusion, async_grad_allreduce, sequence_parallel):
        ctx.use_bias = bias is not None
        ctx.gradient_accumulation_fusion = gradient_accumulation_fusion
        ctx.async_grad_allreduce = async_grad_allreduce
        ctx.sequence_parallel = sequence_parallel and ps.get_tensor_model_parallel_world_size() > 1
        from . import fused_tp
        fused = ctx.sequence_parallel and fused_tp.active_column(input, weight)
        ctx.fused = fused
        if ctx.sequence_parallel:
            if fused:
                output = torch.empty(input.size(0) * ps.get_tensor_model_parallel_world_size(), *input.shape[1:-1],
                                     weight.size(0), dtype=input.dtype, device=input.device)
                _, total_input = fused_tp.ag_gemm(input, weight, out=output.view(-1, weight.size(0)))
            else:
                total_input, _ = _all_gather_first(input)
                output = _linear_fwd(total_input, weight)
            if _keep_gathered():
                ctx.save_for_backward(total_input, weight)
                ctx.saved_gathered = True
            else:
                ctx.save_for_backward(input, weight)
                ctx.saved_gathered = False
        else:
            ctx.save_for_backward(input, weight)
            ctx.saved_gathered = True
            output = _linear_fwd(input, weight)
        if bias is not None:
            output = output + bias
        return output

    @staticmethod
    def backward(ctx, grad_output):
        saved_input, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        g2d = grad_output.reshape(-1, grad_output.size(-1))
        handle = None
        if ctx.sequence_parallel and not ctx.saved_gathered:
            total_input, handle = _all_gather_first(saved_input, async_op=True)
        else:
            total_input = saved_input
        grad_bias = g2d.sum(dim=0) if ctx.use_bias else None

        from . import fused_tp
        if ctx.sequence_parallel and ctx.fused and fused_tp.column_backward_fused():
            grad_input = fused_tp.gemm_rs(g2d, weight, transposed_weight=True).view(
                grad_output.size(0) // ps.get_tensor_model_parallel_world_size(), *grad_output.shape[1:-1],
                weight.size(1))
            if handle is not None:
                handle.wait()
            grad_weight = _wgrad(g2d, total_input.reshape(-1, total_input.size(-1)), weight,
                                 ctx.gradient_accumulation_fusion)
            return grad_input, grad_weight, grad_bias, None, None, None

        if (ctx.async_grad_allreduce and ps.get_tensor_model_parallel_world_size() > 1 and weight.dtype == g2d.dtype
                and fused_tp.active_all_reduce(g2d, weight.size(1))):
            # dX = all_reduce(dY @ W): one fused GEMM -> all-reduce kernel over peer memory (reference: cuBLAS GEMM +
            # async NCCL all-reduce overlapped with wgrad, layers.py:278-283)
            grad_input = fused_tp.gemm_ar(g2d, weight, transposed_weight=True).view(*grad_output.shape[:-1],
                                                                                     weight.size(1))
            grad_weight = _wgrad(g2d, total_input.reshape(-1, total_input.size(-1)), weight,
                                 ctx.gradient_accumulation_fusion)
            return grad_input, grad_weight, grad_bias, None, None, None
        grad_input = ops.gemm_nn(g2d, weight).view(*grad_output.shape[:-1], weight.size(1))
        if handle is not None:
            handle.wait()
        rs_out = None
        if ctx.async_grad_allreduce and ps.get_tensor_model_parallel_world_size() > 1:
            handle = dist.all_reduce(grad_input, group=_tp_group(), async_op=True)
        elif ctx.sequence_parallel:
            assert not ctx.async_grad_allreduce
            rs_out, handle = _reduce_scatter_first(grad_input, async_op=True)
        else:
            handle = None
        grad_weight = _wgrad(g2d, total_input.reshape(-1, total_input.size(-1)), weight,
                             ctx.gradient_accumulation_fusion)
        if handle is not None:
            handle.wait()
        if ctx.sequence_parallel:
            return rs_out, grad_weight, grad_bias, None, None, None
        return grad_input, grad_weight, grad_bias, None, None, None


def linear_with_grad_accumulation_and_async_allreduce(input: torch.Tensor, weight: torch.Tensor,
                                                      bias: Optional[torch.Tensor],
                                                      gradient_accumulation_fusion: bool,
                                                      async_grad_allreduce: bool,
                                                      sequence_parallel_enabled: bool) -> torch.Tensor:
    """Linear layer execution with asynchronous communication and gradient accumulation fusion in backprop
    (same contract as the reference, layers.py:320-407)."""
    args = [input, weight, bias, gradient_accumulation_fusion, async_grad_allreduce, sequence_parallel_enabled]
    with torch.amp.autocast(device_type="cuda" if input.is_cuda else "cpu", enabled=False):
        return LinearWithGradAccumulationAndAsyncCommunication.apply(*args)


class ColumnParallelLinear(torch.nn.Module):
    """Y = X A + b with A split along its output (column) dimension: weight shard is [out/tp, in].

    Returns ``(output, bias_if_skip_bias_add)``."""

    def __init__(self, input_size, output_size, *, bias=True, gather_output=True, init_method=init.xavier_normal_,
                 stride=1, keep_master_weight_for_test=False, skip_bias_add=False,
                 async_tensor_model_parallel_allreduce=True, params_dtype=torch.float32,
                 use_cpu_initialization=False, perform_initialization=True, gradient_accumulation_fusion=False,
                 sequence_parallel_enabled: bool = False, world_size: int = None):
        super().__init__()
        self.input_size, self.output_size, self.gather_output = input_size, output_size, gather_output
        world_size = world_size if world_size is not None else ps.get_tensor_model_parallel_world_size()
        self.world_size = world_size
        self.output_size_per_partition = divide(output_size, world_size)
        self.skip_bias_add = skip_bias_add
        if use_cpu_initialization:
            self.weight = Parameter(torch.empty(self.output_size_per_partition, input_size, dtype=params_dtype))
            if perform_initialization:
                self.master_weight = _initialize_affine_weight_cpu(
                    self.weight, output_size, input_size, self.output_size_per_partition, 0, init_method,
                    stride=stride, return_master_weight=keep_master_weight_for_test, params_dtype=params_dtype)
        else:
            self.weight = Parameter(torch.empty(self.output_size_per_partition, input_size,
                                                device=current_device(), dtype=params_dtype))
            if perform_initialization:
                _initialize_affine_weight_gpu(self.weight, init_method, partition_dim=0, stride=stride)
        if bias:
            dev = None if use_cpu_initialization else current_device()
            self.bias = Parameter(torch.zeros(self.output_size_per_partition, dtype=params_dtype, device=dev))
            set_tensor_model_parallel_attributes(self.bias, True, 0, stride)
        else:
            self.register_parameter("bias", None)
        self.async_tensor_model_parallel_allreduce = async_tensor_model_parallel_allreduce and world_size > 1
        if sequence_parallel_enabled and world_size <= 1:
            warnings.warn(f"`sequence_parallel_enabled` is set to `True`, but tensor model parallel size is "
                          f"{world_size}. Disabling sequence parallel.")
            sequence_parallel_enabled = False
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.gradient_accumulation_fusion = gradient_accumulation_fusion
        if self.async_tensor_model_parallel_allreduce and self.sequence_parallel_enabled:
            raise RuntimeError("`async_tensor_model_parallel_allreduce` and `sequence_parallel_enabled` "
                               "cannot be enabled at the same time.")

    def forward(self, input_):
        bias = self.bias if not self.skip_bias_add else None
        if self.async_tensor_model_parallel_allreduce or self.sequence_parallel_enabled:
            input_parallel = input_
        else:
            input_parallel = copy_to_tensor_model_parallel_region(input_)
        output_parallel = linear_with_grad_accumulation_and_async_allreduce(
            input_parallel, self.weight, bias, self.gradient_accumulation_fusion,
            self.async_tensor_model_parallel_allreduce, self.sequence_parallel_enabled)
        if self.gather_output:
            assert not self.sequence_parallel_enabled
            output = gather_from_tensor_model_parallel_region(output_parallel)
        else:
            output = output_parallel
        return output, (self.bias if self.skip_bias_add else None)


class _RowLinearFusedRS(torch.autograd.Function):
    """Row-parallel forward with the GEMM -> reduce-scatter fused (SP); backward = all-gather -> dgrad GEMM."""

    @staticmethod
    def forward(ctx, input, weight, gradient_accumulation_fusion):
        from . import fused_tp
        ctx.gradient_accumulation_fusion = gradient_accumulation_fusion
        ctx.save_for_backward(input, weight)
        x2d = input.reshape(-1, input.size(-1))
        out2d = fused_tp.gemm_rs(x2d, weight, transposed_weight=False)
        world = ps.get_tensor_model_parallel_world_size()
        return out2d.view(input.size(0) // world, *input.shape[1:-1], weight.size(0))

    @staticmethod
    def backward(ctx, grad_output):
        from . import fused_tp
        input, weight = ctx.saved_tensors
        # dX = AG(dY) @ W ; dW = AG(dY)^T @ X
        gi2d, total_g = fused_tp.ag_gemm(grad_output.contiguous(), weight, transposed_weight=True, keep=False)
        grad_input = gi2d.view(*input.shape)
        grad_weight = _wgrad(total_g.reshape(-1, total_g.size(-1)), input.reshape(-1, input.size(-1)), weight,
                             ctx.gradient_accumulation_fusion)
        return grad_input, grad_weight, None


class _RowLinearFusedAR(torch.autograd.Function):
    """Row-parallel forward without sequence parallelism: GEMM -> all-reduce fused (reference: GEMM, then
    ``reduce_from_tensor_model_parallel_region``, layers.py:694 / mappings.py:13-23); backward = plain dgrad + wgrad
    (the incoming gradient is already replicated over the TP group)."""

    @staticmethod
    def forward(ctx, input, weight, gradient_accumulation_fusion):
        from . import fused_tp
        ctx.gradient_accumulation_fusion = gradient_accumulation_fusion
        ctx.save_for_backward(input, weight)
        out2d = fused_tp.gemm_ar(input.reshape(-1, input.size(-1)), weight, transposed_weight=False)
        return out2d.view(*input.shape[:-1], weight.size(0))

    @staticmethod
    def backward(ctx, grad_output):
        input, weight = ctx.saved_tensors
        g2d = grad_output.contiguous().reshape(-1, grad_output.size(-1))
        grad_input = ops.gemm_nn(g2d, weight).view(*input.shape)
        grad_weight = _wgrad(g2d, input.reshape(-1, input.size(-1)), weight, ctx.gradient_accumulation_fusion)
        return grad_input, grad_weight, None


class RowParallelLinear(torch.nn.Module):
    """Y = X A + b with A split along its input (row) dimension: weight shard is [out, in/tp].

    SP: GEMM -> reduce-scatter(seq); otherwise GEMM -> all-reduce.  The bias is unsharded."""

    def __init__(self, input_size, output_size, *, bias=True, input_is_parallel=False,
                 init_method=init.xavier_normal_, stride=1, keep_master_weight_for_test=False,
                 skip_bias_add=False, params_dtype=torch.float32, use_cpu_initialization=False,
                 perform_initialization=True, gradient_accumulation_fusion=False,
                 sequence_parallel_enabled: bool = False, world_size: int = None):
        super().__init__()
        self.input_size, self.output_size, self.input_is_parallel = input_size, output_size, input_is_parallel
        world_size = world_size if world_size is not None else ps.get_tensor_model_parallel_world_size()
        self.world_size = world_size
        self.input_size_per_partition = divide(input_size, world_size)
        self.skip_bias_add = skip_bias_add
        self.gradient_accumulation_fusion = gradient_accumulation_fusion
        self.sequence_parallel_enabled = sequence_parallel_enabled
        if self.sequence_parallel_enabled and not self.input_is_parallel:
            raise RuntimeError("To enable `sequence_parallel_enabled`, `input_is_parallel` must be `True`")
        if use_cpu_initialization:
            self.weight = Parameter(torch.empty(output_size, self.input_size_per_partition, dtype=params_dtype))
            if perform_initialization:
                self.master_weight = _initialize_affine_weight_cpu(
                    self.weight, output_size, input_size, self.input_size_per_partition, 1, init_method,
                    stride=stride, return_master_weight=keep_master_weight_for_test, params_dtype=params_dtype)
        else:
            self.weight = Parameter(torch.empty(output_size, self.input_size_per_partition,
                                                device=current_device(), dtype=params_dtype))
            if perform_initialization:
                _initialize_affine_weight_gpu(self.weight, init_method, partition_dim=1, stride=stride)
        if bias:
            dev = None if use_cpu_initialization else current_device()
            self.bias = Parameter(torch.zeros(output_size, dtype=params_dtype, device=dev))
            setattr(self.bias, "sequence_parallel", sequence_parallel_enabled)
        else:
            self.register_parameter("bias", None)

    def forward(self, input_):
        if self.input_is_parallel:
            input_parallel = input_
        else:
            assert not self.sequence_parallel_enabled
            input_parallel = scatter_to_tensor_model_parallel_region(input_)
        from . import fused_tp
        if self.sequence_parallel_enabled and self.world_size > 1 and fused_tp.active_row(input_parallel, self.weight):
            output_ = _RowLinearFusedRS.apply(input_parallel, self.weight, self.gradient_accumulation_fusion)
        elif (not self.sequence_parallel_enabled and self.world_size > 1 and input_parallel.dtype == self.weight.dtype
              and fused_tp.active_all_reduce(input_parallel.reshape(-1, input_parallel.size(-1)), self.weight.size(0))):
            output_ = _RowLinearFusedAR.apply(input_parallel, self.weight, self.gradient_accumulation_fusion)
        else:
            output_parallel = linear_with_grad_accumulation_and_async_allreduce(
                input_parallel, self.weight, None, self.gradient_accumulation_fusion, False, False)
            if self.sequence_parallel_enabled:
                output_ = reduce_scatter_to_sequence_parallel_region(output_parallel)
            else:
                output_ = reduce_from_tensor_model_parallel_region(output_parallel)
        if not self.skip_bias_add:
            output = output_ + self.bias if self.bias is not None else output_
            output_bias = None
        else:
            output, output_bias = output_, self.bias
        return output, output_bias

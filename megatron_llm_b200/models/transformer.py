"""Transformer blocks.

Parity target: megatron/model/transformer.py -- DropPath :43-63, ParallelMLP :77-141, CoreAttention :144-277,
ParallelAttention :280-560, ParallelTransformerLayer :612-846, NoopTransformerLayer :849-872,
_get_num_layers :875-924, ParallelTransformer :927-1282.

B200-first differences (same math, same parameter names / checkpoint keys):
  * GQA/MQA is native: K/V are never broadcast to the query head count (reference :458-465 materialises
    copies); the KV cache stores ``n_kv/tp`` heads.
  * RoPE is applied in place on the packed QKV GEMM output by one kernel (``ops.rope_qkv_``) from a
    device-resident (cos,sin) table; the reference rebuilds/moves a complex table per call (:383-389).
  * residual-add + norm are one kernel; the add at the end of layer *i* is deferred into the first norm of
    layer *i+1* (``deferred residual``), so the residual stream is read/written once per sub-block.
  * SwiGLU/GeLU, bias+dropout+add are single kernels; attention is the tcgen05 kernel (flash library only
    as the out-of-envelope fallback).
  * RoPE models are servable: keys are rotated at their absolute positions *before* entering the KV cache
    (the reference re-applies ``position_ids``-indexed frequencies to the whole cache, :492-511, which cannot
    work for incremental decoding).
"""
from __future__ import annotations

import math
from contextlib import nullcontext
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .. import ops
from ..ops import attention_sm100
from ..parallel import layers as tp_layers
from ..parallel import state as ps
from ..parallel.random import checkpoint as tp_checkpoint, get_cuda_rng_tracker
from ..utils.core_utils import divide, make_viewless_tensor
from ..utils.device import current_device
from .activations import GLU_ACTIVATIONS, attention_mask_func, bias_gelu_impl
from .enums import AttnMaskType, AttnType, LayerType, ModelType, PositionEmbeddingType
from .fused_softmax import FusedScaleMaskSoftmax
from .module import MegatronModule
from .norms import LayerNorm, RMSNorm


_ROPE_TABLES = {}


def _shared_rope_table(cfg, device):
    """One fp32 [len, hn/2, 2] (cos, sin) table per (config, device), built once and kept on the device."""
    key = (cfg, str(device))
    tab = _ROPE_TABLES.get(key)
    if tab is None:
        hn, length, theta, scaling = cfg
        tab = ops.rope_table(hn, length, theta=theta, scaling_factor=scaling, device=device)
        _ROPE_TABLES[key] = tab
    return tab


class DropPath(MegatronModule):
    """Stochastic depth per sample (input is [s, b, h]; one Bernoulli draw per batch element)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, hidden_state):
        if self.drop_prob == 0.0 or not self.training:
            return hidden_state
        keep_prob = 1 - self.drop_prob
        shape = (1,) + (hidden_state.shape[1],) + (1,) * (hidden_state.ndim - 2)
        random_tensor = keep_prob + torch.rand(shape, dtype=hidden_state.dtype, device=hidden_state.device)
        random_tensor.floor_()
        return hidden_state.div(keep_prob) * random_tensor


class ParallelMLP(MegatronModule):
    """h -> ffn (x2 for GLU) -> activation -> h, Column then Row parallel, both ``skip_bias_add``."""

    def __init__(self, init_method, output_layer_init_method, args, world_size):
        super().__init__()
        self.add_bias = args.use_bias
        self.dense_h_to_4h = tp_layers.ColumnParallelLinear(
            args.hidden_size, 2 * args.ffn_hidden_size if args.glu_activation else args.ffn_hidden_size,
            bias=self.add_bias, gather_output=False, init_method=init_method, skip_bias_add=True,
            async_tensor_model_parallel_allreduce=args.async_tensor_model_parallel_allreduce,
            params_dtype=args.params_dtype, use_cpu_initialization=args.use_cpu_initialization,
            perform_initialization=args.perform_initialization,
            gradient_accumulation_fusion=args.gradient_accumulation_fusion,
            sequence_parallel_enabled=args.sequence_parallel, world_size=world_size)
        self.use_bias = args.use_bias
        self.bias_gelu_fusion = args.bias_gelu_fusion
        self.glu_kind = args.glu_activation
        if args.glu_activation:
            self.activation_func = GLU_ACTIVATIONS[args.glu_activation]
        elif getattr(args, "onnx_safe", False):
            from .activations import erf_gelu
            self.activation_func = erf_gelu
        else:
            self.activation_func = lambda x: ops.gelu(x, None, approximate=False)
        self.dense_4h_to_h = tp_layers.RowParallelLinear(
            args.ffn_hidden_size, args.hidden_size, bias=self.add_bias, input_is_parallel=True,
            init_method=output_layer_init_method, skip_bias_add=True, params_dtype=args.params_dtype,
            use_cpu_initialization=args.use_cpu_initialization, perform_initialization=args.perform_initialization,
            gradient_accumulation_fusion=args.gradient_accumulation_fusion,
            sequence_parallel_enabled=args.sequence_parallel, world_size=world_size)

    def forward(self, hidden_states):
        intermediate_parallel, bias_parallel = self.dense_h_to_4h(hidden_states)
        if self.bias_gelu_fusion and bias_parallel is not None and not self.glu_kind:
            intermediate_parallel = bias_gelu_impl(intermediate_parallel, bias_parallel)
        else:
            if bias_parallel is not None:
                intermediate_parallel = intermediate_parallel + bias_parallel
            intermediate_parallel = self.activation_func(intermediate_parallel)
        return self.dense_4h_to_h(intermediate_parallel)


class CoreAttention(MegatronModule):
    """Unfused attention: QK^T -> scale/mask/softmax kernel -> dropout -> PV.  q/k/v: [s, b, n, hn]
    (k/v may carry fewer heads; they are expanded by view, never copied, for the batched matmul)."""

    def __init__(self, layer_number, attn_mask_type=AttnMaskType.padding, args=None, world_size=None):
        super().__init__()
        self.fp16, self.bf16 = args.fp16, args.bf16
        self.apply_query_key_layer_scaling = args.apply_query_key_layer_scaling
        self.attention_softmax_in_fp32 = args.attention_softmax_in_fp32
        if self.apply_query_key_layer_scaling:
            self.attention_softmax_in_fp32 = True
        self.layer_number = max(1, layer_number)
        self.attn_mask_type = attn_mask_type
        self.sequence_parallel = args.sequence_parallel
        projection_size = args.kv_channels * args.num_attention_heads
        world_size = world_size if world_size is not None else ps.get_tensor_model_parallel_world_size()
        self.hidden_size_per_partition = divide(projection_size, world_size)
        self.hidden_size_per_attention_head = divide(projection_size, args.num_attention_heads)
        self.num_attention_heads_per_partition = divide(args.num_attention_heads, world_size)
        coeff = None
        self.norm_factor = math.sqrt(self.hidden_size_per_attention_head)
        if self.apply_query_key_layer_scaling:
            coeff = self.layer_number
            self.norm_factor *= coeff
        self.scale_mask_softmax = FusedScaleMaskSoftmax(
            self.fp16, self.bf16, self.attn_mask_type, args.masked_softmax_fusion, attention_mask_func,
            self.attention_softmax_in_fp32, coeff)
        self.attention_dropout = torch.nn.Dropout(args.attention_dropout)

    def forward(self, query_layer, key_layer, value_layer, attention_mask):
        sq, b, np_, hn = query_layer.shape
        sk, nkv = key_layer.size(0), key_layer.size(2)
        if nkv != np_:
            g = np_ // nkv
            key_layer = key_layer.unsqueeze(3).expand(sk, b, nkv, g, hn).reshape(sk, b, np_, hn)
            value_layer = value_layer.unsqueeze(3).expand(sk, b, nkv, g, hn).reshape(sk, b, np_, hn)
        q = query_layer.reshape(sq, b * np_, hn).transpose(0, 1)       # [b*np, sq, hn]
        k = key_layer.reshape(sk, b * np_, hn).transpose(0, 1)          # [b*np, sk, hn]
        scores = torch.bmm(q, k.transpose(1, 2)) * (1.0 / self.norm_factor)
        attention_scores = scores.view(b, np_, sq, sk)
        attention_probs = self.scale_mask_softmax(attention_scores, attention_mask)
        if not self.sequence_parallel:
            with get_cuda_rng_tracker().fork():
                attention_probs = self.attention_dropout(attention_probs)
        else:
            attention_probs = self.attention_dropout(attention_probs)
        v = value_layer.reshape(sk, b * np_, hn).transpose(0, 1)        # [b*np, sk, hn]
        ctx = torch.bmm(attention_probs.view(b * np_, sq, sk).to(v.dtype), v)  # [b*np, sq, hn]
        ctx = ctx.view(b, np_, sq, hn).permute(2, 0, 1, 3).contiguous()
        return ctx.view(sq, b, self.hidden_size_per_partition)


class ParallelAttention(MegatronModule):
    """Self / cross attention: fused QKV Column linear (per-KV-group layout [q_0..q_{g-1}, k, v]), RoPE,
    attention core, output Row linear.  Input/output: [s, b, h]."""

    def __init__(self, init_method, output_layer_init_method, layer_number, attention_type=AttnType.self_attn,
                 attn_mask_type=AttnMaskType.padding, world_size: int = None, args=None):
        super().__init__()
        self.layer_number = max(1, layer_number)
        self.attention_type = attention_type
        self.attn_mask_type = attn_mask_type
        self.params_dtype = args.params_dtype
        self.sequence_parallel = args.sequence_parallel
        self.use_flash_attn = args.use_flash_attn
        self.sliding_window_size = args.sliding_window_size
        self.num_attention_heads_kv = args.num_attention_heads_kv
        self.num_attention_heads = args.num_attention_heads
        self.seq_length = args.seq_length
        self.attention_dropout_p = args.attention_dropout
        if self.use_flash_attn:
            assert attention_type == AttnType.self_attn, "FlashAttention code path only supports self-attention for now"
            assert self.attn_mask_type == AttnMaskType.causal, "FlashAttention code path only supports causal mask for now"
        projection_size = args.kv_channels * args.num_attention_heads
        qkv_projection_size = args.kv_channels * args.num_attention_heads + 2 * args.kv_channels * args.num_attention_heads_kv
        world_size = world_size if world_size is not None else ps.get_tensor_model_parallel_world_size()
        self.hidden_size_per_attention_head = divide(projection_size, args.num_attention_heads)
        self.num_attention_heads_per_partition = divide(args.num_attention_heads, world_size)
        self.num_kv_heads_per_partition = divide(args.num_attention_heads_kv, world_size)
        self.q_per_kv = divide(args.num_attention_heads, args.num_attention_heads_kv)
        common = dict(gather_output=False, init_method=init_method,
                      async_tensor_model_parallel_allreduce=args.async_tensor_model_parallel_allreduce,
                      params_dtype=args.params_dtype, use_cpu_initialization=args.use_cpu_initialization,
                      perform_initialization=args.perform_initialization,
                      gradient_accumulation_fusion=args.gradient_accumulation_fusion,
                      sequence_parallel_enabled=args.sequence_parallel, world_size=world_size)
        if attention_type == AttnType.self_attn:
            self.query_key_value = tp_layers.ColumnParallelLinear(args.hidden_size, qkv_projection_size,
                                                                  bias=args.use_bias, **common)
        else:
            assert attention_type == AttnType.cross_attn
            self.query = tp_layers.ColumnParallelLinear(args.hidden_size, projection_size, bias=args.use_bias, **common)
            self.key_value = tp_layers.ColumnParallelLinear(args.hidden_size, 2 * projection_size, bias=args.use_bias,
                                                            **common)
        self.core_attention = CoreAttention(self.layer_number, self.attn_mask_type, args, world_size)
        self.checkpoint_core_attention = args.recompute_granularity == "selective"
        self.dense = tp_layers.RowParallelLinear(
            projection_size, args.hidden_size, bias=args.use_bias, input_is_parallel=True,
            init_method=output_layer_init_method, skip_bias_add=True, params_dtype=args.params_dtype,
            use_cpu_initialization=args.use_cpu_initialization, perform_initialization=args.perform_initialization,
            gradient_accumulation_fusion=args.gradient_accumulation_fusion,
            sequence_parallel_enabled=args.sequence_parallel, world_size=world_size)
        self.position_embedding_type = args.position_embedding_type
        if self.position_embedding_type == PositionEmbeddingType.rotary:
            # fp32 (cos, sin) table shared by all layers of the process.  NOT a module buffer: ``module.bfloat16()``
            # (Float16Module) would silently down-cast a buffer and the kernel reads it as float2.
            self._rope_cfg = (self.hidden_size_per_attention_head, max(args.max_position_embeddings, args.seq_length),
                              float(args.rope_theta), float(args.rope_scaling_factor))
            # complex table kept for API parity (tools / verify scripts)
            self.freqs_cis = None

    def _checkpointed_attention_forward(self, query_layer, key_layer, value_layer, attention_mask):
        def custom_forward(*inputs):
            return self.core_attention(inputs[0], inputs[1], inputs[2], inputs[3])
        return tp_checkpoint(custom_forward, False, query_layer, key_layer, value_layer, attention_mask)

    def _allocate_memory(self, inference_max_sequence_len, batch_size):
        return torch.empty(inference_max_sequence_len, batch_size, self.num_kv_heads_per_partition,
                           self.hidden_size_per_attention_head, dtype=self.params_dtype, device=current_device())

    def _attention_dropout_rng(self, p):
        """RNG scope of the attention dropout (reference transformer.py:540-553): the model-parallel stream, so TP ranks
        -- which hold different heads -- draw different masks; sequence-parallel runs use the default stream."""
        return get_cuda_rng_tracker().fork() if (not self.sequence_parallel and p > 0) else nullcontext()

    def forward(self, hidden_states, attention_mask, encoder_output=None, inference_params=None, position_ids=None):
        hn = self.hidden_size_per_attention_head
        pos_offset = 0
        if inference_params is not None:
            if self.layer_number not in inference_params.key_value_memory_dict:
                kmem = self._allocate_memory(inference_params.max_sequence_len, inference_params.max_batch_size)
                vmem = self._allocate_memory(inference_params.max_sequence_len, inference_params.max_batch_size)
                inference_params.key_value_memory_dict[self.layer_number] = (kmem, vmem)
            else:
                kmem, vmem = inference_params.key_value_memory_dict[self.layer_number]
            pos_offset = inference_params.sequence_len_offset

        if self.attention_type == AttnType.self_attn:
            mixed, _ = self.query_key_value(hidden_states)      # [sq, b, (np + 2 nkv) hn]
            sq, b = mixed.shape[:2]
            if self.position_embedding_type == PositionEmbeddingType.rotary:
                pid = position_ids if inference_params is None else None
                mixed = ops.rope_qkv_(mixed, self.num_kv_heads_per_partition, self.q_per_kv + 2, hn,
                                      _shared_rope_table(self._rope_cfg, mixed.device), pid, pos_offset)
            # fast path: the tcgen05 attention kernels read Q/K/V in place from the packed projection output and the
            # backward writes one packed gradient (no splits / transposes / view-gradient accumulation)
            p_drop = self.attention_dropout_p if self.training else 0.0
            if (inference_params is None and self.use_flash_attn
                    and attention_sm100.packed_supported(mixed, self.num_kv_heads_per_partition, self.q_per_kv, hn,
                                                         p_drop)):
                window = None
                if self.sliding_window_size is not None and sq > self.sliding_window_size:
                    window = self.sliding_window_size
                with self._attention_dropout_rng(p_drop):
                    context_layer = attention_sm100.packed_attention(mixed, self.num_kv_heads_per_partition,
                                                                     self.q_per_kv, window, None, hn, p_drop)
                return self.dense(context_layer)
            qkv = mixed.view(sq, b, self.num_kv_heads_per_partition, self.q_per_kv + 2, hn)
            if self.q_per_kv == 1:
                query_layer = qkv[:, :, :, 0]                                     # view [sq,b,np,hn]
            else:
                query_layer = qkv[:, :, :, :self.q_per_kv].reshape(sq, b, self.num_attention_heads_per_partition, hn)
            key_layer = qkv[:, :, :, self.q_per_kv]                               # view [sq,b,nkv,hn]
            value_layer = qkv[:, :, :, self.q_per_kv + 1]
        else:
            mixed_kv, _ = self.key_value(encoder_output)
            mixed_kv = mixed_kv.view(*mixed_kv.shape[:-1], self.num_attention_heads_per_partition, 2 * hn)
            key_layer, value_layer = mixed_kv[..., :hn], mixed_kv[..., hn:]
            query_layer, _ = self.query(hidden_states)
            query_layer = query_layer.view(*query_layer.shape[:-1], self.num_attention_heads_per_partition, hn)
            sq, b = query_layer.shape[:2]

        if inference_params is not None:
            b0 = inference_params.batch_size_offset
            b1 = b0 + key_layer.size(1)
            s0 = inference_params.sequence_len_offset
            s1 = s0 + key_layer.size(0)
            assert b1 <= kmem.size(1) and s1 <= kmem.size(0)
            kmem[s0:s1, b0:b1] = key_layer
            vmem[s0:s1, b0:b1] = value_layer
            key_layer, value_layer = kmem[:s1, b0:b1], vmem[:s1, b0:b1]

        if not self.use_flash_attn:
            if self.checkpoint_core_attention:
                context_layer = self._checkpointed_attention_forward(query_layer, key_layer, value_layer, attention_mask)
            else:
                context_layer = self.core_attention(query_layer, key_layer, value_layer, attention_mask)
        else:
            window = None
            if self.sliding_window_size is not None and key_layer.shape[0] > self.sliding_window_size:
                window = self.sliding_window_size
            q = query_layer.transpose(0, 1)   # [b, s, n, hn] strided views, no copies
            k = key_layer.transpose(0, 1)
            v = value_layer.transpose(0, 1)
            p = self.attention_dropout_p if self.training else 0.0
            with self._attention_dropout_rng(p):
                ctx = ops.flash_attention(q, k, v, causal=True, window=window, dropout_p=p)
            # [b, s, n, hn] -> [s, b, n*hn]
            context_layer = ctx.transpose(0, 1).reshape(sq, b, -1)
        return self.dense(context_layer)


def _bias_dropout_add(x, bias, residual, prob, training):
    """residual + dropout(x + bias): one kernel forward, one backward (``ops.bias_dropout_add``)."""
    return ops.bias_dropout_add(x, bias, residual, prob, training)


def bias_dropout_add(x, bias, residual, prob, training):
    """residual + dropout(x + bias) (reference transformer.py:529-535)."""
    return ops.bias_dropout_add(x, bias, residual, prob, training)


def dropout_add(x, bias, residual, prob, training):
    """residual + dropout(x); ``bias`` must be None (reference transformer.py:503-511, the bias-free models)."""
    assert bias is None
    return ops.bias_dropout_add(x, None, residual, prob, training)


def get_bias_dropout_add(training):
    return lambda x, bias, residual, prob: bias_dropout_add(x, bias, residual, prob, training)


def get_dropout_add(training):
    return lambda x, bias, residual, prob: dropout_add(x, bias, residual, prob, training)


def bias_dropout_add_fused_train(x, bias, residual, prob):
    """(the reference needs a jit-scripted variant per mode, transformer.py:544-560; here one kernel serves both)"""
    return bias_dropout_add(x, bias, residual, prob, True)


def bias_dropout_add_fused_inference(x, bias, residual, prob):
    return bias_dropout_add(x, bias, residual, prob, False)


class ParallelTransformerLayer(MegatronModule):
    """One transformer layer, [s, b, h] -> [s, b, h].

    Variants: pre-LN (default), ``use_post_ln``, ``apply_residual_connection_post_layernorm``, Falcon
    ``parallel_attn`` (attention and MLP read the same normed input, outputs summed) with optional
    ``parallel_layernorm`` (separate MLP norm), decoder layers with cross attention."""

    def __init__(self, init_method: Callable, output_layer_init_method: Callable, layer_number: int,
                 layer_type=LayerType.encoder, self_attn_mask_type=AttnMaskType.padding, drop_path_rate: float = 0.0,
                 world_size: int = None, hidden_dropout: float = 0.0, args=None):
        super().__init__()
        self.layer_number = layer_number
        self.layer_type = layer_type
        self.apply_residual_connection_post_layernorm = args.apply_residual_connection_post_layernorm
        self.bf16 = args.bf16
        self.fp32_residual_connection = args.fp32_residual_connection
        self.parallel_layernorm = args.parallel_layernorm

        def make_norm():
            if args.use_rms_norm:
                return RMSNorm(args.hidden_size, eps=args.layernorm_epsilon, sequence_parallel=args.sequence_parallel)
            return LayerNorm(args.hidden_size, eps=args.layernorm_epsilon,
                             no_persist_layer_norm=args.no_persist_layer_norm, sequence_parallel=args.sequence_parallel)

        self.use_post_ln = args.use_post_ln
        if args.use_post_ln:
            self.input_layernorm = torch.nn.Identity()
            self.output_layernorm = make_norm()
        else:
            self.input_layernorm = make_norm()
            self.output_layernorm = torch.nn.Identity()
        if self.parallel_layernorm:
            self.mlp_layernorm = make_norm()
        self.self_attention = ParallelAttention(init_method, output_layer_init_method, layer_number,
                                                attention_type=AttnType.self_attn, attn_mask_type=self_attn_mask_type,
                                                world_size=world_size, args=args)
        self.hidden_dropout = hidden_dropout
        self.bias_dropout_fusion = args.bias_dropout_fusion
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else None
        self.parallel_attn = args.parallel_attn
        self.use_bias = args.use_bias
        if not args.parallel_attn:
            self.post_attention_layernorm = make_norm()
        if self.layer_type == LayerType.decoder:
            self.inter_attention = ParallelAttention(init_method, output_layer_init_method, layer_number,
                                                     attention_type=AttnType.cross_attn, world_size=world_size,
                                                     args=args)
            self.post_inter_attention_layernorm = make_norm()
        self.mlp = ParallelMLP(init_method, output_layer_init_method, args, world_size)
        # fast path: plain pre-LN decoder-only layer whose residual adds can be fused into the norms
        self._fusable = (not args.use_post_ln and not args.parallel_attn and not self.use_bias
                         and not self.apply_residual_connection_post_layernorm
                         and self.layer_type == LayerType.encoder and self.drop_path is None
                         and not self.fp32_residual_connection)

    def _add_dropout(self, x, bias, residual, prob):
        if self.drop_path is None:
            return _bias_dropout_add(x, bias if self.use_bias else None, residual, prob, self.training)
        out = x + bias if (self.use_bias and bias is not None) else x
        out = F.dropout(out, p=prob, training=self.training)
        return residual + self.drop_path(out)

    def supports_deferred_residual(self) -> bool:
        return self._fusable and not (self.training and self.hidden_dropout > 0.0)

    def forward(self, hidden_states, attention_mask, encoder_output=None, enc_dec_attn_mask=None,
                inference_params=None, position_ids=None, deferred_residual=None, return_deferred=False):
        if self.supports_deferred_residual():
            # hidden = hidden_states (+ deferred_residual), all adds fused into the norm kernels
            if deferred_residual is not None:
                ln1, hidden = self.input_layernorm(hidden_states, residual=deferred_residual)
            else:
                ln1, hidden = self.input_layernorm(hidden_states), hidden_states
            attn_out, _ = self.self_attention(ln1, attention_mask, inference_params=inference_params,
                                              position_ids=position_ids)
            ln2, hidden2 = self.post_attention_layernorm(attn_out, residual=hidden)
            mlp_out, _ = self.mlp(ln2)
            if return_deferred:
                return mlp_out, hidden2
            return make_viewless_tensor(mlp_out + hidden2, requires_grad=mlp_out.requires_grad, keep_graph=True)

        if deferred_residual is not None:
            hidden_states = hidden_states + deferred_residual
        layernorm_output = self.input_layernorm(hidden_states)
        attention_output, attention_bias = self.self_attention(layernorm_output, attention_mask,
                                                               inference_params=inference_params,
                                                               position_ids=position_ids)
        residual = layernorm_output if self.apply_residual_connection_post_layernorm else hidden_states
        if self.parallel_layernorm:
            layernorm_output = self.mlp_layernorm(hidden_states)
        if self.parallel_attn:
            layernorm_input = attention_output
        else:
            layernorm_input = self._add_dropout(attention_output, attention_bias, residual, self.hidden_dropout)
            layernorm_output = self.post_attention_layernorm(layernorm_input)
        if self.layer_type == LayerType.decoder:
            attention_output, attention_bias = self.inter_attention(layernorm_output, enc_dec_attn_mask,
                                                                    encoder_output=encoder_output)
            residual = layernorm_output if self.apply_residual_connection_post_layernorm else layernorm_input
            layernorm_input = self._add_dropout(attention_output, attention_bias, residual, self.hidden_dropout)
            layernorm_output = self.post_inter_attention_layernorm(layernorm_input)
        mlp_output, mlp_bias = self.mlp(layernorm_output)
        if self.parallel_attn:
            mlp_output = mlp_output + attention_output
        elif self.apply_residual_connection_post_layernorm:
            residual = layernorm_output
        else:
            residual = layernorm_input
        output = self._add_dropout(mlp_output, mlp_bias, residual, self.hidden_dropout)
        output = make_viewless_tensor(output, requires_grad=output.requires_grad, keep_graph=True)
        output = self.output_layernorm(output)
        if return_deferred:
            return output, None
        return output


class NoopTransformerLayer(MegatronModule):
    """Placeholder for a stage that owns zero layers (standalone embedding stage)."""

    def __init__(self, layer_number):
        super().__init__()
        self.layer_number = layer_number

    def forward(self, hidden_states, attention_mask, encoder_output=None, enc_dec_attn_mask=None,
                inference_params=None, **kw):
        return hidden_states.clone()


def _get_num_layers(args, is_encoder_and_decoder_model, is_decoder=False):
    """Number of layers owned by this pipeline rank."""
    pp = ps.get_pipeline_model_parallel_world_size()
    # (the reference leaves decoder_num_layers unset when only --num_layers is given and then fails on it; default it)
    dec_layers = args.decoder_num_layers if args.decoder_num_layers is not None else args.num_layers
    if pp > 1:
        if is_encoder_and_decoder_model:
            assert args.pipeline_model_parallel_split_rank is not None
            num_ranks_in_encoder = (args.pipeline_model_parallel_split_rank - 1 if args.standalone_embedding_stage
                                    else args.pipeline_model_parallel_split_rank)
            num_ranks_in_decoder = args.transformer_pipeline_model_parallel_size - num_ranks_in_encoder
            assert args.encoder_num_layers % num_ranks_in_encoder == 0
            assert dec_layers % num_ranks_in_decoder == 0
            if ps.is_pipeline_stage_before_split():
                num_layers = (0 if args.standalone_embedding_stage and ps.get_pipeline_model_parallel_rank() == 0
                              else args.encoder_num_layers // num_ranks_in_encoder)
            else:
                num_layers = dec_layers // num_ranks_in_decoder
        else:
            assert args.num_layers == args.encoder_num_layers
            assert args.num_layers % args.transformer_pipeline_model_parallel_size == 0, \
                "num_layers must be divisible by transformer_pipeline_model_parallel_size"
            num_layers = (0 if args.standalone_embedding_stage and ps.get_pipeline_model_parallel_rank() == 0
                          else args.num_layers // args.transformer_pipeline_model_parallel_size)
    else:
        num_layers = dec_layers if is_decoder else args.encoder_num_layers
    return num_layers


class ParallelTransformer(MegatronModule):
    """Stack of transformer layers owned by this (virtual) pipeline stage."""

    def __init__(self, init_method: Callable, output_layer_init_method, layer_type=LayerType.encoder,
                 self_attn_mask_type=AttnMaskType.padding, post_layer_norm=True, pre_process=True, post_process=True,
                 drop_path_rate=0.0, args=None, model_type=None):
        super().__init__()
        world_size = ps.get_tensor_model_parallel_world_size()
        assert args is not None
        if model_type is None:       # the reference reads args.model_type (set by get_model); default = decoder-only
            model_type = getattr(args, "model_type", None) or ModelType.encoder_or_decoder
        self.layer_type = layer_type
        self.model_type = model_type
        self.bf16 = args.bf16
        self.fp32_residual_connection = args.fp32_residual_connection
        self.post_layer_norm = post_layer_norm
        self.pre_process, self.post_process = pre_process, post_process
        self.input_tensor = None
        self.drop_path_rate = drop_path_rate
        self.transformer_impl = args.transformer_impl
        self.recompute_granularity = args.recompute_granularity
        self.recompute_method = args.recompute_method
        self.recompute_num_layers = args.recompute_num_layers
        self.distribute_saved_activations = args.distribute_saved_activations and not args.sequence_parallel
        self.sequence_parallel = args.sequence_parallel
        self._any_dropout = args.hidden_dropout > 0.0 or args.attention_dropout > 0.0 or drop_path_rate > 0.0
        assert args.transformer_impl == "local", \
            "transformer_engine fp8 layers are not part of this build (fp8 is never on a Llama/Falcon/Mistral path)"
        self.num_microbatches_in_previous_step = -1
        self.microbatch_count = 0
        self.checkpoint_core_attention = args.recompute_granularity == "selective"
        self.num_layers = _get_num_layers(args, model_type == ModelType.encoder_and_decoder,
                                          layer_type == LayerType.decoder)
        self.drop_path_rates = [r.item() for r in torch.linspace(0, self.drop_path_rate, args.num_layers)]

        if args.lima_dropout:
            # LIMA: per-layer hidden dropout ramps linearly from 0 (bottom) to hidden_dropout (top)
            ramp = [args.hidden_dropout * i / max(1, args.num_layers - 1) for i in range(args.num_layers)]
        else:
            ramp = [args.hidden_dropout] * args.num_layers

        def build_layer(layer_number):
            return ParallelTransformerLayer(
                init_method, output_layer_init_method, layer_number, layer_type=layer_type,
                self_attn_mask_type=self_attn_mask_type, drop_path_rate=self.drop_path_rates[layer_number - 1],
                world_size=world_size, hidden_dropout=ramp[layer_number - 1], args=args)

        if args.virtual_pipeline_model_parallel_size is not None:
            assert args.num_layers % args.virtual_pipeline_model_parallel_size == 0, \
                "num_layers_per_stage must be divisible by virtual_pipeline_model_parallel_size"
            assert model_type != ModelType.encoder_and_decoder
            self.num_layers = self.num_layers // args.virtual_pipeline_model_parallel_size
            first = 1 if args.standalone_embedding_stage else 0
            offset = ps.get_virtual_pipeline_model_parallel_rank() * (
                args.num_layers // args.virtual_pipeline_model_parallel_size) + \
                (max(0, ps.get_pipeline_model_parallel_rank() - first) * self.num_layers)
        else:
            if model_type == ModelType.encoder_and_decoder and ps.get_pipeline_model_parallel_world_size() > 1:
                pipeline_rank = ps.get_pipeline_model_parallel_rank()
                if layer_type == LayerType.encoder:
                    offset = pipeline_rank * self.num_layers
                else:
                    offset = (pipeline_rank - args.pipeline_model_parallel_split_rank) * self.num_layers
            else:
                # with a standalone embedding stage, stage 0 owns no layer and stage r >= 1 owns the (r-1)-th slice
                # (the reference numbers the layers as if stage 0 owned one as well and then indexes its per-layer
                # tables out of range, transformer.py:1075)
                first = 1 if args.standalone_embedding_stage else 0
                offset = max(0, ps.get_pipeline_model_parallel_rank() - first) * self.num_layers

        if self.num_layers == 0:
            self.num_layers = 1
            self.layers = torch.nn.ModuleList([NoopTransformerLayer(1)])
        else:
            self.layers = torch.nn.ModuleList([build_layer(i + 1 + offset) for i in range(self.num_layers)])

        if self.post_process and self.post_layer_norm:
            if args.use_rms_norm:
                self.final_layernorm = RMSNorm(args.hidden_size, eps=args.layernorm_epsilon,
                                               sequence_parallel=args.sequence_parallel)
            else:
                self.final_layernorm = LayerNorm(args.hidden_size, eps=args.layernorm_epsilon,
                                                 no_persist_layer_norm=args.no_persist_layer_norm,
                                                 sequence_parallel=args.sequence_parallel)

    def _get_layer(self, layer_number):
        return self.layers[layer_number]

    def _checkpointed_forward(self, hidden_states, attention_mask, encoder_output, enc_dec_attn_mask, position_ids):
        """Full activation recompute: ``uniform`` checkpoints every chunk of ``recompute_num_layers`` layers,
        ``block`` checkpoints only the first ``recompute_num_layers`` layers of the stage.  (Unlike the
        reference :1141-1144, ``position_ids`` is forwarded.)"""
        def custom(start, end):
            def custom_forward(x, mask, enc_out, enc_dec_mask, pos):
                for index in range(start, end):
                    x = self._get_layer(index)(x, mask, enc_out, enc_dec_mask, position_ids=pos)
                return x
            return custom_forward

        if self.recompute_method == "uniform":
            l = 0
            while l < self.num_layers:
                hidden_states = tp_checkpoint(custom(l, min(l + self.recompute_num_layers, self.num_layers)),
                                              self.distribute_saved_activations, hidden_states, attention_mask,
                                              encoder_output, enc_dec_attn_mask, position_ids)
                l += self.recompute_num_layers
        elif self.recompute_method == "block":
            for l in range(self.num_layers):
                if l < self.recompute_num_layers:
                    hidden_states = tp_checkpoint(custom(l, l + 1), self.distribute_saved_activations, hidden_states,
                                                  attention_mask, encoder_output, enc_dec_attn_mask, position_ids)
                else:
                    hidden_states = custom(l, l + 1)(hidden_states, attention_mask, encoder_output,
                                                     enc_dec_attn_mask, position_ids)
        else:
            raise ValueError("Invalid activation recompute method.")
        return hidden_states

    def set_input_tensor(self, input_tensor):
        """Pipeline stages > 0 receive their input from the previous stage instead of the embedding."""
        self.input_tensor = input_tensor

    def forward(self, hidden_states, attention_mask, encoder_output=None, enc_dec_attn_mask=None,
                inference_params=None, position_ids=None):
        if inference_params:
            assert self.recompute_granularity is None, "inference does not work with activation checkpointing"
        if not self.pre_process:
            hidden_states = self.input_tensor
        hidden_states = make_viewless_tensor(hidden_states, requires_grad=True, keep_graph=True)
        # the forked (per-TP-rank) RNG stream only matters if something draws random numbers
        rng_context = get_cuda_rng_tracker().fork() if (self.sequence_parallel and self.training and self._any_dropout) \
            else nullcontext()
        deferred = None
        with rng_context:
            if self.recompute_granularity == "full":
                hidden_states = self._checkpointed_forward(hidden_states, attention_mask, encoder_output,
                                                           enc_dec_attn_mask, position_ids)
            else:
                n = self.num_layers
                for index in range(n):
                    layer = self._get_layer(index)
                    can_defer = isinstance(layer, ParallelTransformerLayer) and layer.supports_deferred_residual()
                    # only defer when someone downstream (next fusable layer or the final norm) can absorb it
                    nxt = self._get_layer(index + 1) if index + 1 < n else None
                    absorb = (isinstance(nxt, ParallelTransformerLayer) and nxt.supports_deferred_residual()) or \
                             (nxt is None and self.post_process and self.post_layer_norm)
                    if can_defer and absorb:
                        hidden_states, deferred = layer(hidden_states, attention_mask, encoder_output=encoder_output,
                                                        enc_dec_attn_mask=enc_dec_attn_mask,
                                                        inference_params=inference_params, position_ids=position_ids,
                                                        deferred_residual=deferred, return_deferred=True)
                    else:
                        hidden_states = layer(hidden_states, attention_mask, encoder_output=encoder_output,
                                              enc_dec_attn_mask=enc_dec_attn_mask, inference_params=inference_params,
                                              position_ids=position_ids, deferred_residual=deferred) \
                            if isinstance(layer, ParallelTransformerLayer) else \
                            layer(hidden_states, attention_mask, encoder_output=encoder_output,
                                  enc_dec_attn_mask=enc_dec_attn_mask, inference_params=inference_params)
                        deferred = None
        if self.post_process and self.post_layer_norm:
            if deferred is not None:
                hidden_states, _ = self.final_layernorm(hidden_states, residual=deferred)
            else:
                hidden_states = self.final_layernorm(hidden_states)
        elif deferred is not None:
            hidden_states = hidden_states + deferred
        return hidden_states

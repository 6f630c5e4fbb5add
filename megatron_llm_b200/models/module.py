"""Module base classes (parity: megatron/model/module.py: MegatronModule :21-121, Float16Module :160-202)."""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.nn.parameter import Parameter

from ..parallel import state as ps
from ..parallel.layers import VocabParallelEmbedding

_FLOAT_TYPES = (torch.float32,)
_HALF_TYPES = (torch.float16,)
_BF16_TYPES = (torch.bfloat16,)


def param_is_not_shared(param) -> bool:
    return not getattr(param, "shared", False)


class MegatronModule(torch.nn.Module):
    """torch Module plus the tied-embedding machinery for pipeline parallelism."""

    def __init__(self, share_word_embeddings: bool = True):
        super().__init__()
        self.share_word_embeddings = share_word_embeddings

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.state_dict(prefix=prefix, keep_vars=keep_vars)

    def word_embeddings_weight(self):
        if self.pre_process:
            return self.language_model.embedding.word_embeddings.weight
        if not self.share_word_embeddings:
            raise Exception("word_embeddings_weight() called for last stage, but share_word_embeddings is false")
        return self.word_embeddings.weight

    def initialize_word_embeddings(self, init_method_normal, args):
        """With tied embeddings and PP>1 the last stage owns a second, zero-initialised copy of the word
        embedding (flagged ``shared``) that is synchronised with the first stage by an all-reduce at init and
        whose gradient is all-reduced over the embedding group every step."""
        if not self.share_word_embeddings:
            raise Exception("initialize_word_embeddings() was called but share_word_embeddings is false")
        if args.pipeline_model_parallel_size == 1:
            return
        if ps.is_pipeline_last_stage() and not self.pre_process:
            assert not ps.is_pipeline_first_stage()
            self._word_embeddings_for_head_key = "word_embeddings_for_head"
            self.word_embeddings = VocabParallelEmbedding(
                args.padded_vocab_size, args.hidden_size, init_method=init_method_normal(args.init_method_std),
                params_dtype=args.params_dtype, use_cpu_initialization=args.use_cpu_initialization,
                perform_initialization=args.perform_initialization)
            self.word_embeddings.weight.data.fill_(0)
            self.word_embeddings.weight.shared = True
        if not ps.is_pipeline_first_stage(ignore_virtual=True) and self.pre_process:
            self.language_model.embedding.zero_parameters()
        if not dist.is_initialized():
            if not getattr(MegatronModule, "embedding_warning_printed", False):
                print("WARNING! Distributed processes aren't initialized, so word embeddings in the last layer are "
                      "not initialized. If you are just manipulating a model this is fine, but this needs to be "
                      "handled manually. If you are training something is definitely wrong.")
                MegatronModule.embedding_warning_printed = True
            return
        if ps.is_rank_in_embedding_group():
            dist.all_reduce(self.word_embeddings_weight().data, group=ps.get_embedding_group())
        if ps.is_rank_in_position_embedding_group() and args.pipeline_model_parallel_split_rank is not None:
            self.language_model.embedding.to(self.word_embeddings_weight().device)
            pos = self.language_model.embedding.position_embeddings
            dist.all_reduce(pos.weight.data, group=ps.get_position_embedding_group())


def conversion_helper(val, conversion):
    if not isinstance(val, (tuple, list)):
        return conversion(val)
    rtn = [conversion_helper(v, conversion) for v in val]
    return tuple(rtn) if isinstance(val, tuple) else rtn


def fp32_to_float16(val, float16_convertor):
    def convert(v):
        t = v.data if isinstance(v, (Parameter, torch.autograd.Variable)) else v
        if isinstance(t, torch.Tensor) and t.dtype in _FLOAT_TYPES:
            v = float16_convertor(v)
        return v
    return conversion_helper(val, convert)


def float16_to_fp32(val):
    def convert(v):
        t = v.data if isinstance(v, (Parameter, torch.autograd.Variable)) else v
        if isinstance(t, torch.Tensor) and t.dtype in _HALF_TYPES + _BF16_TYPES:
            v = v.float()
        return v
    return conversion_helper(val, convert)


class Float16Module(MegatronModule):
    """Holds the model in fp16/bf16; casts inputs on the first pipeline stage and outputs (loss/logits) to fp32
    on the last."""

    def __init__(self, module, args):
        super().__init__()
        if args.fp16:
            self.add_module("module", module.half())
            self.float16_convertor = lambda v: v.half()
        elif args.bf16:
            self.add_module("module", module.bfloat16())
            self.float16_convertor = lambda v: v.bfloat16()
        else:
            raise Exception("should not be here")

    def set_input_tensor(self, input_tensor):
        return self.module.set_input_tensor(input_tensor)

    def forward(self, *inputs, **kwargs):
        if ps.is_pipeline_first_stage():
            inputs = fp32_to_float16(inputs, self.float16_convertor)
        outputs = self.module(*inputs, **kwargs)
        if ps.is_pipeline_last_stage():
            outputs = float16_to_fp32(outputs)
        return outputs

    def state_dict(self, prefix="", keep_vars=False):
        return self.module.state_dict(prefix=prefix, keep_vars=keep_vars)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def load_state_dict(self, state_dict, strict=True):
        self.module.load_state_dict(state_dict, strict=strict)

"""BERT (masked LM + optional next-sentence head).  Parity: megatron/model/bert_model.py:21-242.

Checkpoint keys: ``language_model``, ``lm_head`` {dense, layernorm, bias}, ``binary_head``,
``word_embeddings_for_head`` (last PP stage)."""
from __future__ import annotations

import torch

from ..parallel import layers as tp_layers
from ..parallel.cross_entropy import vocab_parallel_cross_entropy
from .activations import erf_gelu, get_linear_layer, init_method_normal, scaled_init_method_normal
from .enums import AttnMaskType
from .language_model import get_language_model, parallel_lm_logits
from .module import MegatronModule
from .norms import MixedFusedLayerNorm as LayerNorm


def _args():
    from ..global_vars import get_args
    return get_args()


def bert_extended_attention_mask(attention_mask):
    """[b, s] keep-mask -> [b, 1, s, s] boolean mask, True = masked out."""
    m = attention_mask.unsqueeze(1) * attention_mask.unsqueeze(2)
    return (m.unsqueeze(1) < 0.5)


def bert_position_ids(token_ids):
    s = token_ids.size(1)
    return torch.arange(s, dtype=torch.long, device=token_ids.device).unsqueeze(0).expand_as(token_ids)


class BertLMHead(MegatronModule):
    """dense -> gelu -> layernorm -> tied vocab projection + vocab-parallel bias."""

    def __init__(self, mpu_vocab_size, hidden_size, init_method, layernorm_epsilon, parallel_output, args=None):
        super().__init__()
        args = args or _args()
        self.bias = torch.nn.Parameter(torch.zeros(mpu_vocab_size))
        tp_layers.set_tensor_model_parallel_attributes(self.bias, True, 0, 1)
        self.parallel_output = parallel_output
        self.dense = get_linear_layer(hidden_size, hidden_size, init_method, args.perform_initialization)
        self.dense.weight.sequence_parallel = args.sequence_parallel
        self.dense.bias.sequence_parallel = args.sequence_parallel
        self.layernorm = LayerNorm(hidden_size, eps=layernorm_epsilon, sequence_parallel=args.sequence_parallel)
        self.gelu = erf_gelu if args.onnx_safe else torch.nn.functional.gelu

    def forward(self, hidden_states, word_embeddings_weight):
        h = self.layernorm(self.gelu(self.dense(hidden_states)))
        return parallel_lm_logits(h, word_embeddings_weight, self.parallel_output, bias=self.bias)


def post_language_model_processing(lm_output, pooled_output, lm_head, binary_head, lm_labels, logit_weights,
                                   fp16_lm_cross_entropy):
    lm_logits = lm_head(lm_output, logit_weights)                               # [s, b, V/tp]
    binary_logits = binary_head(pooled_output) if binary_head is not None else None
    if lm_labels is None:
        return lm_logits.transpose(0, 1).contiguous(), binary_logits
    labels = lm_labels.transpose(0, 1).contiguous()
    if fp16_lm_cross_entropy:
        assert lm_logits.dtype == torch.half
    lm_loss = vocab_parallel_cross_entropy(lm_logits, labels)                   # fp32 accumulation inside
    return lm_loss.transpose(0, 1).contiguous(), binary_logits


class HeadsMixin:
    """state-dict plumbing shared by the models that bolt heads on a ``language_model``: subclasses list
    ``(checkpoint key, attribute name, uses state_dict_for_save_checkpoint)`` in ``_head_specs()``."""

    def _head_specs(self):
        return []

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        sd = {self._language_model_key: self.language_model.state_dict_for_save_checkpoint(prefix=prefix,
                                                                                         keep_vars=keep_vars)}
        for key, attr, custom in self._head_specs():
            mod = getattr(self, attr)
            sd[key] = (mod.state_dict_for_save_checkpoint if custom else mod.state_dict)(prefix=prefix,
                                                                                        keep_vars=keep_vars)
        return sd

    def load_state_dict(self, state_dict, strict=True):
        self.language_model.load_state_dict(state_dict[self._language_model_key], strict=strict)
        for key, attr, _ in self._head_specs():
            if key in state_dict:
                getattr(self, attr).load_state_dict(state_dict[key], strict=strict)
            elif strict and key != "word_embeddings_for_head":
                print(f"***WARNING*** could not find {key} in the checkpoint, initializing to random")


class BertModel(HeadsMixin, MegatronModule):
    def __init__(self, num_tokentypes: int = 2, add_binary_head=True, parallel_output=True, pre_process=True,
                 post_process=True, model_type=None, args=None):
        super().__init__()
        args = args or _args()
        self.fp16_lm_cross_entropy = args.fp16_lm_cross_entropy
        self.add_binary_head, self.parallel_output = add_binary_head, parallel_output
        self.pre_process, self.post_process = pre_process, post_process
        init_method = init_method_normal(args.init_method_std)
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=add_binary_head, encoder_attn_mask_type=AttnMaskType.padding,
            init_method=init_method, scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
        self.initialize_word_embeddings(init_method_normal, args)
        if post_process:
            self.lm_head = BertLMHead(self.word_embeddings_weight().size(0), args.hidden_size, init_method,
                                      args.layernorm_epsilon, parallel_output, args)
            self.binary_head = get_linear_layer(args.hidden_size, 2, init_method, args.perform_initialization) \
                if add_binary_head else None

    def _head_specs(self):
        specs = []
        if self.post_process:
            specs.append(("lm_head", "lm_head", True))
            if self.add_binary_head:
                specs.append(("binary_head", "binary_head", False))
            if not self.pre_process:
                specs.append(("word_embeddings_for_head", "word_embeddings", False))
        return specs

    def set_input_tensor(self, input_tensor):
        self.language_model.set_input_tensor(input_tensor)

    def forward(self, bert_model_input, attention_mask, tokentype_ids=None, lm_labels=None):
        lm_output = self.language_model(bert_model_input, bert_position_ids(bert_model_input),
                                        bert_extended_attention_mask(attention_mask), tokentype_ids=tokentype_ids)
        pooled = None
        if self.post_process and self.add_binary_head:
            lm_output, pooled = lm_output
        if not self.post_process:
            return lm_output
        return post_language_model_processing(lm_output, pooled, self.lm_head, self.binary_head, lm_labels,
                                              self.word_embeddings_weight(), self.fp16_lm_cross_entropy)

"""GPT-family model (parity: megatron/model/gpt_model.py:18-123).

``forward(input_ids, position_ids, attention_mask, labels=None, ...)`` returns on the last pipeline stage:
labels given -> ``(loss [b, s], logits [b, s, V/tp])`` if ``return_logits`` else ``loss``; no labels ->
logits ``[b, s, V/tp]``.  Other stages return their hidden states.
"""
from __future__ import annotations

import torch

from ..parallel import state as ps
from ..parallel.cross_entropy import vocab_parallel_cross_entropy
from .activations import init_method_normal, scaled_init_method_normal
from .enums import AttnMaskType
from .language_model import get_language_model, parallel_lm_logits
from .module import MegatronModule


def post_language_model_processing(lm_output, labels, logit_weights, parallel_output, fp16_lm_cross_entropy):
    output = parallel_lm_logits(lm_output, logit_weights, parallel_output)   # [s, b, V/tp]
    if labels is None:
        return output.transpose(0, 1).contiguous()                         # [b, s, V/tp]
    labels_sb = labels.transpose(0, 1).contiguous()                        # [s, b]
    if fp16_lm_cross_entropy:
        assert output.dtype == torch.half
        loss = vocab_parallel_cross_entropy(output, labels_sb)
    else:
        # the fused CE kernel reads the bf16/fp16 logits directly and accumulates in fp32: no fp32 copy of
        # [s, b, V/tp] is ever materialised (the reference calls ``output.float()`` first)
        loss = vocab_parallel_cross_entropy(output, labels_sb)
    loss = loss.transpose(0, 1).contiguous()                               # [b, s]
    return loss, output.transpose(0, 1)


class GPTModel(MegatronModule):
    """GPT-2 style language model (also the base of Llama / Falcon / Mistral)."""

    def __init__(self, num_tokentypes=0, parallel_output=True, pre_process=True, post_process=True, args=None,
                 model_type=None):
        if args is None:
            from ..global_vars import get_args
            args = get_args()
        super().__init__(share_word_embeddings=args.tie_embed_logits)
        self.tie_embed_logits = args.tie_embed_logits
        self.parallel_output = parallel_output
        self.pre_process, self.post_process = pre_process, post_process
        self.fp16_lm_cross_entropy = args.fp16_lm_cross_entropy
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=False, encoder_attn_mask_type=AttnMaskType.causal,
            init_method=init_method_normal(args.init_method_std),
            scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=self.pre_process, post_process=self.post_process, args=args, model_type=model_type)
        if self.tie_embed_logits:
            self.initialize_word_embeddings(init_method_normal, args)

    def set_input_tensor(self, input_tensor):
        self.language_model.set_input_tensor(input_tensor)

    def forward(self, input_ids, position_ids, attention_mask, labels=None, tokentype_ids=None,
                inference_params=None):
        lm_output = self.language_model(input_ids, position_ids, attention_mask, inference_params=inference_params)
        if self.post_process:
            weights = self.word_embeddings_weight() if self.tie_embed_logits else self.language_model.lm_head
            return post_language_model_processing(lm_output, labels, weights, self.parallel_output,
                                                  self.fp16_lm_cross_entropy)
        return lm_output

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        sd = {self._language_model_key: self.language_model.state_dict_for_save_checkpoint(prefix=prefix,
                                                                                         keep_vars=keep_vars)}
        if self.post_process and not self.pre_process and self.tie_embed_logits:
            sd[self._word_embeddings_for_head_key] = self.word_embeddings.state_dict(prefix=prefix,
                                                                                    keep_vars=keep_vars)
        return sd

    def load_state_dict(self, state_dict, strict=True):
        if self.post_process and not self.pre_process and self.tie_embed_logits:
            self.word_embeddings.load_state_dict(state_dict[self._word_embeddings_for_head_key], strict=strict)
        if self._language_model_key in state_dict:
            state_dict = state_dict[self._language_model_key]
        self.language_model.load_state_dict(state_dict, strict=strict)

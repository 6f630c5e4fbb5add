"""Sequence classification head on the BERT encoder.  Parity: megatron/model/classification.py:17-107."""
from __future__ import annotations

import torch

from .activations import get_linear_layer, init_method_normal, scaled_init_method_normal
from .bert_model import HeadsMixin, _args, bert_extended_attention_mask, bert_position_ids
from .enums import AttnMaskType
from .language_model import get_language_model
from .module import MegatronModule


class Classification(HeadsMixin, MegatronModule):
    def __init__(self, num_classes, num_tokentypes=2, pre_process=True, post_process=True, model_type=None, args=None):
        super().__init__(share_word_embeddings=False)
        args = args or _args()
        self.num_classes = num_classes
        self.pre_process, self.post_process = pre_process, post_process
        init_method = init_method_normal(args.init_method_std)
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=True, encoder_attn_mask_type=AttnMaskType.padding,
            init_method=init_method, scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
        if post_process:
            self.classification_dropout = torch.nn.Dropout(args.hidden_dropout)
            self.classification_head = get_linear_layer(args.hidden_size, num_classes, init_method,
                                                        args.perform_initialization)

    def _head_specs(self):
        return [("classification_head", "classification_head", False)] if self.post_process else []

    def set_input_tensor(self, input_tensor):
        self.language_model.set_input_tensor(input_tensor)

    def forward(self, model_input, attention_mask, tokentype_ids=None):
        lm_output = self.language_model(model_input, bert_position_ids(model_input),
                                        bert_extended_attention_mask(attention_mask), tokentype_ids=tokentype_ids)
        if not self.post_process:
            return lm_output
        _, pooled = lm_output
        logits = self.classification_head(self.classification_dropout(pooled))
        return logits.view(-1, self.num_classes)

"""Multiple-choice head (RACE): one logit per choice.  Parity: megatron/model/multiple_choice.py:18-120."""
from __future__ import annotations

import torch

from .activations import get_linear_layer, init_method_normal, scaled_init_method_normal
from .bert_model import HeadsMixin, _args, bert_extended_attention_mask, bert_position_ids
from .enums import AttnMaskType
from .language_model import get_language_model
from .module import MegatronModule


class MultipleChoice(HeadsMixin, MegatronModule):
    def __init__(self, num_tokentypes=2, pre_process=True, post_process=True, model_type=None, args=None):
        super().__init__(share_word_embeddings=False)
        args = args or _args()
        self.pre_process, self.post_process = pre_process, post_process
        init_method = init_method_normal(args.init_method_std)
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=True, encoder_attn_mask_type=AttnMaskType.padding,
            init_method=init_method, scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
        if post_process:
            self.multichoice_dropout = torch.nn.Dropout(args.hidden_dropout)
            self.multichoice_head = get_linear_layer(args.hidden_size, 1, init_method, args.perform_initialization)

    def _head_specs(self):
        return [("multichoice_head", "multichoice_head", False)] if self.post_process else []

    def set_input_tensor(self, input_tensor):
        self.language_model.set_input_tensor(input_tensor)

    def forward(self, model_input, attention_mask, tokentype_ids=None):
        """inputs are [batch, choices, seq]; choices are folded into the batch for the encoder."""
        num_choices = model_input.shape[1]
        flat = lambda t: t.view(-1, t.size(-1)) if t is not None else None   # noqa: E731
        input_ids, attention_mask, tokentype_ids = flat(model_input), flat(attention_mask), flat(tokentype_ids)
        lm_output = self.language_model(input_ids, bert_position_ids(input_ids),
                                        bert_extended_attention_mask(attention_mask), tokentype_ids=tokentype_ids)
        if not self.post_process:
            return lm_output
        _, pooled = lm_output
        return self.multichoice_head(self.multichoice_dropout(pooled)).view(-1, num_choices)

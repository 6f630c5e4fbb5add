"""T5 encoder-decoder.  Parity: megatron/model/t5_model.py:20-198 (``ModelType.encoder_and_decoder``; with PP the
encoder occupies the stages before ``--pipeline_model_parallel_split_rank``)."""
from __future__ import annotations

import torch

from ..parallel.cross_entropy import vocab_parallel_cross_entropy
from .activations import init_method_normal, scaled_init_method_normal
from .bert_model import HeadsMixin, _args, bert_position_ids as t5_position_ids
from .enums import AttnMaskType
from .language_model import get_language_model, parallel_lm_logits
from .module import MegatronModule


def t5_extended_attention_mask(attention_mask_list):
    """[b, s, s] -> [b, 1, s, s] for each mask."""
    return [m.unsqueeze(1) if m is not None else None for m in attention_mask_list]


class T5LMHead(MegatronModule):
    """Tied vocab projection + vocab-parallel bias."""

    def __init__(self, mpu_vocab_size, parallel_output):
        super().__init__()
        self.bias = torch.nn.Parameter(torch.zeros(mpu_vocab_size))
        self.bias.model_parallel, self.bias.partition_dim, self.bias.stride = True, 0, 1
        self.bias.tensor_model_parallel, self.bias.partition_stride = True, 1
        self.parallel_output = parallel_output

    def forward(self, hidden_states, word_embeddings_weight):
        return parallel_lm_logits(hidden_states, word_embeddings_weight, self.parallel_output, bias=self.bias)


class T5Model(HeadsMixin, MegatronModule):
    def __init__(self, num_tokentypes=0, parallel_output=True, pre_process=True, post_process=True, add_encoder=True,
                 add_decoder=True, model_type=None, args=None):
        super().__init__()
        args = args or _args()
        self.fp16_lm_cross_entropy = args.fp16_lm_cross_entropy
        self.parallel_output = parallel_output
        self.pre_process, self.post_process = pre_process, post_process
        self.add_encoder, self.add_decoder = add_encoder, add_decoder
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=False, add_encoder=add_encoder, add_decoder=add_decoder,
            encoder_attn_mask_type=AttnMaskType.padding, init_method=init_method_normal(args.init_method_std),
            scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
        self.initialize_word_embeddings(init_method_normal, args)
        if post_process and add_decoder:
            self.lm_head = T5LMHead(self.word_embeddings_weight().size(0), parallel_output)

    def _head_specs(self):
        specs = []
        if self.post_process and self.add_decoder:
            specs.append(("lm_head", "lm_head", True))
            if not self.pre_process:
                specs.append(("word_embeddings_for_head", "word_embeddings", False))
        return specs

    def set_input_tensor(self, input_tensor):
        self.language_model.set_input_tensor(input_tensor)

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                encoder_decoder_attn_mask, tokentype_ids=None, lm_labels=None, enc_hidden_states=None):
        enc_mask, dec_mask, cross_mask = t5_extended_attention_mask(
            [encoder_attn_mask, decoder_attn_mask, encoder_decoder_attn_mask])
        lm_output = self.language_model(encoder_input_ids, t5_position_ids(encoder_input_ids), enc_mask,
                                        decoder_input_ids, t5_position_ids(decoder_input_ids), dec_mask, cross_mask,
                                        tokentype_ids=tokentype_ids, enc_hidden_states=enc_hidden_states)
        if self.post_process and self.add_decoder:
            decoder_output, _ = lm_output
            lm_logits = self.lm_head(decoder_output, self.word_embeddings_weight())   # [s, b, V/tp]
            if lm_labels is None:
                return lm_logits.transpose(0, 1).contiguous()
            if self.fp16_lm_cross_entropy:
                assert lm_logits.dtype == torch.half
            loss = vocab_parallel_cross_entropy(lm_logits, lm_labels.transpose(0, 1).contiguous())
            return loss.transpose(0, 1).contiguous()
        if self.add_decoder and not self.add_encoder:
            return lm_output[0]
        return lm_output

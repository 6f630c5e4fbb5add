"""ICT / REALM bi-encoder: BERT-based query and context encoders producing [CLS] embeddings.

Parity: megatron/model/biencoder_model.py:19-345.  Checkpoint keys: ``shared_model`` or ``query_model`` /
``context_model``, each {``language_model``, ``projection_enc``}."""
from __future__ import annotations

import os
import sys

import torch

from ..parallel import state as ps
from .activations import get_linear_layer, init_method_normal, scaled_init_method_normal
from .bert_model import HeadsMixin, _args, bert_position_ids
from .enums import AttnMaskType
from .language_model import get_language_model
from .module import MegatronModule


def _print0(msg):
    from ..utils import print_rank_0
    print_rank_0(msg)


def biencoder_model_provider(only_query_model=False, only_context_model=False,
                             biencoder_shared_query_context_model=False, pre_process=True, post_process=True,
                             model_type=None):
    assert ps.get_tensor_model_parallel_world_size() == 1 and ps.get_pipeline_model_parallel_world_size() == 1, \
        "Model parallel size > 1 not supported for ICT"
    _print0("building BiEncoderModel...")
    return BiEncoderModel(num_tokentypes=2, parallel_output=False, only_query_model=only_query_model,
                          only_context_model=only_context_model,
                          biencoder_shared_query_context_model=biencoder_shared_query_context_model,
                          pre_process=pre_process, post_process=post_process, model_type=model_type)


def get_model_provider(only_query_model=False, only_context_model=False, biencoder_shared_query_context_model=False,
                       model_type=None):
    def model_provider(pre_process=True, post_process=True):
        _print0("building Bienoder model ...")
        return biencoder_model_provider(only_query_model, only_context_model, biencoder_shared_query_context_model,
                                        pre_process, post_process, model_type)
    return model_provider


class PretrainedBertModel(HeadsMixin, MegatronModule):
    """BERT encoder whose output is the (optionally projected) hidden state of the first token."""

    def __init__(self, num_tokentypes=2, parallel_output=True, pre_process=True, post_process=True, model_type=None,
                 args=None):
        super().__init__(share_word_embeddings=False)
        args = args or _args()
        from ..global_vars import get_tokenizer
        self.pad_id = get_tokenizer().pad
        self.biencoder_projection_dim = args.biencoder_projection_dim
        self.parallel_output = parallel_output
        self.pre_process, self.post_process = pre_process, post_process
        init_method = init_method_normal(args.init_method_std)
        self.language_model, self._language_model_key = get_language_model(
            num_tokentypes=num_tokentypes, add_pooler=False, encoder_attn_mask_type=AttnMaskType.padding,
            init_method=init_method, scaled_init_method=scaled_init_method_normal(args.init_method_std, args.num_layers),
            pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
        if self.biencoder_projection_dim > 0:
            self.projection_enc = get_linear_layer(args.hidden_size, self.biencoder_projection_dim, init_method,
                                                   args.perform_initialization)

    def _head_specs(self):
        return [("projection_enc", "projection_enc", False)] if self.biencoder_projection_dim > 0 else []

    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        lm_output = self.language_model(input_ids, bert_position_ids(input_ids), attention_mask.unsqueeze(1),
                                        tokentype_ids=tokentype_ids)
        pooled = lm_output[0, :, :]                       # [CLS] of every sample ([s, b, h] layout)
        if self.biencoder_projection_dim:
            pooled = self.projection_enc(pooled)
        return pooled


class BiEncoderModel(MegatronModule):
    def __init__(self, num_tokentypes=1, parallel_output=True, only_query_model=False, only_context_model=False,
                 biencoder_shared_query_context_model=False, pre_process=True, post_process=True, model_type=None,
                 args=None):
        super().__init__(share_word_embeddings=False)
        args = args or _args()
        kw = dict(num_tokentypes=num_tokentypes, parallel_output=parallel_output, pre_process=pre_process,
                  post_process=post_process, model_type=model_type, args=args)
        assert not (only_context_model and only_query_model)
        self.biencoder_shared_query_context_model = biencoder_shared_query_context_model
        self.use_context_model, self.use_query_model = not only_query_model, not only_context_model
        self.biencoder_projection_dim = args.biencoder_projection_dim
        self.query_model = self.context_model = None
        if biencoder_shared_query_context_model:
            self.model = PretrainedBertModel(**kw)
            self._model_key = "shared_model"
            self.query_model = self.context_model = self.model
        else:
            if self.use_query_model:
                self.query_model = PretrainedBertModel(**kw)
                self._query_key = "query_model"
            if self.use_context_model:
                self.context_model = PretrainedBertModel(**kw)
                self._context_key = "context_model"

    def set_input_tensor(self, input_tensor):
        return

    def forward(self, query_tokens, query_attention_mask, query_types, context_tokens, context_attention_mask,
                context_types):
        if not self.use_query_model:
            raise ValueError("Cannot embed query without the query model.")
        if not self.use_context_model:
            raise ValueError("Cannot embed block without the block model.")
        return (self.embed_text(self.query_model, query_tokens, query_attention_mask, query_types),
                self.embed_text(self.context_model, context_tokens, context_attention_mask, context_types))

    @staticmethod
    def embed_text(model, tokens, attention_mask, token_types):
        return model(tokens, attention_mask, token_types)

    def _parts(self):
        if self.biencoder_shared_query_context_model:
            return [(self._model_key, self.model)]
        parts = []
        if self.use_query_model:
            parts.append((self._query_key, self.query_model))
        if self.use_context_model:
            parts.append((self._context_key, self.context_model))
        return parts

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return {k: m.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars) for k, m in self._parts()}

    def load_state_dict(self, state_dict, strict=True):
        for k, m in self._parts():
            _print0(f"Loading {k}")
            m.load_state_dict(state_dict[k], strict=strict)

    def init_state_dict_from_bert(self):
        """Initialise both towers from a pretrained BERT checkpoint (``--bert_load``) at iteration 0 of ICT."""
        from ..checkpointing import (fix_query_key_value_ordering, get_checkpoint_names,
                                     get_checkpoint_tracker_filename)
        args = _args()
        if args.bert_load is None:
            _print0("bert_load argument is None")
            return
        tracker = get_checkpoint_tracker_filename(args.bert_load)
        if not os.path.isfile(tracker):
            raise FileNotFoundError("Could not find BERT checkpoint")
        with open(tracker) as f:
            iteration = int(f.read().strip())
            assert iteration > 0
        name, _ = get_checkpoint_names(args.bert_load, iteration, False)
        try:
            state_dict = torch.load(name, map_location="cpu", weights_only=False)
        except ModuleNotFoundError:
            from ..checkpointing import legacy_pickle_modules
            with legacy_pickle_modules():
                state_dict = torch.load(name, map_location="cpu", weights_only=False)
        version = state_dict.get("checkpoint_version", 0)
        lm = state_dict["model"]["language_model"]
        towers = [self.model] if self.biencoder_shared_query_context_model else \
            [m for m in (self.query_model, self.context_model) if m is not None]
        for m in towers:
            m.language_model.load_state_dict(lm)
            fix_query_key_value_ordering(m, version)
        if not self.biencoder_shared_query_context_model and self.biencoder_projection_dim > 0 \
                and self.query_model is not None and self.context_model is not None:
            self.context_model.projection_enc.load_state_dict(self.query_model.projection_enc.state_dict())

"""Embedding + transformer stack + (untied) LM head.

Parity target: megatron/model/language_model.py -- parallel_lm_logits :24-53, get_language_model :56-94,
Pooler :97-130, Embedding :133-326, TransformerLanguageModel :329-638 (incl. checkpoint keys
``embedding/encoder/decoder/pooler/lm_head`` and the legacy-key loader :582-638).
"""
from __future__ import annotations

import math

from typing import Callable

import torch
from torch import nn

from ..parallel import layers as tp_layers
from ..parallel import mappings
from ..parallel import state as ps
from ..parallel.layers import _initialize_affine_weight_cpu, _initialize_affine_weight_gpu
from ..parallel.random import get_cuda_rng_tracker
from ..parallel.tp_utils import VocabUtility
from ..utils.device import current_device
from .activations import get_linear_layer, init_method_normal, scaled_init_method_normal
from .enums import AttnMaskType, LayerType, PositionEmbeddingType
from .module import MegatronModule
from .transformer import ParallelTransformer


def _get_args():
    from ..global_vars import get_args
    return get_args()


def parallel_lm_logits(input_, word_embeddings_weight, parallel_output, bias=None):
    """LM head: the same fused linear as ColumnParallelLinear (SP all-gather -> GEMM) over the vocab shard."""
    args = _get_args()
    if args.async_tensor_model_parallel_allreduce or args.sequence_parallel:
        input_parallel = input_
        model_parallel = ps.get_tensor_model_parallel_world_size() > 1
        async_grad_allreduce = args.async_tensor_model_parallel_allreduce and model_parallel and \
            not args.sequence_parallel
    else:
        input_parallel = mappings.copy_to_tensor_model_parallel_region(input_)
        async_grad_allreduce = False
    logits_parallel = tp_layers.linear_with_grad_accumulation_and_async_allreduce(
        input=input_parallel, weight=word_embeddings_weight, bias=bias,
        gradient_accumulation_fusion=args.gradient_accumulation_fusion,
        async_grad_allreduce=async_grad_allreduce, sequence_parallel_enabled=args.sequence_parallel)
    if parallel_output:
        return logits_parallel
    return mappings.gather_from_tensor_model_parallel_region(logits_parallel)


def get_language_model(num_tokentypes, add_pooler: bool, encoder_attn_mask_type, init_method=None,
                       scaled_init_method=None, add_encoder=True, add_decoder=False,
                       decoder_attn_mask_type=AttnMaskType.causal, pre_process=True, post_process=True, args=None,
                       model_type=None):
    assert args is not None
    if init_method is None:
        init_method = init_method_normal(args.init_method_std)
    if scaled_init_method is None:
        scaled_init_method = scaled_init_method_normal(args.init_method_std, args.num_layers)
    language_model = TransformerLanguageModel(
        init_method, scaled_init_method, encoder_attn_mask_type, num_tokentypes=num_tokentypes,
        add_encoder=add_encoder, add_decoder=add_decoder, decoder_attn_mask_type=decoder_attn_mask_type,
        add_pooler=add_pooler, pre_process=pre_process, post_process=post_process, args=args, model_type=model_type)
    return language_model, "language_model"


class Pooler(MegatronModule):
    """tanh(W h[sequence_index]) over the (gathered) sequence."""

    def __init__(self, hidden_size, init_method, args):
        super().__init__()
        self.dense = get_linear_layer(hidden_size, hidden_size, init_method, args.perform_initialization)
        self.sequence_parallel = args.sequence_parallel

    def forward(self, hidden_states, sequence_index=0):
        if self.sequence_parallel:
            hidden_states = mappings.gather_from_sequence_parallel_region(hidden_states,
                                                                          tensor_parallel_output_grad=False)
        pooled = hidden_states[sequence_index, :, :]
        return torch.tanh(self.dense(pooled))


class Embedding(MegatronModule):
    """word (vocab-parallel) + optional absolute position + optional token-type embeddings; output [s, b, h]
    (scattered along s under sequence parallelism) followed by dropout."""

    def __init__(self, hidden_size, vocab_size, max_position_embeddings, embedding_dropout_prob, init_method,
                 num_tokentypes=0):
        super().__init__()
        self.hidden_size, self.init_method, self.num_tokentypes = hidden_size, init_method, num_tokentypes
        args = _get_args()
        self.word_embeddings = tp_layers.VocabParallelEmbedding(
            vocab_size, hidden_size, init_method=init_method, params_dtype=args.params_dtype,
            gradient_accumulation_fusion=args.gradient_accumulation_fusion,
            use_cpu_initialization=args.use_cpu_initialization, perform_initialization=args.perform_initialization)
        self._word_embeddings_key = "word_embeddings"
        self.position_embedding_type = args.position_embedding_type
        if self.position_embedding_type == PositionEmbeddingType.absolute:
            assert max_position_embeddings is not None
            self.position_embeddings = torch.nn.Embedding(max_position_embeddings, hidden_size)
            self._position_embeddings_key = "position_embeddings"
            self.init_method(self.position_embeddings.weight)
        else:
            self.position_embeddings = None
        self._tokentype_embeddings_key = "tokentype_embeddings"
        if num_tokentypes > 0:
            self.tokentype_embeddings = torch.nn.Embedding(num_tokentypes, hidden_size)
            if args.perform_initialization:
                self.init_method(self.tokentype_embeddings.weight)
        else:
            self.tokentype_embeddings = None
        self.fp32_residual_connection = args.fp32_residual_connection
        self.sequence_parallel = args.sequence_parallel
        self.embedding_dropout = torch.nn.Dropout(embedding_dropout_prob)

    def zero_parameters(self):
        self.word_embeddings.weight.data.fill_(0)
        self.word_embeddings.weight.shared = True
        if self.position_embedding_type == PositionEmbeddingType.absolute:
            self.position_embeddings.weight.data.fill_(0)
            self.position_embeddings.weight.shared = True
        if self.num_tokentypes > 0:
            self.tokentype_embeddings.weight.data.fill_(0)
            self.tokentype_embeddings.weight.shared = True

    def add_tokentype_embeddings(self, num_tokentypes):
        if self.tokentype_embeddings is not None:
            raise Exception("tokentype embeddings is already initialized")
        if torch.distributed.is_initialized() and torch.distributed.get_rank() == 0:
            print("adding embedding for {} tokentypes".format(num_tokentypes), flush=True)
        self.num_tokentypes = num_tokentypes
        self.tokentype_embeddings = torch.nn.Embedding(num_tokentypes, self.hidden_size)
        self.init_method(self.tokentype_embeddings.weight)

    def forward(self, input_ids, position_ids, tokentype_ids=None):
        drop = self.training and self.embedding_dropout.p > 0.0      # p = 0: no RNG use (and CUDA-graph capturable)
        if (self.position_embeddings is None and tokentype_ids is None and not self.fp32_residual_connection
                and input_ids.dim() == 2):
            # word embeddings only (Llama / Falcon / Mistral): gather straight into the [s, b, h] layout and, under
            # sequence parallelism, reduce-scatter the partial lookups instead of all-reduce + split
            embeddings = self.word_embeddings.forward_sbh(input_ids, self.sequence_parallel)
            if drop:
                if self.sequence_parallel:
                    with get_cuda_rng_tracker().fork():
                        embeddings = self.embedding_dropout(embeddings)
                else:
                    embeddings = self.embedding_dropout(embeddings)
            return embeddings
        embeddings = self.word_embeddings(input_ids)
        if self.position_embedding_type == PositionEmbeddingType.absolute:
            assert self.position_embeddings is not None
            embeddings = embeddings + self.position_embeddings(position_ids)
        else:
            assert self.position_embeddings is None
        if tokentype_ids is not None:
            assert self.tokentype_embeddings is not None
            embeddings = embeddings + self.tokentype_embeddings(tokentype_ids)
        else:
            assert self.tokentype_embeddings is None
        embeddings = embeddings.transpose(0, 1).contiguous()  # [b,s,h] -> [s,b,h]
        if self.fp32_residual_connection:
            embeddings = embeddings.float()
        if self.sequence_parallel:
            embeddings = mappings.scatter_to_sequence_parallel_region(embeddings)
            if drop:
                with get_cuda_rng_tracker().fork():
                    embeddings = self.embedding_dropout(embeddings)
        elif drop:
            embeddings = self.embedding_dropout(embeddings)
        return embeddings

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        sd = {self._word_embeddings_key: self.word_embeddings.state_dict(prefix=prefix, keep_vars=keep_vars)}
        if self.position_embedding_type == PositionEmbeddingType.absolute:
            sd[self._position_embeddings_key] = self.position_embeddings.state_dict(prefix=prefix, keep_vars=keep_vars)
        if self.num_tokentypes > 0:
            sd[self._tokentype_embeddings_key] = self.tokentype_embeddings.state_dict(prefix=prefix,
                                                                                      keep_vars=keep_vars)
        return sd

    def load_state_dict(self, state_dict, strict=True):
        def pick(key, marker):
            if key in state_dict:
                return state_dict[key]
            # flat legacy names, e.g. "word_embeddings.weight"
            return {k.split(marker + ".")[1]: v for k, v in state_dict.items() if marker in k}

        self.word_embeddings.load_state_dict(pick(self._word_embeddings_key, "word_embeddings"), strict=strict)
        if self.position_embedding_type == PositionEmbeddingType.absolute:
            self.position_embeddings.load_state_dict(pick(self._position_embeddings_key, "position_embeddings"),
                                                     strict=strict)
        if self.num_tokentypes > 0:
            sd = pick(self._tokentype_embeddings_key, "tokentype_embeddings")
            if len(sd) > 0:
                self.tokentype_embeddings.load_state_dict(sd, strict=strict)
            else:
                print("***WARNING*** expected tokentype embeddings in the checkpoint but could not find it",
                      flush=True)


class TransformerLanguageModel(MegatronModule):
    def __init__(self, init_method: Callable, output_layer_init_method, encoder_attn_mask_type, num_tokentypes=0,
                 add_encoder=True, add_decoder=False, decoder_attn_mask_type=AttnMaskType.causal, add_pooler=False,
                 pre_process=True, post_process=True, args=None, model_type=None):
        super().__init__()
        assert args is not None
        self.pre_process, self.post_process = pre_process, post_process
        self.hidden_size = args.hidden_size
        self.num_tokentypes = num_tokentypes
        self.init_method = init_method
        self.add_encoder, self.add_decoder, self.add_pooler = add_encoder, add_decoder, add_pooler
        self.encoder_attn_mask_type, self.decoder_attn_mask_type = encoder_attn_mask_type, decoder_attn_mask_type
        self.encoder_hidden_state = None
        self.flop_estimate = self._estimate_flops(args)

        if self.pre_process:
            self.embedding = Embedding(self.hidden_size, args.padded_vocab_size, args.max_position_embeddings,
                                       args.hidden_dropout if not args.lima_dropout else 0.0, self.init_method,
                                       self.num_tokentypes)
            self._embedding_key = "embedding"
        if self.add_encoder:
            self.encoder = ParallelTransformer(self.init_method, output_layer_init_method,
                                               self_attn_mask_type=self.encoder_attn_mask_type,
                                               pre_process=self.pre_process, post_process=self.post_process, args=args,
                                               model_type=model_type)
            self._encoder_key = "encoder"
        else:
            self.encoder = None
        if self.add_decoder:
            self.decoder = ParallelTransformer(self.init_method, output_layer_init_method,
                                               layer_type=LayerType.decoder,
                                               self_attn_mask_type=self.decoder_attn_mask_type,
                                               pre_process=self.pre_process, post_process=self.post_process, args=args,
                                               model_type=model_type)
            self._decoder_key = "decoder"
        else:
            self.decoder = None
        if self.post_process and self.add_pooler:
            self.pooler = Pooler(self.hidden_size, self.init_method, args)
            self._pooler_key = "pooler"

        self.tie_embed_logits = args.tie_embed_logits
        if self.post_process and not self.tie_embed_logits:
            start, end = VocabUtility.vocab_range_from_global_vocab_size(
                args.padded_vocab_size, ps.get_tensor_model_parallel_rank(), args.tensor_model_parallel_size)
            num_embeds = end - start
            data = torch.empty(num_embeds, self.hidden_size, dtype=args.params_dtype,
                               device=None if args.use_cpu_initialization else current_device())
            self.lm_head = nn.Parameter(data)
            self._lm_key = "lm_head"
            head_init = nn.init.xavier_uniform_ if args.init_method_xavier_uniform else nn.init.xavier_normal_
            # Xavier scales with fan_in + fan_out.  The reference applies it to the TP *shard* (:448-457), so its LM-head
            # std grows with the TP size (0.0074 at TP=1 -> 0.0156 at TP=8 for Llama-2-7B): the initial loss and the
            # whole loss curve of a from-scratch run then depend on the parallel layout.  Here the fans are those of
            # the full [padded_vocab, hidden] matrix, whatever the layout.
            full_fans = float(args.padded_vocab_size + self.hidden_size)

            def head_init_shard(w, _init=head_init, _fans=full_fans):
                return _init(w, gain=math.sqrt((w.size(0) + w.size(1)) / _fans))
            if args.perform_initialization:
                if args.use_cpu_initialization:
                    _initialize_affine_weight_cpu(self.lm_head, args.padded_vocab_size, self.hidden_size, num_embeds,
                                                  0, head_init, params_dtype=args.params_dtype)
                else:
                    _initialize_affine_weight_gpu(self.lm_head, head_init_shard, partition_dim=0, stride=1)
            else:
                tp_layers.set_tensor_model_parallel_attributes(self.lm_head, True, 0, 1)

    @staticmethod
    def _estimate_flops(args):
        """Analytical forward flop count per sequence (reference :370-384)."""
        s, ell, v, h = args.max_position_embeddings, args.num_layers, getattr(args, "padded_vocab_size", 0), \
            args.hidden_size
        mlp_mult = 64 if args.glu_activation else 16
        per_layer = 6 * s * h ** 2 + 2 * s ** 2 * h + 2 * s ** 2 * h + 2 * s * h ** 2 + mlp_mult * s * h ** 2
        return ell * per_layer + 6 * s * h * (v or 0)

    def set_input_tensor(self, input_tensor):
        if not isinstance(input_tensor, list):
            input_tensor = [input_tensor]
        if self.add_encoder and self.add_decoder:
            assert len(input_tensor) == 1, "input_tensor should only be length 1 for stage with both encoder and decoder"
            self.encoder.set_input_tensor(input_tensor[0])
        elif self.add_encoder:
            assert len(input_tensor) == 1, "input_tensor should only be length 1 for stage with only encoder"
            self.encoder.set_input_tensor(input_tensor[0])
        elif self.add_decoder:
            if len(input_tensor) == 2:
                self.decoder.set_input_tensor(input_tensor[0])
                self.encoder_hidden_state = input_tensor[1]
            elif len(input_tensor) == 1:
                self.decoder.set_input_tensor(None)
                self.encoder_hidden_state = input_tensor[0]
            else:
                raise Exception("input_tensor must have either length 1 or 2")
        else:
            raise Exception("Stage must have at least either encoder or decoder")

    def forward(self, enc_input_ids, enc_position_ids, enc_attn_mask, dec_input_ids=None, dec_position_ids=None,
                dec_attn_mask=None, enc_dec_attn_mask=None, tokentype_ids=None, inference_params=None,
                pooling_sequence_index=0, enc_hidden_states=None, output_enc_hidden=False):
        encoder_input = self.embedding(enc_input_ids, enc_position_ids, tokentype_ids=tokentype_ids) \
            if self.pre_process else None
        if enc_hidden_states is None:
            if self.encoder is not None:
                encoder_output = self.encoder(encoder_input, enc_attn_mask, inference_params=inference_params,
                                              position_ids=enc_position_ids)
            else:
                encoder_output = self.encoder_hidden_state
        else:
            encoder_output = enc_hidden_states.to(encoder_input.dtype)
        pooled_output = None
        if self.post_process and self.add_pooler:
            pooled_output = self.pooler(encoder_output, pooling_sequence_index)
        if not self.add_decoder or output_enc_hidden:
            if self.add_pooler and self.post_process:
                return encoder_output, pooled_output
            return encoder_output
        decoder_input = self.embedding(dec_input_ids, dec_position_ids) if self.pre_process else None
        decoder_output = self.decoder(decoder_input, dec_attn_mask, encoder_output=encoder_output,
                                      enc_dec_attn_mask=enc_dec_attn_mask, inference_params=inference_params)
        if self.add_pooler and self.post_process:
            return decoder_output, encoder_output, pooled_output
        return decoder_output, encoder_output

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        sd = {}
        if self.pre_process:
            sd[self._embedding_key] = self.embedding.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)
        if self.add_encoder:
            sd[self._encoder_key] = self.encoder.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)
        if self.post_process:
            if self.add_pooler:
                sd[self._pooler_key] = self.pooler.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)
            if not self.tie_embed_logits:
                sd[self._lm_key] = self.lm_head.data
        if self.add_decoder:
            sd[self._decoder_key] = self.decoder.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)
        return sd

    def load_state_dict(self, state_dict, strict=True):
        if self.pre_process:
            if self._embedding_key in state_dict:
                emb = state_dict[self._embedding_key]
            else:  # legacy flat layout
                emb = {k: v for k, v in state_dict.items() if "_embeddings" in k}
            self.embedding.load_state_dict(emb, strict=strict)
        if self.post_process and not self.tie_embed_logits:
            self.lm_head.data.copy_(state_dict[self._lm_key])
        if self.add_encoder:
            if self._encoder_key in state_dict:
                enc = state_dict[self._encoder_key]
            elif "transformer" in state_dict:
                enc = state_dict["transformer"]
            else:
                enc = {k.split("transformer.")[1]: v for k, v in state_dict.items() if "transformer." in k}
            enc = {k.replace(".attention.", ".self_attention."): v for k, v in enc.items()}
            self.encoder.load_state_dict(enc, strict=strict)
        if self.post_process and self.add_pooler:
            assert "pooler" in state_dict, "could not find data for pooler in the checkpoint"
            self.pooler.load_state_dict(state_dict[self._pooler_key], strict=strict)
        if self.add_decoder:
            assert "decoder" in state_dict, "could not find data for decoder in the checkpoint"
            self.decoder.load_state_dict(state_dict[self._decoder_key], strict=strict)

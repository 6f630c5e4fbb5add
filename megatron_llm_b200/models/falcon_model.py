"""Falcon (parity: megatron/model/falcon_model.py:10-41): rotary + parallel attention/MLP + GeLU, no bias,
LayerNorm, tied embeddings; Falcon-40B adds ``parallel_layernorm`` and 8 KV heads."""
import warnings

from .enums import PositionEmbeddingType
from .gpt_model import GPTModel


class FalconModel(GPTModel):
    def __init__(self, num_tokentypes: int = 0, parallel_output: bool = True, pre_process: bool = True,
                 post_process: bool = True, model_type=None, args=None):
        if args is None:
            from ..global_vars import get_args
            args = get_args()
        assert args.position_embedding_type == PositionEmbeddingType.rotary, \
            f"Falcon uses rotary embedding, not {args.position_embedding_type}"
        assert isinstance(args.num_attention_heads_kv, int), "Falcon needs an explicit num_attention_heads_kv"
        assert not args.use_post_ln, "FalconModel requires pre-normalization, not use_post_ln"
        assert args.glu_activation is None, "FalconModel requires gelu activation (set glu_activation=None)"
        assert not args.use_bias, "Falcon does not use bias"
        assert args.parallel_attn, "Falcon uses parallel_attn"
        if not args.parallel_layernorm:
            warnings.warn("Falcon-40B uses parallel_layernorm (Falcon-7B does not)")
        if not args.use_flash_attn:
            warnings.warn("Falcon specifies --use_flash_attn; the unfused attention path will be used instead")
        if args.bias_gelu_fusion:
            warnings.warn("Falcon has no bias: disabling bias_gelu_fusion")
            args.bias_gelu_fusion = False
        if args.bias_dropout_fusion:
            args.bias_dropout_fusion = False
        if args.hidden_dropout != 0.0 and not args.lima_dropout:
            warnings.warn(f"Falcon is trained without hidden dropout, not {args.hidden_dropout}")
        super().__init__(num_tokentypes=num_tokentypes, parallel_output=parallel_output, pre_process=pre_process,
                         post_process=post_process, args=args, model_type=model_type)

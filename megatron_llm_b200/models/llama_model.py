"""Llama / Llama-2 / CodeLlama (parity: megatron/model/llama_model.py:11-43): a GPTModel whose flags must spell
rotary + SwiGLU + RMSNorm + no bias + untied embeddings."""
import warnings

from .enums import PositionEmbeddingType
from .gpt_model import GPTModel


class LlamaModel(GPTModel):
    def __init__(self, num_tokentypes: int = 0, parallel_output: bool = True, pre_process: bool = True,
                 post_process: bool = True, model_type=None, version: int = 2, args=None):
        if args is None:
            from ..global_vars import get_args
            args = get_args()
        assert version in {1, 2}, f"Unknown llama version {version}"
        assert args.position_embedding_type == PositionEmbeddingType.rotary, \
            f"Llama uses rotary embedding, not {args.position_embedding_type}"
        assert not args.use_post_ln, "Llama does not use post_ln"
        assert args.glu_activation == "swiglu", "Llama works with swiglu activation"
        assert not args.use_bias, "Llama does not use bias"
        assert not args.parallel_attn, "Llama does not use parallel_attn"
        assert args.use_rms_norm, "Llama uses rms_norm"
        assert not args.tie_embed_logits, "Llama unties embedding and lm_head weights"
        if not args.use_flash_attn:
            warnings.warn("Llama specifies --use_flash_attn; the unfused attention path will be used instead")
        if version == 1 and args.layernorm_epsilon != 1e-6:
            warnings.warn(f"Llama-1 uses layernorm_epsilon=1e-6, not {args.layernorm_epsilon}")
        if version == 2 and args.layernorm_epsilon != 1e-5:
            warnings.warn(f"Llama-2 uses layernorm_epsilon=1e-5, not {args.layernorm_epsilon}")
        if args.hidden_dropout != 0.0 and not args.lima_dropout:
            warnings.warn(f"Llama is trained without hidden dropout, not {args.hidden_dropout}")
        if args.attention_dropout != 0.0:
            warnings.warn(f"Llama is trained without attention dropout, not {args.attention_dropout}")
        super().__init__(num_tokentypes=num_tokentypes, parallel_output=parallel_output, pre_process=pre_process,
                         post_process=post_process, args=args, model_type=model_type)

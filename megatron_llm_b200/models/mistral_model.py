"""Mistral (parity: megatron/model/mistral_model.py:11-46): Llama flags + GQA + sliding window 4096."""
import warnings

from .enums import PositionEmbeddingType
from .gpt_model import GPTModel


class MistralModel(GPTModel):
    def __init__(self, num_tokentypes: int = 0, parallel_output: bool = True, pre_process: bool = True,
                 post_process: bool = True, model_type=None, args=None):
        if args is None:
            from ..global_vars import get_args
            args = get_args()
        assert args.position_embedding_type == PositionEmbeddingType.rotary, \
            f"Mistral uses rotary embedding, not {args.position_embedding_type}"
        assert not args.use_post_ln, "Mistral does not use post_ln"
        assert args.glu_activation == "swiglu", "Mistral works with swiglu activation"
        assert not args.use_bias, "Mistral does not use bias"
        assert not args.parallel_attn, "Mistral does not use parallel_attn"
        assert args.use_rms_norm, "Mistral uses rms_norm"
        assert not args.tie_embed_logits, "Mistral unties embedding and lm_head weights"
        assert args.sliding_window_size == 4096, "Mistral uses sliding window attention (sliding_window=4096)"
        if not args.use_flash_attn:
            warnings.warn("Mistral specifies --use_flash_attn; the unfused attention path will be used instead")
        if args.layernorm_epsilon != 1e-5:
            warnings.warn(f"Mistral uses layernorm_epsilon=1e-5, not {args.layernorm_epsilon}")
        if args.hidden_dropout != 0.0 and not args.lima_dropout:
            warnings.warn(f"Mistral is trained without hidden dropout, not {args.hidden_dropout}")
        if args.attention_dropout != 0.0:
            warnings.warn(f"Mistral is trained without attention dropout, not {args.attention_dropout}")
        super().__init__(num_tokentypes=num_tokentypes, parallel_output=parallel_output, pre_process=pre_process,
                         post_process=post_process, args=args, model_type=model_type)

"""Convert Hugging Face / Meta checkpoints (Llama, Llama-2, CodeLlama, Mistral, Falcon) to the Megatron layout.

Parity target: weights_conversion/hf_to_megatron.py.  Same rules:
  * fused QKV rows grouped per KV head ``[q_0..q_{g-1}, k, v]``;
  * HF -> Meta rotary permutation of every Q/K head's rows (``permute_qkv``);
  * ``dense_h_to_4h = concat([up (w3), gate (w1)])`` for SwiGLU;
  * legacy key names ``transformer`` / ``.attention.`` (accepted by ``TransformerLanguageModel.load_state_dict``);
  * output: ``<out>/release/mp_rank_00/model_optim_rng.pt`` + ``latest_checkpointed_iteration.txt`` = ``release``,
    with the architecture stored in ``args`` so training can use ``--use_checkpoint_args``.
Implementation: one table of per-architecture name maps + three generic transforms, instead of one hand-written
function per model family."""
from __future__ import annotations

import os
import re
import sys
import warnings
from argparse import ArgumentParser, Namespace
from pathlib import Path

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))

from weights_conversion.utils.merge_llama import (llama_s2dense, llama_s2heads, llama_s2hidden, llama_s2layer,  # noqa
                                                  merge_llama)
from weights_conversion.utils.permute_qkv import permute_qkv  # noqa: E402

falcon_s2layer = {7: 32, 40: 60}
falcon_s2heads = {7: 71, 40: 128}
falcon_s2hidden = {7: 4544, 40: 8192}


def group_qkv(wq, wk, wv, n_heads, n_kv_heads, head_dim):
    """[n*hn, h], [nkv*hn, h], [nkv*hn, h] -> fused [(n + 2 nkv) hn, h] grouped per KV head."""
    g = n_heads // n_kv_heads
    q = wq.reshape(n_kv_heads, g, head_dim, -1)
    k = wk.reshape(n_kv_heads, 1, head_dim, -1)
    v = wv.reshape(n_kv_heads, 1, head_dim, -1)
    return torch.cat([q, k, v], dim=1).reshape((n_heads + 2 * n_kv_heads) * head_dim, -1)


def llama_like_to_megatron(weights: dict, n_layer, hidden, n_heads, n_kv_heads, source: str) -> dict:
    """``weights`` uses Meta's names (see merge_llama); ``source`` = 'hf' applies the rotary permutation."""
    hn = hidden // n_heads
    embedding = {"word_embeddings.weight": weights["tok_embeddings.weight"]}
    transformer = {"final_layernorm.weight": weights["norm.weight"]}
    for layer in range(n_layer):
        p = f"layers.{layer}"
        transformer[f"{p}.attention.dense.weight"] = weights[f"{p}.attention.wo.weight"]
        transformer[f"{p}.post_attention_layernorm.weight"] = weights[f"{p}.ffn_norm.weight"]
        transformer[f"{p}.input_layernorm.weight"] = weights[f"{p}.attention_norm.weight"]
        transformer[f"{p}.mlp.dense_4h_to_h.weight"] = weights[f"{p}.feed_forward.w2.weight"]
        transformer[f"{p}.mlp.dense_h_to_4h.weight"] = torch.cat([weights.pop(f"{p}.feed_forward.w3.weight"),
                                                                 weights.pop(f"{p}.feed_forward.w1.weight")])
        qkv = group_qkv(weights.pop(f"{p}.attention.wq.weight"), weights.pop(f"{p}.attention.wk.weight"),
                        weights.pop(f"{p}.attention.wv.weight"), n_heads, n_kv_heads, hn)
        if source == "hf":
            qkv = permute_qkv(qkv, hidden, n_heads, n_kv_heads)
        transformer[f"{p}.attention.query_key_value.weight"] = qkv
    return {"embedding": embedding, "transformer": transformer, "lm_head": weights["output.weight"]}


def llama_to_megatron(weights: dict, size: int, source: str = "meta", version: int = 1) -> dict:
    """Size-table front end of ``llama_like_to_megatron`` (the reference's entry point, hf_to_megatron.py:117)."""
    n_heads = llama_s2heads[size]
    n_kv = n_heads if (version == 1 or size <= 13) else 8
    return llama_like_to_megatron(weights, llama_s2layer[size], llama_s2hidden[size], n_heads, n_kv, source)


def mistral_to_megatron(weights: dict, size: int = 7) -> dict:
    """Mistral-7B (Meta-style names, e.g. from ``hf_llama_state_to_meta_names``): 32 layers, 32 heads over 8 KV heads,
    Hugging Face rotary layout (reference hf_to_megatron.py:185)."""
    assert size == 7
    return llama_like_to_megatron(weights, 32, 4096, 32, 8, "hf")


def falcon_to_megatron(weights: dict, size: int) -> dict:
    """HF Falcon already fuses QKV per KV group; only the rotary permutation and the key names change.
    Embeddings are tied: the LM head must equal the word embeddings."""
    n_layer, hidden, n_heads = falcon_s2layer[size], falcon_s2hidden[size], falcon_s2heads[size]
    n_kv = 1 if size == 7 else 8
    assert torch.allclose(weights["lm_head.weight"], weights["transformer.word_embeddings.weight"])
    embedding = {"word_embeddings.weight": weights["transformer.word_embeddings.weight"]}
    transformer = {"final_layernorm.weight": weights["transformer.ln_f.weight"],
                   "final_layernorm.bias": weights["transformer.ln_f.bias"]}
    for layer in range(n_layer):
        a, b = f"layers.{layer}", f"transformer.h.{layer}"
        transformer[f"{a}.mlp.dense_h_to_4h.weight"] = weights[f"{b}.mlp.dense_h_to_4h.weight"]
        transformer[f"{a}.mlp.dense_4h_to_h.weight"] = weights[f"{b}.mlp.dense_4h_to_h.weight"]
        transformer[f"{a}.attention.query_key_value.weight"] = permute_qkv(
            weights[f"{b}.self_attention.query_key_value.weight"], hidden, n_heads, n_kv)
        transformer[f"{a}.attention.dense.weight"] = weights[f"{b}.self_attention.dense.weight"]
        if size == 7:
            transformer[f"{a}.input_layernorm.weight"] = weights[f"{b}.input_layernorm.weight"]
            transformer[f"{a}.input_layernorm.bias"] = weights[f"{b}.input_layernorm.bias"]
        else:
            transformer[f"{a}.input_layernorm.weight"] = weights[f"{b}.ln_attn.weight"]
            transformer[f"{a}.input_layernorm.bias"] = weights[f"{b}.ln_attn.bias"]
            transformer[f"{a}.mlp_layernorm.weight"] = weights[f"{b}.ln_mlp.weight"]
            transformer[f"{a}.mlp_layernorm.bias"] = weights[f"{b}.ln_mlp.bias"]
    return {"embedding": embedding, "transformer": transformer}


def hf_llama_state_to_meta_names(sd: dict) -> dict:
    out = {"tok_embeddings.weight": sd["model.embed_tokens.weight"], "norm.weight": sd["model.norm.weight"],
           "output.weight": sd["lm_head.weight"]}
    ren = {"self_attn.q_proj": "attention.wq", "self_attn.k_proj": "attention.wk", "self_attn.v_proj": "attention.wv",
           "self_attn.o_proj": "attention.wo", "mlp.gate_proj": "feed_forward.w1", "mlp.down_proj": "feed_forward.w2",
           "mlp.up_proj": "feed_forward.w3", "input_layernorm": "attention_norm",
           "post_attention_layernorm": "ffn_norm"}
    for key, w in sd.items():
        m = re.match(r"^model\.layers\.([0-9]+)\.(.+)\.weight$", key)
        if m and m.group(2) in ren:
            out[f"layers.{m.group(1)}.{ren[m.group(2)]}.weight"] = w
    return out


def architecture_args(model_name: str, size: int, n_layer=None, hidden=None, n_heads=None, n_kv=None, ffn=None,
                      vocab=None, config=None) -> dict:
    """The ``args`` stored in the checkpoint (consumed by ``--use_checkpoint_args``)."""
    if model_name == "falcon":
        a = {"num_layers": falcon_s2layer[size], "hidden_size": falcon_s2hidden[size],
             "num_attention_heads": falcon_s2heads[size], "num_attention_heads_kv": 1 if size == 7 else 8,
             "ffn_hidden_size": 4 * falcon_s2hidden[size], "parallel_attn": True, "parallel_layernorm": size != 7,
             "max_position_embeddings": 2048, "seq_length": 2048, "glu_activation": None, "use_rms_norm": False,
             "tie_embed_logits": True, "make_vocab_size_divisible_by": 128, "tokenizer_type": "FalconTokenizer",
             "layernorm_epsilon": 1e-5}
    else:
        a = {"num_layers": n_layer, "hidden_size": hidden, "num_attention_heads": n_heads,
             "num_attention_heads_kv": n_kv, "ffn_hidden_size": ffn, "parallel_attn": False,
             "parallel_layernorm": False, "make_vocab_size_divisible_by": 1, "glu_activation": "swiglu",
             "use_rms_norm": True, "tie_embed_logits": False, "tokenizer_type": "SentencePieceTokenizer",
             "max_position_embeddings": 2048 if model_name == "llama" else 4096,
             "seq_length": 2048 if model_name == "llama" else 4096,
             "layernorm_epsilon": 1e-6 if model_name == "llama" else 1e-5}
        if model_name == "codellama":
            a.update({"max_position_embeddings": 16384, "seq_length": 16384, "rope_theta": 1e6})
        if model_name == "mistral":
            a.update({"max_position_embeddings": 32768, "seq_length": 32768, "sliding_window_size": 4096})
    a.update({"padded_vocab_size": vocab, "use_bias": False, "use_post_ln": False,
              "tensor_model_parallel_size": 1, "pipeline_model_parallel_size": 1, "iteration": "release",
              "bias_gelu_fusion": False, "bias_droput_fusion": False, "position_embedding_type": "rotary"})
    return a


def save_megatron(out: Path, megatron_weights: dict, args: dict, dtype):
    from megatron_llm_b200.models.enums import PositionEmbeddingType
    args = dict(args)
    args["position_embedding_type"] = PositionEmbeddingType.rotary
    def cast(x):
        return {k: cast(v) for k, v in x.items()} if isinstance(x, dict) else x.to(dtype)
    final = {"iteration": "release", "model": {"language_model": cast(megatron_weights)}, "checkpoint_version": 3.0,
             "args": Namespace(**args)}
    (out / "release" / "mp_rank_00").mkdir(parents=True, exist_ok=True)
    (out / "latest_checkpointed_iteration.txt").write_text("release")
    torch.save(final, out / "release" / "mp_rank_00" / "model_optim_rng.pt")
    print("Saved weights in", out)


def main(model_name: str = "falcon", size: int = 7, out: Path = None, cache_dir: Path = None, model_path: str = None,
         dtype=torch.bfloat16):
    out = Path(out or f"{model_name}-{size}b-megatron").absolute()
    if model_name == "falcon":
        from transformers import AutoModelForCausalLM
        path = model_path or f"tiiuae/falcon-{size}b"
        sd = AutoModelForCausalLM.from_pretrained(path, trust_remote_code=True, cache_dir=cache_dir).state_dict()
        mw = falcon_to_megatron(sd, size)
        vocab = mw["embedding"]["word_embeddings.weight"].size(0)
        save_megatron(out, mw, architecture_args("falcon", size, vocab=vocab), dtype)
        return
    version = 1 if model_name == "llama" else 2
    if model_name == "mistral":
        from transformers import AutoModelForCausalLM
        model = AutoModelForCausalLM.from_pretrained(model_path or "mistralai/Mistral-7B-v0.1", cache_dir=cache_dir)
        weights, source, cfg = hf_llama_state_to_meta_names(model.state_dict()), "hf", model.config
        n_layer, hidden, n_heads, n_kv, ffn = (cfg.num_hidden_layers, cfg.hidden_size, cfg.num_attention_heads,
                                               cfg.num_key_value_heads, cfg.intermediate_size)
    else:
        weights, source = merge_llama(size, version, root_dir=cache_dir, model_path=model_path)
        n_layer, hidden, n_heads, ffn = llama_s2layer[size], llama_s2hidden[size], llama_s2heads[size], \
            llama_s2dense[size]
        n_kv = n_heads if (version == 1 or size <= 13) else 8
    mw = llama_like_to_megatron(weights, n_layer, hidden, n_heads, n_kv, source)
    vocab = mw["embedding"]["word_embeddings.weight"].size(0)
    save_megatron(out, mw, architecture_args(model_name, size, n_layer, hidden, n_heads, n_kv, ffn, vocab), dtype)


if __name__ == "__main__":
    parser = ArgumentParser(description="Convert Huggingface llama/mistral/falcon weights to the megatron layout")
    parser.add_argument("model", choices={"falcon", "llama", "llama2", "codellama", "mistral"})
    parser.add_argument("--size", default=7, choices={7, 13, 30, 34, 40, 65, 70}, type=int)
    parser.add_argument("--out", type=Path)
    parser.add_argument("--cache-dir", type=Path)
    parser.add_argument("--model-path", type=str, help="local HF checkpoint directory (no network available)")
    a = parser.parse_args()
    if a.model == "falcon":
        assert a.size in {7, 40}
    elif a.model == "llama":
        assert a.size in {7, 13, 30, 65}
    elif a.model == "codellama":
        assert a.size in {7, 13, 34}
    elif a.model == "mistral":
        assert a.size in {7}
    else:
        assert a.size in {7, 13, 70}
    main(a.model, a.size, a.out, a.cache_dir, a.model_path)

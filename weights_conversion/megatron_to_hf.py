"""Convert a Megatron checkpoint (TP=PP=1, e.g. produced by tools/checkpoint_util.py) back to Hugging Face format.

Parity target: weights_conversion/megatron_to_hf.py.  Inverse of ``hf_to_megatron``: un-group the fused QKV, revert the
rotary permutation, split the GLU weight into up/gate, re-tie Falcon's embeddings; writes a ``transformers`` model
directory (config + safetensors) and, if given, the tokenizer."""
from __future__ import annotations

import os
import sys
from argparse import ArgumentParser
from pathlib import Path

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))
from weights_conversion.utils.permute_qkv import permute_qkv  # noqa: E402


def ungroup_qkv(qkv, n_heads, n_kv_heads, head_dim):
    """fused [(n + 2 nkv) hn, h] -> (wq [n hn, h], wk [nkv hn, h], wv [nkv hn, h])."""
    g = n_heads // n_kv_heads
    x = qkv.reshape(n_kv_heads, g + 2, head_dim, -1)
    wq = x[:, :g].reshape(n_heads * head_dim, -1)
    wk = x[:, g].reshape(n_kv_heads * head_dim, -1)
    wv = x[:, g + 1].reshape(n_kv_heads * head_dim, -1)
    return wq, wk, wv


def load_megatron(input_dir: Path):
    it = (input_dir / "latest_checkpointed_iteration.txt").read_text().strip()
    sub = it if it == "release" else f"iter_{int(it):07d}"
    base = input_dir / sub
    ranks = sorted(p.name for p in base.iterdir())
    assert len(ranks) == 1, f"merge the checkpoint to tp=pp=1 first (found {ranks}); see tools/checkpoint_util.py"
    ckpt = torch.load(base / ranks[0] / "model_optim_rng.pt", map_location="cpu", weights_only=False)
    lm = ckpt["model"]["language_model"]
    enc = lm["encoder"] if "encoder" in lm else lm["transformer"]
    enc = {k.replace(".attention.", ".self_attention."): v for k, v in enc.items()}
    emb = lm["embedding"]["word_embeddings"]["weight"] if "word_embeddings" in lm["embedding"] else \
        lm["embedding"]["word_embeddings.weight"]
    return ckpt["args"], emb, enc, lm.get("lm_head")


def llama_like_to_hf(args, emb, enc, lm_head, vocab_size=None):
    n, nkv, h = args.num_attention_heads, args.num_attention_heads_kv, args.hidden_size
    hn = h // n
    vocab_size = vocab_size or emb.size(0)
    sd = {"model.embed_tokens.weight": emb[:vocab_size], "model.norm.weight": enc["final_layernorm.weight"],
          "lm_head.weight": lm_head[:vocab_size]}
    for layer in range(args.num_layers):
        a, b = f"layers.{layer}", f"model.layers.{layer}"
        qkv = permute_qkv(enc[f"{a}.self_attention.query_key_value.weight"], h, n, nkv, revert=True)
        wq, wk, wv = ungroup_qkv(qkv, n, nkv, hn)
        sd[f"{b}.self_attn.q_proj.weight"], sd[f"{b}.self_attn.k_proj.weight"], sd[f"{b}.self_attn.v_proj.weight"] = \
            wq, wk, wv
        sd[f"{b}.self_attn.o_proj.weight"] = enc[f"{a}.self_attention.dense.weight"]
        up, gate = torch.chunk(enc[f"{a}.mlp.dense_h_to_4h.weight"], 2, dim=0)
        sd[f"{b}.mlp.up_proj.weight"], sd[f"{b}.mlp.gate_proj.weight"] = up, gate
        sd[f"{b}.mlp.down_proj.weight"] = enc[f"{a}.mlp.dense_4h_to_h.weight"]
        sd[f"{b}.input_layernorm.weight"] = enc[f"{a}.input_layernorm.weight"]
        sd[f"{b}.post_attention_layernorm.weight"] = enc[f"{a}.post_attention_layernorm.weight"]
    return sd


def falcon_to_hf(args, emb, enc, vocab_size=None):
    n, nkv, h = args.num_attention_heads, args.num_attention_heads_kv, args.hidden_size
    vocab_size = vocab_size or emb.size(0)
    sd = {"transformer.word_embeddings.weight": emb[:vocab_size], "lm_head.weight": emb[:vocab_size],
          "transformer.ln_f.weight": enc["final_layernorm.weight"], "transformer.ln_f.bias": enc["final_layernorm.bias"]}
    for layer in range(args.num_layers):
        a, b = f"layers.{layer}", f"transformer.h.{layer}"
        sd[f"{b}.mlp.dense_h_to_4h.weight"] = enc[f"{a}.mlp.dense_h_to_4h.weight"]
        sd[f"{b}.mlp.dense_4h_to_h.weight"] = enc[f"{a}.mlp.dense_4h_to_h.weight"]
        sd[f"{b}.self_attention.query_key_value.weight"] = permute_qkv(
            enc[f"{a}.self_attention.query_key_value.weight"], h, n, nkv, revert=True)
        sd[f"{b}.self_attention.dense.weight"] = enc[f"{a}.self_attention.dense.weight"]
        if getattr(args, "parallel_layernorm", False):
            sd[f"{b}.ln_attn.weight"], sd[f"{b}.ln_attn.bias"] = enc[f"{a}.input_layernorm.weight"], \
                enc[f"{a}.input_layernorm.bias"]
            sd[f"{b}.ln_mlp.weight"], sd[f"{b}.ln_mlp.bias"] = enc[f"{a}.mlp_layernorm.weight"], \
                enc[f"{a}.mlp_layernorm.bias"]
        else:
            sd[f"{b}.input_layernorm.weight"], sd[f"{b}.input_layernorm.bias"] = enc[f"{a}.input_layernorm.weight"], \
                enc[f"{a}.input_layernorm.bias"]
    return sd


def build_hf_config(model: str, args, vocab_size: int):
    from transformers import FalconConfig, LlamaConfig, MistralConfig
    if model == "falcon":
        return FalconConfig(vocab_size=vocab_size, hidden_size=args.hidden_size, num_hidden_layers=args.num_layers,
                            num_attention_heads=args.num_attention_heads, num_kv_heads=args.num_attention_heads_kv,
                            new_decoder_architecture=bool(getattr(args, "parallel_layernorm", False)),
                            parallel_attn=True, bias=False, layer_norm_epsilon=args.layernorm_epsilon)
    common = dict(vocab_size=vocab_size, hidden_size=args.hidden_size, intermediate_size=args.ffn_hidden_size,
                  num_hidden_layers=args.num_layers, num_attention_heads=args.num_attention_heads,
                  num_key_value_heads=args.num_attention_heads_kv, rms_norm_eps=args.layernorm_epsilon,
                  max_position_embeddings=args.max_position_embeddings, tie_word_embeddings=False,
                  rope_theta=getattr(args, "rope_theta", 10000.0))
    if model == "mistral":
        return MistralConfig(sliding_window=getattr(args, "sliding_window_size", 4096), **common)
    cfg = LlamaConfig(**common)
    scaling = getattr(args, "rope_scaling_factor", 1.0)
    if scaling and scaling != 1.0:
        cfg.rope_scaling = {"type": "linear", "factor": scaling}
    return cfg


def main(model: str, input_dir: Path, output_dir: Path, vocab_file=None, no_new_tokens=True, dtype=torch.bfloat16,
         override_special_tokens=None):
    from transformers import AutoModelForCausalLM
    args, emb, enc, lm_head = load_megatron(Path(input_dir))
    vocab_size = emb.size(0)
    if vocab_file is not None:
        from megatron_llm_b200.tokenizer.tokenizer import _SentencePieceTokenizer
        tok = _SentencePieceTokenizer(str(vocab_file), new_tokens=not no_new_tokens)
        vocab_size = tok.vocab_size
    sd = falcon_to_hf(args, emb, enc, vocab_size) if model == "falcon" else \
        llama_like_to_hf(args, emb, enc, lm_head, vocab_size)
    cfg = build_hf_config(model, args, vocab_size)
    with torch.device("meta"):
        hf = AutoModelForCausalLM.from_config(cfg)
    hf = hf.to_empty(device="cpu").to(dtype)
    missing, unexpected = hf.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=False)
    missing = [m for m in missing if "rotary_emb" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    hf.save_pretrained(output_dir)
    print("Saved Hugging Face model in", output_dir)


if __name__ == "__main__":
    parser = ArgumentParser(description="Convert megatron weights back to the Hugging Face format")
    parser.add_argument("--model", type=str, default="llama2", choices={"falcon", "llama", "llama2", "codellama", "mistral"})
    parser.add_argument("--input_dir", type=Path, required=True)
    parser.add_argument("--output_dir", type=Path, required=True)
    parser.add_argument("--vocab_file", type=Path, default=None)
    parser.add_argument("--no_new_tokens", action="store_false", dest="new_tokens")
    a = parser.parse_args()
    main(a.model, a.input_dir, a.output_dir, a.vocab_file, not a.new_tokens)

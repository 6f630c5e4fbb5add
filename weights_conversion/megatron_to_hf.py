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


def _layer_weight(ckpt, layer_idx, name):
    """``ckpt``: a language-model state dict (``transformer`` or ``encoder`` section, ``attention`` or
    ``self_attention`` naming) as stored in the checkpoint."""
    section = ckpt["transformer"] if "transformer" in ckpt else ckpt["encoder"]
    for att in ("attention", "self_attention"):
        key = f"layers.{layer_idx}." + name.replace("{att}", att)
        if key in section:
            return section[key]
    raise KeyError(name)


def convert_wqkv(llama_mega, layer_idx=0, n_heads=32, n_heads_kv=8):
    """(wq, wk, wv) in the Hugging Face layout for one layer (reference megatron_to_hf.py:47-71)."""
    qkv = _layer_weight(llama_mega, layer_idx, "{att}.query_key_value.weight")
    hidden = qkv.size(1)
    return ungroup_qkv(permute_qkv(qkv, hidden, n_heads, n_heads_kv, revert=True), n_heads, n_heads_kv,
                       hidden // n_heads)


def convert_ffn(llama_mega, layer_idx=0, n_dense=11008):
    """(w1 = gate, w3 = up) of one layer's fused ``dense_h_to_4h`` = [up; gate] (reference :74-77)."""
    w3, w1 = _layer_weight(llama_mega, layer_idx, "mlp.dense_h_to_4h.weight").split(n_dense, dim=0)
    return w1, w3


def write_json(obj, path):
    import json
    with open(path, "w") as f:
        json.dump(obj, f)


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


_HUB_TOKENIZERS = {"codellama": "TheBloke/CodeLlama-13B-fp16", "mistral": "mistralai/Mistral-7B-v0.1",
                   "falcon": "tiiuae/falcon-40b"}          # (reference megatron_to_hf.py:489-507)


def _megatron_tokenizer(tokenizer_type, vocab_file, new_tokens, vocab_extra_ids_list):
    from argparse import Namespace
    from megatron_llm_b200.tokenizer import build_tokenizer
    return build_tokenizer(Namespace(tokenizer_type=tokenizer_type, vocab_file=vocab_file, merge_file=None, rank=0,
                                     vocab_extra_ids=0, vocab_extra_ids_list=vocab_extra_ids_list,
                                     new_tokens=new_tokens, make_vocab_size_divisible_by=128,
                                     tensor_model_parallel_size=1,
                                     tokenizer_model=vocab_file if tokenizer_type == "FalconTokenizer" else None))


def write_tokenizer(model: str, output_dir, vocab_file=None, new_tokens=True, vocab_extra_ids_list=None,
                    override_special_tokens=(), cache_dir=None):
    """Save a Hugging Face tokenizer whose ids agree with the Megatron tokenizer the model was trained with
    (parity: megatron_to_hf.py::write_tokenizer).  Llama-family: ``tokenizer.model`` (``--vocab_file``, or the hub
    copy when reachable) plus the special tokens Megatron appends (<CLS> <SEP> <EOD> <MASK> <PAD>, the
    ``--vocab_extra_ids_list`` entries), every id cross-checked.  Falcon: the wrapped HF tokenizer itself.
    ``override_special_tokens``: ``key=token`` with key in bos|cls|eos|mask|pad|sep|unk."""
    import warnings
    if model == "falcon":
        mt = _megatron_tokenizer("FalconTokenizer", str(vocab_file) if vocab_file else None, new_tokens,
                                 vocab_extra_ids_list)
        hf_tok = mt.tokenizer
    else:
        from transformers import LlamaTokenizerFast
        try:
            if vocab_file:
                src = Path(vocab_file)
                hf_tok = LlamaTokenizerFast.from_pretrained(src.parent if src.suffix == ".model" else src)
            else:
                hf_tok = LlamaTokenizerFast.from_pretrained(_HUB_TOKENIZERS.get(model, "meta-llama/Llama-2-7b-hf"),
                                                            cache_dir=cache_dir)
            vocab_file = hf_tok.vocab_file
        except OSError as e:
            print("ERROR: could not load the tokenizer ({}); no tokenizer written.".format(e))
            return None
        mt = _megatron_tokenizer("SentencePieceTokenizer", str(vocab_file), new_tokens, vocab_extra_ids_list)
        # Megatron appends its special tokens after the SentencePiece pieces in a fixed order; appending the same
        # tokens in the same order gives the same ids
        appended = sorted((i, t) for t, i in mt.vocab.items() if i >= len(hf_tok))
        for _, tok in appended:
            hf_tok.add_tokens(tok, special_tokens=True)
        for attr, name in (("cls", "cls_token"), ("sep", "sep_token"), ("mask", "mask_token"), ("pad", "pad_token")):
            tid = getattr(mt, attr, None)
            if tid is not None and new_tokens:
                setattr(hf_tok, name, mt.inv_vocab[tid])
        extra = [t for t in (vocab_extra_ids_list.split(",") if vocab_extra_ids_list else [])]
        if extra:
            hf_tok.add_special_tokens({"additional_special_tokens": extra})
        hf_vocab = hf_tok.get_vocab()
        named = [v for k, v in hf_tok.special_tokens_map.items() if k != "additional_special_tokens"]
        for tok in named + extra + [t for _, t in appended]:
            a, b = mt.vocab.get(tok), hf_vocab.get(tok)
            assert a == b, f"Megatron and Hugging Face tokenizers disagree on {tok!r}: {a} vs {b}"
    for item in override_special_tokens or ():
        key, sep, value = item.partition("=")
        if not sep:
            warnings.warn(f"Illegal override string {item}")
        elif key not in {"bos", "cls", "eos", "mask", "pad", "sep", "unk"}:
            warnings.warn(f"Cannot override key {key}")
        elif value not in mt.vocab:
            warnings.warn(f"Token {value} not found in megatron tokenizer")
        else:
            setattr(hf_tok, f"{key}_token", value)
            assert getattr(hf_tok, f"{key}_token_id") == mt.vocab[value]
    print("Final HF tokenizer configuration:")
    print(hf_tok)
    hf_tok.save_pretrained(output_dir)
    return hf_tok


def main(model: str, input_dir: Path, output_dir: Path, vocab_file=None, no_new_tokens=True, dtype=torch.bfloat16,
         override_special_tokens=None, num_output_shards=1, vocab_extra_ids_list=None, cache_dir=None,
         tokenizer=True):
    from transformers import AutoModelForCausalLM
    args, emb, enc, lm_head = load_megatron(Path(input_dir))
    vocab_size = emb.size(0)
    if vocab_file is not None and model != "falcon":
        from megatron_llm_b200.tokenizer.tokenizer import _SentencePieceTokenizer
        tok = _SentencePieceTokenizer(str(vocab_file), vocab_extra_ids_list=vocab_extra_ids_list,
                                      new_tokens=not no_new_tokens)
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
    kw = {}
    if num_output_shards and num_output_shards > 1:
        total = sum(p.numel() * p.element_size() for p in hf.parameters())
        kw["max_shard_size"] = -(-total // num_output_shards)          # bytes per safetensors shard
    hf.save_pretrained(output_dir, **kw)
    print("Saved Hugging Face model in", output_dir)
    if tokenizer and (vocab_file is not None or cache_dir is not None):
        write_tokenizer(model, output_dir, vocab_file, not no_new_tokens, vocab_extra_ids_list,
                        override_special_tokens or (), cache_dir)


def write_llama_model(model_path, input_base_path, num_output_shards=2, norm_eps=None, rope_theta=None):
    """The reference's per-family writers (megatron_to_hf.py:80, :196, :333): weights + config, no tokenizer.  The
    architecture (incl. the norm epsilon and RoPE base) comes from the checkpoint's stored arguments."""
    main("llama2", Path(input_base_path), Path(model_path), num_output_shards=num_output_shards, tokenizer=False)


def write_mistral_model(model_path, input_base_path, num_output_shards=2, **_):
    main("mistral", Path(input_base_path), Path(model_path), num_output_shards=num_output_shards, tokenizer=False)


def write_falcon_model(model_path, input_base_path, num_output_shards=2, safe_serialization=True):
    main("falcon", Path(input_base_path), Path(model_path), num_output_shards=num_output_shards, tokenizer=False)


if __name__ == "__main__":
    parser = ArgumentParser(description="Convert megatron weights back to the Hugging Face format")
    parser.add_argument("--model", type=str, default="llama2", choices={"falcon", "llama", "llama2", "codellama", "mistral"})
    parser.add_argument("--input_dir", type=Path, required=True, help="Megatron checkpoint directory (tp = pp = 1)")
    parser.add_argument("--output_dir", type=Path, required=True, help="where the HF model and tokenizer are written")
    parser.add_argument("--num_output_shards", type=int, default=1, help="number of safetensors shards")
    parser.add_argument("--cache_dir", help="Hugging Face cache_dir (tokenizer download when --vocab_file is not given)")
    parser.add_argument("--vocab_file", type=Path, default=None,
                        help="tokenizer.model (Llama family) or a local tokenizer directory (Falcon)")
    parser.add_argument("--vocab_extra_ids_list", help="comma separated list of special tokens added to the tokenizer")
    parser.add_argument("--override_special_tokens", nargs="*", default=[],
                        help="key=token pairs, key in bos|cls|eos|mask|pad|sep|unk, e.g. eos=<|im_end|>")
    parser.add_argument("--no_new_tokens", action="store_false", dest="new_tokens")
    a = parser.parse_args()
    main(a.model, a.input_dir, a.output_dir, a.vocab_file, not a.new_tokens,
         override_special_tokens=a.override_special_tokens, num_output_shards=a.num_output_shards,
         vocab_extra_ids_list=a.vocab_extra_ids_list, cache_dir=a.cache_dir)

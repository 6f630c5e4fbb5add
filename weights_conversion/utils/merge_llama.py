"""Merge Meta's sharded Llama checkpoints (``consolidated.NN.pth``) into one state dict
(parity: weights_conversion/utils/merge_llama.py).  Column-parallel weights are concatenated on dim 0, row-parallel
ones on dim 1, replicated ones (norms, rope) are taken from shard 0."""
from __future__ import annotations

import json
import re
from pathlib import Path

import torch

_DIM_BY_SUFFIX = {
    "attention.wq.weight": 0, "attention.wk.weight": 0, "attention.wv.weight": 0, "attention.wo.weight": 1,
    "feed_forward.w1.weight": 0, "feed_forward.w3.weight": 0, "feed_forward.w2.weight": 1,
    "tok_embeddings.weight": 1, "output.weight": 0,
}
llama_s2layer = {7: 32, 13: 40, 30: 60, 34: 48, 65: 80, 70: 80}
llama_s2heads = {7: 32, 13: 40, 30: 52, 34: 64, 65: 64, 70: 64}
llama_s2dense = {7: 11008, 13: 13824, 30: 17920, 34: 22016, 65: 22016, 70: 28672}
llama_s2hidden = {7: 4096, 13: 5120, 30: 6656, 34: 8192, 65: 8192, 70: 8192}


def _concat_dim(key: str):
    for suffix, dim in _DIM_BY_SUFFIX.items():
        if key.endswith(suffix):
            return dim
    return None


def merge_meta_llama(size: int, root_dir: Path) -> dict:
    paths = sorted(p for p in Path(root_dir).iterdir() if re.match(r"^consolidated\.[0-9]+\.pth$", p.name))
    if len(paths) == 1:
        return torch.load(paths[0], map_location="cpu")
    shards = [torch.load(p, map_location="cpu") for p in paths]
    merged = {}
    for key in shards[0]:
        if key.endswith("rope.freqs"):
            continue
        dim = _concat_dim(key)
        merged[key] = shards[0][key] if dim is None else torch.cat([s[key] for s in shards], dim=dim)
    return merged


def merge_hf_llama(size: int, version: int, cache_dir=None, model_path=None, tokenizer_len=None):
    """Load a Hugging Face Llama and rename its weights to Meta's names (QKV stay in the HF rotary layout)."""
    from transformers import AutoModelForCausalLM
    assert model_path is not None, "offline environment: pass --model-path to a local HF checkpoint"
    model = AutoModelForCausalLM.from_pretrained(model_path, cache_dir=cache_dir, torch_dtype=torch.float32)
    sd = model.state_dict()
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
    return out, model.config


def merge_llama(size: int, version: int, root_dir=None, tokenizer_len=None, model_path=None):
    if root_dir is not None and any(Path(root_dir).glob("consolidated.*.pth")):
        return merge_meta_llama(size, Path(root_dir)), "meta"
    weights, _ = merge_hf_llama(size, version, cache_dir=root_dir, model_path=model_path, tokenizer_len=tokenizer_len)
    return weights, "hf"

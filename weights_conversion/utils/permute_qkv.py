"""Rotary-layout permutation of the fused QKV weight (parity: weights_conversion/utils/permute_qkv.py).

Hugging Face stores every Q/K head with the two rotary halves contiguous (``rotate_half`` convention); Megatron / Meta
use interleaved pairs.  ``permute_qkv`` converts HF -> interleaved for the Q heads and the K head of every KV group of
the fused ``[q_0..q_{g-1}, k, v]`` weight (V untouched); ``revert=True`` is the inverse.  ``update_checkpoint``
applies it to a whole Megatron checkpoint directory."""
from __future__ import annotations

import shutil
import sys
from argparse import ArgumentParser
from pathlib import Path

import torch


def _permute_head(x: torch.Tensor, revert: bool) -> torch.Tensor:
    head_dim = x.size(0)
    rest = x.shape[1:]
    if revert:   # interleaved -> halves
        return x.reshape(head_dim // 2, 2, *rest).transpose(0, 1).reshape(head_dim, *rest)
    return x.reshape(2, head_dim // 2, *rest).transpose(0, 1).reshape(head_dim, *rest)


def permute_qkv(qkv_w: torch.Tensor, dim: int, n_heads: int, n_heads_kv: int, revert: bool = False) -> torch.Tensor:
    head_dim = dim // n_heads
    g = n_heads // n_heads_kv
    grouped = qkv_w.reshape(-1, g + 2, head_dim, *qkv_w.shape[1:])      # [kv groups, g+2, hn, ...]
    out = grouped.clone()
    for idx in range(g + 1):                                             # q heads and the k head
        for grp in range(grouped.size(0)):
            out[grp, idx] = _permute_head(grouped[grp, idx], revert)
    return out.reshape(qkv_w.shape)


def update_checkpoint(input_dir: Path, output_dir: Path, overwrite_ok: bool = False):
    if output_dir.exists():
        if not overwrite_ok:
            raise FileExistsError(f"Output directory {output_dir} already exists")
        print(f"Removing {output_dir}")
        shutil.rmtree(output_dir)
    output_dir.mkdir(parents=True)
    it = (input_dir / "latest_checkpointed_iteration.txt").read_text().strip()
    print("Updating weights of iteration", it)
    (output_dir / "latest_checkpointed_iteration.txt").write_text(it)
    sub = it if it == "release" else f"iter_{int(it):07d}"
    (output_dir / sub).mkdir()
    for rank_dir in sorted((input_dir / sub).iterdir()):
        ckpt = torch.load(rank_dir / "model_optim_rng.pt", map_location="cpu", weights_only=False)
        args = ckpt["args"]
        lm = ckpt["model"]["language_model"]
        enc = lm["encoder"] if "encoder" in lm else lm["transformer"]
        # global head counts (as the reference): only their ratio and hidden/heads = head_dim are used, and the
        # number of KV groups in this TP shard follows from the shard's row count
        for key, w in list(enc.items()):
            if "query_key_value.weight" in key:
                enc[key] = permute_qkv(w, args.hidden_size, args.num_attention_heads, args.num_attention_heads_kv)
        (output_dir / sub / rank_dir.name).mkdir()
        torch.save(ckpt, output_dir / sub / rank_dir.name / "model_optim_rng.pt")


if __name__ == "__main__":
    parser = ArgumentParser(description="Fix permutation of the QKV weights of a Megatron checkpoint")
    parser.add_argument("--input-dir", type=Path, required=True)
    parser.add_argument("--output-dir", type=Path, required=True)
    parser.add_argument("--overwrite-ok", action="store_true")
    a = parser.parse_args()
    update_checkpoint(a.input_dir, a.output_dir, a.overwrite_ok)

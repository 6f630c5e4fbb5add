"""Checkpoint saver plugin: consume unsharded per-layer messages and write a (target TP x PP)-sharded checkpoint.

Parity: tools/checkpoint_saver_megatron.py.  Works on plain state dicts (no model construction); writes one pipeline
stage at a time so host memory stays bounded."""
from __future__ import annotations

import copy
import os
import sys

import torch

try:
    from checkpoint_loader_megatron import GLU_KEYS, LAYER_RULES
except ImportError:  # imported as tools.checkpoint_saver_megatron
    from tools.checkpoint_loader_megatron import GLU_KEYS, LAYER_RULES


def add_arguments(parser):
    group = parser.add_argument_group(title="Megatron saver")
    try:
        group.add_argument("--megatron_path", type=str, default=None, help="Base directory of the framework repository")
    except Exception:
        pass  # the loader plugin registered the same flag
    group.add_argument("--target_tensor_parallel_size", type=int,
                       help="Target tensor model parallel size, defaults to the tensor parallel size in the input "
                            "checkpoint if provided by the loader, otherwise to 1")
    group.add_argument("--target_pipeline_parallel_size", type=int,
                       help="Target pipeline model parallel size, default to the pipeline parallel size in the input "
                            "checkpoint if provided by the loader, otherwise to 1")


def split_glu(full, tp, dim=0):
    up, gate = torch.chunk(full, 2, dim=dim)
    return [torch.cat([u, g], dim=dim) for u, g in zip(torch.chunk(up, tp, dim=dim), torch.chunk(gate, tp, dim=dim))]


def padded_vocab(orig, divisible_by, tp):
    mult = divisible_by * tp
    return ((orig + mult - 1) // mult) * mult


def resize_vocab(w, true_size, target):
    """Trim the padding rows to ``true_size`` (if known) then pad/trim to ``target`` rows."""
    if true_size is not None:
        w = w[:true_size]
    if w.size(0) > target:
        return w[:target]
    if w.size(0) < target:
        pad = w[-1:].expand(target - w.size(0), *w.shape[1:])
        return torch.cat([w, pad], dim=0)
    return w


def save_checkpoint(queue, args):
    def get(name=None):
        val = queue.get()
        if val == "exit":
            sys.exit("Loader exited, exiting saver")
        if name is not None and args.checking and val["name"] != name:
            sys.exit(f'Unexpected message. Expecting "{name}" but got "{val["name"]}". Exiting saver.')
        if name is not None:
            print(f"received {name}")
        return val

    def check(msg):
        if not args.checking:
            return
        msg.pop("name", None)
        if len(msg) > 0:
            print(f"Unexpected values in the message: {list(msg.keys())}; pass --no_checking to ignore.")
            sys.exit(1)

    md = get()
    tp = args.target_tensor_parallel_size or getattr(md, "previous_tensor_parallel_size", None) or 1
    pp = args.target_pipeline_parallel_size or getattr(md, "previous_pipeline_parallel_size", None) or 1
    assert md.num_layers % pp == 0, "num_layers must be divisible by the target pipeline parallel size"
    assert md.num_attention_heads_kv % tp == 0 or md.num_attention_heads_kv == 1, \
        "KV heads must be divisible by the target tensor parallel size"
    dtype = md.params_dtype
    glu = md.glu_activation is not None

    ck_args = copy.deepcopy(md.checkpoint_args)
    ck_args.tensor_model_parallel_size, ck_args.pipeline_model_parallel_size = tp, pp
    ck_args.params_dtype = dtype
    ck_args.consumed_train_samples, ck_args.consumed_valid_samples = md.consumed_train_samples, md.consumed_valid_samples
    if hasattr(ck_args, "world_size"):
        ck_args.world_size = tp * pp
    ck_args.num_layers_per_virtual_pipeline_stage = None

    it = md.iteration
    sub = "release" if it == "release" else f"iter_{int(it):07d}"
    os.makedirs(args.save_dir, exist_ok=True)

    def write_stage(pp_rank, states):
        for t, lm in enumerate(states):
            name = f"mp_rank_{t:02d}" if pp == 1 else f"mp_rank_{t:02d}_{pp_rank:03d}"
            d = os.path.join(args.save_dir, sub, name)
            os.makedirs(d, exist_ok=True)
            model = {"language_model": lm["language_model"]}
            for extra in ("lm_head", "binary_head", "word_embeddings_for_head"):
                if extra in lm:
                    model[extra] = lm[extra]
            torch.save({"args": ck_args, "checkpoint_version": 3.0, "iteration": it, "model": model},
                       os.path.join(d, "model_optim_rng.pt"))

    # ---- embeddings
    emb_msg = get("embeddings")
    word = emb_msg.pop("word embeddings")
    pos = emb_msg.pop("position embeddings", None)
    tokentype = emb_msg.pop("tokentype embeddings", None)
    check(emb_msg)
    true_size = md.true_vocab_size
    target_vocab = padded_vocab(true_size, md.make_vocab_size_divisible_by, tp) if true_size is not None \
        else word.size(0)
    if true_size is None and word.size(0) % tp != 0:
        sys.exit("the vocabulary is not divisible by the target TP size: pass --true_vocab_size / --vocab_file")
    ck_args.padded_vocab_size = target_vocab
    word_shards = torch.chunk(resize_vocab(word, true_size, target_vocab).to(dtype), tp, dim=0)
    head_shards = None
    if not md.tie_embed_logits:
        m = get("lm_head")
        head_shards = torch.chunk(resize_vocab(m.pop("lm_head"), true_size, target_vocab).to(dtype), tp, dim=0)
        check(m)

    def new_stage(pp_rank):
        states = []
        for t in range(tp):
            lm = {"encoder": {}}
            if pp_rank == 0:
                emb = {"word_embeddings": {"weight": word_shards[t].clone()}}
                if pos is not None:
                    emb["position_embeddings"] = {"weight": pos.to(dtype)}
                if tokentype is not None:
                    emb["tokentype_embeddings"] = {"weight": tokentype.to(dtype)}
                lm["embedding"] = emb
            states.append({"language_model": lm})
        return states

    per_stage = md.num_layers // pp
    total = 0
    for pp_rank in range(pp):
        states = new_stage(pp_rank)
        for li in range(per_stage):
            msg = get(f"transformer layer {total}")
            for suffix, mkey, dim in LAYER_RULES:
                if mkey not in msg:
                    continue
                full = msg.pop(mkey).to(dtype)
                if dim is None:
                    shards = [full] * tp
                elif glu and mkey in GLU_KEYS:
                    shards = split_glu(full, tp, dim)
                else:
                    shards = torch.chunk(full, tp, dim=dim)
                for t in range(tp):
                    states[t]["language_model"]["encoder"][f"layers.{li}.{suffix}"] = shards[t].clone()
            check(msg)
            total += 1
        if pp_rank == pp - 1:
            msg = get("final layernorm")
            for t in range(tp):
                enc = states[t]["language_model"]["encoder"]
                enc["final_layernorm.weight"] = msg["weight"].to(dtype)
                if "bias" in msg:
                    enc["final_layernorm.bias"] = msg["bias"].to(dtype)
                if head_shards is not None:
                    states[t]["language_model"]["lm_head"] = head_shards[t].clone()
                elif pp > 1:
                    states[t]["word_embeddings_for_head"] = {"weight": word_shards[t].clone()}
            msg.pop("weight"), msg.pop("bias", None)
            check(msg)
            # optional BERT tail
            msg = queue.get()
            while msg != "done":
                if msg == "exit":
                    sys.exit("Loader exited, exiting saver")
                name = msg["name"]
                print(f"received {name}")
                bias_shards = None
                if name == "lm head" and "vocab bias" in msg:
                    b = resize_vocab(msg["vocab bias"].unsqueeze(1), true_size, target_vocab).squeeze(1)
                    bias_shards = torch.chunk(b.to(dtype), tp, dim=0)
                for t in range(tp):
                    if name == "pooler":
                        states[t]["language_model"]["pooler"] = {"dense.weight": msg["weight"].to(dtype),
                                                                 "dense.bias": msg["bias"].to(dtype)}
                    elif name == "lm head":
                        states[t]["lm_head"] = {"dense.weight": msg["dense weight"].to(dtype),
                                                "dense.bias": msg["dense bias"].to(dtype),
                                                "layernorm.weight": msg["layernorm weight"].to(dtype),
                                                "layernorm.bias": msg["layernorm bias"].to(dtype)}
                        if bias_shards is not None:
                            states[t]["lm_head"]["bias"] = bias_shards[t].clone()
                    elif name == "binary head":
                        states[t]["binary_head"] = {"weight": msg["weight"].to(dtype), "bias": msg["bias"].to(dtype)}
                    elif args.checking:
                        sys.exit(f"unexpected message {name}")
                msg = queue.get()
        write_stage(pp_rank, states)
        del states
    with open(os.path.join(args.save_dir, "latest_checkpointed_iteration.txt"), "w") as f:
        f.write(str(it))
    print("Done!")

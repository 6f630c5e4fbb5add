"""Checkpoint loader plugin: stream a (TP x PP)-sharded Megatron checkpoint as unsharded per-layer messages.

Parity: tools/checkpoint_loader_megatron.py (message protocol documented in tools/checkpoint_util.py).  The reference
instantiates full models on every rank slot just to read their tensors; this loader never builds a model -- it works on
the saved state dicts with a table of shard rules, so resharding a 70B checkpoint needs host memory for one pipeline
stage only and no GPU."""
from __future__ import annotations

import os
import sys
import types

import torch

# ------------------------------------------------------------------------------------------------ shard rules
# suffix -> (message key, dim the tensor is split on across TP ranks; None = replicated)
LAYER_RULES = [
    ("input_layernorm.weight", "input layernorm weight", None),
    ("input_layernorm.bias", "input layernorm bias", None),
    ("mlp_layernorm.weight", "mlp layernorm weight", None),
    ("mlp_layernorm.bias", "mlp layernorm bias", None),
    ("post_attention_layernorm.weight", "post layernorm weight", None),
    ("post_attention_layernorm.bias", "post layernorm bias", None),
    ("output_layernorm.weight", "output layernorm weight", None),
    ("output_layernorm.bias", "output layernorm bias", None),
    ("self_attention.query_key_value.weight", "qkv weight", 0),
    ("self_attention.query_key_value.bias", "qkv bias", 0),
    ("self_attention.dense.weight", "dense weight", 1),
    ("self_attention.dense.bias", "dense bias", None),
    ("mlp.dense_h_to_4h.weight", "mlp l0 weight", 0),
    ("mlp.dense_h_to_4h.bias", "mlp l0 bias", 0),
    ("mlp.dense_4h_to_h.weight", "mlp l1 weight", 1),
    ("mlp.dense_4h_to_h.bias", "mlp l1 bias", None),
]
GLU_KEYS = ("mlp l0 weight", "mlp l0 bias")


def add_arguments(parser):
    group = parser.add_argument_group(title="Megatron loader")
    group.add_argument("--true_vocab_size", type=int, default=None,
                       help="original size of vocab, if specified will trim padding from embedding table.")
    group.add_argument("--vocab_file", type=str, default=None,
                       help="Path to the vocab file. If specified will use this to get vocab size and trim padding "
                            "from the embedding table.")
    group.add_argument("--megatron_path", type=str, default=None, help="Base directory of the framework repository")


def _iteration_dir(load_dir, load_iters=None):
    tracker = os.path.join(load_dir, "latest_checkpointed_iteration.txt")
    if load_iters is not None:
        it = str(load_iters)
    else:
        with open(tracker) as f:
            it = f.read().strip()
    if it == "release":
        return os.path.join(load_dir, "release"), "release"
    return os.path.join(load_dir, f"iter_{int(it):07d}"), int(it)


def _rank_file(base, tp_rank, pp_rank, pp_size):
    name = f"mp_rank_{tp_rank:02d}" if pp_size == 1 else f"mp_rank_{tp_rank:02d}_{pp_rank:03d}"
    for fn in ("model_optim_rng.pt", "model_rng.pt"):
        p = os.path.join(base, name, fn)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(os.path.join(base, name))


def _normalise(lm: dict):
    """language_model dict -> (embedding dict with flat keys, encoder dict with new-style names, lm_head, pooler)."""
    emb = {}
    for k, v in lm.get("embedding", {}).items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                emb[f"{k}.{kk}"] = vv
        else:
            emb[k] = v
    enc = lm.get("encoder", lm.get("transformer", {}))
    enc = {k.replace(".attention.", ".self_attention."): v for k, v in enc.items()}
    return emb, enc, lm.get("lm_head"), lm.get("pooler")


def merge_glu(shards, dim=0):
    """Every TP shard of a GLU projection is [up_shard; gate_shard]: un-interleave before concatenating."""
    ups, gates = zip(*(torch.chunk(s, 2, dim=dim) for s in shards))
    return torch.cat(list(ups) + list(gates), dim=dim)


def true_vocab_size(args):
    if args.true_vocab_size is not None and args.vocab_file is not None:
        sys.exit("only one of --true_vocab_size / --vocab_file may be given")
    if args.true_vocab_size is not None:
        return args.true_vocab_size
    if args.vocab_file is not None:
        import json
        try:
            with open(args.vocab_file) as f:
                return len(json.load(f))
        except Exception:
            import sentencepiece as spm
            return spm.SentencePieceProcessor(model_file=args.vocab_file).vocab_size()
    return None


def load_checkpoint(queue, args):
    """``queue`` only needs ``put``; messages follow the protocol in checkpoint_util.py."""
    try:
        _load_checkpoint(queue, args)
    except BaseException:
        queue.put("exit")
        raise


def _load_checkpoint(queue, args):
    base, iteration = _iteration_dir(args.load_dir, getattr(args, "load_iters", None))
    # rank (0, 0) tells the parallel layout
    names = sorted(os.listdir(base))
    pp_size = 1 + max((int(n.split("_")[3]) for n in names if n.count("_") >= 3 and n.startswith("mp_rank_")), default=0)
    tp_size = 1 + max(int(n.split("_")[2]) for n in names if n.startswith("mp_rank_"))
    first = torch.load(_rank_file(base, 0, 0, pp_size), map_location="cpu", weights_only=False)
    margs = first["args"]
    if getattr(margs, "num_layers_per_virtual_pipeline_stage", None) is not None or "model0" in first:
        sys.exit("checkpoints written with the interleaved schedule are not supported by the loader")
    assert tp_size == margs.tensor_model_parallel_size and pp_size == margs.pipeline_model_parallel_size, \
        (tp_size, pp_size, margs.tensor_model_parallel_size, margs.pipeline_model_parallel_size)

    md = types.SimpleNamespace()
    md.model_type = args.model_type
    for k, default in (("num_layers", None), ("hidden_size", None), ("seq_length", None),
                       ("num_attention_heads", None), ("max_position_embeddings", None), ("tokenizer_type", None),
                       ("make_vocab_size_divisible_by", 128), ("num_attention_heads_kv", None),
                       ("parallel_attn", False), ("parallel_layernorm", False), ("use_flash_attn", False),
                       ("hidden_dropout", 0.1), ("lima_dropout", False), ("use_bias", True), ("use_rms_norm", False),
                       ("ffn_hidden_size", None), ("glu_activation", None), ("tie_embed_logits", True),
                       ("params_dtype", torch.float32), ("sliding_window_size", None), ("layernorm_epsilon", 1e-5),
                       ("rope_theta", 10000.0), ("rope_scaling_factor", 1.0), ("use_post_ln", False),
                       ("bert_binary_head", True)):
        setattr(md, k, getattr(margs, k, default))
    if md.num_attention_heads_kv is None:
        md.num_attention_heads_kv = md.num_attention_heads
    pet = getattr(margs, "position_embedding_type", "absolute")
    md.position_embedding_type = getattr(pet, "name", str(pet))
    md.iteration = iteration
    md.previous_tensor_parallel_size = tp_size
    md.previous_pipeline_parallel_size = pp_size
    md.true_vocab_size = true_vocab_size(args)
    md.consumed_train_samples = getattr(margs, "consumed_train_samples", 0)
    md.consumed_valid_samples = getattr(margs, "consumed_valid_samples", 0)
    md.checkpoint_args = margs
    if getattr(args, "bf16", False):
        md.params_dtype = torch.bfloat16
    queue.put(md)

    def put(name, msg):
        print(f"sending {name}")
        msg["name"] = name
        queue.put(msg)

    def stage(pp_rank):
        if pp_rank == 0:
            files = [first] + [torch.load(_rank_file(base, t, 0, pp_size), map_location="cpu", weights_only=False)
                               for t in range(1, tp_size)]
        else:
            files = [torch.load(_rank_file(base, t, pp_rank, pp_size), map_location="cpu", weights_only=False)
                     for t in range(tp_size)]
        return files, [_normalise(f["model"]["language_model"]) for f in files]

    glu = md.glu_activation is not None
    total = 0
    files, parts = stage(0)
    # ---- embeddings
    msg = {"word embeddings": torch.cat([p[0]["word_embeddings.weight"] for p in parts], dim=0)}
    if md.position_embedding_type == "absolute" and "position_embeddings.weight" in parts[0][0]:
        msg["position embeddings"] = parts[0][0]["position_embeddings.weight"]
    if "tokentype_embeddings.weight" in parts[0][0]:
        msg["tokentype embeddings"] = parts[0][0]["tokentype_embeddings.weight"]
    put("embeddings", msg)
    # ---- untied lm head lives on the last stage; the protocol sends it right after the embeddings
    last_files, last_parts = (files, parts) if pp_size == 1 else stage(pp_size - 1)
    if not md.tie_embed_logits:
        put("lm_head", {"lm_head": torch.cat([p[2] for p in last_parts], dim=0)})
    # ---- transformer layers, stage by stage
    for pp_rank in range(pp_size):
        if pp_rank > 0:
            files, parts = (last_files, last_parts) if pp_rank == pp_size - 1 else stage(pp_rank)
        encs = [p[1] for p in parts]
        n_local = 1 + max(int(k.split(".")[1]) for k in encs[0] if k.startswith("layers."))
        for li in range(n_local):
            msg = {}
            for suffix, mkey, dim in LAYER_RULES:
                key = f"layers.{li}.{suffix}"
                if key not in encs[0]:
                    continue
                shards = [e[key] for e in encs]
                if dim is None:
                    msg[mkey] = shards[0]
                elif glu and mkey in GLU_KEYS:
                    msg[mkey] = merge_glu(shards, dim)
                else:
                    msg[mkey] = torch.cat(shards, dim=dim)
            put(f"transformer layer {total}", msg)
            total += 1
    assert total == md.num_layers, (total, md.num_layers)
    # ---- tail
    enc_last = last_parts[0][1]
    msg = {"weight": enc_last["final_layernorm.weight"]}
    if "final_layernorm.bias" in enc_last:
        msg["bias"] = enc_last["final_layernorm.bias"]
    put("final layernorm", msg)
    last_model = last_files[0]["model"]
    if md.model_type == "BERT":
        pooler = last_parts[0][3]
        if pooler is not None:
            put("pooler", {"weight": pooler["dense.weight"], "bias": pooler["dense.bias"]})
        if "lm_head" in last_model:
            h = last_model["lm_head"]
            m = {"dense weight": h["dense.weight"], "dense bias": h["dense.bias"],
                 "layernorm weight": h["layernorm.weight"], "layernorm bias": h["layernorm.bias"]}
            if "bias" in h:
                # vocab-parallel output bias (the reference's loader drops it and the saver re-initialises it to
                # zero, checkpoint_loader_megatron.py:318-331; it is a trained parameter, so it travels here)
                m["vocab bias"] = torch.cat([f["model"]["lm_head"]["bias"] for f in last_files], dim=0)
            put("lm head", m)
        if md.bert_binary_head and "binary_head" in last_model:
            put("binary head", {"weight": last_model["binary_head"]["weight"],
                                "bias": last_model["binary_head"]["bias"]})
    queue.put("done")

"""Training equivalence across parallel layouts on CPU/Gloo (the reference has no such test; SURVEY 4 asks for
TP/PP/DP invariance): one tp1/pp1 run saves its initial weights, the checkpoint resharder converts them to every other
layout, and the first optimizer steps must produce the same loss trajectory (fp32).  Also: optimizer-state
save -> load round trip for the plain and the distributed optimizer."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
from tests.dist_utils import run_distributed  # noqa: E402

MODEL = ["--model_name", "llama2", "--num_layers", "8", "--hidden_size", "32", "--num_attention_heads", "4",
         "--num_attention_heads_kv", "2", "--ffn_hidden_size", "64", "--seq_length", "16",
         "--max_position_embeddings", "16", "--micro_batch_size", "1", "--global_batch_size", "4",
         "--tokenizer_type", "NullTokenizer", "--vocab_file", "64", "--make_vocab_size_divisible_by", "8",
         "--data_type", "synthetic", "--train_iters", "10", "--lr", "1e-2", "--min_lr", "1e-2", "--lr_decay_style",
         "constant", "--weight_decay", "0.01", "--clip_grad", "1.0", "--hidden_dropout", "0.0", "--attention_dropout",
         "0.0", "--use_rms_norm", "--glu_activation", "swiglu", "--position_embedding_type", "rotary", "--no_bias_gelu_fusion",
         "--no_tie_embed_logits", "--eval_iters", "0", "--seed", "7", "--save_interval", "1000"]


def _train_worker(rank, world, argv, out_path, n_steps, save_first, save_last):
    import finetune
    from megatron_llm_b200 import get_args
    from megatron_llm_b200.checkpointing import save_checkpoint
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models.enums import ModelType
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.training import (_setup_model_and_optimizer, build_train_valid_test_data_iterators,
                                            train_step)
    initialize_megatron(extra_args_provider=finetune.extra_args, args_list=argv)
    args = get_args()
    model, optimizer, sched = _setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
    if save_first:
        save_checkpoint(1, model, optimizer, sched)   # iteration 0 is not a loadable checkpoint (same as the reference)
    if args.virtual_pipeline_model_parallel_size is not None:
        it = [build_train_valid_test_data_iterators(finetune.data_provider, args)[0] for _ in model]
    else:
        it = build_train_valid_test_data_iterators(finetune.data_provider, args)[0]
    for m in model:
        m.train()
    losses = []
    for step in range(n_steps):
        loss_dict, skipped, grad_norm, _ = train_step(finetune.forward_step, it, model, optimizer, sched)
        assert skipped == 0
        if ps.is_pipeline_last_stage(ignore_virtual=True):
            losses.append(float(loss_dict["lm loss"]))
        args.iteration = args.iteration + 1 if hasattr(args, "iteration") else step + 1
        args.consumed_train_samples += args.global_batch_size
    if save_last:
        save_checkpoint(args.iteration, model, optimizer, sched)
    if losses and ps.get_tensor_model_parallel_rank() == 0 and ps.get_data_parallel_rank() == 0:
        with open(out_path, "w") as f:
            json.dump(losses, f)
    if getattr(args, "cuda_graph_microbatch", False) and torch.cuda.is_available():
        # (GPU runs from tests/test_tp_model_gpu.py) NCCL teardown under live CUDA graphs can block: result is on disk
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


def _run(world, extra, out_path, n_steps=3, save_first=False, save_last=False):
    run_distributed(_train_worker, world, MODEL + extra, str(out_path), n_steps, save_first, save_last)
    with open(out_path) as f:
        return json.load(f)


def _reshard(src, dst, tp, pp):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util
    checkpoint_util.main(["--model_type", "llama2", "--load_dir", str(src), "--save_dir", str(dst),
                          "--target_tensor_parallel_size", str(tp), "--target_pipeline_parallel_size", str(pp),
                          "--true_vocab_size", "64"])


@pytest.fixture(scope="module")
def baseline(tmp_path_factory):
    d = tmp_path_factory.mktemp("baseline")
    losses = _run(1, ["--save", str(d / "ckpt")], d / "loss.json", save_first=True)
    assert len(losses) == 3 and all(l > 0 for l in losses)
    return d / "ckpt", losses


LAYOUTS = [
    ("tp2", 2, 2, 1, []),
    ("tp2_sp", 2, 2, 1, ["--sequence_parallel"]),
    ("pp2", 2, 1, 2, []),
    ("dp2", 2, 1, 1, []),
    ("dp2_distopt", 2, 1, 1, ["--use_distributed_optimizer"]),
    ("tp2_pp2", 4, 2, 2, []),
    # SP norm grads are TP-reduced after backward while DP buckets may already be in flight (deferred buckets, ddp.py)
    ("tp2_sp_dp2", 4, 2, 1, ["--sequence_parallel"]),
    ("tp2_sp_dp2_distopt", 4, 2, 1, ["--sequence_parallel", "--use_distributed_optimizer"]),
    ("pp2_dp2_recompute", 4, 1, 2, ["--recompute_granularity", "full", "--recompute_method", "uniform"]),
]


@pytest.mark.parametrize("name,world,tp,pp,extra", LAYOUTS, ids=[l[0] for l in LAYOUTS])
def test_layout_matches_baseline(baseline, tmp_path, name, world, tp, pp, extra):
    ckpt, ref = baseline
    load = ckpt
    if tp > 1 or pp > 1:
        load = tmp_path / "resharded"
        _reshard(ckpt, load, tp, pp)
    losses = _run(world, ["--tensor_model_parallel_size", str(tp), "--pipeline_model_parallel_size", str(pp),
                          "--load", str(load), "--finetune", "--no_load_optim", "--no_load_rng"] + extra,
                  tmp_path / "loss.json")
    assert losses == pytest.approx(ref, rel=2e-4, abs=2e-4), (name, losses, ref)


def test_tied_embeddings_pipeline_and_zero(tmp_path):
    """Tied embeddings with PP=2 x DP=2 and ZeRO-1: the last stage owns a ``shared`` copy of the word embedding whose
    gradient is all-reduced over the embedding group after backward -- its bucket must be a deferred one, and the
    reduce-scatter must come after that all-reduce."""
    tied = [("gpt" if a == "llama2" else a) for a in MODEL if a != "--no_tie_embed_logits"]

    def run(world, more, out, save_first=False):
        run_distributed(_train_worker, world, tied + more, str(out), 3, save_first, False)
        with open(out) as f:
            return json.load(f)
    ref = run(1, ["--save", str(tmp_path / "ckpt")], tmp_path / "ref.json", True)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util
    load = tmp_path / "pp2"
    checkpoint_util.main(["--model_type", "GPT", "--load_dir", str(tmp_path / "ckpt"), "--save_dir", str(load),
                          "--target_tensor_parallel_size", "1", "--target_pipeline_parallel_size", "2",
                          "--true_vocab_size", "64"])
    got = run(4, ["--pipeline_model_parallel_size", "2", "--use_distributed_optimizer", "--load", str(load),
                  "--finetune", "--no_load_optim", "--no_load_rng"], tmp_path / "pp2_dp2_zero.json")
    assert got == pytest.approx(ref, rel=2e-4, abs=2e-4), (got, ref)


@pytest.mark.parametrize("extra", [[], ["--use_distributed_optimizer"]], ids=["dp2", "dp2_distopt"])
def test_tied_embeddings_data_parallel(tmp_path, extra):
    """Tied embedding + fused wgrad accumulation: the word-embedding weight gets the LM-head wgrad (GEMM epilogue,
    reported through the ready callback) and the lookup gradient (autograd hook); its DP bucket must not be reduced
    before both have landed."""
    tied = [("gpt" if a == "llama2" else a) for a in MODEL if a != "--no_tie_embed_logits"]   # (Llama never ties)

    def run(world, more, out):
        run_distributed(_train_worker, world, tied + more, str(out), 3, world == 1, False)
        with open(out) as f:
            return json.load(f)
    ref = run(1, ["--save", str(tmp_path / "ckpt")], tmp_path / "ref.json")
    got = run(2, ["--load", str(tmp_path / "ckpt"), "--finetune", "--no_load_optim", "--no_load_rng"] + extra,
              tmp_path / "dp2.json")
    assert got == pytest.approx(ref, rel=2e-4, abs=2e-4), (got, ref)


def test_fp32_residual_connection_with_bf16_params(tmp_path):
    """--fp32_residual_connection + --bf16: the residual stream (and the pipeline p2p tensors) are fp32, the norms hand
    parameter-dtype activations to the GEMMs.  Must train, stay close to the plain bf16 run, and agree between PP=1 and
    PP=2 (the p2p buffers take the fp32 dtype)."""
    def run(world, more, out, save_first=False):
        run_distributed(_train_worker, world, MODEL + ["--bf16"] + more, str(out), 3, save_first, False)
        with open(out) as f:
            return json.load(f)
    plain = run(1, ["--save", str(tmp_path / "ckpt")], tmp_path / "plain.json", True)
    load = ["--load", str(tmp_path / "ckpt"), "--finetune", "--no_load_optim", "--no_load_rng"]
    fp32res = run(1, ["--fp32_residual_connection"] + load, tmp_path / "fp32res.json")
    assert fp32res == pytest.approx(plain, rel=2e-2), (fp32res, plain)
    assert fp32res != plain                                   # the flag does something
    pp2 = tmp_path / "pp2"
    _reshard(tmp_path / "ckpt", pp2, 1, 2)
    got = run(2, ["--fp32_residual_connection", "--pipeline_model_parallel_size", "2", "--load", str(pp2), "--finetune",
                  "--no_load_optim", "--no_load_rng"], tmp_path / "pp2.json")
    assert got == pytest.approx(fp32res, rel=5e-3), (got, fp32res)


def test_interleaved_schedule_matches_baseline(baseline, tmp_path):
    """pp=4 with 2 virtual chunks per stage (8 layers, 1 layer per chunk): the resharder has no interleaved layout
    (like the reference), so the pp=8-style per-layer files are regrouped into model0/model1 here."""
    ckpt, ref = baseline
    flat = tmp_path / "pp8"
    _reshard(ckpt, flat, 1, 8)
    dst = tmp_path / "vpp"
    for stage in range(4):
        chunks = []
        for v in range(2):
            layer = v * 4 + stage          # virtual chunk v of stage s owns layer v * pp + s
            sd = torch.load(flat / "iter_0000001" / f"mp_rank_00_{layer:03d}" / "model_optim_rng.pt", weights_only=False)
            chunks.append(sd)
        out = dict(chunks[0])
        out["model0"], out["model1"] = chunks[0]["model"], chunks[1]["model"]
        del out["model"]
        out["args"].pipeline_model_parallel_size = 4
        os.makedirs(dst / "iter_0000001" / f"mp_rank_00_{stage:03d}")
        torch.save(out, dst / "iter_0000001" / f"mp_rank_00_{stage:03d}" / "model_optim_rng.pt")
    (dst / "latest_checkpointed_iteration.txt").write_text("1")
    losses = _run(4, ["--pipeline_model_parallel_size", "4", "--num_layers_per_virtual_pipeline_stage", "1", "--load",
                      str(dst), "--finetune", "--no_load_optim", "--no_load_rng"], tmp_path / "loss.json")
    assert losses == pytest.approx(ref, rel=2e-4, abs=2e-4), (losses, ref)


@pytest.mark.parametrize("distopt", [False, True], ids=["optimizer", "distributed_optimizer"])
def test_resume_reproduces_training(tmp_path, distopt):
    """2 steps + save + 2 steps  ==  load + 2 steps (weights, optimizer moments, LR schedule, data position)."""
    extra = ["--use_distributed_optimizer"] if distopt else []
    world = 2
    full = _run(world, extra + ["--save", str(tmp_path / "unused")], tmp_path / "a.json", n_steps=4)
    first = _run(world, extra + ["--save", str(tmp_path / "ckpt")], tmp_path / "b.json", n_steps=2, save_last=True)
    assert first == pytest.approx(full[:2], rel=1e-6)
    resumed = _run(world, extra + ["--load", str(tmp_path / "ckpt")], tmp_path / "c.json", n_steps=2)
    assert resumed == pytest.approx(full[2:], rel=2e-5, abs=2e-5), (resumed, full)


def _flat_tensors(sd, prefix=""):
    out = {}
    for k, v in sd.items():
        if isinstance(v, dict):
            out.update(_flat_tensors(v, prefix + k + "."))
        elif torch.is_tensor(v):
            out[prefix + k] = v
    return out


FAMILIES = {
    "falcon": (["--model_name", "falcon", "--num_layers", "4", "--hidden_size", "32", "--num_attention_heads", "4",
                "--num_attention_heads_kv", "2", "--parallel_attn", "--parallel_layernorm", "--position_embedding_type",
                "rotary", "--no_bias_gelu_fusion"], "falcon"),
    "gpt": (["--model_name", "gpt", "--num_layers", "4", "--hidden_size", "32", "--num_attention_heads", "4"], "GPT"),
}


@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_reshard_other_families_tp2_pp2_and_back(tmp_path, family):
    """tools/checkpoint_util.py for the Falcon (parallel attention + two norms, GQA, tied head) and GPT (biases, learned
    positions, tied head) layouts: 1x1 -> TP2 x PP2 trains like the original (with sequence parallelism), and merging
    back to 1x1 returns every tensor bit for bit."""
    model, model_type = FAMILIES[family]
    common = MODEL[MODEL.index("--seq_length"):]
    common = [a for a in common if a not in ("--use_rms_norm", "--no_tie_embed_logits", "--no_bias_gelu_fusion")]
    for flag in ("--glu_activation", "--position_embedding_type"):
        i = common.index(flag)
        del common[i:i + 2]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util

    def run(world, more, out, save_first=False):
        run_distributed(_train_worker, world, model + common + more, str(out), 3, save_first, False)
        with open(out) as f:
            return json.load(f)
    ckpt = tmp_path / "ckpt"
    ref = run(1, ["--save", str(ckpt)], tmp_path / "ref.json", True)
    sharded, back = tmp_path / "tp2pp2", tmp_path / "back"
    checkpoint_util.main(["--model_type", model_type, "--load_dir", str(ckpt), "--save_dir", str(sharded),
                          "--target_tensor_parallel_size", "2", "--target_pipeline_parallel_size", "2",
                          "--true_vocab_size", "64"])
    got = run(4, ["--tensor_model_parallel_size", "2", "--pipeline_model_parallel_size", "2", "--sequence_parallel",
                  "--load", str(sharded), "--finetune", "--no_load_optim", "--no_load_rng"], tmp_path / "got.json")
    assert got == pytest.approx(ref, rel=2e-4, abs=2e-4), (got, ref)
    checkpoint_util.main(["--model_type", model_type, "--load_dir", str(sharded), "--save_dir", str(back),
                          "--target_tensor_parallel_size", "1", "--target_pipeline_parallel_size", "1",
                          "--true_vocab_size", "64"])
    a = _flat_tensors(torch.load(ckpt / "iter_0000001" / "mp_rank_00" / "model_optim_rng.pt", weights_only=False)["model"])
    sub = [p for p in os.listdir(back) if p.startswith("iter") or p == "release"][0]
    b = _flat_tensors(torch.load(back / sub / "mp_rank_00" / "model_optim_rng.pt", weights_only=False)["model"])
    assert set(a) == set(b)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_standalone_embedding_stage_trains(tmp_path):
    """--standalone_embedding_stage with PP=3: stage 0 holds only the embedding (a no-op layer), the 8 layers are split
    over stages 1 and 2 and numbered 1..8 (the reference numbers them 5..12 and indexes its per-layer tables out of
    range).  Smoke test: three steps, finite decreasing-ish losses."""
    losses = _run(3, ["--pipeline_model_parallel_size", "3", "--standalone_embedding_stage"], tmp_path / "loss.json")
    assert len(losses) == 3 and all(l == l and 3.0 < l < 6.0 for l in losses), losses

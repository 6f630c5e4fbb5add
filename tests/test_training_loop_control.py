"""Training-loop control of the real entry point (``python finetune.py`` on CPU, synthetic data): periodic saves and
the tracker file, ``--exit_interval``, resume from ``--load`` (iteration, consumed samples, loss continuity),
``--load_iters``, ``--finetune``, ``--skip_iters``, batch-size ramp-up, ``--exit_duration_in_mins``, the SIGTERM
handler, and ``--use_checkpoint_args``.  (the reference has no tests for any of these.)"""
import os
import re
import signal
import subprocess
import sys
import time

import pytest

from tests.dist_utils import free_port

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
ARCH = ["--model_name", "llama2", "--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4",
        "--num_attention_heads_kv", "2", "--ffn_hidden_size", "64", "--use_rms_norm", "--glu_activation", "swiglu",
        "--position_embedding_type", "rotary", "--no_bias_gelu_fusion", "--no_tie_embed_logits"]
RUN = ["--seq_length", "16", "--max_position_embeddings", "16", "--micro_batch_size", "1", "--global_batch_size", "2",
       "--tokenizer_type", "NullTokenizer", "--vocab_file", "64", "--make_vocab_size_divisible_by", "8",
       "--data_type", "synthetic", "--lr", "1e-2", "--min_lr", "1e-2", "--lr_decay_style", "constant",
       "--hidden_dropout", "0", "--attention_dropout", "0", "--eval_iters", "1", "--eval_interval", "1000",
       "--log_interval", "1", "--seed", "7", "--num_workers", "0"]
LINE = re.compile(r"iteration\s+(\d+)/\s*(\d+) \| consumed samples:\s+(\d+) .*?global batch size:\s+(\d+) \| "
                  r"lm loss: ([0-9.E+-]+)")


def _env():
    return dict(os.environ, MLB200_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1",
                MASTER_PORT=str(free_port()), RANK="0", WORLD_SIZE="1")


def _run(extra, arch=ARCH, expect_ok=True):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "finetune.py")] + arch + RUN + extra, env=_env(),
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    if expect_ok:
        assert r.returncode == 0, r.stdout[-2500:] + "\n" + r.stderr[-2500:]
    return r


def _iters(out):
    """[(iteration, consumed samples, global batch size, loss)] from the log lines."""
    return [(int(m[1]), int(m[3]), int(m[4]), float(m[5])) for m in LINE.finditer(out)]


def test_periodic_save_exit_interval_and_resume(tmp_path):
    ckpt = str(tmp_path / "ckpt")
    first = _run(["--train_iters", "6", "--save", ckpt, "--save_interval", "2", "--exit_interval", "4"])
    assert [i for i, *_ in _iters(first.stdout)] == [1, 2, 3, 4] and "exiting program at iteration 4" in first.stdout
    assert sorted(os.listdir(ckpt)) == ["iter_0000002", "iter_0000004", "latest_checkpointed_iteration.txt"]
    assert open(os.path.join(ckpt, "latest_checkpointed_iteration.txt")).read().strip() == "4"
    # uninterrupted reference run: the resumed run must continue its loss trajectory exactly (optimizer state, RNG and
    # the position in the data all come from the checkpoint)
    straight = _iters(_run(["--train_iters", "6"]).stdout)
    resumed = _run(["--train_iters", "6", "--load", ckpt, "--save", ckpt, "--save_interval", "100"])
    got = _iters(resumed.stdout)
    assert [i for i, *_ in got] == [5, 6] and [c for _, c, *_ in got] == [10, 12]
    for (_, _, _, a), (_, _, _, b) in zip(got, straight[4:]):
        assert a == pytest.approx(b, rel=1e-5)
    # --load_iters picks an older checkpoint; --finetune loads weights only and restarts the counters
    # (the schedule length is part of the checkpoint and is checked: resumed runs keep --train_iters)
    older = _iters(_run(["--train_iters", "6", "--load", ckpt, "--load_iters", "2", "--exit_interval", "3"]).stdout)
    assert [i for i, *_ in older] == [3] and older[0][3] == pytest.approx(straight[2][3], rel=1e-5)
    mismatch = _run(["--train_iters", "3", "--load", ckpt], expect_ok=False)
    assert mismatch.returncode != 0 and "total number of iterations" in mismatch.stderr
    fresh = _iters(_run(["--train_iters", "2", "--load", ckpt, "--finetune"]).stdout)
    assert [(i, c) for i, c, *_ in fresh] == [(1, 2), (2, 4)]


def test_skip_iters_and_batch_size_rampup():
    out = _run(["--train_iters", "3", "--skip_iters", "1"]).stdout
    assert "skipping iteration 1" in out
    assert "number of skipped iterations:   1" in out
    # (batch-size ramp-up belongs to sample-based training: --train_samples, not --train_iters)
    ramp = _iters(_run(["--train_samples", "20", "--global_batch_size", "6", "--rampup_batch_size", "2", "2",
                        "8"]).stdout)
    # 2 -> 6 in steps of 2 over 8 samples: 4 samples per increment
    assert [g for _, _, g, _ in ramp] == [2, 2, 4, 6, 6]
    assert [c for _, c, _, _ in ramp] == [2, 4, 8, 14, 20]


def test_exit_duration_saves_and_exits(tmp_path):
    ckpt = str(tmp_path / "ckpt")
    out = _run(["--train_iters", "100000", "--save", ckpt, "--save_interval", "100000", "--exit_duration_in_mins",
                "0.02"]).stdout
    done = _iters(out)[-1][0]
    assert 1 <= done < 100000 and "exiting program after" in out
    assert open(os.path.join(ckpt, "latest_checkpointed_iteration.txt")).read().strip() == str(done)


def test_sigterm_checkpoints_and_exits(tmp_path):
    ckpt = str(tmp_path / "ckpt")
    cmd = [sys.executable, "-u", os.path.join(ROOT, "finetune.py")] + ARCH + RUN + [
        "--train_iters", "100000", "--save", ckpt, "--save_interval", "100000", "--exit_signal_handler"]
    p = subprocess.Popen(cmd, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT)
    seen, deadline = [], time.time() + 300
    try:
        for line in p.stdout:
            seen.append(line)
            if " iteration        3/" in line:
                p.send_signal(signal.SIGTERM)
                break
            assert time.time() < deadline, "training did not start"
        rest, _ = p.communicate(timeout=120)
    finally:
        if p.poll() is None:
            p.kill()
    out = "".join(seen) + rest
    assert p.returncode == 0, out[-3000:]
    assert "exiting program after receiving SIGTERM" in out
    saved = int(open(os.path.join(ckpt, "latest_checkpointed_iteration.txt")).read().strip())
    assert saved >= 3 and os.path.isdir(os.path.join(ckpt, f"iter_{saved:07d}"))


def test_checkpoint_args_are_checked_and_can_be_adopted(tmp_path):
    ckpt = str(tmp_path / "ckpt")
    _run(["--train_iters", "2", "--save", ckpt, "--save_interval", "1", "--exit_interval", "1"])
    wrong = [a if a != "2" or ARCH[i - 1] != "--num_layers" else "3" for i, a in enumerate(ARCH)]
    bad = _run(["--train_iters", "2", "--load", ckpt], arch=wrong, expect_ok=False)
    assert bad.returncode != 0 and "num_layers value from checkpoint (2)" in bad.stderr
    # --use_checkpoint_args: architecture flags that are NOT given come from the checkpoint (sizes, heads, GLU, RoPE),
    # the normalisation / embedding-tying / position-embedding choices are always taken from it
    minimal = ["--model_name", "llama2", "--no_bias_gelu_fusion"]
    ok = _run(["--train_iters", "2", "--load", ckpt, "--use_checkpoint_args"], arch=minimal)
    assert "Setting num_layers to 2 from checkpoint" in ok.stdout and "Setting use_rms_norm to True" in ok.stdout
    assert [i for i, *_ in _iters(ok.stdout)] == [2]


def test_validation_metrics_tensorboard_and_timers(tmp_path):
    """--metrics plug-ins are evaluated during validation; the TensorBoard writer gets the scalars (loss, lr, timers,
    memory, validation metrics) and the per-interval timer report is printed at --timing_log_level 2."""
    tb = tmp_path / "tb"
    out = _run(["--train_iters", "4", "--eval_interval", "2", "--eval_iters", "2", "--metrics", "perplexity", "accuracy",
                "count_loss_mask", "--tensorboard_dir", str(tb), "--log_timers_to_tensorboard",
                "--log_memory_to_tensorboard", "--log_params_norm", "--timing_log_level", "2"]).stdout
    val = [l for l in out.splitlines() if "validation loss at iteration" in l]
    assert len(val) == 2
    for key in ("lm loss value", "lm loss PPL", "ppl value", "lm accuracy value", "count loss mask value: 1.600000E+01"):
        assert key in val[0], (key, val[0])
    assert "validation loss at the end of training for test data" in out and "params norm:" in out
    assert "forward-backward" in out and "optimizer" in out            # timer report
    events = [f for f in os.listdir(tb) if f.startswith("events.out.tfevents")]
    assert events
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(tb))
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    for tag in ("learning-rate", "lm loss", "lm loss validation", "ppl validation", "lm accuracy validation"):
        assert tag in tags, (tag, sorted(tags)[:40])
    assert [e.step for e in acc.Scalars("lm loss")] == [1, 2, 3, 4]


def test_checkpoints_pickled_with_legacy_module_paths_load(tmp_path):
    """Checkpoints written by the reference (``megatron.*`` class paths, e.g. the args' position-embedding enum) and by
    its ancestors (a loss scaler pickled as ``fp16.loss_scaler.DynamicLossScaler`` / ``megatron.fp16_deprecated...``)
    must unpickle: the file written here is re-pickled with those module paths and then resumed from."""
    import pickle
    import types
    import torch
    ckpt = str(tmp_path / "ckpt")
    _run(["--train_iters", "3", "--save", ckpt, "--save_interval", "2", "--exit_interval", "2"])
    path = os.path.join(ckpt, "iter_0000002", "mp_rank_00", "model_optim_rng.pt")
    state = torch.load(path, weights_only=False)
    sys.path.insert(0, ROOT)
    from megatron_llm_b200.fp16_deprecated import loss_scaler as shim
    # objects whose classes claim to live in the legacy modules
    old = types.ModuleType("fp16.loss_scaler")
    old.DynamicLossScaler = type("DynamicLossScaler", (), {"__module__": "fp16.loss_scaler"})
    mid = types.ModuleType("megatron.fp16_deprecated.loss_scaler")
    mid.LossScaler = type("LossScaler", (), {"__module__": "megatron.fp16_deprecated.loss_scaler"})
    sys.modules["fp16"], sys.modules["fp16.loss_scaler"] = types.ModuleType("fp16"), old
    saved_mid = sys.modules.get("megatron.fp16_deprecated.loss_scaler")
    sys.modules["megatron.fp16_deprecated.loss_scaler"] = mid
    try:
        a, b = old.DynamicLossScaler(), mid.LossScaler()
        a.cur_scale, b.cur_scale = 4096.0, 1.0
        state["legacy_scaler"], state["legacy_scaler2"] = a, b
        torch.save(state, path)
    finally:
        sys.modules.pop("fp16", None), sys.modules.pop("fp16.loss_scaler", None)
        if saved_mid is None:
            sys.modules.pop("megatron.fp16_deprecated.loss_scaler", None)
        else:
            sys.modules["megatron.fp16_deprecated.loss_scaler"] = saved_mid
    with open(path, "rb") as f:
        assert b"fp16.loss_scaler" in f.read()
    assert hasattr(shim, "DynamicLossScaler") and hasattr(shim, "LossScaler")
    r = _run(["--train_iters", "3", "--load", ckpt])
    assert [i for i, *_ in _iters(r.stdout)] == [3]

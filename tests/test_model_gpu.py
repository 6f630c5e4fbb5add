"""Model-level parity on a B200: the same tiny Llama / Falcon / GPT training steps with the sm_100a kernels in bf16 vs.
the plain PyTorch operator path (MLB200_DISABLE_KERNELS=1) in FP32 must agree on loss (1 %) and gradient norm (3 %).
Both runs build their weights from the same fp32 CPU random stream (--use_cpu_initialization), so the oracle's
weights are the un-rounded values of the bf16 run's, and both see the same fixed batch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, json, torch
sys.path.insert(0, %(root)r)
os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=%(port)r)
import finetune
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.training import setup_model_and_optimizer, train_step
argv = %(argv)r.split()
initialize_megatron(finetune.extra_args, {}, args_list=argv)
model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
assert next(model[0].parameters()).is_cuda, "the model must live on the GPU in this test"
def it():
    g = torch.Generator().manual_seed(0)
    batches = [torch.randint(0, 1000, (2, %(seq)d + 1), generator=g) for _ in range(2)]   # one step = 2 micro-batches
    while True:                                      # the SAME step every time: the loss must go down
        for b in batches:
            yield {"text": b}
data = it()
out = []
for step in range(3):
    loss, skipped, gnorm, _ = train_step(finetune.forward_step, data, model, opt, sched)
    out.append((loss["lm loss"].item(), gnorm.item()))
print("RESULT " + json.dumps(out))
'''

COMMON = ("--num_layers 2 --hidden_size 256 --num_attention_heads 2 --seq_length 256 --max_position_embeddings 256 "
          "--micro_batch_size 2 --global_batch_size 4 --train_iters 10 --lr 3e-4 --bf16 --hidden_dropout 0 "
          "--attention_dropout 0 --tokenizer_type NullTokenizer --vocab_file 1024 --data_type synthetic "
          "--log_interval 100 --eval_iters 0 --eval_interval 1000 --num_workers 0 --lr_decay_style constant "
          "--use_flash_attn --position_embedding_type rotary --use_cpu_initialization --clip_grad 0 ")
CONFIGS = {
    "llama": COMMON + "--model_name llama2 --use_rms_norm --glu_activation swiglu --no_tie_embed_logits "
                      "--ffn_hidden_size 704 --num_attention_heads_kv 1",
    "falcon": COMMON + "--model_name falcon --parallel_attn --parallel_layernorm --num_attention_heads_kv 1",
    "gpt": COMMON.replace("--position_embedding_type rotary", "").replace("--use_flash_attn", "") +
           "--model_name gpt --use_bias",
}


def _run(argv, disable, port, fp32=False):
    env = dict(os.environ, MLB200_FORCE_CPU="0")
    env["MLB200_DISABLE_KERNELS"] = "1" if disable else "0"
    if fp32:
        argv = argv.replace("--bf16", "")
    code = SCRIPT % {"root": ROOT, "port": str(port), "argv": argv, "seq": 256}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("name", list(CONFIGS))
def test_kernels_match_torch_path(name):
    a = _run(CONFIGS[name], disable=False, port=29610)
    b = _run(CONFIGS[name], disable=True, port=29611, fp32=True)
    for (la, ga), (lb, gb) in zip(a, b):
        assert la == la and ga == ga, "nan"
        assert abs(la - lb) < 1e-2 * max(1.0, abs(lb)), (a, b)
        assert abs(ga - gb) < 3e-2 * max(1e-3, abs(gb)), (a, b)
    assert a[-1][0] < a[0][0]          # same batch every step: the loss goes down
    assert b[-1][0] < b[0][0]


def test_cuda_graph_microbatch_matches_eager():
    """Replaying the micro-batch from a CUDA graph (--cuda_graph_microbatch) trains exactly like the eager schedule:
    same kernels, same order -> identical loss / grad-norm trajectory (first steps are eager warm-up + capture)."""
    argv = CONFIGS["llama"].replace("--train_iters 10", "--train_iters 20")
    script = SCRIPT.replace("for step in range(3):", "for step in range(6):")
    def run(extra, port):
        env = dict(os.environ, MLB200_DISABLE_KERNELS="0", MLB200_FORCE_CPU="0")
        code = script % {"root": ROOT, "port": str(port), "argv": argv + extra, "seq": 256}
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        import json
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
        return json.loads(line[len("RESULT "):])
    eager = run("", 29612)
    graph = run(" --cuda_graph_microbatch", 29613)
    for (la, ga), (lb, gb) in zip(eager, graph):
        assert abs(la - lb) < 2e-3 * max(1.0, abs(la)), (eager, graph)
        assert abs(ga - gb) < 2e-2 * max(1e-3, abs(ga)), (eager, graph)


def test_fp16_with_dynamic_loss_scaling_trains():
    """--fp16: the tcgen05 GEMMs run with fp16 operands, every other kernel has an fp16 instantiation, and the flat
    optimizer unscales / skips on overflow with the dynamic scaler.  The loss must stay finite and go down, with
    kernels on and off."""
    argv = CONFIGS["llama"].replace("--bf16", "--fp16 --initial_loss_scale 4096 --loss_scale_window 2")
    script = SCRIPT.replace("for step in range(3):", "for step in range(5):")
    for disable in (False, True):
        # (fp16 ATTENTION stays on the path this test was validated with on hardware, the library: the tcgen05 fp16
        # attention variant was written after the last GPU run and has its own, isolated first-run checks in
        # tests/test_z_attention_variants_gpu.py -- this test is about the fp16 GEMMs / optimizer / loss scaler)
        env = dict(os.environ, MLB200_DISABLE_KERNELS="1" if disable else "0", MLB200_FORCE_CPU="0",
                   MLB200_ATTN_FP16="0")
        code = script % {"root": ROOT, "port": "29614", "argv": argv, "seq": 256}
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        import json
        out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
        losses = [l for l, _ in out]
        assert all(l == l and l < 20 for l in losses), out
        assert losses[-1] < losses[0] + 0.05, out

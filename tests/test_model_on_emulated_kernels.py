"""Whole training steps of tiny models (Llama in the suite; MLB200_EMU_MODELS=llama,falcon,gpt for all three families)
through the REAL kernel sources, on the CPU.

``tests/test_model_gpu.py`` checks on a B200 that the model trains the same with the hand-written kernels as with the
plain PyTorch operators.  This is the same comparison without a GPU: ``ops._ext.load()`` returns the emulated extension
(tests/emu/emu_extension.py: tcgen05 GEMMs and attention on the functional model, SIMT kernels on the thread shim) and
the ``is_cuda`` predicates of ``ops/__init__.py`` are lifted, so every operator of the step -- embedding gather /
scatter-add, RMSNorm with fused residual, RoPE on the packed QKV, NT / NN / TN GEMMs (wgrad accumulated into fp32
main_grad), packed-QKV attention forward and backward, SwiGLU, the vocab-parallel cross-entropy kernels, the flat
grad-norm + clip + AdamW -- calls the kernel the GPU build calls, with the arguments the Python stack really passes.
(One layer, two steps of one micro-batch in the suite; MLB200_EMU_MODEL_LAYERS / _STEPS / _MICROBATCHES widen it: two layers
x three steps x two micro-batches of two sequences take six minutes and pass as well.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, os, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests", "emu"))
os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=%(port)r, MLB200_FORCE_CPU="1")
EMULATE = %(emulate)r
from megatron_llm_b200 import ops
from megatron_llm_b200.ops import _ext, attention_sm100
ext = None
if EMULATE:
    import emu_extension
    ext = emu_extension.FullEmuExtension(%(build)r)
    _ext.load = lambda: ext
    ops.cuda_ops_available = lambda t: True
    _16 = (torch.bfloat16, torch.float16)
    ops._gemm_ok = lambda *ts: all(t.dtype in _16 and t.dtype == ts[0].dtype for t in ts)
    ops._norm_kernel_ok = lambda x, w: (x.dtype == w.dtype and x.dtype in _16 + (torch.float32,) and x.size(-1) %% 8 == 0
                                        and x.size(-1) <= 8192)
    ops._embedding_kernel_ok = lambda ids, weight: (ids.dim() == 2 and weight.dim() == 2 and weight.is_contiguous()
                                                    and weight.size(1) %% 8 == 0)
    attention_sm100.packed_supported = lambda mixed, nkv, g, hn, p: (mixed.dtype in _16 and hn in (64, 128) and mixed.dim() == 3
                                                                     and mixed.stride(2) == 1 and mixed.size(0) %% 128 == 0
                                                                     and p == 0.0)
import finetune
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.training import setup_model_and_optimizer, train_step
COMMON = ("--num_layers %(layers)d --hidden_size 128 --num_attention_heads 2 --seq_length 128 --max_position_embeddings 128 "
          "--micro_batch_size 1 --global_batch_size %(gbs)d --train_iters 10 --lr 3e-3 --bf16 --hidden_dropout 0 "
          "--attention_dropout 0 --tokenizer_type NullTokenizer --vocab_file 248 --data_type synthetic --log_interval 100 "
          "--eval_iters 0 --eval_interval 1000 --num_workers 0 --lr_decay_style constant --use_cpu_initialization "
          "--clip_grad 1.0 --kv_channels 64 ")
MODELS = {
    "llama": "--model_name llama2 --use_rms_norm --glu_activation swiglu --no_tie_embed_logits --ffn_hidden_size 256 "
             "--num_attention_heads_kv 1 --use_flash_attn --position_embedding_type rotary",
    "falcon": "--model_name falcon --parallel_attn --parallel_layernorm --num_attention_heads_kv 1 --use_flash_attn "
              "--position_embedding_type rotary",
    "gpt": "--model_name gpt --use_bias",        # LayerNorm + bias, GeLU, learned positions, tied embeddings, softmax kernels
}
argv = (COMMON + MODELS[%(model)r]).split()
initialize_megatron(finetune.extra_args, {}, args_list=argv)
model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
def it():
    g = torch.Generator().manual_seed(0)
    batches = [torch.randint(0, 240, (1, 129), generator=g) for _ in range(%(gbs)d)]      # one step = all of them
    while True:
        for b in batches:
            yield {"text": b}
data = it()
out = []
for step in range(%(steps)d):
    loss, skipped, gnorm, _ = train_step(finetune.forward_step, data, model, opt, sched)
    out.append((loss["lm loss"].item(), float(gnorm)))
print("RESULT " + json.dumps({"steps": out, "calls": ext.calls if ext else {}}))
'''


LAYERS, STEPS = int(os.environ.get("MLB200_EMU_MODEL_LAYERS", "1")), int(os.environ.get("MLB200_EMU_MODEL_STEPS", "2"))


def _run(model, emulate, port, build):
    code = SCRIPT % {"root": ROOT, "port": str(port), "emulate": emulate, "build": build, "layers": LAYERS, "steps": STEPS,
                     "model": model,
                     "gbs": int(os.environ.get("MLB200_EMU_MODEL_MICROBATCHES", "1"))}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-4000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


EXPECTED = {
    "llama": ("gemm", "norm_fwd", "norm_bwd", "rope_qkv", "glu_fwd", "glu_bwd", "embedding_fwd", "embedding_bwd", "ce_stats",
              "ce_bwd", "attn_fwd_packed", "attn_bwd_packed", "sqnorm_flat", "clip_coef", "adamw_flat"),
    "falcon": ("gemm", "norm_fwd", "norm_bwd", "rope_qkv", "gelu", "embedding_fwd", "ce_stats", "attn_fwd_packed",
               "attn_bwd_packed", "adamw_flat"),
    "gpt": ("gemm", "norm_fwd", "norm_bwd", "gelu", "bias_dropout_add", "softmax_fwd", "softmax_bwd", "embedding_fwd",
            "ce_stats", "ce_bwd", "adamw_flat"),
}


import pytest  # noqa: E402


@pytest.mark.parametrize("model", os.environ.get("MLB200_EMU_MODELS", "llama").split(","))
def test_training_steps_through_the_emulated_kernels(tmp_path, model):
    emu = _run(model, True, 29731, str(tmp_path / "ext"))
    ref = _run(model, False, 29732, "")
    for name in EXPECTED[model]:                     # every kernel family of the step was really called
        assert emu["calls"].get(name, 0) > 0, (name, emu["calls"])
    for (la, ga), (lb, gb) in zip(emu["steps"], ref["steps"]):
        assert la == la and abs(la - lb) < 2e-2 * max(1.0, abs(lb)), (emu["steps"], ref["steps"])
        assert abs(ga - gb) < 6e-2 * max(1e-3, abs(gb)), (emu["steps"], ref["steps"])
    assert emu["steps"][-1][0] < emu["steps"][0][0]            # the same step every time: the loss goes down

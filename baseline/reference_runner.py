"""Reference arm of bench.py: runs the UNMODIFIED reference checkout in ``baseline/_ref`` through its own public
API (``initialize_megatron`` + ``pretrain`` with the stock ``finetune.py`` model_provider / forward_step) on the
same metric and config, and prints the same JSON line with ``"impl": "reference"``.

None of this repo's models, kernels or engine are on that path.  The only shims are environmental:
  * ``apex`` is not installed      -> stub modules: FusedAdam/FusedSGD = torch.optim.AdamW(fused=True)/SGD,
                                      amp_C.multi_tensor_l2norm/scale = torch._foreach_* ; fused_layer_norm_cuda =
                                      ATen native_layer_norm (fwd/bwd); run with
                                      ``--no_gradient_accumulation_fusion`` (apex wgrad extension absent)
  * nvFuser flags removed in torch -> ``set_jit_fusion_options`` is a no-op
  * no tokenizer / dataset files   -> synthetic tokenizer object + synthetic dataset via the data_provider hook
Numbers produced here must be labelled "reference (shimmed: apex->torch fused AdamW)".

Timing: ``megatron.training.train_step`` is wrapped to drop CUDA events after W warm-up steps and after W+K steps
(barrier + synchronize on both sides, max over ranks).  The reference's stock loop already copies every micro-batch
host->device (pinned DataLoader + ``broadcast_data``) and reads the loss back every step at ``--log_interval 1``,
so its device-timed and end-to-end numbers coincide.
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _install_apex_stub():
    import torch

    apex = types.ModuleType("apex")
    optimizers = types.ModuleType("apex.optimizers")
    mta = types.ModuleType("apex.multi_tensor_apply")
    amp_C = types.ModuleType("amp_C")

    class FusedAdam(torch.optim.AdamW):
        def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                     weight_decay=0.0, amsgrad=False, set_grad_none=True, **kw):
            super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                             fused=torch.cuda.is_available())

    class FusedSGD(torch.optim.SGD):
        def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, **kw):
            super().__init__(params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                             nesterov=nesterov)

    def multi_tensor_applier(op, noop_flag, tensor_lists, *args):
        return op(2048 * 32, noop_flag, tensor_lists, *args)

    def multi_tensor_l2norm(chunk, noop, tensor_lists, per_tensor=False):
        ts = tensor_lists[0]
        norms = torch._foreach_norm(ts, 2.0)
        stacked = torch.stack([n.float() for n in norms])
        return torch.linalg.vector_norm(stacked, 2.0).view(1), stacked if per_tensor else None

    def multi_tensor_scale(chunk, noop, tensor_lists, scale):
        src, dst = tensor_lists
        for s, d in zip(src, dst):
            if s.data_ptr() == d.data_ptr() and s.dtype == d.dtype:
                d.mul_(scale)
            else:
                d.copy_(s.to(d.dtype) * scale if scale != 1.0 else s)

    # apex's ``fused_layer_norm_cuda`` extension (imported by the reference's MixedFusedLayerNorm: Falcon / GPT / BERT):
    # same contract on ATen's native LayerNorm kernels
    fln = types.ModuleType("fused_layer_norm_cuda")

    def forward_affine(input_, normalized_shape, weight, bias, eps):
        out, mean, rstd = torch.native_layer_norm(input_, list(normalized_shape), weight, bias, eps)
        return out, mean, rstd

    def backward_affine(grad_output, mean, invvar, input_, normalized_shape, weight, bias, eps):
        return torch.ops.aten.native_layer_norm_backward(grad_output, input_, list(normalized_shape), mean, invvar,
                                                         weight, bias, [True, True, True])

    fln.forward_affine, fln.backward_affine = forward_affine, backward_affine
    sys.modules["fused_layer_norm_cuda"] = fln

    optimizers.FusedAdam, optimizers.FusedSGD = FusedAdam, FusedSGD
    mta.multi_tensor_applier = multi_tensor_applier
    amp_C.multi_tensor_l2norm, amp_C.multi_tensor_scale = multi_tensor_l2norm, multi_tensor_scale
    apex.optimizers, apex.multi_tensor_apply = optimizers, mta
    sys.modules.update({"apex": apex, "apex.optimizers": optimizers, "apex.multi_tensor_apply": mta, "amp_C": amp_C})


class _SyntheticTokenizer:
    def __init__(self, vocab_size):
        self._v = vocab_size

    vocab_size = property(lambda self: self._v)
    eod = property(lambda self: 2)
    pad = property(lambda self: 0)
    vocab = property(lambda self: {})
    inv_vocab = property(lambda self: {})

    def tokenize(self, text):
        return [int(t) for t in text.split()]

    def detokenize(self, ids):
        return " ".join(map(str, ids))


def _install_torch_shims():
    """torch 2.11 removed private storage helpers the reference's distributed optimizer calls
    (optimizer/distrib_optimizer.py:384 ``storage()._untyped()``); BASELINE.md section 3."""
    import torch
    ts = torch.storage.TypedStorage
    if not hasattr(ts, "_untyped"):
        ts._untyped = lambda self: self.untyped()


class _Tee:
    """Keeps what the reference prints (its training log goes to stdout on the LAST rank) so the loss of the final
    iteration can be reported; nothing is forwarded to the real stdout (the JSON line must be the only output)."""

    def __init__(self):
        self.buf = []

    def write(self, s):
        self.buf.append(s)
        if len(self.buf) > 4096:
            del self.buf[:2048]

    def flush(self):
        pass

    def last_loss(self):
        import re
        for line in reversed("".join(self.buf).splitlines()):
            m = re.search(r"lm loss: ([0-9.eE+-]+)", line)
            if m:
                return float(m.group(1))
        return None


def main(a, MODELS, ClockSampler, bench):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29578")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("LOCAL_WORLD_SIZE", str(world))
    os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "1"
    assert world == a.gpus
    # the reference must import ITS OWN ``megatron`` package, not this repo's compat alias
    for k in [k for k in sys.modules if k == "megatron" or k.startswith("megatron.") or k == "finetune"]:
        del sys.modules[k]
    sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(HERE)]
    _install_apex_stub()
    _install_torch_shims()

    layers, hidden, heads, kv, ffn, vocab = MODELS[a.model]
    if a.layers:
        layers = a.layers
    tp, pp, dp = bench.parallel_layout(a, a.gpus)
    family = a.model.split("-")[0].rstrip("2")          # llama / mistral / falcon
    W, K = a.warmup, a.steps
    argv = ["reference_bench", "--model_name", {"llama": "llama2", "mistral": "mistral", "falcon": "falcon"}[family],
            "--num_layers", str(layers), "--hidden_size", str(hidden),
            "--num_attention_heads", str(heads), "--num_attention_heads_kv", str(kv), "--ffn_hidden_size", str(ffn),
            "--seq_length", str(a.seq), "--max_position_embeddings", str(a.seq), "--micro_batch_size",
            str(a.micro_batch), "--global_batch_size", str(a.global_batch), "--tensor_model_parallel_size",
            str(tp), "--pipeline_model_parallel_size", str(pp), "--train_iters", str(W + K), "--lr", "1e-5",
            "--min_lr", "1e-6", "--lr_decay_style", "cosine", "--weight_decay", "0.1", "--clip_grad", "1.0",
            "--adam_beta1", "0.9", "--adam_beta2", "0.95", "--adam_eps", "1e-5", "--bf16", "--use_flash_attn",
            "--position_embedding_type",
            "rotary", "--hidden_dropout", "0.0", "--attention_dropout", "0.0", "--layernorm_epsilon", "1e-5",
            "--no_bias_gelu_fusion", "--no_bias_dropout_fusion", "--no_gradient_accumulation_fusion",
            "--log_interval", "1", "--eval_iters", "0", "--eval_interval", "1000000", "--num_workers", "0",
            "--seed", "1234", "--tokenizer_type", "SentencePieceTokenizer", "--vocab_file", "synthetic",
            "--data_type", "gpt", "--data_path", "synthetic"]
    if family == "falcon":
        argv += ["--parallel_attn", "--parallel_layernorm"]
    else:
        argv += ["--use_rms_norm", "--glu_activation", "swiglu", "--no_tie_embed_logits"]
    if family == "mistral":
        argv += ["--sliding_window_size", "4096"]
    if tp > 1:
        argv.append("--sequence_parallel")
    if a.recompute:
        argv += ["--recompute_granularity", "full", "--recompute_method", "uniform", "--recompute_num_layers", "1"]
    if a.dist_opt:
        argv.append("--use_distributed_optimizer")
    sys.argv = argv

    import megatron  # the reference package (baseline/_ref/megatron)
    assert os.path.abspath(megatron.__file__).startswith(os.path.abspath(REF)), megatron.__file__
    import megatron.global_vars as gv
    import megatron.initialize as init
    import megatron.tokenizer.tokenizer as tok
    import megatron.training as training

    def build_tokenizer(args):
        t = _SyntheticTokenizer(vocab)
        args.padded_vocab_size = tok._vocab_size_with_padding(t.vocab_size, args)
        return t

    gv.build_tokenizer = build_tokenizer          # no tokenizer files offline
    init.set_jit_fusion_options = lambda args: None  # nvFuser API no longer exists
    training.megatron.initialize.set_jit_fusion_options = init.set_jit_fusion_options

    import finetune as ref_finetune   # baseline/_ref/finetune.py (stock model_provider / forward_step / loss)
    from megatron.model import ModelType

    class _Synthetic(torch.utils.data.Dataset):
        def __init__(self, n, seed):
            self.n, self.seed = max(n, 1), seed

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(self.seed + i)
            return {"text": torch.randint(0, vocab, (a.seq + 1,), generator=g, dtype=torch.int64)}

    def data_provider(nums):
        return _Synthetic(nums[0], 1), _Synthetic(nums[1], 2), _Synthetic(nums[2], 3)

    dev_state = {"calls": 0, "start": None, "end": None}
    orig_train_step = training.train_step
    sampler = ClockSampler(index=int(os.environ.get("LOCAL_RANK", "0")))

    def timed_train_step(*args, **kw):
        import torch.distributed as dist
        if dev_state["calls"] == W:
            dist.barrier()
            torch.cuda.synchronize()
            if rank == 0:
                sampler.start()
            dev_state["start"] = torch.cuda.Event(enable_timing=True)
            dev_state["start"].record()
        out = orig_train_step(*args, **kw)
        dev_state["calls"] += 1
        if dev_state["calls"] == W + K:
            dev_state["end"] = torch.cuda.Event(enable_timing=True)
            dev_state["end"].record()
            torch.cuda.synchronize()
            dist.barrier()
        return out

    training.train_step = timed_train_step

    real_stdout = sys.stdout
    tee = _Tee()
    sys.stdout = tee
    try:
        init.initialize_megatron(ref_finetune.extra_args, {"tokenizer_type": "SentencePieceTokenizer"})
        args = megatron.get_args()
        training.pretrain(args, data_provider, ref_finetune.model_provider, ModelType.encoder_or_decoder,
                          ref_finetune.forward_step, collate_fn=None)
    finally:
        sys.stdout = real_stdout
    import torch.distributed as dist
    ms = torch.tensor([dev_state["start"].elapsed_time(dev_state["end"])], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the reference logs on the last rank: fetch its final "lm loss" (NaN = not found)
    ll = tee.last_loss()
    loss_t = torch.tensor([ll if ll is not None else float("nan")], device="cuda")
    dist.broadcast(loss_t, src=world - 1)
    peak = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], device="cuda")
    dist.all_reduce(peak, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        tokens = K * a.global_batch * a.seq
        val = tokens / (ms.item() / 1e3)
        last_loss = loss_t.item()
        print(json.dumps({
            "metric": bench.METRIC.format(model=a.model, par=bench.parallelism_string(a, a.gpus), seq=a.seq),
            "value": val, "unit": "tokens/s", "n_gpus": a.gpus, "steps": K, "warmup": W,
            "ms_per_step": ms.item() / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic tokens, random-init weights",
            "impl": "reference", "label": "reference (shimmed: apex->torch fused AdamW, no wgrad-accum fusion)",
            "config": bench.bench_config(a, a.gpus),
            "details": {"peak_mem_gb": round(peak.item(), 2)},
            "clocks": clocks,
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": a.global_batch * (a.seq + 1) * 8,
                    "d2h_bytes_per_step": 4, "last_loss": None if last_loss != last_loss else last_loss,
                    "note": "the reference's stock loop copies inputs H2D and reads the loss every step"},
            "gpu_launches": 0}), flush=True)

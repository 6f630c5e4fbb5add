"""Tensor-parallel model parity on two B200s: a tiny Llama trained with TP=2 + sequence parallelism must follow the
same loss / grad-norm trajectory with (a) every sm_100a kernel, the fused GEMM+collective kernels and the micro-batch
CUDA graph as with (b) the plain PyTorch operator path over NCCL.  (The layout invariance of the algorithm itself --
TP=2 vs TP=1 from the same weights -- is checked on CPU in test_training_parallel.py.)"""
import json
import os

import pytest

from tests.dist_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

ARGV = ("--model_name llama2 --num_layers 2 --hidden_size 512 --num_attention_heads 4 --num_attention_heads_kv 2 "
        "--ffn_hidden_size 1408 --seq_length 512 --max_position_embeddings 512 --micro_batch_size 2 "
        "--global_batch_size 4 --train_iters 20 --lr 1e-3 --lr_decay_style constant --bf16 --hidden_dropout 0 "
        "--attention_dropout 0 --tokenizer_type NullTokenizer --vocab_file 1024 --data_type synthetic "
        "--log_interval 100 --eval_iters 0 --eval_interval 1000 --num_workers 0 --use_flash_attn --use_rms_norm "
        "--glu_activation swiglu --no_tie_embed_logits --position_embedding_type rotary "
        "--tensor_model_parallel_size 2 --sequence_parallel --seed 11")
STEPS = 6


def _worker(rank, world, mode, out_path):
    import torch
    import torch.distributed as dist
    fast = mode == "fast"
    os.environ["MLB200_DISABLE_KERNELS"] = "0" if fast else "1"
    os.environ["MLB200_FUSED_TP"] = "1" if fast else "0"
    import finetune
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models import ModelType
    from megatron_llm_b200.parallel import fused_tp
    from megatron_llm_b200.training import setup_model_and_optimizer, train_step
    argv = ARGV.split() + (["--cuda_graph_microbatch"] if fast else [])
    initialize_megatron(finetune.extra_args, {}, args_list=argv)
    model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
    comm = fused_tp.communicator()
    assert (comm is not None) == fast, "fused TP communicator binding does not follow MLB200_FUSED_TP"

    def it():
        g = torch.Generator().manual_seed(0)
        while True:
            yield {"text": torch.randint(0, 1000, (2, 512 + 1), generator=g)}
    data = it()
    out = []
    for _ in range(STEPS):
        loss, skipped, gnorm, _ = train_step(finetune.forward_step, data, model, opt, sched)
        out.append((loss["lm loss"].item(), gnorm.item()))
    if comm is not None:
        assert comm.error_flag() == 0, "a fused-kernel spin-wait timed out"
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(out, f)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)      # (tearing NCCL down after graph capture can hang; the result is on disk)


def test_tp2_kernels_fused_graph_match_torch_nccl_path(tmp_path):
    res = {}
    for mode in ("fast", "plain"):
        path = tmp_path / f"{mode}.json"
        try:
            run_distributed(_worker, 2, mode, str(path), backend="nccl")
        except RuntimeError:
            if not path.exists():
                raise
        res[mode] = json.load(open(path))
    fast, plain = res["fast"], res["plain"]
    for (la, ga), (lb, gb) in zip(fast, plain):
        assert la == la and ga == ga, "nan"
        assert abs(la - lb) < 3e-2 * max(1.0, abs(lb)), (fast, plain)
        assert abs(ga - gb) < 0.15 * max(1e-3, abs(gb)), (fast, plain)
    assert plain[0][0] > 6.0, "first loss should be ~ln(vocab)"


# ------------------------------------------------------------------------------------------------------------------
# Layout invariance on hardware: TP = 1 / 2 / 4 / 8 from the SAME weights (--use_cpu_initialization builds the full
# fp32 master weight on every rank and keeps the rank's shard, so no resharding is involved), the same fixed batches,
# every kernel + fused collectives + the micro-batch graph on.  bf16 tolerance.
INV_ARGV = ("--model_name llama2 --num_layers 2 --hidden_size 1024 --num_attention_heads 8 --num_attention_heads_kv 8 "
            "--ffn_hidden_size 2816 --seq_length 1024 --max_position_embeddings 1024 --micro_batch_size 2 "
            "--global_batch_size 4 --train_iters 20 --lr 3e-4 --lr_decay_style constant --bf16 --hidden_dropout 0 "
            "--attention_dropout 0 --tokenizer_type NullTokenizer --vocab_file 1024 --data_type synthetic "
            "--log_interval 100 --eval_iters 0 --eval_interval 1000 --num_workers 0 --use_flash_attn --use_rms_norm "
            "--glu_activation swiglu --no_tie_embed_logits --position_embedding_type rotary --use_cpu_initialization "
            "--clip_grad 1.0 --seed 11")


def _invariance_worker(rank, world, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MLB200_DISABLE_KERNELS"] = "0"
    os.environ["MLB200_FUSED_TP"] = "1"
    import finetune
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models import ModelType
    from megatron_llm_b200.parallel import fused_tp
    from megatron_llm_b200.training import setup_model_and_optimizer, train_step
    argv = INV_ARGV.split() + ["--tensor_model_parallel_size", str(world)]
    if world > 1:
        argv += ["--sequence_parallel", "--cuda_graph_microbatch"]
    initialize_megatron(finetune.extra_args, {}, args_list=argv)
    model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
    comm = fused_tp.communicator()
    assert (comm is not None) == (world > 1)

    def it():
        g = torch.Generator().manual_seed(0)
        batches = [torch.randint(0, 1000, (2, 1024 + 1), generator=g) for _ in range(4)]
        while True:
            for b in batches:
                yield {"text": b}
    data = it()
    out = []
    for _ in range(STEPS):
        loss, skipped, gnorm, _ = train_step(finetune.forward_step, data, model, opt, sched)
        out.append((loss["lm loss"].item(), gnorm.item()))
    if comm is not None:
        assert comm.error_flag() == 0, "a fused-kernel spin-wait timed out"
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(out, f)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


def test_loss_trajectory_is_invariant_to_the_tp_size(tmp_path):
    import torch
    sizes = [n for n in (1, 2, 4, 8) if n <= torch.cuda.device_count()]
    res = {}
    for n in sizes:
        path = tmp_path / f"tp{n}.json"
        try:
            run_distributed(_invariance_worker, n, str(path), backend="nccl")
        except RuntimeError:
            if not path.exists():
                raise
        res[n] = json.load(open(path))
    keep = os.environ.get("MLB200_TEST_RECORD")          # e.g. gpurun_out/tp_invariance.json
    if keep:
        with open(keep, "w") as f:
            json.dump({f"tp{n}": v for n, v in res.items()}, f)
    ref = res[1]
    assert ref[0][0] > 6.0 and ref[-1][0] < ref[0][0]
    for n in sizes[1:]:
        for (la, ga), (lb, gb) in zip(res[n], ref):
            assert la == la and ga == ga, "nan"
            assert abs(la - lb) < 1.5e-2 * max(1.0, abs(lb)), (n, res[n], ref)
            assert abs(ga - gb) < 0.10 * max(1e-3, abs(gb)), (n, res[n], ref)

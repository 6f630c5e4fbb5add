"""Command-line / config system.

Parity target: megatron/arguments.py (16 argument groups ~230 flags :372-1103, ``validate_args`` :53-350).
The flag names, ``dest`` names and defaults match the reference so launch scripts carry over unchanged.
Flags are declared in compact tables (name, kwargs) per group instead of one ``add_argument`` call each.
"""
from __future__ import annotations

import argparse
import os

import torch

from .models.enums import PositionEmbeddingType
from .metrics import METRICS


def _S(**kw):  # store_true
    return dict(action="store_true", **kw)


def _SF(dest, **kw):  # store_false with explicit dest
    return dict(action="store_false", dest=dest, **kw)


_GROUPS = {
    "network size": [
        ("--num_layers", dict(type=int, default=None)),
        ("--encoder_num_layers", dict(type=int, default=None)),
        ("--decoder_num_layers", dict(type=int, default=None)),
        ("--hidden_size", dict(type=int, default=None)),
        ("--ffn_hidden_size", dict(type=int, default=None)),
        ("--num_attention_heads", dict(type=int, default=None)),
        ("--num_attention_heads_kv", dict(type=int, default=None)),
        ("--kv_channels", dict(type=int, default=None)),
        ("--max_position_embeddings", dict(type=int, default=None)),
        ("--make_vocab_size_divisible_by", dict(type=int, default=128)),
        ("--layernorm_epsilon", dict(type=float, default=1e-5)),
        ("--apply_residual_connection_post_layernorm", _S()),
        ("--use_bias", _S()),
        ("--use_rms_norm", _S()),
        ("--use_post_ln", _S()),
        ("--onnx_safe", dict(type=bool, required=False)),
        ("--glu_activation", dict(type=str, choices=["liglu", "geglu", "reglu", "swiglu"], default=None)),
        ("--position_embedding_type", dict(type=lambda x: PositionEmbeddingType[x],
                                           choices=list(PositionEmbeddingType),
                                           default=PositionEmbeddingType.absolute)),
        ("--rope_scaling_factor", dict(type=float, default=1.0)),
        ("--rope_theta", dict(type=float, default=10000.0)),
        ("--parallel_attn", _S()),
        ("--parallel_layernorm", _S()),
        ("--no_tie_embed_logits", _SF("tie_embed_logits")),
        ("--sliding_window_size", dict(type=int, default=None)),
    ],
    "logging": [
        ("--log_params_norm", _S()),
        ("--log_num_zeros_in_grad", _S()),
        ("--timing_log_level", dict(type=int, default=0, choices=range(0, 3))),
        ("--barrier_with_L1_time", _SF("barrier_with_L1_time")),
        ("--timing_log_option", dict(type=str, default="minmax", choices=["max", "minmax", "all"])),
        ("--tensorboard_log_interval", dict(type=int, default=1)),
        ("--tensorboard_queue_size", dict(type=int, default=1000)),
        ("--log_timers_to_tensorboard", _S()),
        ("--log_batch_size_to_tensorboard", _S()),
        ("--log_validation_ppl_to_tensorboard", _S()),
        ("--log_memory_to_tensorboard", _S()),
        ("--log_world_size_to_tensorboard", _S()),
        ("--wandb_logger", _S()),
        ("--wandb_project", dict(type=str, default=None)),
        ("--wandb_entity", dict(type=str, default="meditron")),
        ("--wandb_name", dict(type=str, default=None)),
        ("--wandb_id", dict(type=str, default=None)),
        ("--wandb_resume", dict(type=str, default="allow")),
        ("--wandb_api_key", dict(type=str, default=None)),
        ("--metrics", dict(default=[], nargs="+", choices=list(METRICS) + ["all"])),
    ],
    "regularization": [
        ("--attention_dropout", dict(type=float, default=0.1)),
        ("--hidden_dropout", dict(type=float, default=0.1)),
        ("--lima_dropout", _S()),
        ("--weight_decay", dict(type=float, default=0.01)),
        ("--start_weight_decay", dict(type=float)),
        ("--end_weight_decay", dict(type=float)),
        ("--weight_decay_incr_style", dict(type=str, default="constant", choices=["constant", "linear", "cosine"])),
        ("--clip_grad", dict(type=float, default=1.0)),
        ("--adam_beta1", dict(type=float, default=0.9)),
        ("--adam_beta2", dict(type=float, default=0.999)),
        ("--adam_eps", dict(type=float, default=1e-08)),
        ("--sgd_momentum", dict(type=float, default=0.9)),
    ],
    "training": [
        ("--micro_batch_size", dict(type=int, default=None)),
        ("--global_batch_size", dict(type=int, default=None)),
        ("--rampup_batch_size", dict(nargs="*", default=None)),
        ("--recompute_activations", _S()),
        ("--recompute_granularity", dict(type=str, default=None, choices=["full", "selective"])),
        ("--distribute_saved_activations", _S()),
        ("--recompute_method", dict(type=str, default=None, choices=["uniform", "block"])),
        ("--recompute_num_layers", dict(type=int, default=1)),
        ("--train_iters", dict(type=int, default=None)),
        ("--skip_iters", dict(type=int, nargs="*", default=[])),
        ("--train_samples", dict(type=int, default=None)),
        ("--log_interval", dict(type=int, default=100)),
        ("--exit_interval", dict(type=int, default=None)),
        ("--exit_duration_in_mins", dict(type=float, default=None)),
        ("--exit_signal_handler", _S()),
        ("--tensorboard_dir", dict(type=str, default=None)),
        ("--no_masked_softmax_fusion", _SF("masked_softmax_fusion")),
        ("--no_bias_gelu_fusion", _SF("bias_gelu_fusion")),
        ("--no_bias_dropout_fusion", _SF("bias_dropout_fusion")),
        ("--use_flash_attn", _S()),
        ("--optimizer", dict(type=str, default="adam", choices=["adam", "sgd"])),
        ("--dataloader_type", dict(type=str, default=None, choices=["single", "cyclic"])),
        ("--no_async_tensor_model_parallel_allreduce", _SF("async_tensor_model_parallel_allreduce")),
        ("--no_persist_layer_norm", _S()),
        ("--sequence_parallel", _S()),
        ("--no_gradient_accumulation_fusion", _SF("gradient_accumulation_fusion")),
        # B200-native switches (not in the reference)
        # fused GEMM+collective kernels over NVLink peer memory (all-gather->GEMM, GEMM->reduce-scatter).  Unset =
        # automatic: on when one tensor-parallel group spans the whole job (the configuration measured faster than
        # NCCL at TP=2 and TP=8, profiles/README.md), off (NCCL) when there are several TP groups.
        ("--fused_tp_comm", dict(action="store_const", const=True, default=None)),
        ("--no_fused_tp_comm", dict(action="store_const", const=False, dest="fused_tp_comm", default=None)),
        ("--no_fused_dp_comm", _SF("fused_dp_comm")),
        # step-range profiling (megatron_llm_b200/profiler.py): torch.profiler trace + cudaProfilerStart/Stop + NVTX
        ("--profile", _S()),
        ("--profile_step_start", dict(type=int, default=10)),
        ("--profile_step_end", dict(type=int, default=12)),
        ("--profile_ranks", dict(type=int, nargs="+", default=[0])),
        ("--profile_dir", dict(type=str, default=None)),
        ("--ddp_bucket_size_mb", dict(type=int, default=256)),
    ],
    "initialization": [
        ("--seed", dict(type=int, default=1234)),
        ("--data_parallel_random_init", _S()),
        ("--init_method_std", dict(type=float, default=0.02)),
        ("--init_method_xavier_uniform", _S()),
    ],
    "learning rate": [
        ("--lr", dict(type=float, default=None)),
        ("--lr_decay_style", dict(type=str, default="linear",
                                  choices=["constant", "linear", "cosine", "inverse-square-root"])),
        ("--lr_decay_iters", dict(type=int, default=None)),
        ("--lr_decay_samples", dict(type=int, default=None)),
        ("--lr_warmup_fraction", dict(type=float, default=None)),
        ("--lr_warmup_iters", dict(type=int, default=0)),
        ("--lr_warmup_samples", dict(type=int, default=0)),
        ("--min_lr", dict(type=float, default=0.0)),
        ("--override_opt_param_scheduler", _S()),
        ("--use_checkpoint_opt_param_scheduler", _S()),
    ],
    "checkpointing": [
        ("--save", dict(type=str, default=None)),
        ("--save_interval", dict(type=int, default=None)),
        ("--no_save_optim", dict(action="store_true", default=None)),
        ("--no_save_rng", dict(action="store_true", default=None)),
        ("--load", dict(type=str, default=None)),
        ("--load_iters", dict(type=int, default=None)),
        ("--no_load_optim", dict(action="store_true", default=None)),
        ("--no_load_rng", dict(action="store_true", default=None)),
        ("--finetune", _S()),
        ("--no_initialization", _SF("perform_initialization")),
        ("--use_checkpoint_args", _S()),
    ],
    "mixed precision": [
        ("--fp16", _S()),
        ("--bf16", _S()),
        ("--loss_scale", dict(type=float, default=None)),
        ("--initial_loss_scale", dict(type=float, default=2 ** 32)),
        ("--min_loss_scale", dict(type=float, default=1.0)),
        ("--loss_scale_window", dict(type=float, default=1000)),
        ("--hysteresis", dict(type=int, default=2)),
        ("--fp32_residual_connection", _S()),
        ("--no_query_key_layer_scaling", _SF("apply_query_key_layer_scaling")),
        ("--attention_softmax_in_fp32", _S()),
        ("--accumulate_allreduce_grads_in_fp32", _S()),
        ("--fp16_lm_cross_entropy", _S()),
    ],
    "distributed": [
        ("--tensor_model_parallel_size", dict(type=int, default=1)),
        ("--pipeline_model_parallel_size", dict(type=int, default=1)),
        ("--pipeline_model_parallel_split_rank", dict(type=int, default=None)),
        ("--num_layers_per_virtual_pipeline_stage", dict(type=int, default=None)),
        ("--distributed_backend", dict(default=None, choices=["nccl", "gloo"])),
        ("--DDP_impl", dict(default="local", choices=["local", "torch"])),
        ("--no_contiguous_buffers_in_local_ddp", _SF("use_contiguous_buffers_in_local_ddp")),
        ("--no_scatter_gather_tensors_in_pipeline", _SF("scatter_gather_tensors_in_pipeline")),
        ("--use_ring_exchange_p2p", _S()),
        ("--local_rank", dict(type=int, default=None)),
        ("--use_cpu_initialization", dict(action="store_true", default=None)),
        ("--empty_unused_memory_level", dict(default=0, type=int, choices=[0, 1, 2])),
        ("--standalone_embedding_stage", _S()),
        ("--use_distributed_optimizer", _S()),
    ],
    "validation": [
        ("--eval_only", _S()),
        ("--eval_iters", dict(type=int, default=100)),
        ("--eval_interval", dict(type=int, default=1000)),
    ],
    "data and dataloader": [
        ("--data_path", dict(nargs="*", default=None)),
        ("--split", dict(type=str, default="969, 30, 1")),
        ("--train_data_path", dict(nargs="*", default=None)),
        ("--valid_data_path", dict(nargs="*", default=None)),
        ("--test_data_path", dict(nargs="*", default=None)),
        ("--vocab_file", dict(type=str, default=None)),
        ("--merge_file", dict(type=str, default=None)),
        ("--vocab_extra_ids", dict(type=int, default=0)),
        ("--vocab_extra_ids_list", dict(type=str, default=None)),
        ("--no_new_tokens", _SF("new_tokens")),
        ("--seq_length", dict(type=int, default=None)),
        ("--variable_seq_lengths", dict(action="store_true", default=None)),
        ("--scalar_loss_mask", dict(type=float, default=0.0)),
        ("--encoder_seq_length", dict(type=int, default=None)),
        ("--decoder_seq_length", dict(type=int, default=None)),
        ("--retriever_seq_length", dict(type=int, default=256)),
        ("--sample_rate", dict(type=float, default=1.0)),
        ("--mask_prob", dict(type=float, default=0.15)),
        ("--short_seq_prob", dict(type=float, default=0.1)),
        ("--mmap_warmup", _S()),
        ("--num_workers", dict(type=int, default=2)),
        ("--tokenizer_type", dict(type=str, default=None,
                                  choices=["BertWordPieceLowerCase", "BertWordPieceCase", "GPT2BPETokenizer",
                                           "SentencePieceTokenizer", "FalconTokenizer", "NullTokenizer"])),
        ("--tokenizer_model", dict(type=str, default=None)),
        ("--data_impl", dict(type=str, default="infer", choices=["lazy", "cached", "mmap", "infer"])),
        ("--reset_position_ids", _S()),
        ("--reset_attention_mask", _S()),
        ("--eod_mask_loss", _S()),
    ],
    "autoresume": [
        ("--adlr_autoresume", _S()),
        ("--adlr_autoresume_interval", dict(type=int, default=1000)),
    ],
    "biencoder": [
        ("--ict_head_size", dict(type=int, default=None)),
        ("--biencoder_projection_dim", dict(type=int, default=0)),
        ("--biencoder_shared_query_context_model", _S()),
        ("--ict_load", dict(type=str, default=None)),
        ("--cuda_graph_microbatch", dict(action="store_true")),
        ("--bert_load", dict(type=str, default=None)),
        # (missing from the reference's parser although pretrain_bert.py reads it; restored from upstream Megatron-LM)
        ("--bert_no_binary_head", dict(action="store_false", dest="bert_binary_head")),
        ("--titles_data_path", dict(type=str, default=None)),
        ("--query_in_block_prob", dict(type=float, default=0.1)),
        ("--use_one_sent_docs", _S()),
        ("--evidence_data_path", dict(type=str, default=None)),
        ("--retriever_report_topk_accuracies", dict(nargs="+", type=int, default=[])),
        ("--retriever_score_scaling", _S()),
        ("--block_data_path", dict(type=str, default=None)),
        ("--embedding_path", dict(type=str, default=None)),
        ("--indexer_batch_size", dict(type=int, default=128)),
        ("--indexer_log_interval", dict(type=int, default=1000)),
    ],
    "vision": [
        ("--num_classes", dict(type=int, default=1000)),
        ("--img_h", dict(type=int, default=224)),
        ("--img_w", dict(type=int, default=224)),
        ("--num_channels", dict(type=int, default=3)),
        ("--patch_dim", dict(type=int, default=16)),
        ("--classes_fraction", dict(type=float, default=1.0)),
        ("--data_per_class_fraction", dict(type=float, default=1.0)),
        ("--no_data_sharding", _SF("data_sharding")),
        ("--head_lr_mult", dict(type=float, default=1.0)),
        ("--iter_per_epoch", dict(type=int, default=1250)),
        # DINO options: parsed for command-line compatibility (the reference ships the flags, arguments.py:1082-1102,
        # but no DINO model reads them)
        ("--dino_local_img_size", dict(type=int, default=96)),
        ("--dino_local_crops_number", dict(type=int, default=10)),
        ("--dino_head_hidden_size", dict(type=int, default=2048)),
        ("--dino_bottleneck_size", dict(type=int, default=256)),
        ("--dino_freeze_last_layer", dict(type=float, default=1)),
        ("--dino_norm_last_layer", _S()),
        ("--dino_warmup_teacher_temp", dict(type=float, default=0.04)),
        ("--dino_teacher_temp", dict(type=float, default=0.07)),
        ("--dino_warmup_teacher_temp_epochs", dict(type=int, default=30)),
    ],
    "inference": [
        ("--inference_batch_times_seqlen_threshold", dict(type=int, default=512)),
        ("--max_tokens_to_oom", dict(type=int, default=12000)),
    ],
    "transformer-engine": [
        ("--fp8_e4m3", dict(action="store_true", dest="fp8_e4m3")),
        ("--fp8_hybrid", _S()),
        ("--no_fp8_wgrad", _SF("fp8_wgrad")),
        ("--fp8_margin", dict(type=int, default=0)),
        ("--fp8_interval", dict(type=int, default=1)),
        ("--transformer_impl", dict(default="local", choices=["local", "transformer_engine"])),
        ("--fp8_amax_history_len", dict(type=int, default=1)),
        ("--fp8_amax_compute_algo", dict(default="most_recent", choices=["most_recent", "max"])),
    ],
}


def build_base_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Megatron-LLM-B200 Arguments", allow_abbrev=False)
    for title, flags in _GROUPS.items():
        group = parser.add_argument_group(title=title)
        for name, kw in flags:
            group.add_argument(name, **kw)
    return parser


def parse_args(extra_args_provider=None, ignore_unknown_args=False, args_list=None):
    parser = build_base_parser()
    if extra_args_provider is not None:
        parser = extra_args_provider(parser)
    if ignore_unknown_args:
        args, _ = parser.parse_known_args(args_list)
    else:
        args = parser.parse_args(args_list)
    args.rank = int(os.getenv("RANK", "0"))
    args.world_size = int(os.getenv("WORLD_SIZE", "1"))
    return args


def _print_args(args):
    if args.rank == 0:
        print("------------------------ arguments ------------------------", flush=True)
        for arg in sorted(vars(args)):
            dots = "." * (48 - len(arg))
            print("  {} {} {}".format(arg, dots, getattr(args, arg)), flush=True)
        print("-------------------- end of arguments ---------------------", flush=True)


def _check_arg_is_not_none(args, arg):
    assert getattr(args, arg) is not None, "{} argument is None".format(arg)


def validate_args(args, defaults={}):
    """Derive and cross-check arguments (reference arguments.py:53-350)."""
    from .utils.device import use_cuda

    if args.distributed_backend is None:
        args.distributed_backend = "nccl" if use_cuda() else "gloo"
    # tensor / pipeline / data parallel sizes
    args.tensor_model_parallel_size = min(args.tensor_model_parallel_size, args.world_size)
    assert args.world_size % args.tensor_model_parallel_size == 0, \
        "world size ({}) is not divisible by tensor model parallel size ({})".format(
            args.world_size, args.tensor_model_parallel_size)
    args.pipeline_model_parallel_size = min(args.pipeline_model_parallel_size,
                                            args.world_size // args.tensor_model_parallel_size)
    args.transformer_pipeline_model_parallel_size = (
        args.pipeline_model_parallel_size - 1 if args.standalone_embedding_stage
        else args.pipeline_model_parallel_size)
    model_parallel_size = args.pipeline_model_parallel_size * args.tensor_model_parallel_size
    assert args.world_size % model_parallel_size == 0, \
        "world size is not divisible by tensor parallel size ({}) times pipeline parallel size ({})".format(
            args.tensor_model_parallel_size, args.pipeline_model_parallel_size)
    args.data_parallel_size = args.world_size // model_parallel_size
    if args.rank == 0:
        print("using world size: {}, data-parallel-size: {}, tensor-model-parallel size: {}, "
              "pipeline-model-parallel size: {} ".format(args.world_size, args.data_parallel_size,
                                                         args.tensor_model_parallel_size,
                                                         args.pipeline_model_parallel_size), flush=True)
    if args.pipeline_model_parallel_size > 1 and args.pipeline_model_parallel_split_rank is not None:
        assert args.pipeline_model_parallel_split_rank < args.pipeline_model_parallel_size, \
            "split rank needs to be less than pipeline model parallel size ({})".format(
                args.pipeline_model_parallel_size)

    # defaults only fill Nones
    for key, value in defaults.items():
        if getattr(args, key, None) is not None:
            if args.rank == 0:
                print("WARNING: overriding default arguments for {key}:{v} with {key}:{v2}".format(
                    key=key, v=value, v2=getattr(args, key)), flush=True)
        else:
            setattr(args, key, value)

    assert args.micro_batch_size is not None and args.micro_batch_size > 0
    if args.global_batch_size is None:
        args.global_batch_size = args.micro_batch_size * args.data_parallel_size
        if args.rank == 0:
            print("setting global batch size to {}".format(args.global_batch_size), flush=True)
    assert args.global_batch_size > 0
    if args.num_layers_per_virtual_pipeline_stage is not None:
        assert args.pipeline_model_parallel_size > 2, \
            "pipeline-model-parallel size should be greater than 2 with interleaved schedule"
        assert args.num_layers % args.num_layers_per_virtual_pipeline_stage == 0, \
            "number of layers is not divisible by number of layers per virtual pipeline stage"
        args.virtual_pipeline_model_parallel_size = \
            (args.num_layers // args.transformer_pipeline_model_parallel_size) // \
            args.num_layers_per_virtual_pipeline_stage
    else:
        args.virtual_pipeline_model_parallel_size = None

    # dtypes
    args.params_dtype = torch.float
    if args.fp16:
        assert not args.bf16
        args.params_dtype = torch.half
    if args.bf16:
        assert not args.fp16
        args.params_dtype = torch.bfloat16
        if not args.accumulate_allreduce_grads_in_fp32:
            args.accumulate_allreduce_grads_in_fp32 = True
            if args.rank == 0:
                print("accumulate and all-reduce gradients in fp32 for bfloat16 data type.", flush=True)
    if args.rank == 0:
        print("using {} for parameters ...".format(args.params_dtype), flush=True)
    if args.accumulate_allreduce_grads_in_fp32:
        assert args.DDP_impl == "local"
        assert args.use_contiguous_buffers_in_local_ddp
    if args.use_distributed_optimizer:
        assert args.DDP_impl == "local"
        assert args.use_contiguous_buffers_in_local_ddp
    if args.DDP_impl == "torch":
        args.use_contiguous_buffers_in_local_ddp = False
    if args.dataloader_type is None:
        args.dataloader_type = "single"

    args.consumed_train_samples = 0
    args.consumed_valid_samples = 0

    # recompute
    if args.recompute_activations:
        args.recompute_granularity = "selective"
    del args.recompute_activations

    # iteration vs sample based training
    if args.train_iters:
        assert args.train_samples is None, "expected iteration-based training"
        assert args.lr_decay_samples is None, "expected iteration-based learning rate decay"
        assert args.lr_warmup_samples == 0, "expected iteration-based learning rate warmup"
        assert args.rampup_batch_size is None, "expected no batch-size rampup for iteration-based training"
        if args.lr_warmup_fraction is not None:
            assert args.lr_warmup_iters == 0, "can only specify one of lr_warmup_fraction and lr_warmup_iters"
    if args.train_samples:
        assert args.train_iters is None, "expected sample-based training"
        assert args.lr_decay_iters is None, "expected sample-based learning rate decay"
        assert args.lr_warmup_iters == 0, "expected sample-based learnig rate warmup"
        if args.lr_warmup_fraction is not None:
            assert args.lr_warmup_samples == 0, "can only specify one of lr_warmup_fraction and lr_warmup_samples"

    if args.num_layers is not None:
        assert args.encoder_num_layers is None, "cannot have both num_layers and encoder_num_layers specified"
        args.encoder_num_layers = args.num_layers
    else:
        assert args.encoder_num_layers is not None, "either num_layers or encoder_num_layers should be specified"
        args.num_layers = args.encoder_num_layers

    for req in ["num_layers", "hidden_size", "num_attention_heads", "max_position_embeddings"]:
        _check_arg_is_not_none(args, req)

    if args.ffn_hidden_size is None:
        args.ffn_hidden_size = 4 * args.hidden_size
    if args.kv_channels is None:
        assert args.hidden_size % args.num_attention_heads == 0
        args.kv_channels = args.hidden_size // args.num_attention_heads
    if args.num_attention_heads_kv is None:
        args.num_attention_heads_kv = args.num_attention_heads

    if args.seq_length is not None:
        assert args.encoder_seq_length is None
        args.encoder_seq_length = args.seq_length
    else:
        assert args.encoder_seq_length is not None
        args.seq_length = args.encoder_seq_length

    if args.position_embedding_type == PositionEmbeddingType.absolute:
        assert args.max_position_embeddings >= args.seq_length
        if args.decoder_seq_length is not None:
            assert args.max_position_embeddings >= args.decoder_seq_length
    else:
        assert args.rope_scaling_factor >= 1, "rope_scaling_factor must be >= 1"
        assert args.max_position_embeddings >= args.seq_length, \
            "max_position_embeddings must be >= seq_length for rotary embeddings"
    if args.lr is not None:
        assert args.min_lr <= args.lr
    if args.save is not None:
        assert args.save_interval is not None
    if args.fp16_lm_cross_entropy:
        assert args.fp16, "lm cross entropy in fp16 only support in fp16 mode."
    if args.fp32_residual_connection:
        assert args.fp16 or args.bf16, "residual connection in fp32 only supported when using fp16 or bf16."
    if args.weight_decay_incr_style == "constant":
        assert args.start_weight_decay is None
        assert args.end_weight_decay is None
        args.start_weight_decay = args.weight_decay
        args.end_weight_decay = args.weight_decay
    else:
        assert args.start_weight_decay is not None
        assert args.end_weight_decay is not None

    # persistent layer norm / recompute checks
    if args.distribute_saved_activations:
        assert args.tensor_model_parallel_size > 1, \
            "can distribute recomputed activations only across tensor model parallel groups"
        assert args.recompute_granularity == "full", \
            "distributed recompute activations is only application to full recompute granularity"
        assert args.recompute_method is not None, \
            "for distributed recompute activations to work you need to use a recompute method"
    if args.recompute_granularity == "selective":
        assert args.recompute_method is None, \
            "recompute method is not yet supported for selective recomputing granularity"

    # sequence parallelism is meaningless with TP=1
    if args.tensor_model_parallel_size == 1:
        args.sequence_parallel = False
    if args.sequence_parallel:
        args.async_tensor_model_parallel_allreduce = False
    # (the reference requires CUDA_DEVICE_MAX_CONNECTIONS=1 for overlap, arguments.py:340-348; this framework
    #  schedules overlap with explicit streams/events and fused kernels, so no such requirement)

    if args.variable_seq_lengths is None:
        args.variable_seq_lengths = False
    if getattr(args, "data_type", None) == "instruction" and args.variable_seq_lengths:
        pass
    if args.use_flash_attn:
        assert not args.reset_attention_mask or True  # flash path ignores attention_mask (reference quirk)
    if args.glu_activation is not None and args.bias_gelu_fusion:
        args.bias_gelu_fusion = False
    if not args.use_bias:
        args.bias_gelu_fusion = False
        args.bias_dropout_fusion = False

    _print_args(args)
    return args


# legacy helper names kept for API parity --------------------------------------------------------
def _add_network_size_args(parser):
    return parser

"""Training runtime: ``pretrain`` / ``train_step`` / logging / evaluation / data iterators.

Parity target: megatron/training.py (pretrain :55-169, get_model :199-304, scheduler :307-350,
_setup_model_and_optimizer :353-390, train_step :393-459, training_log :462-641, _train :654-770,
evaluate :773-868, build_train_valid_test_data_iterators :877-966).

Host-sync hygiene (the reference ``.item()``s several values per micro-batch / step): losses and the grad norm
stay on the device and are only read at log intervals.
"""
from __future__ import annotations

import math
import sys
import time
from datetime import datetime
from typing import Callable

import torch
import torch.distributed as dist

from . import initialize as _initialize
from .checkpointing import load_checkpoint, save_checkpoint
from .data.data_samplers import build_pretraining_data_loader
from .global_vars import (get_args, get_counters, get_current_global_batch_size, get_num_microbatches,
                          get_signal_handler, get_tensorboard_writer, get_timers, update_num_microbatches)
from .models import Float16Module
from .models.enums import ModelType
from .optimizer import get_megatron_optimizer
from .optimizer_param_scheduler import OptimizerParamScheduler
from .parallel import state as ps
from .parallel.ddp import DistributedDataParallel as LocalDDP
from .parallel.layers import set_defaults_if_not_set_tensor_model_parallel_attributes
from .parallel.schedules import get_forward_backward_func
from .utils import (calc_params_l2_norm, check_adlr_autoresume_termination, is_last_rank, print_rank_0,
                    print_rank_last, report_memory, unwrap_model)
from .utils.device import current_device, use_cuda

_TRAIN_START_TIME = time.time()


def print_datetime(string):
    if dist.is_initialized():
        dist.barrier()
    print_rank_0("[" + string + "] datetime: {} ".format(datetime.now().strftime("%Y-%m-%d %H:%M:%S")))


def pretrain(args, train_valid_test_dataset_provider, model_provider_func, model_type: ModelType,
             forward_step_func, process_non_loss_data_func=None, collate_fn=None):
    """Main training program: build model/optimizer/scheduler, build data iterators, train, evaluate, save."""
    _initialize.set_jit_fusion_options(args)
    global _TRAIN_START_TIME
    start = torch.tensor([_TRAIN_START_TIME], dtype=torch.float64, device=current_device())
    if dist.is_initialized():
        dist.all_reduce(start, op=dist.ReduceOp.MIN)
    _TRAIN_START_TIME = start.item()
    print_rank_0("time to initialize megatron (seconds): {:.3f}".format(time.time() - _TRAIN_START_TIME))
    print_datetime("after megatron is initialized")
    timers = get_timers()

    timers("model-and-optimizer-setup", log_level=0).start(barrier=True)
    model, optimizer, opt_param_scheduler = _setup_model_and_optimizer(model_provider_func, model_type, args=args)
    timers("model-and-optimizer-setup").stop()
    print_datetime("after model, optimizer, and learning rate scheduler are built")

    timers("train/valid/test-data-iterators-setup", log_level=0).start(barrier=True)
    if args.virtual_pipeline_model_parallel_size is not None:
        its = [build_train_valid_test_data_iterators(train_valid_test_dataset_provider, args, collate_fn=collate_fn)
               for _ in range(len(model))]
        train_data_iterator = [i[0] for i in its]
        valid_data_iterator = [i[1] for i in its]
        test_data_iterator = [i[2] for i in its]
    else:
        train_data_iterator, valid_data_iterator, test_data_iterator = build_train_valid_test_data_iterators(
            train_valid_test_dataset_provider, args, collate_fn=collate_fn)
    timers("train/valid/test-data-iterators-setup").stop()
    print_datetime("after dataloaders are built")
    print_rank_0("done with setup ...")
    timers.log(["model-and-optimizer-setup", "train/valid/test-data-iterators-setup"], barrier=True)
    print_rank_0("training ...")

    iteration = 0
    if args.do_train and args.train_iters > 0:
        iteration = _train(args, forward_step_func, model, optimizer, opt_param_scheduler, train_data_iterator,
                           valid_data_iterator, process_non_loss_data_func)
    print_datetime("after training is done")
    if args.do_valid:
        evaluate_and_print_results("the end of training for val data", forward_step_func, valid_data_iterator, model,
                                   iteration, process_non_loss_data_func, verbose=False, args=args)
    if args.save and iteration != 0:
        save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
    if args.do_test:
        evaluate_and_print_results("the end of training for test data", forward_step_func, test_data_iterator, model,
                                   0, process_non_loss_data_func, verbose=True, args=args)
    return iteration


def update_train_iters(args):
    """Sample-based training: derive train_iters (accounting for batch-size ramp-up)."""
    if args.train_iters:
        return
    if args.rampup_batch_size is None:
        args.train_iters = args.train_samples // args.global_batch_size
    else:
        iterations, consumed = 0, 0
        while consumed <= int(args.rampup_batch_size[2]):
            update_num_microbatches(consumed, consistency_check=False)
            consumed += get_current_global_batch_size()
            iterations += 1
        update_num_microbatches(0, consistency_check=False)
        iterations += (args.train_samples - consumed) // args.global_batch_size
        args.train_iters = iterations
    print_rank_0("setting training iterations to {}".format(args.train_iters))


def get_model(model_provider_func: Callable, model_type=ModelType.encoder_or_decoder, wrap_with_ddp: bool = True,
              args=None):
    """Build the (virtual-)stage model chunk(s), move to device, wrap in Float16Module and DDP."""
    if args is None:
        args = get_args()
    args.model_type = model_type
    if ps.get_pipeline_model_parallel_world_size() > 1 and args.virtual_pipeline_model_parallel_size is not None:
        assert model_type != ModelType.encoder_and_decoder, \
            "Interleaved schedule not supported for model with both encoder and decoder"
        model = []
        for i in range(args.virtual_pipeline_model_parallel_size):
            ps.set_virtual_pipeline_model_parallel_rank(i)
            pre_process, post_process = ps.is_pipeline_first_stage(), ps.is_pipeline_last_stage()
            this_model = model_provider_func(pre_process=pre_process, post_process=post_process)
            this_model.model_type = model_type
            model.append(this_model)
    else:
        pre_process, post_process = ps.is_pipeline_first_stage(), ps.is_pipeline_last_stage()
        add_encoder, add_decoder = True, True
        if model_type == ModelType.encoder_and_decoder:
            if ps.get_pipeline_model_parallel_world_size() > 1:
                assert args.pipeline_model_parallel_split_rank is not None, \
                    "Split rank needs to be specified for model with both encoder and decoder"
                rank = ps.get_pipeline_model_parallel_rank()
                split_rank = args.pipeline_model_parallel_split_rank
                world_size = ps.get_pipeline_model_parallel_world_size()
                pre_process = rank == 0 or rank == split_rank
                post_process = (rank == (split_rank - 1)) or (rank == (world_size - 1))
                add_encoder = ps.is_pipeline_stage_before_split()
                add_decoder = ps.is_pipeline_stage_after_split()
            model = model_provider_func(pre_process=pre_process, post_process=post_process,
                                        add_encoder=add_encoder, add_decoder=add_decoder)
        else:
            model = model_provider_func(pre_process=pre_process, post_process=post_process)
        model.model_type = model_type
    if not isinstance(model, list):
        model = [model]
    for m in model:
        for param in m.parameters():
            set_defaults_if_not_set_tensor_model_parallel_attributes(param)
    if ps.get_data_parallel_rank() == 0:
        print(" > number of parameters on (tensor, pipeline) model parallel rank ({}, {}): {}".format(
            ps.get_tensor_model_parallel_rank(), ps.get_pipeline_model_parallel_rank(),
            sum(sum(p.nelement() for p in m.parameters()) for m in model)), flush=True)
    for m in model:
        m.to(current_device())
    if args.fp16 or args.bf16:
        model = [Float16Module(m, args) for m in model]
    if wrap_with_ddp:
        if args.DDP_impl in ("local", "torch"):
            # the local wrapper is already bucketed + overlapped, so ``--DDP_impl torch`` maps onto it as well
            model = [LocalDDP(m, args.accumulate_allreduce_grads_in_fp32 or args.fp16,
                              args.use_contiguous_buffers_in_local_ddp,
                              bucket_size_mb=getattr(args, "ddp_bucket_size_mb", 256),
                              use_distributed_optimizer=args.use_distributed_optimizer) for m in model]
            if args.data_parallel_random_init:
                for m in model:
                    m.broadcast_params()
            if use_cuda() and args.distributed_backend == "nccl" and getattr(args, "fused_dp_comm", True) \
                    and ps.get_data_parallel_world_size() > 1:
                try:
                    from .parallel import symm
                    for m in model:
                        symm.bind_dp_communicator(m)
                except Exception as e:  # the NCCL bucketed path is the checked fallback
                    print_rank_0(f"WARNING: peer-memory DP reduction unavailable ({e!r}); using NCCL")
        else:
            raise NotImplementedError("Unknown DDP implementation specified: {}. Exiting.".format(args.DDP_impl))
    return model


def _get_optimizer_param_scheduler(optimizer, args):
    if args.train_iters:
        if args.lr_decay_iters is None:
            args.lr_decay_iters = args.train_iters
        lr_decay_steps = args.lr_decay_iters * args.global_batch_size
        wd_incr_steps = args.train_iters * args.global_batch_size
        if args.lr_warmup_fraction is not None:
            lr_warmup_steps = args.lr_warmup_fraction * lr_decay_steps
        else:
            lr_warmup_steps = args.lr_warmup_iters * args.global_batch_size
    elif args.train_samples:
        update_train_iters(args)
        if args.lr_decay_samples is None:
            args.lr_decay_samples = args.train_samples
        lr_decay_steps = args.lr_decay_samples
        wd_incr_steps = args.train_samples
        if args.lr_warmup_fraction is not None:
            lr_warmup_steps = args.lr_warmup_fraction * lr_decay_steps
        else:
            lr_warmup_steps = args.lr_warmup_samples
    else:
        raise Exception("either train_iters or train_samples should be provided.")
    return OptimizerParamScheduler(
        optimizer, max_lr=args.lr, min_lr=args.min_lr, lr_warmup_steps=lr_warmup_steps, lr_decay_steps=lr_decay_steps,
        lr_decay_style=args.lr_decay_style, start_wd=args.start_weight_decay, end_wd=args.end_weight_decay,
        wd_incr_steps=wd_incr_steps, wd_incr_style=args.weight_decay_incr_style,
        use_checkpoint_opt_param_scheduler=args.use_checkpoint_opt_param_scheduler,
        override_opt_param_scheduler=args.override_opt_param_scheduler)


get_optimizer_param_scheduler = _get_optimizer_param_scheduler


def _setup_model_and_optimizer(model_provider_func, model_type, no_wd_decay_cond=None, scale_lr_cond=None,
                               lr_mult=1.0, args=None):
    if args is None:
        args = get_args()
    model = get_model(model_provider_func, model_type, args=args)
    optimizer = get_megatron_optimizer(model, no_wd_decay_cond, scale_lr_cond, lr_mult)
    opt_param_scheduler = _get_optimizer_param_scheduler(optimizer, args)
    if args.load is not None:
        timers = get_timers()
        timers("load-checkpoint", log_level=0).start(barrier=True)
        args.iteration = load_checkpoint(model, optimizer, opt_param_scheduler)
        timers("load-checkpoint").stop(barrier=True)
        timers.log(["load-checkpoint"])
    else:
        args.iteration = 0
    assert args.DDP_impl == "local" or len(model) == 1
    unwrapped = unwrap_model(model)
    if args.iteration == 0 and len(unwrapped) == 1 and hasattr(unwrapped[0], "init_state_dict_from_bert"):
        print_rank_0("Initializing ICT from pretrained BERT model")
        unwrapped[0].init_state_dict_from_bert()
        if args.fp16:
            optimizer.reload_model_params()
    return model, optimizer, opt_param_scheduler


setup_model_and_optimizer = _setup_model_and_optimizer


def train_step(forward_step_func, data_iterator, model, optimizer, opt_param_scheduler):
    """One optimizer step: zero grads -> fwd/bwd over all micro-batches -> reduce -> step -> (gather) -> LR."""
    args = get_args()
    timers = get_timers()
    if args.DDP_impl == "local" and args.use_contiguous_buffers_in_local_ddp:
        for partition in model:
            partition.zero_grad_buffer()
    optimizer.zero_grad()

    timers("forward-backward", log_level=1).start(barrier=args.barrier_with_L1_time)
    forward_backward_func = get_forward_backward_func()
    fwd_bwd_timers = timers if args.timing_log_level > 1 else None
    losses_reduced = forward_backward_func(forward_step_func, data_iterator, model, optimizer, fwd_bwd_timers,
                                           forward_only=False)
    timers("forward-backward").stop()
    if args.empty_unused_memory_level >= 1 and use_cuda():
        torch.cuda.empty_cache()

    optimizer.reduce_model_grads(args, timers)
    timers("optimizer", log_level=1).start(barrier=args.barrier_with_L1_time)
    update_successful, grad_norm, num_zeros_in_grad = optimizer.step(args, timers)
    timers("optimizer").stop()
    if update_successful:
        optimizer.gather_model_params(args, timers)
        increment = get_num_microbatches() * args.micro_batch_size * args.data_parallel_size
        opt_param_scheduler.step(increment=increment)
        skipped_iter = 0
    else:
        skipped_iter = 1
    if use_cuda():
        from .parallel import symm
        symm.check_timeouts()      # a timed-out peer-memory handshake is fatal, not a silent fallback
    if args.empty_unused_memory_level >= 2 and use_cuda():
        torch.cuda.empty_cache()

    if ps.is_pipeline_last_stage(ignore_virtual=True):
        loss_reduced = {}
        for key in losses_reduced[0]:
            vals = [x[key] for x in losses_reduced]
            loss_reduced[key] = sum(vals) / len(vals)
        return loss_reduced, skipped_iter, grad_norm, num_zeros_in_grad
    return {}, skipped_iter, grad_norm, num_zeros_in_grad


def _to_float(x):
    if isinstance(x, torch.Tensor):
        return x.float().sum().item() if x.numel() == 1 else x.float().mean().item()
    return float(x) if x is not None else None


def training_log(loss_dict, total_loss_dict, learning_rate, iteration, loss_scale, report_memory_flag, skipped_iter,
                 grad_norm, params_norm, num_zeros_in_grad):
    """Accumulate losses (on device), and every ``log_interval`` print / write TB+W&B scalars and timers."""
    args = get_args()
    timers = get_timers()
    writer = get_tensorboard_writer()
    counters = get_counters()
    advanced_iters_key, skipped_iters_key, nan_iters_key = "advanced iterations", "skipped iterations", "nan iterations"
    if not skipped_iter:
        total_loss_dict[advanced_iters_key] = total_loss_dict.get(advanced_iters_key, 0) + 1
    else:
        total_loss_dict.setdefault(advanced_iters_key, 0)
    total_loss_dict[skipped_iters_key] = total_loss_dict.get(skipped_iters_key, 0) + skipped_iter
    got_nan_t = None
    for key in loss_dict:
        if not skipped_iter:
            zero = torch.zeros(1, dtype=torch.float32, device=current_device())
            total_loss_dict[key] = total_loss_dict.get(key, zero) + loss_dict[key].detach().float().view(-1)[:1]
        else:
            v = loss_dict[key].float().sum().item()
            if v == float("inf") or v == -float("inf") or v != v:
                got_nan_t = True
    total_loss_dict[nan_iters_key] = total_loss_dict.get(nan_iters_key, 0) + int(bool(got_nan_t))

    timers_to_log = ["forward-backward", "forward-compute", "backward-compute", "batch-generator", "forward-recv",
                     "forward-send", "backward-recv", "backward-send", "forward-send-forward-recv",
                     "forward-send-backward-recv", "backward-send-forward-recv", "backward-send-backward-recv",
                     "forward-backward-send-forward-backward-recv", "layernorm-grads-all-reduce",
                     "embedding-grads-all-reduce", "grads-all-reduce", "grads-reduce-scatter", "params-all-gather",
                     "optimizer-copy-to-main-grad", "optimizer-unscale-and-check-inf", "optimizer-clip-main-grad",
                     "optimizer-count-zeros", "optimizer-inner-step", "optimizer-copy-main-to-model-params",
                     "optimizer"]
    batch_size = args.micro_batch_size * args.data_parallel_size * get_num_microbatches()
    total_iterations = total_loss_dict[advanced_iters_key] + total_loss_dict[skipped_iters_key]

    log_now = iteration % args.log_interval == 0
    tb_now = writer is not None and (iteration % args.tensorboard_log_interval == 0)
    if tb_now:
        samples = args.consumed_train_samples
        if getattr(args, "log_learning_rate_to_tensorboard", True):
            writer.add_scalar("learning-rate", learning_rate, iteration)
            writer.add_scalar("learning-rate vs samples", learning_rate, samples)
        if args.log_batch_size_to_tensorboard:
            writer.add_scalar("batch-size", batch_size, iteration)
            writer.add_scalar("batch-size vs samples", batch_size, samples)
        for key in loss_dict:
            v = _to_float(loss_dict[key])
            writer.add_scalar(key, v, iteration)
            writer.add_scalar(key + " vs samples", v, samples)
        if getattr(args, "log_loss_scale_to_tensorboard", True):
            writer.add_scalar("loss-scale", _to_float(loss_scale), iteration)
            writer.add_scalar("loss-scale vs samples", _to_float(loss_scale), samples)
        if args.log_world_size_to_tensorboard:
            writer.add_scalar("world-size", args.world_size, iteration)
            writer.add_scalar("world-size vs samples", args.world_size, samples)
        if grad_norm is not None:
            writer.add_scalar("grad-norm", _to_float(grad_norm), iteration)
            writer.add_scalar("grad-norm vs samples", _to_float(grad_norm), samples)
        if num_zeros_in_grad is not None:
            writer.add_scalar("num-zeros", num_zeros_in_grad, iteration)
            writer.add_scalar("num-zeros vs samples", num_zeros_in_grad, samples)
        if params_norm is not None:
            writer.add_scalar("params-norm", params_norm, iteration)
            writer.add_scalar("params-norm vs samples", params_norm, samples)
        if args.log_memory_to_tensorboard and use_cuda():
            mem_stats = torch.cuda.memory_stats()
            writer.add_scalar("mem-reserved-bytes", mem_stats["reserved_bytes.all.current"], iteration)
            writer.add_scalar("mem-allocated-bytes", mem_stats["allocated_bytes.all.current"], iteration)
            writer.add_scalar("mem-allocated-count", mem_stats["allocation.all.current"], iteration)

    if log_now:
        elapsed_time = timers("interval-time").elapsed(barrier=True)
        elapsed_time_per_iteration = elapsed_time / max(1, total_iterations)
        tokens_per_sec = counters["tokens"] / max(elapsed_time, 1e-9)
        counters["tokens"] = 0
        if writer:
            if args.log_timers_to_tensorboard:
                writer.add_scalar("iteration-time", elapsed_time_per_iteration, iteration)
            writer.add_scalar("tokens-per-sec", tokens_per_sec, iteration)
        log_string = " iteration {:8d}/{:8d} |".format(iteration, args.train_iters)
        log_string += " consumed samples: {:12d} |".format(args.consumed_train_samples)
        log_string += " elapsed time per iteration (ms): {:.1f} |".format(elapsed_time_per_iteration * 1000.0)
        log_string += " tokens/sec: {:.1f} |".format(tokens_per_sec)
        log_string += " learning rate: {:.3E} |".format(learning_rate)
        log_string += " global batch size: {:5d} |".format(batch_size)
        for key in list(total_loss_dict.keys()):
            if key not in [advanced_iters_key, skipped_iters_key, nan_iters_key]:
                avg = total_loss_dict[key].item() / float(max(1, total_loss_dict[advanced_iters_key]))
                if avg > 0.0:
                    log_string += " {}: {:.6E} |".format(key, avg)
                total_loss_dict[key] = torch.zeros(1, dtype=torch.float32, device=current_device())
        log_string += " loss scale: {:.1f} |".format(_to_float(loss_scale))
        if grad_norm is not None:
            log_string += " grad norm: {:.3f} |".format(_to_float(grad_norm))
        if num_zeros_in_grad is not None:
            log_string += " num zeros: {:.1f} |".format(num_zeros_in_grad)
        if params_norm is not None:
            log_string += " params norm: {:.3f} |".format(params_norm)
        log_string += " number of skipped iterations: {:3d} |".format(total_loss_dict[skipped_iters_key])
        log_string += " number of nan iterations: {:3d} |".format(total_loss_dict[nan_iters_key])
        total_loss_dict[advanced_iters_key] = 0
        total_loss_dict[skipped_iters_key] = 0
        total_loss_dict[nan_iters_key] = 0
        print_rank_last(log_string)
        if report_memory_flag and learning_rate > 0.0:
            report_memory("(after {} iterations)".format(iteration))
            report_memory_flag = False
        timers.log(timers_to_log, normalizer=args.log_interval)
    if tb_now and args.log_timers_to_tensorboard:
        timers.write(timers_to_log, writer, iteration, normalizer=total_iterations or 1)
    return report_memory_flag


def save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler):
    timers = get_timers()
    timers("save-checkpoint", log_level=0).start(barrier=True)
    save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
    timers("save-checkpoint").stop(barrier=True)
    timers.log(["save-checkpoint"])


def _train(args, forward_step_func, model, optimizer, opt_param_scheduler, train_data_iterator, valid_data_iterator,
           process_non_loss_data_func):
    timers = get_timers()
    _initialize.write_args_to_tensorboard()
    for m in model:
        m.train()
    total_loss_dict = {}
    iteration = args.iteration
    timers("interval-time", log_level=0).start(barrier=True)
    print_datetime("before the start of training step")
    report_memory_flag = True
    from .profiler import StepProfiler
    step_profiler = StepProfiler(args, dist.get_rank() if dist.is_initialized() else 0)
    while iteration < args.train_iters:
        update_num_microbatches(args.consumed_train_samples)
        args.curr_iteration = iteration
        step_profiler.step_begin(iteration)
        if iteration in args.skip_iters:
            print_rank_0(f"=== skipping iteration {iteration} (forward only) ===")
            fwd = get_forward_backward_func()
            fwd(forward_step_func, train_data_iterator, model, optimizer, None, forward_only=True)
            loss_dict, skipped_iter, grad_norm, num_zeros_in_grad = {}, 1, None, None
        else:
            loss_dict, skipped_iter, grad_norm, num_zeros_in_grad = train_step(
                forward_step_func, train_data_iterator, model, optimizer, opt_param_scheduler)
        trace = step_profiler.step_end(iteration)
        if trace:
            print(f"> profiler trace of iterations [{step_profiler.start}, {step_profiler.end}) written to {trace}",
                  flush=True)
        iteration += 1
        args.consumed_train_samples += ps.get_data_parallel_world_size() * args.micro_batch_size * \
            get_num_microbatches()
        loss_scale = optimizer.get_loss_scale()
        params_norm = calc_params_l2_norm(model) if args.log_params_norm else None
        lr = max(g["lr"] for g in optimizer.param_groups)
        report_memory_flag = training_log(loss_dict, total_loss_dict, lr, iteration, loss_scale, report_memory_flag,
                                          skipped_iter, grad_norm, params_norm, num_zeros_in_grad)
        if args.adlr_autoresume and (iteration % args.adlr_autoresume_interval == 0):
            check_adlr_autoresume_termination(iteration, model, optimizer, opt_param_scheduler)
        if args.eval_interval and iteration % args.eval_interval == 0 and args.do_valid:
            evaluate_and_print_results("iteration {}".format(iteration), forward_step_func, valid_data_iterator, model,
                                       iteration, process_non_loss_data_func, verbose=False, args=args)
        saved_checkpoint = False
        if args.exit_signal_handler:
            if any(get_signal_handler().signals_received()):
                save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler)
                print_datetime("exiting program after receiving SIGTERM.")
                sys.exit()
        if args.save and args.save_interval and iteration % args.save_interval == 0:
            save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler)
            saved_checkpoint = True
        if args.exit_duration_in_mins:
            train_time = (time.time() - _TRAIN_START_TIME) / 60.0
            done = torch.tensor([train_time > args.exit_duration_in_mins], dtype=torch.int, device=current_device())
            if dist.is_initialized():
                dist.all_reduce(done, op=dist.ReduceOp.MAX)
            if done.item():
                if not saved_checkpoint:
                    save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler)
                print_datetime("exiting program after {} minutes".format(train_time))
                sys.exit()
        if args.exit_interval and iteration % args.exit_interval == 0:
            if args.save and not saved_checkpoint:
                save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler)
            if dist.is_initialized():
                dist.barrier()
            print_datetime("exiting program at iteration {}".format(iteration))
            sys.exit()
    return iteration


def evaluate(forward_step_func, data_iterator, model, process_non_loss_data_func, verbose=False, args=None):
    if args is None:
        args = get_args()
    for m in model:
        m.eval()
    total_loss_dict = {}
    with torch.no_grad():
        iteration = 0
        while iteration < args.eval_iters:
            iteration += 1
            if verbose and iteration % args.log_interval == 0:
                print_rank_0("Evaluating iter {}/{}".format(iteration, args.eval_iters))
            forward_backward_func = get_forward_backward_func()
            loss_dicts = forward_backward_func(forward_step_func, data_iterator, model, optimizer=None, timers=None,
                                               forward_only=True)
            if args.empty_unused_memory_level >= 1 and use_cuda():
                torch.cuda.empty_cache()
            if ps.is_pipeline_last_stage(ignore_virtual=True):
                for loss_dict in loss_dicts:
                    for key in loss_dict:
                        v = loss_dict[key]
                        v = v.detach().float().view(-1)[:1] if isinstance(v, torch.Tensor) else \
                            torch.tensor([float(v)], device=current_device())
                        total_loss_dict[key] = total_loss_dict.get(
                            key, torch.zeros(1, dtype=torch.float32, device=current_device())) + v
            args.consumed_valid_samples += ps.get_data_parallel_world_size() * args.micro_batch_size * \
                get_num_microbatches()
        collected_non_loss_data = None
        if process_non_loss_data_func is not None and is_last_rank():
            collected_non_loss_data = forward_backward_func(forward_step_func, data_iterator, model, optimizer=None,
                                                            timers=None, forward_only=True,
                                                            collect_non_loss_data=True)
    for m in model:
        m.train()
    for key in total_loss_dict:
        total_loss_dict[key] /= args.eval_iters * get_num_microbatches()
    return total_loss_dict, collected_non_loss_data


def evaluate_and_print_results(prefix, forward_step_func, data_iterator, model, iteration, process_non_loss_data_func,
                               verbose=False, args=None):
    if args is None:
        args = get_args()
    writer = get_tensorboard_writer()
    total_loss_dict, collected = evaluate(forward_step_func, data_iterator, model, process_non_loss_data_func,
                                          verbose, args=args)
    string = " validation loss at {} | ".format(prefix)
    for key in total_loss_dict:
        val = total_loss_dict[key].item()
        string += "{} value: {:.6E} | ".format(key, val)
        if key == "lm loss":
            ppl = math.exp(min(20, val))
            string += "{} PPL: {:.6E} | ".format(key, ppl)
        if writer:
            writer.add_scalar("{} validation".format(key), val, iteration)
            writer.add_scalar("{} validation vs samples".format(key), val, args.consumed_train_samples)
            if args.log_validation_ppl_to_tensorboard and key == "lm loss":
                writer.add_scalar("{} validation ppl".format(key), ppl, iteration)
                writer.add_scalar("{} validation ppl vs samples".format(key), ppl, args.consumed_train_samples)
    if process_non_loss_data_func is not None and writer and is_last_rank():
        process_non_loss_data_func(collected, iteration, writer)
    length = len(string) + 1
    print_rank_last("-" * length)
    print_rank_last(string)
    print_rank_last("-" * length)
    if writer and hasattr(writer, "flush_all"):
        writer.flush_all()


def cyclic_iter(it):
    while True:
        for x in it:
            yield x


def build_train_valid_test_data_iterators(build_train_valid_test_datasets_provider, args, collate_fn=None):
    """Datasets + loaders are built on TP-rank 0 only (other TP ranks receive batches through
    ``broadcast_data``); the do_train/valid/test flags are broadcast over the TP group."""
    train_dataloader = valid_dataloader = test_dataloader = None
    print_rank_0("> building train, validation, and test datasets ...")
    if args.iteration > 0 and args.consumed_train_samples == 0:
        assert args.train_samples is None, "only backward compatiblity support for iteration-based training"
        args.consumed_train_samples = args.iteration * args.global_batch_size
    if args.iteration > 0 and args.consumed_valid_samples == 0:
        if args.train_samples is None:
            args.consumed_valid_samples = (args.iteration // args.eval_interval) * args.eval_iters * \
                args.global_batch_size
    if ps.get_tensor_model_parallel_rank() == 0:
        train_samples = args.train_samples if args.train_samples else args.train_iters * args.global_batch_size
        eval_iters = (args.train_iters // args.eval_interval + 1) * args.eval_iters if args.eval_interval else 0
        test_iters = args.eval_iters
        nums = [train_samples, eval_iters * args.global_batch_size, test_iters * args.global_batch_size]
        print_rank_0(" > datasets target sizes (minimum size):")
        print_rank_0("    train:      {}".format(nums[0]))
        print_rank_0("    validation: {}".format(nums[1]))
        print_rank_0("    test:       {}".format(nums[2]))
        train_ds, valid_ds, test_ds = build_train_valid_test_datasets_provider(nums)
        train_dataloader = build_pretraining_data_loader(train_ds, args.consumed_train_samples, collate_fn=collate_fn)
        valid_dataloader = build_pretraining_data_loader(valid_ds, args.consumed_valid_samples, collate_fn=collate_fn)
        test_dataloader = build_pretraining_data_loader(test_ds, 0, collate_fn=collate_fn)
        do_train = train_dataloader is not None and args.train_iters > 0
        do_valid = valid_dataloader is not None and args.eval_iters > 0
        do_test = test_dataloader is not None and args.eval_iters > 0
        flags = torch.tensor([int(do_train), int(do_valid), int(do_test)], dtype=torch.long, device=current_device())
    else:
        flags = torch.tensor([0, 0, 0], dtype=torch.long, device=current_device())
    if dist.is_initialized() and ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(flags, ps.get_tensor_model_parallel_src_rank(), group=ps.get_tensor_model_parallel_group())
    args.do_train, args.do_valid, args.do_test = (bool(f) for f in flags.tolist())
    dl_type = args.dataloader_type
    assert dl_type in ["single", "cyclic"]

    def make_iter(dl):
        if dl is None:
            return None
        return iter(dl) if dl_type == "single" else iter(cyclic_iter(dl))

    return make_iter(train_dataloader), make_iter(valid_dataloader), make_iter(test_dataloader)

"""Epoch-based fine-tuning loop shared by the downstream tasks (parity: tasks/finetune_utils.py).

The reference calls ``train_step`` with a pre-fork signature (bit-rotted); this one uses the current
``training.train_step`` / ``training_log`` / ``evaluate_and_print_results`` APIs."""
import sys
from functools import partial

import torch
import torch.distributed as dist

from megatron_llm_b200 import get_args, get_num_microbatches, get_timers, print_rank_0
from megatron_llm_b200 import training
from megatron_llm_b200.checkpointing import load_checkpoint, save_checkpoint
from megatron_llm_b200.models.enums import ModelType
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.utils import average_losses_across_data_parallel_group, calc_params_l2_norm
from megatron_llm_b200.utils.device import current_device


def process_batch(batch, is_fp16=False):
    dev = current_device()
    tokens = batch["text"].long().to(dev).contiguous()
    types = batch["types"].long().to(dev).contiguous()
    labels = batch["label"].long().to(dev).contiguous()
    attention_mask = batch["padding_mask"].float().to(dev).contiguous()
    if is_fp16:
        attention_mask = attention_mask.half()
    return tokens, types, labels, attention_mask


def cross_entropy_loss_func(labels, output_tensor):
    loss = torch.nn.functional.cross_entropy(output_tensor.contiguous().float(), labels)
    avg = average_losses_across_data_parallel_group([loss])
    return loss, {"lm loss": avg[0]}


def _cross_entropy_forward_step(batch, model):
    timers = get_timers()
    args = get_args()
    timers("batch-generator", log_level=2).start()
    try:
        batch_ = next(batch)
    except TypeError:          # already a batch, not an iterator
        batch_ = batch
    tokens, types, labels, attention_mask = process_batch(batch_, args.fp16)
    timers("batch-generator").stop()
    return model(tokens, attention_mask, tokentype_ids=types), partial(cross_entropy_loss_func, labels)


def build_data_loader(dataset, micro_batch_size, num_workers, drop_last, task_collate_fn=None):
    """``micro_batch_size`` is per data-parallel rank."""
    sampler = torch.utils.data.distributed.DistributedSampler(
        dataset, num_replicas=mpu.get_data_parallel_world_size(), rank=mpu.get_data_parallel_rank())
    return torch.utils.data.DataLoader(dataset, batch_size=micro_batch_size, sampler=sampler, shuffle=False,
                                       num_workers=num_workers, drop_last=drop_last,
                                       pin_memory=torch.cuda.is_available(), collate_fn=task_collate_fn)


def _build_infinite_size_dataloader(dataloader):
    while True:
        yield from dataloader


def _build_train_valid_dataloaders(train_dataset, valid_dataset, task_collate_fn=None):
    args = get_args()
    print_rank_0("building train and validation dataloaders ...")
    train = build_data_loader(train_dataset, args.micro_batch_size, args.num_workers, not args.keep_last,
                              task_collate_fn)
    args.train_iters_per_epoch = len(train)
    args.train_iters = args.epochs * args.train_iters_per_epoch
    valid = _build_infinite_size_dataloader(build_data_loader(valid_dataset, args.micro_batch_size, args.num_workers,
                                                              not args.keep_last, task_collate_fn))
    # multiple-choice style datasets expand every sample into ``sample_multiplier`` sequences
    args.orig_micro_batch_size, args.orig_global_batch_size = args.micro_batch_size, args.global_batch_size
    if hasattr(train_dataset, "sample_multiplier"):
        args.micro_batch_size *= train_dataset.sample_multiplier
        args.global_batch_size *= train_dataset.sample_multiplier
    return train, valid


def _train(model, optimizer, opt_param_scheduler, forward_step, train_dataloader, valid_dataloader,
           end_of_epoch_callback, args):
    timers = get_timers()
    assert get_num_microbatches() == 1, "finetuning with gradient accumulation doesn't currently work"
    for m in model:
        m.train()
    losses_dict_sum = {}
    start_epoch = args.iteration // args.train_iters_per_epoch
    start_iteration = args.iteration % args.train_iters_per_epoch
    iteration = args.iteration
    report_memory_flag = True
    timers("interval-time", log_level=0).start(barrier=True)
    for epoch in range(start_epoch, args.epochs):
        print_rank_0("working on epoch {} ...".format(epoch + 1))
        train_dataloader.sampler.set_epoch(args.seed + epoch)
        for it_, batch in enumerate(train_dataloader):
            if it_ < start_iteration:
                continue
            start_iteration = 0
            losses_dict, skipped_iter, grad_norm, num_zeros = training.train_step(
                forward_step, batch, model, optimizer, opt_param_scheduler)
            iteration += 1
            params_norm = calc_params_l2_norm(model) if args.log_params_norm else None
            report_memory_flag = training.training_log(
                losses_dict, losses_dict_sum, optimizer.param_groups[0]["lr"], iteration,
                optimizer.get_loss_scale().item(), report_memory_flag, skipped_iter, grad_norm, params_norm, num_zeros)
            saved = False
            if args.save and args.save_interval and iteration % args.save_interval == 0:
                save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
                saved = True
            if args.eval_interval and iteration % args.eval_interval == 0:
                training.evaluate_and_print_results("iteration {}".format(iteration), forward_step, valid_dataloader,
                                                    model, iteration, None, False, args=args)
            if args.exit_interval and iteration % args.exit_interval == 0:
                if not saved:
                    save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
                dist.barrier()
                print_rank_0("exiting program at iteration {}".format(iteration))
                sys.exit()
        if args.save:
            save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
        if end_of_epoch_callback is not None:
            end_of_epoch_callback(model, epoch)


def finetune(train_valid_datasets_provider, model_provider, model_type=ModelType.encoder_or_decoder,
             forward_step=_cross_entropy_forward_step, end_of_epoch_callback_provider=None, task_collate_fn=None):
    """Main entry used by every fine-tuning task; ``--epochs 0`` = evaluation only."""
    args = get_args()
    timers = get_timers()
    assert args.rampup_batch_size is None, "batch size scaling is not supported for finetuning"
    timers("train/valid/test dataset/dataloder", log_level=0).start()
    if args.epochs > 0:
        train_dataset, valid_dataset = train_valid_datasets_provider()
        train_dataloader, valid_dataloader = _build_train_valid_dataloaders(train_dataset, valid_dataset,
                                                                            task_collate_fn)
    else:
        args.train_iters = 0
        args.orig_micro_batch_size, args.orig_global_batch_size = args.micro_batch_size, args.global_batch_size
    timers("train/valid/test dataset/dataloder").stop()
    end_of_epoch_callback = end_of_epoch_callback_provider() if end_of_epoch_callback_provider is not None else None
    timers("model and optimizer", log_level=0).start()
    model, optimizer, opt_param_scheduler = training.setup_model_and_optimizer(model_provider, model_type, args=args)
    timers("model and optimizer").stop()
    timers("pretrained checkpoint", log_level=0).start(barrier=True)
    if args.iteration == 0 and args.pretrained_checkpoint is not None:
        saved = args.load, args.no_load_rng
        args.load, args.no_load_rng = args.pretrained_checkpoint, True
        load_checkpoint(model, None, None, strict=False)
        args.load, args.no_load_rng = saved
        optimizer.reload_model_params()
    timers("pretrained checkpoint").stop()
    print_rank_0("done with setups ...")
    timers.log(["train/valid/test dataset/dataloder", "model and optimizer", "pretrained checkpoint"], barrier=True)
    print_rank_0("training ...")
    if args.epochs > 0:
        _train(model, optimizer, opt_param_scheduler, forward_step, train_dataloader, valid_dataloader,
               end_of_epoch_callback, args)
    elif end_of_epoch_callback is not None:
        print_rank_0("evaluation only mode, setting epoch to -1")
        end_of_epoch_callback(model, epoch=-1, output_predictions=True)
    print_rank_0("done :-)")

"""Accuracy evaluation over one or more validation sets (parity: tasks/eval_utils.py)."""
import os
import time
from functools import partial

import torch
import torch.distributed as dist

from megatron_llm_b200 import get_args, print_rank_last
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.parallel.schedules import get_forward_backward_func
from megatron_llm_b200.utils import is_last_rank
from megatron_llm_b200.utils.device import current_device

from tasks import finetune_utils


def accuracy_func_provider(single_dataset_provider):
    """Returns ``metrics_func(model, epoch, output_predictions=False)`` over all ``--valid_data`` paths."""
    args = get_args()
    dataloaders = []
    for datapath in args.valid_data:
        dataset = single_dataset_provider(datapath)
        mbs = getattr(args, "orig_micro_batch_size", args.micro_batch_size)
        loader = finetune_utils.build_data_loader(dataset, mbs, num_workers=args.num_workers,
                                                  drop_last=(mpu.get_data_parallel_world_size() > 1))
        dataloaders.append((dataset.dataset_name, loader))

    def metrics_func(model, epoch, output_predictions=False):
        print_rank_last("calculating metrics ...")
        correct = total = 0
        named_predictions, names = [], "predictions"
        if output_predictions:
            assert mpu.get_data_parallel_world_size() == 1
        for name, loader in dataloaders:
            out = calculate_correct_answers(name, model, loader, epoch, output_predictions)
            if output_predictions:
                c, t, predictions = out
                named_predictions.append((name, predictions))
                names += "_" + name
            else:
                c, t = out
            correct += c
            total += t
        if is_last_rank():
            print(" >> |epoch: {}| overall: correct / total = {} / {} = {:.4f} %".format(
                epoch, correct, total, float(correct) * 100.0 / max(float(total), 1.0)))
        if output_predictions and is_last_rank():
            assert args.load is not None
            torch.save(named_predictions, os.path.join(args.load, names + ".pt"))

    return metrics_func


def calculate_correct_answers(name, model, dataloader, epoch, output_predictions):
    args = get_args()
    forward_backward_func = get_forward_backward_func()
    t0 = time.time()
    for m in model:
        m.eval()
    saved = args.micro_batch_size, args.global_batch_size
    multiplier = getattr(dataloader.dataset, "sample_multiplier", 1)
    orig_mbs = getattr(args, "orig_micro_batch_size", args.micro_batch_size)
    orig_gbs = getattr(args, "orig_global_batch_size", args.global_batch_size)
    num_micro_batches = max(1, orig_gbs // (orig_mbs * args.data_parallel_size))

    def loss_func(labels, uids, output_tensor):
        logits = output_tensor
        info = {}
        if output_predictions:
            info["softmaxes"] = torch.softmax(logits.float(), dim=-1).cpu().numpy().tolist()
            info["labels"] = labels.cpu().numpy().tolist()
            info["ids"] = uids.cpu().numpy().tolist()
        info["total"] = labels.size(0)
        info["correct"] = (torch.argmax(logits, dim=-1) == labels).sum().item()
        return 0, info

    def forward_step(batch, model):
        try:
            batch_ = next(batch)
        except TypeError:
            batch_ = batch
        tokens, types, labels, attention_mask = finetune_utils.process_batch(batch_, args.fp16)
        return model(tokens, attention_mask, tokentype_ids=types), partial(loss_func, labels, batch_["uid"])

    total = correct = 0
    softmaxes, labels, ids = [], [], []
    with torch.no_grad():
        for batch in dataloader:
            n = len(batch["label"])
            args.micro_batch_size = n * multiplier
            args.global_batch_size = n * multiplier * num_micro_batches
            for info in forward_backward_func(forward_step, batch, model, optimizer=None, timers=None,
                                              forward_only=True):
                if output_predictions:
                    softmaxes.extend(info["softmaxes"]), labels.extend(info["labels"]), ids.extend(info["ids"])
                total += info["total"]
                correct += info["correct"]
    for m in model:
        m.train()
    args.micro_batch_size, args.global_batch_size = saved
    if mpu.is_pipeline_last_stage():
        t = torch.tensor([correct, total], dtype=torch.long, device=current_device())
        dist.all_reduce(t, group=mpu.get_data_parallel_group())
        c, n = t[0].item(), t[1].item()
        print_rank_last(" > |epoch: {}| metrics for {}: correct / total = {} / {} = {:.4f} %, elapsed time (sec): "
                        "{:.3f}".format(epoch, name, c, n, float(c) * 100.0 / max(float(n), 1.0), time.time() - t0))
        return (c, n, (softmaxes, labels, ids)) if output_predictions else (c, n)
    return (0, 0, ()) if output_predictions else (0, 0)

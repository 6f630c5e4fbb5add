"""Validation metrics of the supervised retriever: top-k accuracy and mean gold rank against each question's own
negative pool (parity: tasks/orqa/supervised/eval_utils.py)."""
import math
import time
from collections import OrderedDict

import numpy as np
import torch

from megatron_llm_b200 import get_args, print_rank_0
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.utils import average_losses_across_data_parallel_group
from megatron_llm_b200.utils.device import current_device
from tasks import finetune_utils

_TENSOR_KEYS = ("query", "query_mask", "query_types", "query_pad_mask", "context", "context_mask", "context_types",
                "context_pad_mask")


def task_collate_fn(batch_data):
    out = OrderedDict()
    for d in batch_data:
        for k, v in d.items():
            out.setdefault(k, []).append(v)
    for k in _TENSOR_KEYS:
        out[k] = torch.from_numpy(np.asarray(out[k])).long()
    for k in ("neg_context", "neg_context_mask", "neg_context_types"):
        if k in out:       # negatives of all samples are stacked along the batch dimension
            out[k] = torch.from_numpy(np.concatenate(out[k])).long()
    return out


def process_batch(batch):
    dev = current_device()
    ids = lambda k: batch[k].long().to(dev)                 # noqa: E731
    mask = lambda k: (batch[k] < 0.5).to(dev)               # noqa: E731
    neg = "neg_context" in batch
    return (ids("query"), mask("query_mask"), ids("query_types"), ids("query_pad_mask"), ids("context"),
            mask("context_mask"), ids("context_types"), ids("context_pad_mask"),
            ids("neg_context") if neg else None, mask("neg_context_mask") if neg else None,
            ids("neg_context_types") if neg else None, batch["reference"])


def accuracy_func_provider(single_dataset_provider, rank0sampler=False):
    args = get_args()
    dataset = single_dataset_provider(args.valid_data)
    drop_last = mpu.get_data_parallel_world_size() > 1 and not rank0sampler
    loader = finetune_utils.build_data_loader(dataset, args.eval_micro_batch_size or args.micro_batch_size,
                                              num_workers=args.num_workers, drop_last=drop_last,
                                              task_collate_fn=task_collate_fn)

    def metrics_func(model, epoch, output_predictions=False):
        print_rank_0("calculating metrics by accuracy func in ORQA...")
        if args.task != "RET-FINETUNE-NQ":
            raise AssertionError("{} Task not supported".format(args.task))
        t0 = time.time()
        stats, total = retrieval_loss(model, loader)
        print_rank_0("epoch:{}".format(epoch) + "".join("|{} = {:.2f}".format(k, float(v) / max(total, 1))
                                                        for k, v in stats.items()))
        print_rank_0("taken time to calcuate metrics {:.3f}".format(time.time() - t0))

    return metrics_func


@torch.no_grad()
def retrieval_loss(model, dataloader):
    args = get_args()
    assert len(model) == 1
    net = model[0]
    net.eval()
    stats = {"rank": 0.0, **{"top{}_acc".format(k): 0.0 for k in args.retriever_report_topk_accuracies}}
    total = 0
    for batch in dataloader:
        q_tok, q_mask, q_types, _, c_tok, c_mask, c_types, _, n_tok, n_mask, n_types, _ = process_batch(batch)
        q, c = net(q_tok, q_mask, q_types, torch.cat([c_tok, n_tok]), torch.cat([c_mask, n_mask]),
                   torch.cat([c_types, n_types]))
        scores = torch.matmul(q, c.t()).float()
        if args.retriever_score_scaling:
            scores = scores / math.sqrt(args.hidden_size)
        n = q.size(0)
        gold = scores.diagonal()[:, None]            # the positive of question i is context i
        gold_rank = (scores > gold).sum(dim=1).float()
        vals = [gold_rank.sum().reshape(1)] + [(gold_rank < k).float().sum().reshape(1)
                                               for k in args.retriever_report_topk_accuracies]
        red = average_losses_across_data_parallel_group(vals)
        stats["rank"] += red[0].item()
        for k, v in zip(args.retriever_report_topk_accuracies, red[1:]):
            stats["top{}_acc".format(k)] += v.item() * 100
        total += n
    net.train()
    return stats, total

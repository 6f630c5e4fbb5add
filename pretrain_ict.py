"""Pre-train the bi-encoder retriever with the inverse cloze task.  Parity: pretrain_ict.py (same CLI).

In-batch negatives over the whole data-parallel group: query / context embeddings are all-gathered (with a backward
that keeps only the local slice) and scored against each other; the label of query i is context i."""
import math
from functools import partial

import torch
import torch.distributed as dist
import torch.nn.functional as F

from megatron_llm_b200 import get_args, get_timers, print_rank_0
from megatron_llm_b200.data.biencoder_dataset_utils import get_ict_batch
from megatron_llm_b200.data.dataset_utils import build_train_valid_test_datasets
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.models.biencoder_model import biencoder_model_provider
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.training import pretrain
from megatron_llm_b200.utils import average_losses_across_data_parallel_group


def pretrain_ict_model_provider(pre_process=True, post_process=True):
    args = get_args()
    return biencoder_model_provider(only_context_model=False, only_query_model=False,
                                    biencoder_shared_query_context_model=args.biencoder_shared_query_context_model,
                                    pre_process=pre_process, post_process=post_process,
                                    model_type=ModelType.encoder_or_decoder)


def get_group_world_size_rank():
    group = mpu.get_data_parallel_group()
    return group, dist.get_rank(group=group), dist.get_world_size(group=group)


class AllgatherFromDataParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_):
        assert input_.dim() == 2
        group, rank, world = get_group_world_size_rank()
        out = torch.empty((world * input_.size(0), input_.size(1)), dtype=input_.dtype, device=input_.device)
        dist.all_gather_into_tensor(out, input_.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        group, rank, world = get_group_world_size_rank()
        assert grad_output.shape[0] % world == 0
        n = grad_output.shape[0] // world
        return grad_output[rank * n:(rank + 1) * n].contiguous()


def loss_func(output_tensor):
    args = get_args()
    query_logits, context_logits = output_tensor
    assert mpu.get_tensor_model_parallel_world_size() == 1, "Model parallel size > 1 not supported for ICT"
    all_q = AllgatherFromDataParallelRegion.apply(query_logits)
    all_c = AllgatherFromDataParallelRegion.apply(context_logits)
    n = all_q.size(0)
    scores = torch.matmul(all_q, all_c.t()).float()
    if args.retriever_score_scaling:
        scores = scores / math.sqrt(args.hidden_size)
    log_probs = F.log_softmax(scores, dim=1)
    labels = torch.arange(n, device=scores.device)
    # rank of the gold context among all contexts, computed on the device (no per-sample host loop)
    gold_rank = (log_probs > log_probs.gather(1, labels[:, None])).sum(dim=1)
    accs = [(gold_rank < int(k)).float().mean().reshape(1) for k in args.retriever_report_topk_accuracies]
    loss = F.nll_loss(log_probs, labels, reduction="mean")
    reduced = average_losses_across_data_parallel_group([loss, *accs])
    loss = loss * mpu.get_data_parallel_world_size()     # DDP averages gradients over DP; the loss is global already
    stats = {f"top{k}_acc": v * 100 for k, v in zip(args.retriever_report_topk_accuracies, reduced[1:])}
    return loss, dict(loss=reduced[0], **stats)


def forward_step(data_iterator, model):
    timers = get_timers()
    timers("batch-generator", log_level=2).start()
    query_tokens, query_mask, context_tokens, context_mask, context_indices = get_ict_batch(data_iterator)
    timers("batch-generator").stop()
    output_tensor = model(query_tokens, query_mask, torch.zeros_like(query_tokens), context_tokens, context_mask,
                          torch.zeros_like(context_tokens))
    return output_tensor, partial(loss_func)


def train_valid_test_datasets_provider(train_val_test_num_samples):
    args = get_args()
    print_rank_0("> building train, validation, and test datasets for BERT ICT...")
    ds = build_train_valid_test_datasets(
        data_prefix=args.data_path, data_impl=args.data_impl, splits_string=args.split,
        train_valid_test_num_samples=train_val_test_num_samples, max_seq_length=args.seq_length,
        masked_lm_prob=args.mask_prob, short_seq_prob=args.short_seq_prob, seed=args.seed,
        skip_warmup=(not args.mmap_warmup), binary_head=False, dataset_type="ict")
    print_rank_0("> finished creating BERT ICT datasets ...")
    return ds


def main(args_list=None):
    initialize_megatron(extra_args_provider=None, args_defaults={"tokenizer_type": "BertWordPieceLowerCase"},
                        args_list=args_list)
    pretrain(get_args(), train_valid_test_datasets_provider, pretrain_ict_model_provider,
             ModelType.encoder_or_decoder, forward_step)


if __name__ == "__main__":
    main()

"""Supervised fine-tuning of the bi-encoder retriever on Natural Questions (parity: tasks/orqa/supervised/finetune.py).

In-batch negatives across the data-parallel group; optional per-question hard negatives appended to the context batch
(padded to the same count on every rank before the gather)."""
import math
from functools import partial

import torch
import torch.distributed as dist
import torch.nn.functional as F

from megatron_llm_b200 import get_args, get_timers, get_tokenizer, print_rank_0
from megatron_llm_b200.models.biencoder_model import biencoder_model_provider
from megatron_llm_b200.models.enums import ModelType
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.utils import average_losses_across_data_parallel_group
from tasks import finetune_utils
from tasks.orqa.supervised.eval_utils import accuracy_func_provider, process_batch, task_collate_fn


def get_group_world_size_rank():
    group = mpu.get_data_parallel_group()
    return group, dist.get_rank(group=group), dist.get_world_size(group=group)


def check_and_append_tensor_for_gather(group, rank, world_size, input_):
    """Zero-pad dim 0 to the largest size over the group (ranks can draw different numbers of negatives)."""
    n = torch.tensor([input_.size(0)], device=input_.device)
    sizes = torch.empty(world_size, dtype=n.dtype, device=n.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    longest = int(sizes.max().item())
    if longest > input_.size(0):
        input_ = F.pad(input_, (0, 0) * (input_.dim() - 1) + (0, longest - input_.size(0)))
    return input_


class _GatherKeepLocalGrad(torch.autograd.Function):
    """all-gather along dim 0 whose backward returns the local slice (other ranks' embeddings are constants)."""

    @staticmethod
    def forward(ctx, x):
        group, rank, world = get_group_world_size_rank()
        ctx.n, ctx.rank = x.size(0), rank
        out = torch.empty((world * x.size(0),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n].contiguous()


def orqa(Dataset):
    def cross_entropy_forward_step(batch, model):
        timers = get_timers()
        timers("batch generator", log_level=2).start()
        try:
            batch_ = next(batch)
        except TypeError:
            batch_ = batch
        group, rank, world = get_group_world_size_rank()
        q_tok, q_mask, q_types, _, c_tok, c_mask, c_types, _, n_tok, n_mask, n_types, _ = process_batch(batch_)
        timers("batch generator").stop()
        if n_tok is not None:
            n_tok, n_mask, n_types = (check_and_append_tensor_for_gather(group, rank, world, t)
                                      for t in (n_tok, n_mask, n_types))
            c_tok, c_mask, c_types = torch.cat([c_tok, n_tok]), torch.cat([c_mask, n_mask]), torch.cat([c_types, n_types])
        out = model(q_tok, q_mask, q_types, c_tok, c_mask, c_types)
        return out, partial(cross_entropy_loss_func, q_tok, c_tok)

    def cross_entropy_loss_func(query_tokens, context_tokens, output_tensor):
        args = get_args()
        local_batch = query_tokens.shape[0]
        group, rank, world = get_group_world_size_rank()
        query_logits, context_logits = output_tensor
        if world > 1:
            all_q, all_c = _GatherKeepLocalGrad.apply(query_logits), _GatherKeepLocalGrad.apply(context_logits)
        else:
            all_q, all_c = query_logits, context_logits
        scores = torch.matmul(all_q, all_c.t()).float()
        if args.retriever_score_scaling:
            scores = scores / math.sqrt(args.hidden_size)
        if args.train_with_neg:     # every rank contributes [positives | negatives]: positives start each block
            per_rank = context_tokens.shape[0]
            labels = torch.cat([torch.arange(r * per_rank, r * per_rank + local_batch) for r in range(world)])
            labels = labels.to(scores.device)
        else:
            labels = torch.arange(world * local_batch, device=scores.device)
        log_probs = F.log_softmax(scores, dim=1)
        loss = F.nll_loss(log_probs, labels, reduction="mean")
        correct = (log_probs.argmax(dim=1) == labels).sum().float()
        red = average_losses_across_data_parallel_group([loss, correct])
        return loss * mpu.get_data_parallel_world_size(), {"lm loss": red[0], "correct_prediction_count": red[1]}

    def train_valid_datasets_provider():
        args, tok = get_args(), get_tokenizer()
        return (Dataset("training", args.train_data, tok, args.retriever_seq_length, evaluate=False),
                Dataset("validation", args.valid_data, tok, args.retriever_seq_length, evaluate=True))

    def model_provider(pre_process=True, post_process=True):
        args = get_args()
        print_rank_0("building retriever model for {} ...".format(args.task))
        return biencoder_model_provider(only_context_model=False, only_query_model=False,
                                        biencoder_shared_query_context_model=args.biencoder_shared_query_context_model,
                                        pre_process=pre_process, post_process=post_process,
                                        model_type=ModelType.encoder_or_decoder)

    def single_dataset_provider(datapath):
        args, tok = get_args(), get_tokenizer()
        name = datapath[0].split("/")[-1].split(".")[0]
        return Dataset(name, datapath, tok, args.retriever_seq_length, evaluate=True)

    def metrics_func_provider():
        return accuracy_func_provider(single_dataset_provider)

    finetune_utils.finetune(train_valid_datasets_provider, model_provider, ModelType.encoder_or_decoder,
                            forward_step=cross_entropy_forward_step,
                            end_of_epoch_callback_provider=metrics_func_provider, task_collate_fn=task_collate_fn)


def main():
    args = get_args()
    if args.task != "RET-FINETUNE-NQ":
        raise NotImplementedError("ORQA task {} is not implemented.".format(args.task))
    from tasks.orqa.supervised.data import NQSupervisedDataset
    orqa(NQSupervisedDataset)

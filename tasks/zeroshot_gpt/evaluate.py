"""Zero-shot evaluation of GPT-family models: WIKITEXT103 (perplexity) and LAMBADA (accuracy).

Parity: tasks/zeroshot_gpt/evaluate.py.  ``--model_name`` selects the family (the reference hard-codes GPTModel)."""
import math

import torch
import torch.distributed as dist

from megatron_llm_b200 import get_args, get_tokenizer, is_last_rank, print_rank_0
from megatron_llm_b200 import training
from megatron_llm_b200.checkpointing import load_checkpoint
from megatron_llm_b200.parallel import state as ps
from megatron_llm_b200.parallel.cross_entropy import vocab_parallel_cross_entropy
from megatron_llm_b200.parallel.p2p import recv_forward, send_forward
from megatron_llm_b200.utils import get_ltor_masks_and_position_ids, unwrap_model
from megatron_llm_b200.utils.device import current_device
from tasks import finetune_utils

from .datasets import build_dataset


def _get_model_provider(eval_metric):
    """loss needs vocab-parallel logits (fused CE); accuracy needs the full vocabulary for the arg-max."""
    if eval_metric not in ("loss", "accuracy"):
        raise NotImplementedError("output type for {} evaluation metric is not supported.".format(eval_metric))

    def model_provider(pre_process=True, post_process=True):
        from megatron_llm_b200.models import FalconModel, GPTModel, LlamaModel, MistralModel
        name = getattr(get_args(), "model_name", "gpt")
        cls = {"gpt": GPTModel, "falcon": FalconModel, "mistral": MistralModel}.get(name, LlamaModel)
        print_rank_0("building {} model ...".format(name))
        return cls(num_tokentypes=0, parallel_output=(eval_metric == "loss"), pre_process=pre_process,
                   post_process=post_process)
    return model_provider


def process_batch(batch):
    args, tok = get_args(), get_tokenizer()
    dev = current_device()
    loss_mask = batch["pad_mask"].long().to(dev).contiguous()
    tokens_ = batch["text"].long().to(dev).contiguous()
    labels, tokens = tokens_[:, 1:].contiguous(), tokens_[:, :-1].contiguous()
    attention_mask, _, position_ids = get_ltor_masks_and_position_ids(
        tokens, tok.eod, args.reset_position_ids, args.reset_attention_mask, args.eod_mask_loss)
    return tokens, labels, attention_mask, position_ids, loss_mask


def forward_step(batch, model, eval_metric):
    tokens, labels, attention_mask, position_ids, loss_mask = process_batch(batch)
    args = get_args()
    args.micro_batch_size = len(labels)
    input_tensor = recv_forward()
    unwrap_model(model).set_input_tensor(input_tensor)
    output = model(tokens, position_ids, attention_mask)
    send_forward(output)
    if not ps.is_pipeline_last_stage():
        return None
    if eval_metric == "loss":
        # model output is [b, s, V/tp]; the CE works on any leading shape
        losses = vocab_parallel_cross_entropy(output.contiguous(), labels.contiguous())
        return torch.sum(losses.float().view(-1) * loss_mask.view(-1).float())
    if eval_metric == "accuracy":
        correct = (torch.argmax(output, -1) == labels).float()
        correct[(1 - loss_mask).bool()] = 1           # only the target tokens can make a sample wrong
        return correct.prod(-1).sum()
    raise NotImplementedError("forward method for evaluation metric {} is not implemented.".format(eval_metric))


def evaluate(data_loader, model, eval_metric):
    args = get_args()
    model.eval()
    total = 0.0
    with torch.no_grad():
        for it, batch in enumerate(data_loader):
            if it % args.log_interval == 0:
                print_rank_0("> working on iteration: {}".format(it))
            out = forward_step(batch, model, eval_metric)
            if ps.is_pipeline_last_stage():
                dist.all_reduce(out, group=ps.get_data_parallel_group())
                total += out
    return total


def _evaluate_and_print_results(task, data_loader, model, eval_metric):
    output = evaluate(data_loader, model, eval_metric)
    string = " validation results on {} | ".format(task)
    result = {}
    if is_last_rank():
        output = float(output)
        if eval_metric == "loss":
            ds = data_loader.dataset
            val_loss = output / (ds.num_tokenized_tokens - 1)
            ratio = (ds.num_tokenized_tokens - 1) / (ds.num_original_tokens - 1)
            result = {"loss": val_loss, "ppl": math.exp(min(20, val_loss)),
                      "adjusted_ppl": math.exp(min(20, val_loss * ratio)), "token_ratio": ratio}
            string += "avg loss: {:.4E} | ppl: {:.4E} | adjusted ppl: {:.4E} | token ratio: {} |".format(
                val_loss, result["ppl"], result["adjusted_ppl"], ratio)
        else:
            n = len(data_loader.dataset)
            result = {"correct": output, "total": n, "accuracy": output / n}
            string += "number correct: {:.4E} | total examples: {:.4E} | avg accuracy: {:.4E}".format(
                output, n, output / n)
        print("-" * (len(string) + 1))
        print(string)
        print("-" * (len(string) + 1))
    return result


def main():
    args = get_args()
    if args.num_layers_per_virtual_pipeline_stage is not None:
        print("Interleaved pipeline schedule is not yet supported for text generation.")
        return
    if args.task == "LAMBADA":
        eval_metric = "accuracy"
    elif args.task == "WIKITEXT103":
        eval_metric = "loss"
    else:
        raise NotImplementedError("{} task is not implemented.".format(args.task))
    model = training.get_model(_get_model_provider(eval_metric), wrap_with_ddp=False, args=args)
    if args.load is not None:
        load_checkpoint(model, None, None)
    assert len(model) == 1, "Above condition should have caught this"
    dataset = build_dataset(args.task)
    loader = finetune_utils.build_data_loader(dataset, args.micro_batch_size, args.num_workers, drop_last=False)
    result = _evaluate_and_print_results(args.task, loader, model[0], eval_metric)
    print_rank_0("done :-)")
    return result

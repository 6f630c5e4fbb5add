"""Pre-train / fine-tune / instruction-tune GPT, Llama, Llama-2, CodeLlama, Falcon and Mistral.

Parity target: reference ``finetune.py`` (model_provider :26-62, get_batch :103-166, data_provider :169-193,
loss_func :201-218, forward_step :221-234, extra_args :242-254).  Same CLI; additions: ``--data_type synthetic``
(no files needed) and no host synchronisation per micro-batch (the reference all-reduces and ``.item()``s a token
counter for every micro-batch, :129-140).
"""
import datetime as dt
from functools import partial

import torch

from megatron_llm_b200 import get_args, get_counters, get_timers, get_tokenizer, print_rank_0
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.metrics import MetricInput, get_metric
from megatron_llm_b200.models import FalconModel, GPTModel, LlamaModel, MistralModel, ModelType
from megatron_llm_b200.parallel import broadcast_data
from megatron_llm_b200.training import pretrain
from megatron_llm_b200.utils import average_losses_across_data_parallel_group, get_ltor_masks_and_position_ids


def model_provider(pre_process: bool = True, post_process: bool = True):
    print_rank_0("Building model ...")
    args = get_args()
    name = args.model_name
    if name == "gpt":
        cls = GPTModel
    elif name == "falcon":
        cls = FalconModel
    elif name in {"llama", "llama2", "codellama"}:
        cls = partial(LlamaModel, version=1 if name == "llama" else 2)
    elif name == "mistral":
        cls = MistralModel
        if args.sliding_window_size != 4096:
            print_rank_0("Mistral uses sliding window attention (set sliding_window=4096)")
            args.sliding_window_size = 4096
    else:
        raise KeyError(f"Unkown model {name}")
    if isinstance(args.model_type, ModelType):
        model_type = args.model_type
    elif args.model_type in ("encoder_or_decoder", "encoder_and_decoder"):
        model_type = ModelType[args.model_type]
    else:
        raise KeyError(f"Unsupported model_type {args.model_type}")
    return cls(num_tokentypes=0, parallel_output=True, pre_process=pre_process, post_process=post_process,
               model_type=model_type)


def get_attention_mask_and_position_ids(data, attention_mask, build_mask=True):
    """(b, s) padding mask -> (b, 1, s, s) causal+padding boolean mask (True = masked) and position ids."""
    b, s = data.size()
    mask = None
    if build_mask:
        mask = attention_mask.unsqueeze(1).expand(b, s, s).to(data.device)
        mask = torch.tril(mask).view(b, 1, s, s) < 0.5
    position_ids = torch.arange(s, dtype=torch.long, device=data.device).unsqueeze(0).expand_as(data)
    return mask, position_ids


def get_batch(data_iterator):
    args = get_args()
    tokenizer = get_tokenizer()
    if args.data_type in ("gpt", "synthetic"):
        keys = ["text"]
    elif args.data_type == "instruction":
        keys = ["text", "attention_mask", "assistant_mask", "pad_mask"]
    else:
        raise KeyError(f"Unknown dataset type {args.data_type}")
    data = next(data_iterator) if data_iterator is not None else None
    data_b = broadcast_data(keys, data, torch.int64)
    tokens_ = data_b["text"]
    labels = tokens_[:, 1:].contiguous()
    tokens = tokens_[:, :-1].contiguous()
    # every DP rank processes an identically-shaped micro-batch: no collective / host sync needed
    get_counters()["tokens"] += tokens.numel() * args.data_parallel_size

    need_mask = not args.use_flash_attn   # the flash / tcgen05 attention path never reads the O(s^2) mask
    if args.data_type in ("gpt", "synthetic"):
        attention_mask, loss_mask, position_ids = get_ltor_masks_and_position_ids(
            tokens, tokenizer.eod, args.reset_position_ids, args.reset_attention_mask, args.eod_mask_loss,
            build_attention_mask=need_mask)
        return tokens, labels, loss_mask, attention_mask, position_ids
    attention_mask = data_b["attention_mask"][:, :-1]
    assistant_mask = data_b["assistant_mask"][:, 1:].to(tokens.device)
    pad_mask = data_b["pad_mask"][:, 1:].to(tokens.device)
    loss_mask = torch.full(labels.size(), args.scalar_loss_mask, dtype=torch.float, device=tokens.device)
    loss_mask[assistant_mask == 1] = 1.0
    loss_mask[pad_mask == 1] = 0.0
    attention_mask, position_ids = get_attention_mask_and_position_ids(tokens, attention_mask, build_mask=need_mask)
    return tokens, labels, loss_mask, attention_mask, position_ids


def data_provider(train_val_test_num_samples):
    args = get_args()
    print_rank_0("> building train, validation, and test datasets ...")
    if args.data_type == "synthetic":
        from megatron_llm_b200.data.synthetic import SyntheticGPTDataset
        ds = [SyntheticGPTDataset(max(n, 1), args.seq_length, args.padded_vocab_size, seed=args.seed + 17 * i)
              for i, n in enumerate(train_val_test_num_samples)]
        return tuple(ds)
    if args.data_type == "gpt":
        from megatron_llm_b200.data.gpt_dataset import build_train_valid_test_datasets as builder
    else:
        from megatron_llm_b200.data.instruction_dataset import build_train_valid_test_datasets as builder
    train_ds, valid_ds, test_ds = builder(
        data_prefix=args.data_path, data_impl=args.data_impl, splits_string=args.split,
        train_valid_test_num_samples=train_val_test_num_samples, seq_length=args.seq_length, seed=args.seed,
        skip_warmup=(not args.mmap_warmup), train_data_prefix=args.train_data_path,
        valid_data_prefix=args.valid_data_path, test_data_prefix=args.test_data_path)
    print_rank_0("> finished creating datasets ...")
    return train_ds, valid_ds, test_ds


def loss_func(is_training, batch, outputs):
    loss_mask = batch[2]
    losses, logits = outputs
    losses = losses.float()
    loss_mask = loss_mask.view(-1).float()
    loss = torch.sum(losses.view(-1) * loss_mask) / loss_mask.sum()
    averaged_loss = average_losses_across_data_parallel_group([loss])
    out_dict = {"lm loss": averaged_loss[0]}
    if not is_training:
        inputs = MetricInput(batch, logits, averaged_loss[0])
        args = get_args()
        names = list(args.metrics)
        if "all" in names:
            from megatron_llm_b200.metrics import METRICS
            names = list(METRICS)
        for metric in map(get_metric, names):
            out_dict.update(metric(inputs))
    return loss, out_dict


def run_from_batch(batch, model):
    tokens, labels, loss_mask, attention_mask, position_ids = batch
    output_tensor = model(tokens, position_ids, attention_mask, labels=labels)
    return output_tensor, partial(loss_func, model.training, batch)


def forward_step(data_iterator, model):
    timers = get_timers()
    timers("batch-generator", log_level=2).start()
    batch = get_batch(data_iterator)
    timers("batch-generator").stop()
    return run_from_batch(batch, model)


# the two halves, for schedules that capture the device work of a micro-batch in a CUDA graph (--cuda_graph_microbatch)
forward_step.get_batch = get_batch
forward_step.run = run_from_batch


def extra_args(parser):
    group = parser.add_argument_group(title="validation set")
    group.add_argument("--model_name", choices={"gpt", "llama", "falcon", "llama2", "codellama", "mistral"},
                       default="gpt")
    group.add_argument("--model_type", choices={"encoder_or_decoder", "encoder_and_decoder"},
                       default="encoder_or_decoder")
    group.add_argument("--data_type", choices={"gpt", "instruction", "synthetic"}, default="gpt")
    group.add_argument("--log_learning_rate_to_tensorboard", type=bool, default=True)
    group.add_argument("--log_loss_scale_to_tensorboard", type=bool, default=True)
    return parser


def main(args_list=None):
    initialize_megatron(extra_args, {"tokenizer_type": "GPT2BPETokenizer"}, args_list=args_list)
    args = get_args()
    collate_fn = None
    if args.data_type == "instruction":
        from megatron_llm_b200.data.instruction_dataset import instruction_collator
        collate_fn = instruction_collator
    pretrain(args, data_provider, model_provider, ModelType.encoder_or_decoder, forward_step, collate_fn=collate_fn)
    print(f"Done {dt.datetime.now(dt.timezone.utc)}")


if __name__ == "__main__":
    main()

"""Pre-train T5 (span corruption).  Parity: pretrain_t5.py (same CLI).

Pipeline parallelism: ranks below ``--pipeline_model_parallel_split_rank`` run encoder layers and send one tensor
(the encoder hidden state); ranks at/after it run decoder layers and send two (decoder hidden state + the complete
encoder output that every decoder layer cross-attends to); the schedules accumulate the encoder-output gradient
across those skip connections."""
from functools import partial

import torch

from megatron_llm_b200 import get_args, get_timers, print_rank_0
from megatron_llm_b200.data.dataset_utils import build_train_valid_test_datasets
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType, T5Model
from megatron_llm_b200.parallel.data import broadcast_data
from megatron_llm_b200.training import pretrain
from megatron_llm_b200.utils import average_losses_across_data_parallel_group


def model_provider(pre_process=True, post_process=True, add_encoder=True, add_decoder=True):
    print_rank_0("building T5 model ...")
    return T5Model(num_tokentypes=0, parallel_output=True, pre_process=pre_process, post_process=post_process,
                   add_encoder=add_encoder, add_decoder=add_decoder, model_type=ModelType.encoder_and_decoder)


def get_batch(data_iterator):
    keys = ["text_enc", "text_dec", "labels", "loss_mask", "enc_mask", "dec_mask", "enc_dec_mask"]
    data = next(data_iterator) if data_iterator is not None else None
    d = broadcast_data(keys, data, torch.int64)
    return (d["text_enc"].long(), d["text_dec"].long(), d["loss_mask"].float(), d["labels"].long(),
            d["enc_mask"] < 0.5, d["dec_mask"] < 0.5, d["enc_dec_mask"] < 0.5)


def loss_func(loss_mask, output_tensor):
    lm_loss = torch.sum(output_tensor.float().view(-1) * loss_mask.reshape(-1)) / loss_mask.sum()
    avg = average_losses_across_data_parallel_group([lm_loss])
    return lm_loss, {"lm loss": avg[0]}


def forward_step(data_iterator, model):
    timers = get_timers()
    timers("batch generator", log_level=2).start()
    tokens_enc, tokens_dec, loss_mask, lm_labels, enc_mask, dec_mask, enc_dec_mask = get_batch(data_iterator)
    timers("batch generator").stop()
    output_tensor = model(tokens_enc, tokens_dec, enc_mask, dec_mask, enc_dec_mask, tokentype_ids=None,
                          lm_labels=lm_labels.clamp_min(0))
    return output_tensor, partial(loss_func, loss_mask)


def train_valid_test_datasets_provider(train_val_test_num_samples):
    args = get_args()
    print_rank_0("> building train, validation, and test datasets for T5 ...")
    ds = build_train_valid_test_datasets(
        data_prefix=args.data_path, data_impl=args.data_impl, splits_string=args.split,
        train_valid_test_num_samples=train_val_test_num_samples, max_seq_length=args.encoder_seq_length,
        max_seq_length_dec=args.decoder_seq_length, masked_lm_prob=args.mask_prob,
        short_seq_prob=args.short_seq_prob, seed=args.seed, skip_warmup=(not args.mmap_warmup), dataset_type="t5")
    print_rank_0("> finished creating T5 datasets ...")
    return ds


def main(args_list=None):
    initialize_megatron(extra_args_provider=None, args_defaults={"tokenizer_type": "BertWordPieceLowerCase"},
                        args_list=args_list)
    pretrain(get_args(), train_valid_test_datasets_provider, model_provider, ModelType.encoder_and_decoder,
             forward_step)


if __name__ == "__main__":
    main()

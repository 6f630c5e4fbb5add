"""BERT pre-training dataset: sentence-pair samples with whole-word masking.  Parity: megatron/data/bert_dataset.py."""
from __future__ import annotations

import numpy as np
import torch

from .dataset_utils import (create_masked_lm_predictions, create_tokens_and_tokentypes, get_a_and_b_segments,
                            get_samples_mapping, pad_and_convert_to_numpy, truncate_segments)


class BertDataset(torch.utils.data.Dataset):
    def __init__(self, name, indexed_dataset, data_prefix, num_epochs, max_num_samples, masked_lm_prob,
                 max_seq_length, short_seq_prob, seed, binary_head, tokenizer=None):
        self.name, self.seed, self.masked_lm_prob = name, seed, masked_lm_prob
        self.max_seq_length, self.binary_head = max_seq_length, binary_head
        self.indexed_dataset = indexed_dataset
        # 3 = [CLS] + 2 x [SEP]
        self.samples_mapping = get_samples_mapping(indexed_dataset, data_prefix, num_epochs, max_num_samples,
                                                   max_seq_length - 3, short_seq_prob, seed, name, binary_head)
        if tokenizer is None:
            from ..global_vars import get_tokenizer
            tokenizer = get_tokenizer()
        self.vocab_id_to_token_dict = tokenizer.inv_vocab
        self.vocab_id_list = list(self.vocab_id_to_token_dict.keys())
        self.cls_id, self.sep_id, self.mask_id, self.pad_id = tokenizer.cls, tokenizer.sep, tokenizer.mask, tokenizer.pad

    def __len__(self):
        return self.samples_mapping.shape[0]

    def __getitem__(self, idx):
        start, end, seq_length = self.samples_mapping[idx]
        sample = [self.indexed_dataset[i] for i in range(start, end)]
        rng = np.random.RandomState(seed=((self.seed + idx) % 2 ** 32))   # numpy: randint upper bound exclusive
        return build_training_sample(sample, seq_length, self.max_seq_length, self.vocab_id_list,
                                     self.vocab_id_to_token_dict, self.cls_id, self.sep_id, self.mask_id, self.pad_id,
                                     self.masked_lm_prob, rng, self.binary_head)


def build_training_sample(sample, target_seq_length, max_seq_length, vocab_id_list, vocab_id_to_token_dict, cls_id,
                          sep_id, mask_id, pad_id, masked_lm_prob, np_rng, binary_head):
    """``sample``: list of sentences (token-id arrays) -> dict of padded numpy arrays."""
    if binary_head:
        assert len(sample) > 1
    assert target_seq_length <= max_seq_length
    if binary_head:
        tokens_a, tokens_b, is_next_random = get_a_and_b_segments(sample, np_rng)
    else:
        tokens_a, tokens_b, is_next_random = [t for s in sample for t in s], [], False
    truncated = truncate_segments(tokens_a, tokens_b, len(tokens_a), len(tokens_b), target_seq_length, np_rng)
    tokens, tokentypes = create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id)
    tokens, positions, labels, _, _ = create_masked_lm_predictions(
        tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
        masked_lm_prob * target_seq_length, np_rng)
    tokens_np, types_np, labels_np, padding_mask, loss_mask = pad_and_convert_to_numpy(
        tokens, tokentypes, positions, labels, pad_id, max_seq_length)
    return {"text": tokens_np, "types": types_np, "labels": labels_np, "is_random": int(is_next_random),
            "loss_mask": loss_mask, "padding_mask": padding_mask, "truncated": int(truncated)}

"""Google Natural Questions (DPR tsv: question \\t python-list-of-answers) for retriever evaluation
(parity: tasks/orqa/unsupervised/nq.py)."""
import ast
import csv
from collections import OrderedDict

import numpy as np
import torch
from torch.utils.data import BatchSampler, DataLoader, Dataset

from megatron_llm_b200 import get_args, get_tokenizer, print_rank_0
from megatron_llm_b200.data.biencoder_dataset_utils import make_attention_mask
from megatron_llm_b200.utils.device import current_device


def get_nq_dataset(qa_data, split):
    args = get_args()
    return NQDataset("Google NQ {} Split".format(split), "Google Natural Questions", qa_data, get_tokenizer(),
                     args.retriever_seq_length)


def process_nq_batch(batch):
    dev = current_device()
    return (batch["token_ids"].long().to(dev), (batch["token_mask"] < 0.5).to(dev), batch["token_types"].long().to(dev),
            batch["seq_len"].long().to(dev), batch["reference"])


class CustomDataLoader(DataLoader):
    """Collates the tensor fields and keeps ``reference`` (lists of answer strings) as a python list."""

    def __init__(self, dataset, eval=False, **kwargs):
        kwargs.setdefault("collate_fn", self._collate_fn)
        self.eval = eval
        super().__init__(dataset, **kwargs)

    @staticmethod
    def _collate_fn(batch_data):
        out = OrderedDict()
        for d in batch_data:
            for k, v in d.items():
                out.setdefault(k, []).append(v)
        assert len(out) == 5
        for k in ("token_ids", "token_mask", "token_types", "seq_len"):
            out[k] = torch.from_numpy(np.asarray(out[k])).long()
        return out


def get_one_epoch_nq_dataloader(dataset, micro_batch_size=None):
    """Sequential, not distributed: every rank encodes every question."""
    args = get_args()
    mbs = micro_batch_size if micro_batch_size is not None else args.micro_batch_size
    sampler = BatchSampler(torch.utils.data.SequentialSampler(dataset), batch_size=mbs, drop_last=False)
    return CustomDataLoader(dataset, batch_sampler=sampler, num_workers=args.num_workers,
                            pin_memory=torch.cuda.is_available())


def build_tokens_types_paddings_from_ids(src_ids, max_seq_length, cls_id, sep_id, pad_id):
    ids = ([cls_id] + list(src_ids))[:max_seq_length - 1] + [sep_id]
    n = len(ids)
    pad = max_seq_length - n
    return ids + [pad_id] * pad, [0] * n + [pad_id] * pad, n


def build_tokens_types_paddings_from_text(src_text, tokenizer, max_seq_length):
    return build_tokens_types_paddings_from_ids(tokenizer.tokenize(src_text), max_seq_length, tokenizer.cls,
                                                tokenizer.sep, tokenizer.pad)


def build_sample(token_ids, token_types, num_tokens, reference):
    token_ids = np.array(token_ids, dtype=np.int64)
    return {"token_ids": token_ids, "token_mask": make_attention_mask(token_ids, token_ids),
            "token_types": np.array(token_types, dtype=np.int64), "seq_len": num_tokens, "reference": reference}


class NQDataset(Dataset):
    def __init__(self, task_name, dataset_name, datapath, tokenizer, max_seq_length):
        self.task_name, self.dataset_name = task_name, dataset_name
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        print_rank_0(" > building {} dataset for {}:".format(task_name, dataset_name))
        print_rank_0(datapath)
        self.samples = self.process_samples_from_single_path(datapath)
        print_rank_0("  >> total number of samples: {}".format(len(self.samples)))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        raw = self.samples[idx]
        ids, types, n = build_tokens_types_paddings_from_text(raw["question"], self.tokenizer, self.max_seq_length)
        return build_sample(ids, types, n, raw["answers"])

    @staticmethod
    def process_samples_from_single_path(filename):
        print_rank_0(" > Processing {} ...".format(filename))
        samples = []
        with open(filename, "r") as f:
            for row in csv.reader(f, delimiter="\t"):
                samples.append({"question": row[0], "answers": ast.literal_eval(row[1])})
        print_rank_0(" >> processed {} samples.".format(len(samples)))
        return samples

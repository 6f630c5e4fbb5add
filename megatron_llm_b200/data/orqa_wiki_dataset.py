"""Wikipedia evidence dataset (DPR tsv: id, text, title) for open-retrieval QA.
Parity: megatron/data/orqa_wiki_dataset.py."""
from __future__ import annotations

import csv
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from ..parallel.data import broadcast_data
from ..utils import print_rank_0
from .biencoder_dataset_utils import make_attention_mask


def get_open_retrieval_wiki_dataset():
    from ..global_vars import get_args, get_tokenizer
    args = get_args()
    return OpenRetrievalEvidenceDataset("2018 Wikipedia from DPR codebase", "evidence", args.evidence_data_path,
                                        get_tokenizer(), args.retriever_seq_length)


def get_open_retrieval_batch(data_iterator):
    keys = ["row_id", "context", "context_mask", "context_types", "context_pad_mask"]
    data = None if data_iterator is None else next(data_iterator)
    d = broadcast_data(keys, data, torch.int64)
    return (d["row_id"].long(), d["context"].long(), d["context_mask"] < 0.5, d["context_types"].long(),
            d["context_pad_mask"].long())


def build_tokens_types_paddings_from_ids(text_ids, max_seq_length, cls_id, sep_id, pad_id):
    """[CLS] text [SEP], trimmed to ``max_seq_length`` and padded; token types all 0."""
    ids = ([cls_id] + list(text_ids))[:max_seq_length - 1] + [sep_id]
    n = len(ids)
    pad = max_seq_length - n
    types = [0] * n + [pad_id] * pad
    ids = ids + [pad_id] * pad
    return ids, types, np.array([1] * n + [0] * pad, dtype=np.int64)


def build_tokens_types_paddings_from_text(row, tokenizer, max_seq_length):
    ids = tokenizer.tokenize(row["title"]) + [tokenizer.sep] + tokenizer.tokenize(row["text"])
    return build_tokens_types_paddings_from_ids(ids, max_seq_length, tokenizer.cls, tokenizer.sep, tokenizer.pad)


def build_sample(row_id, context_ids, context_types, context_pad_mask):
    context_ids = np.array(context_ids, dtype=np.int64)
    return {"row_id": row_id, "context": context_ids, "context_mask": make_attention_mask(context_ids, context_ids),
            "context_types": np.array(context_types, dtype=np.int64), "context_pad_mask": context_pad_mask}


class OpenRetrievalEvidenceDataset(Dataset):
    def __init__(self, task_name, dataset_name, datapath, tokenizer, max_seq_length):
        self.task_name, self.dataset_name = task_name, dataset_name
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        print_rank_0(f" > building {task_name} dataset for {dataset_name}:\n{datapath}")
        self.samples, self.id2text = self.process_samples_from_single_path(datapath)
        from ..global_vars import get_args
        args = get_args()
        if args.sample_rate < 1:
            k = int(len(self.samples) * args.sample_rate)
            self.samples = random.sample(self.samples, k)
        print_rank_0(f"  >> total number of samples: {len(self.samples)}")

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        row = self.samples[idx]
        ids, types, pad_mask = build_tokens_types_paddings_from_text(row, self.tokenizer, self.max_seq_length)
        return build_sample(row["doc_id"], ids, types, pad_mask)

    @staticmethod
    def process_samples_from_single_path(filename):
        print_rank_0(f" > Processing {filename} ...")
        rows, id2text = [], {}
        with open(filename) as f:
            reader = csv.reader(f, delimiter="\t")
            next(reader, None)                      # header
            for doc_id, text, title in (r[:3] for r in reader):
                doc_id = int(doc_id)
                assert doc_id not in id2text
                rows.append({"doc_id": doc_id, "text": text, "title": title})
                id2text[doc_id] = (text, title)
                if len(rows) % 100000 == 0:
                    print_rank_0(f"  > processed {len(rows)} rows so far ...")
        print_rank_0(f" >> processed {len(rows)} samples.")
        return rows, id2text

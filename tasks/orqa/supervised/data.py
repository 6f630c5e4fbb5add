"""Supervised retriever data: DPR-format Natural Questions json (question, answers, positive_ctxs,
negative_ctxs, hard_negative_ctxs).  Parity: tasks/orqa/supervised/data.py."""
import json
import random
from abc import ABC, abstractmethod

import numpy as np
from torch.utils.data import Dataset

from megatron_llm_b200 import get_args, print_rank_0
from megatron_llm_b200.data.biencoder_dataset_utils import make_attention_mask


def build_tokens_types_paddings_from_ids(text_ids, max_seq_length, cls_id, sep_id, pad_id):
    """[CLS] text [SEP] trimmed / padded to ``max_seq_length``; returns (ids, token types, pad mask)."""
    ids = ([cls_id] + list(text_ids))[:max_seq_length - 1] + [sep_id]
    n = len(ids)
    pad = max_seq_length - n
    return ids + [pad_id] * pad, [0] * n + [pad_id] * pad, np.array([1] * n + [0] * pad, dtype=np.int64)


def _context_ids(context, tokenizer):
    return tokenizer.tokenize(context["title"]) + [tokenizer.sep] + tokenizer.tokenize(context["text"])


def build_token_types_from_context_list(ctx_list, tokenizer, max_seq_length):
    ids_l, types_l = [], []
    for ctx in ctx_list:
        ids, types, _ = build_tokens_types_paddings_from_ids(_context_ids(ctx, tokenizer), max_seq_length,
                                                             tokenizer.cls, tokenizer.sep, tokenizer.pad)
        ids_l.append(ids)
        types_l.append(types)
    return ids_l, types_l


def build_tokens_types_paddings_from_text(query, context, tokenizer, max_seq_length):
    q = build_tokens_types_paddings_from_ids(tokenizer.tokenize(query), max_seq_length, tokenizer.cls, tokenizer.sep,
                                             tokenizer.pad)
    c = build_tokens_types_paddings_from_ids(_context_ids(context, tokenizer), max_seq_length, tokenizer.cls,
                                             tokenizer.sep, tokenizer.pad)
    return (*q, *c)


def build_sample(query_ids, query_types, query_pad_mask, ctx_ids, ctx_types, ctx_pad_mask, answers,
                 neg_ctx_id_list=None, neg_ctx_types_list=None, include_neg=False):
    query_ids, ctx_ids = np.array(query_ids, dtype=np.int64), np.array(ctx_ids, dtype=np.int64)
    sample = {"query": query_ids, "query_mask": make_attention_mask(query_ids, query_ids),
              "query_types": np.array(query_types, dtype=np.int64), "query_pad_mask": query_pad_mask,
              "context": ctx_ids, "context_mask": make_attention_mask(ctx_ids, ctx_ids),
              "context_types": np.array(ctx_types, dtype=np.int64), "context_pad_mask": ctx_pad_mask,
              "reference": answers}
    if include_neg:
        neg = np.array(neg_ctx_id_list, dtype=np.int64).reshape(-1, ctx_ids.shape[0])
        sample["neg_context"] = neg
        sample["neg_context_types"] = np.array(neg_ctx_types_list, dtype=np.int64).reshape(neg.shape)
        sample["neg_context_mask"] = np.array([make_attention_mask(i, i) for i in neg], dtype=np.int64).reshape(
            neg.shape[0], neg.shape[1], neg.shape[1])
    return sample


class OpenRetrievalAbstractDataset(ABC, Dataset):
    def __init__(self, task_name, dataset_name, datapaths, tokenizer, max_seq_length, evaluate=False):
        args = get_args()
        self.evaluate = evaluate
        self.val_av_rank_hard_neg, self.val_av_rank_other_neg = args.val_av_rank_hard_neg, args.val_av_rank_other_neg
        self.train_with_neg, self.train_hard_neg = args.train_with_neg, args.train_hard_neg
        self.task_name, self.dataset_name = task_name, dataset_name
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        print_rank_0(" > building {} dataset for {}:".format(task_name, dataset_name))
        print_rank_0("  > paths: " + " ".join(datapaths))
        self.samples = []
        for path in datapaths:
            self.samples.extend(self.process_samples_from_single_path(path))
        if args.sample_rate < 1:
            self.samples = random.sample(self.samples, int(len(self.samples) * args.sample_rate))
        print_rank_0("  >> total number of samples: {}".format(len(self.samples)))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        raw = self.samples[idx]
        q_ids, q_types, q_pad, c_ids, c_types, c_pad = build_tokens_types_paddings_from_text(
            raw["question"], raw["pos_context"], self.tokenizer, self.max_seq_length)
        neg_ids = neg_types = None
        if self.evaluate:      # fixed pools for the average-rank metric
            negs = raw["negative_context"][:self.val_av_rank_other_neg] + \
                raw["hard_negative_context"][:self.val_av_rank_hard_neg]
            neg_ids, neg_types = build_token_types_from_context_list(negs, self.tokenizer, self.max_seq_length)
        elif self.train_with_neg:   # hard negatives first, topped up with random ones
            hard, other = list(raw["hard_negative_context"]), list(raw["negative_context"])
            random.shuffle(hard)
            random.shuffle(other)
            negs = hard[:self.train_hard_neg]
            negs += other[:self.train_hard_neg - len(negs)]
            neg_ids, neg_types = build_token_types_from_context_list(negs, self.tokenizer, self.max_seq_length)
        return build_sample(q_ids, q_types, q_pad, c_ids, c_types, c_pad, raw["answers"], neg_ids, neg_types,
                            include_neg=self.evaluate or self.train_with_neg)

    @staticmethod
    @abstractmethod
    def process_samples_from_single_path(filename):
        """file -> list of {'question', 'pos_context', 'hard_negative_context', 'negative_context', 'answers'}."""


def normalize_question(question):
    return question[:-1] if question.endswith("?") else question


class NQSupervisedDataset(OpenRetrievalAbstractDataset):
    def __init__(self, name, datapaths, tokenizer, max_seq_length, evaluate=False):
        super().__init__("natural_questions_ret", name, datapaths, tokenizer, max_seq_length, evaluate=evaluate)

    @staticmethod
    def process_samples_from_single_path(filename):
        print_rank_0(" > Processing {} ...".format(filename))
        with open(filename, "r", encoding="utf-8") as f:
            data = json.load(f)
        samples = [{"question": normalize_question(row["question"]), "pos_context": row["positive_ctxs"][0],
                    "hard_negative_context": list(row.get("hard_negative_ctxs") or []),
                    "negative_context": list(row.get("negative_ctxs") or []), "answers": row["answers"]}
                   for row in data]
        print_rank_0(" >> processed {} samples.".format(len(samples)))
        return samples

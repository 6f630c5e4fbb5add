"""Sample construction shared by the classification / multiple-choice tasks (parity: tasks/data_utils.py)."""
import re

import numpy as np


def clean_text(text):
    """Collapse whitespace and glue sentence-final dots to the preceding word."""
    text = re.sub(r"\s+", " ", text.replace("\n", " "))
    for _ in range(3):
        text = text.replace(" . ", ". ")
    return text


def build_sample(ids, types, paddings, label, unique_id):
    return {"text": np.array(ids, dtype=np.int64), "types": np.array(types, dtype=np.int64),
            "padding_mask": np.array(paddings, dtype=np.int64), "label": int(label), "uid": int(unique_id)}


def build_tokens_types_paddings_from_text(text_a, text_b, tokenizer, max_seq_length):
    a = tokenizer.tokenize(text_a)
    b = tokenizer.tokenize(text_b) if text_b is not None else None
    return build_tokens_types_paddings_from_ids(a, b, max_seq_length, tokenizer.cls, tokenizer.sep, tokenizer.pad)


def build_tokens_types_paddings_from_ids(text_a_ids, text_b_ids, max_seq_length, cls_id, sep_id, pad_id):
    """[CLS] A [SEP] (B [SEP]); trimmed to ``max_seq_length`` (always ending in [SEP] when trimmed) and padded."""
    ids = [cls_id] + list(text_a_ids) + [sep_id]
    types = [0] * len(ids)
    if text_b_ids is not None:
        ids += list(text_b_ids)
        types += [1] * len(text_b_ids)
    trimmed = len(ids) >= max_seq_length
    if trimmed:
        ids, types = ids[:max_seq_length - 1], types[:max_seq_length - 1]
    if text_b_ids is not None or trimmed:
        ids.append(sep_id)
        types.append(0 if text_b_ids is None else 1)
    n = len(ids)
    pad = max_seq_length - n
    return ids + [pad_id] * pad, types + [pad_id] * pad, [1] * n + [0] * pad

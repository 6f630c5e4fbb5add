"""Dataset utilities shared by the GPT / BERT / T5 / ICT datasets (parity: megatron/data/dataset_utils.py)."""
from __future__ import annotations

import collections
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ..utils import print_rank_0

DSET_TYPE_BERT = "standard_bert"
DSET_TYPE_ICT = "ict"
DSET_TYPE_T5 = "t5"
DSET_TYPES = [DSET_TYPE_BERT, DSET_TYPE_ICT, DSET_TYPE_T5]


def get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples):
    """``data_prefix = [w1, prefix1, w2, prefix2, ...]`` -> (prefixes, normalised weights, per-dataset sample
    counts inflated by 0.5% so an uneven blend never runs dry)."""
    assert len(data_prefix) % 2 == 0
    weights = [float(w) for w in data_prefix[0::2]]
    prefixes = [p.strip() for p in data_prefix[1::2]]
    total = sum(weights)
    assert total > 0.0
    weights = [w / total for w in weights]
    if isinstance(train_valid_test_num_samples, list):
        per_ds = [[int(math.ceil(v * w * 1.005)) for v in train_valid_test_num_samples] for w in weights]
    else:
        per_ds = [int(math.ceil(train_valid_test_num_samples * w * 1.005)) for w in weights]
    return prefixes, weights, per_ds


def get_train_valid_test_split_(splits_string, size):
    """'969,30,1' or '90/5/5' -> four document boundaries [0, a, b, size]."""
    sep = "," if "," in splits_string else ("/" if "/" in splits_string else None)
    parts = [float(s) for s in splits_string.split(sep)] if sep else [float(splits_string)]
    parts = (parts + [0.0, 0.0, 0.0])[:3]
    total = sum(parts)
    assert total > 0.0
    bounds = [0]
    for frac in parts:
        bounds.append(bounds[-1] + int(round(frac / total * float(size))))
    diff = bounds[-1] - size
    bounds = [bounds[0]] + [b - diff for b in bounds[1:]]
    assert len(bounds) == 4 and bounds[-1] == size
    return bounds


# =====================================================================================================
# BERT / T5 sample construction (parity: dataset_utils.py:95-418)
# =====================================================================================================
MaskedLmInstance = collections.namedtuple("MaskedLmInstance", ["index", "label"])


def compile_helper():
    """Build the C++ dataset helpers (the reference shells out to ``make``; here: ops/build.py)."""
    from ..ops.build import build_helpers
    build_helpers()


def get_a_and_b_segments(sample, np_rng):
    """Split a list of sentences into segments A | B at a random sentence boundary; swap them half the time
    (``is_next_random`` = swapped, the next-sentence-prediction label)."""
    n = len(sample)
    assert n > 1, "make sure each sample has at least two sentences."
    cut = np_rng.randint(1, n) if n >= 3 else 1
    tokens_a = [t for s in sample[:cut] for t in s]
    tokens_b = [t for s in sample[cut:] for t in s]
    swapped = bool(np_rng.random() < 0.5)
    if swapped:
        tokens_a, tokens_b = tokens_b, tokens_a
    return tokens_a, tokens_b, swapped


def truncate_segments(tokens_a, tokens_b, len_a, len_b, max_num_tokens, np_rng):
    """Drop tokens (front or back, at random) from the longer segment until the pair fits.  In place."""
    assert len_a > 0
    if len_a + len_b <= max_num_tokens:
        return False
    while len_a + len_b > max_num_tokens:
        if len_a > len_b:
            len_a -= 1
            victim = tokens_a
        else:
            len_b -= 1
            victim = tokens_b
        if np_rng.random() < 0.5:
            del victim[0]
        else:
            victim.pop()
    return True


def create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id):
    """[CLS] A [SEP] (B [SEP]) with token types 0 for the A part and 1 for the B part."""
    tokens = [cls_id] + list(tokens_a) + [sep_id]
    types = [0] * len(tokens)
    if tokens_b:
        tokens += list(tokens_b) + [sep_id]
        types += [1] * (len(tokens_b) + 1)
    return tokens, types


def is_start_piece(piece):
    """WordPiece continuation pieces start with '##'."""
    return not piece.startswith("##")


def _draw_ngram_length(np_rng, limit, max_ngrams, pvals, geometric_dist):
    if geometric_dist:       # SpanBERT: geometric(p=0.2) clipped
        return min(int(np_rng.geometric(0.2)), max_ngrams)
    p = pvals[:limit]
    return int(np_rng.choice(np.arange(1, limit + 1), p=p / p.sum()))


def create_masked_lm_predictions(tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id,
                                 mask_id, max_predictions_per_seq, np_rng, max_ngrams=3, do_whole_word_mask=True,
                                 favor_longer_ngram=False, do_permutation=False, geometric_dist=False,
                                 masking_style="bert"):
    """Whole-word n-gram masking.  Returns (output tokens, masked positions, masked labels, token_boundary flags,
    masked spans).  ``bert`` style: 80% [MASK] / 10% keep / 10% random id; ``t5`` style: always the mask id (the
    caller replaces each span by a sentinel)."""
    if masking_style not in ("bert", "t5"):
        raise ValueError("invalid value of masking style")
    # group sub-word pieces into words
    words, boundary = [], [0] * len(tokens)
    for i, tok in enumerate(tokens):
        if tok == cls_id or tok == sep_id:
            boundary[i] = 1
            continue
        start = is_start_piece(vocab_id_to_token_dict[tok])
        if do_whole_word_mask and words and not start:
            words[-1].append(i)
        else:
            words.append([i])
            if start:
                boundary[i] = 1
    output = list(tokens)
    if masked_lm_prob == 0:
        return output, [], [], boundary
    budget = int(min(max_predictions_per_seq, max(1, int(round(len(tokens) * masked_lm_prob)))))
    pvals = 1.0 / np.arange(1, max_ngrams + 1)
    pvals /= pvals.sum()
    if favor_longer_ngram:
        pvals = pvals[::-1]

    def pick_spans(taken_already, budget_left):
        """Random word starts, n-gram length per start, shrink to fit the budget, never overlap."""
        chosen, taken = [], set(taken_already)
        used = 0
        for w in np_rng.permutation(len(words)):
            if used >= budget_left:
                break
            limit = min(max_ngrams, len(words) - w)
            n = min(_draw_ngram_length(np_rng, limit, max_ngrams, pvals, geometric_dist), limit)
            while n > 0:
                idx = [i for word in words[w:w + n] for i in word]
                if used + len(idx) <= budget_left:
                    break
                n -= 1
            if n == 0 or any(i in taken for i in idx):
                continue
            taken.update(idx)
            used += len(idx)
            chosen.append(idx)
        return chosen, taken

    spans, covered = pick_spans((), budget)
    masked, masked_spans = [], []
    for idx in spans:
        for i in idx:
            if masking_style == "t5":
                new = mask_id
            else:
                r = np_rng.random()
                new = mask_id if r < 0.8 else (tokens[i] if r < 0.9 else
                                               vocab_id_list[np_rng.randint(0, len(vocab_id_list))])
            output[i] = new
            masked.append(MaskedLmInstance(index=i, label=tokens[i]))
        masked_spans.append(MaskedLmInstance(index=idx, label=[tokens[i] for i in idx]))
    assert len(masked) <= budget
    if do_permutation:
        perm_spans, _ = pick_spans(covered, budget)
        select = sorted(i for idx in perm_spans for i in idx)
        shuffled = list(select)
        np_rng.shuffle(shuffled)
        before = list(output)
        for src, tgt in zip(select, shuffled):
            output[src] = before[tgt]
            masked.append(MaskedLmInstance(index=src, label=before[src]))
    masked.sort(key=lambda m: m.index)
    masked_spans.sort(key=lambda m: m.index[0])
    return output, [m.index for m in masked], [m.label for m in masked], boundary, masked_spans


def pad_and_convert_to_numpy(tokens, tokentypes, masked_positions, masked_labels, pad_id, max_seq_length):
    n = len(tokens)
    pad = max_seq_length - n
    assert pad >= 0 and len(tokentypes) == n and len(masked_positions) == len(masked_labels)
    tokens_np = np.array(list(tokens) + [pad_id] * pad, dtype=np.int64)
    types_np = np.array(list(tokentypes) + [pad_id] * pad, dtype=np.int64)
    padding_mask = np.array([1] * n + [0] * pad, dtype=np.int64)
    labels = np.full(max_seq_length, -1, dtype=np.int64)
    loss_mask = np.zeros(max_seq_length, dtype=np.int64)
    for pos, lab in zip(masked_positions, masked_labels):
        assert pos < n
        labels[pos] = lab
        loss_mask[pos] = 1
    return tokens_np, types_np, labels, padding_mask, loss_mask


# =====================================================================================================
# dataset builders (parity: dataset_utils.py:421-729)
# =====================================================================================================
def get_indexed_dataset_(data_prefix, data_impl, skip_warmup):
    from .indexed_dataset import make_dataset
    print_rank_0(" > building dataset index ...")
    t0 = time.time()
    ds = make_dataset(data_prefix, data_impl, skip_warmup)
    assert ds.sizes.shape[0] == ds.doc_idx[-1]
    print_rank_0(f" > finished creating indexed dataset in {time.time() - t0:4f} seconds")
    print_rank_0(f" > indexed dataset stats:\n    number of documents: {ds.doc_idx.shape[0] - 1}\n"
                 f"    number of sentences: {ds.sizes.shape[0]}")
    return ds


def _sync_after_index_build():
    """Every rank that builds datasets waits until rank 0 has written the index file.  Only the tensor-parallel rank 0
    of every (DP, PP) coordinate calls the dataset providers (training.py broadcasts the result flags over TP), so a
    world barrier would dead-lock under TP > 1: like the reference (dataset_utils.py:709-717) all-reduce a counter over
    the data-parallel and the pipeline-parallel group, which together connect exactly those ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    from ..parallel import state as ps
    from ..utils.device import current_device
    if not ps.model_parallel_is_initialized():
        dist.barrier()
        return
    counts = torch.ones(1, dtype=torch.long, device=current_device())
    dist.all_reduce(counts, group=ps.get_data_parallel_group())
    dist.all_reduce(counts, group=ps.get_pipeline_model_parallel_group())
    assert counts[0].item() == dist.get_world_size() // dist.get_world_size(group=ps.get_tensor_model_parallel_group())


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_samples_mapping(indexed_dataset, data_prefix, num_epochs, max_num_samples, max_seq_length, short_seq_prob,
                        seed, name, binary_head):
    """[n, 3] int array of (first sentence, last sentence + 1, target length), cached as ``.npy`` next to the data."""
    if not num_epochs:
        if not max_num_samples:
            raise ValueError("Need to specify either max_num_samples or num_epochs")
        num_epochs = np.iinfo(np.int32).max - 1
    if not max_num_samples:
        max_num_samples = np.iinfo(np.int64).max - 1
    fname = f"{data_prefix}_{name}_indexmap"
    if num_epochs != np.iinfo(np.int32).max - 1:
        fname += f"_{num_epochs}ep"
    if max_num_samples != np.iinfo(np.int64).max - 1:
        fname += f"_{max_num_samples}mns"
    fname += f"_{max_seq_length}msl_{short_seq_prob:0.2f}ssp_{seed}s.npy"
    if _rank() == 0 and not os.path.isfile(fname):
        print(f" > WARNING: could not find index map file {fname}, building the indices on rank 0 ...")
        assert indexed_dataset.doc_idx.dtype == np.int64 and indexed_dataset.sizes.dtype == np.int32
        from . import helpers
        t0 = time.time()
        mapping = helpers.build_mapping(indexed_dataset.doc_idx, indexed_dataset.sizes, num_epochs, max_num_samples,
                                        max_seq_length, short_seq_prob, seed, True, 2 if binary_head else 1)
        np.save(fname, mapping, allow_pickle=True)
        print_rank_0(f" > saved the index mapping in {fname} ({time.time() - t0:4f} s)")
    _sync_after_index_build()
    mapping = np.load(fname, allow_pickle=True, mmap_mode="r")
    print_rank_0(f"    total number of samples: {mapping.shape[0]}")
    return mapping


def build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples,
                                    max_seq_length, masked_lm_prob, short_seq_prob, seed, skip_warmup,
                                    binary_head=False, max_seq_length_dec=None, dataset_type="standard_bert"):
    from .blendable_dataset import BlendableDataset
    if len(data_prefix) == 1:
        return _build_train_valid_test_datasets(data_prefix[0], data_impl, splits_string,
                                                train_valid_test_num_samples, max_seq_length, masked_lm_prob,
                                                short_seq_prob, seed, skip_warmup, binary_head, max_seq_length_dec,
                                                dataset_type=dataset_type)
    prefixes, weights, per_ds = get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples)
    columns = ([], [], [])
    for prefix, n in zip(prefixes, per_ds):
        parts = _build_train_valid_test_datasets(prefix, data_impl, splits_string, n, max_seq_length, masked_lm_prob,
                                                 short_seq_prob, seed, skip_warmup, binary_head, max_seq_length_dec,
                                                 dataset_type=dataset_type)
        for col, d in zip(columns, parts):
            if d:
                col.append(d)
    return tuple(BlendableDataset(c, weights) if c else None for c in columns)


def _build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples,
                                     max_seq_length, masked_lm_prob, short_seq_prob, seed, skip_warmup, binary_head,
                                     max_seq_length_dec, dataset_type="standard_bert"):
    if dataset_type not in DSET_TYPES:
        raise ValueError("Invalid dataset_type: ", dataset_type)
    indexed = get_indexed_dataset_(data_prefix, data_impl, skip_warmup)
    titles = None
    if dataset_type == DSET_TYPE_ICT:
        from ..global_vars import get_args
        titles = get_indexed_dataset_(get_args().titles_data_path, data_impl, skip_warmup)
    n_docs = indexed.doc_idx.shape[0] - 1
    splits = get_train_valid_test_split_(splits_string, n_docs)
    print_rank_0(" > dataset split:")
    for i, name in enumerate(("train", "validation", "test")):
        lo, hi = indexed.doc_idx[splits[i]], indexed.doc_idx[splits[i + 1]]
        print_rank_0(f"    {name}:\n     document indices in [{splits[i]}, {splits[i + 1]}) total of "
                     f"{splits[i + 1] - splits[i]} documents\n     sentence indices in [{lo}, {hi}) total of "
                     f"{hi - lo} sentences")

    def build(i, name):
        if splits[i + 1] <= splits[i]:
            return None
        full = indexed.get_doc_idx()
        indexed.set_doc_idx(full[splits[i]:splits[i + 1] + 1])     # the dataset only sees its split's documents
        kw = dict(name=name, data_prefix=data_prefix, num_epochs=None,
                  max_num_samples=train_valid_test_num_samples[i], max_seq_length=max_seq_length, seed=seed)
        if dataset_type == DSET_TYPE_ICT:
            from ..global_vars import get_args
            from .ict_dataset import ICTDataset
            a = get_args()
            ds = ICTDataset(block_dataset=indexed, title_dataset=titles, query_in_block_prob=a.query_in_block_prob,
                            use_one_sent_docs=a.use_one_sent_docs, binary_head=binary_head, **kw)
        elif dataset_type == DSET_TYPE_T5:
            from .t5_dataset import T5Dataset
            ds = T5Dataset(indexed_dataset=indexed, masked_lm_prob=masked_lm_prob,
                           max_seq_length_dec=max_seq_length_dec, short_seq_prob=short_seq_prob, **kw)
        else:
            from .bert_dataset import BertDataset
            ds = BertDataset(indexed_dataset=indexed, masked_lm_prob=masked_lm_prob, short_seq_prob=short_seq_prob,
                             binary_head=binary_head, **kw)
        indexed.set_doc_idx(full)
        assert indexed.doc_idx[0] == 0 and indexed.doc_idx.shape[0] == n_docs + 1
        return ds

    return build(0, "train"), build(1, "valid"), build(2, "test")

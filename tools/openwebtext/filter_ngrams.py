"""Decontaminate a training corpus from downstream-task text (parity: tools/openwebtext/filter_ngrams.py).

1. collect the word n-grams (``--max_ngram_size``, default 13; shorter task texts contribute their longest n-gram
   down to ``--min_ngram_size``) of the evaluation data (``--tasks lambada squad natural_questions triviaqa webqa race
   drop coqa piqa`` or any json-lines file path; LAMBADA via ``--lambada_path``);
2. count how often each n-gram occurs in the corpus and ignore n-grams seen more than ``--key_threshold`` times
   (they are generic phrases, not leaks);
3. for every remaining hit, cut the document at the sentence boundary ``--remove_char_each_side`` characters around
   the match, drop pieces shorter than ``--filter_text_char_len``, and drop documents that shatter into more than
   ``--splits_count`` pieces.
Output: json lines with the cleaned pieces (``<key>`` replaced, ``split_id`` appended to the ids of multi-piece docs)."""
import argparse
import json
import os
import pickle
import re
import time
from collections import Counter

_WORD = re.compile(r"\w+")


def get_words(text):
    words, positions = [], []
    for m in _WORD.finditer(text.lower()):
        words.append(m.group(0))
        positions.append(m.start())
    return words, positions


def task_texts(args, task):
    """Yield the raw evaluation strings of one task."""
    if task == "lambada":
        assert args.lambada_path is not None, "--lambada_path is required for the lambada task"
        with open(args.lambada_path, "r") as f:
            for line in f:
                yield json.loads(line)["text"]
        return
    if os.path.exists(task):                       # any json-lines / text file
        with open(task, "r") as f:
            for line in f:
                line = line.strip()
                if not line:
                    continue
                try:
                    obj = json.loads(line)
                    yield from (v for v in obj.values() if isinstance(v, str)) if isinstance(obj, dict) else [str(obj)]
                except json.JSONDecodeError:
                    yield line
        return
    try:
        from datasets import load_dataset
    except ImportError:
        raise SystemExit(f"task {task}: the `datasets` package (and network access) is needed for hub tasks; "
                         "pass a local file path instead")
    spec = {"squad": ("squad_v2", None, "validation", ["question"]),
            "natural_questions": ("natural_questions", None, "validation", ["question"]),
            "triviaqa": ("trivia_qa", "unfiltered", "test", ["question"]),
            "webqa": ("web_questions", None, "test", ["question"]),
            "race": ("race", "all", "test", ["article", "question"]),
            "drop": ("drop", None, "validation", ["question"]),
            "coqa": ("coqa", None, "validation", ["story"]),
            "piqa": ("piqa", None, "test", ["goal"])}[task]
    for row in load_dataset(spec[0], spec[1], split=spec[2]):
        for field in spec[3]:
            v = row[field]
            yield v["text"] if isinstance(v, dict) else v


def compute_tasks_ngrams(args):
    """n-gram string -> 0 (occurrence counter filled later)."""
    ngrams = {}
    for task in args.tasks:
        t0, before = time.time(), len(ngrams)
        for text in task_texts(args, task):
            words, _ = get_words(text)
            if len(words) < args.min_ngram_size:
                continue
            n = min(len(words), args.max_ngram_size)
            for i in range(len(words) - n + 1):
                ngrams.setdefault(" ".join(words[i:i + n]), 0)
        print(" > {}: {} new ngrams in {:.2f}s".format(task, len(ngrams) - before, time.time() - t0), flush=True)
    return ngrams


def find_matches(words, ngrams, sizes):
    """First (word index, n) whose n-gram is a task n-gram, scanning left to right."""
    for i in range(len(words)):
        for n in sizes:
            if i + n <= len(words) and " ".join(words[i:i + n]) in ngrams:
                return i, n
    return None


def split_text(text, start, length, margin):
    """Pieces before / after the match, cut at the nearest sentence end at least ``margin`` chars away."""
    end_marks = ".!?"
    pos = start - margin
    while pos > 0 and text[pos] not in end_marks:
        pos -= 1
    first = text[:pos + 1] if pos > 0 else ""
    pos = start + length + margin
    while pos < len(text) and text[pos] not in end_marks:
        pos += 1
    second = text[pos + 1:] if pos + 1 < len(text) else ""
    return first, second


def free_ngram(text, args, ngrams, sizes, counter=None):
    """-> list of n-gram-free pieces (or only counts matches when ``counter`` is given)."""
    queue, clean = [text], []
    while queue:
        piece = queue.pop(0)
        words, positions = get_words(piece)
        hit = find_matches(words, ngrams, sizes)
        if hit is None:
            if counter is None:
                clean.append(piece)
            continue
        i, n = hit
        seq = " ".join(words[i:i + n])
        start = positions[i]
        span = positions[i + n - 1] + len(words[i + n - 1]) - start
        if counter is not None:
            counter[seq] += 1
            rest = piece[start + span + 1:]
            if rest:
                queue.append(rest)
            continue
        first, second = split_text(piece, start, span, args.remove_char_each_side)
        if len(first) > args.filter_text_char_len:
            clean.append(first)
        if len(second) > args.filter_text_char_len:
            queue.append(second)
    return clean


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--tasks", nargs="*", required=True, default=None,
                   help="Tasks to use for deduplication: lambada, squad, natural_questions, triviaqa, webqa, race, "
                        "drop, coqa, piqa, or paths of local json-lines / text files")
    p.add_argument("--lambada_path", type=str, default=None, help="Only Lambada task needs the path")
    p.add_argument("--dedup_dataset", nargs="*", default=None, help="Dataset to deduplicate with the key to use, "
                   "e.g. cc.json text")
    p.add_argument("--output", type=str, default=None, help="Output file name to save dedup dataset")
    p.add_argument("--num_threads", type=int, default=40, help="(accepted for CLI parity)")
    p.add_argument("--max_ngram_size", type=int, default=13, help="Maximum size of ngram to use.")
    p.add_argument("--min_ngram_size", type=int, default=8, help="Minimum size of ngram to use.")
    p.add_argument("--filter_text_char_len", type=int, default=200, help="Remove any text below this length.")
    p.add_argument("--key_threshold", type=int, default=10, help="Number of keys to consider as threshold")
    p.add_argument("--save_dictionary", type=str, default=None, help="Save the dictionary")
    p.add_argument("--load_dictionary", type=str, default=None, help="Load the dictionary")
    p.add_argument("--splits_count", type=int, default=10, help="Remove any documents more than this many splits")
    p.add_argument("--remove_char_each_side", type=int, default=200, help="Maximum remove chars from each side of matched ngram")
    args = p.parse_args()
    assert len(args.dedup_dataset) == 2
    dedup_file, key = args.dedup_dataset
    t0 = time.time()
    if args.load_dictionary:
        with open(args.load_dictionary, "rb") as f:
            ngrams = pickle.load(f)
    else:
        ngrams = compute_tasks_ngrams(args)
    sizes = sorted({len(k.split()) for k in ngrams}, reverse=True)
    # pass 1: corpus frequency of every task n-gram
    freq = Counter()
    with open(dedup_file, "r", encoding="utf-8") as f:
        for line in f:
            try:
                free_ngram(json.loads(line)[key], args, ngrams, sizes, counter=freq)
            except Exception as e:
                print("Error:", e)
    rare = {k: 0 for k in ngrams if freq[k] < args.key_threshold}
    print(" > {} of {} task ngrams occur fewer than {} times in the corpus".format(len(rare), len(ngrams),
                                                                                  args.key_threshold), flush=True)
    if args.save_dictionary:
        with open(args.save_dictionary, "wb") as f:
            pickle.dump(rare, f)
    # pass 2: cut them out
    sizes = sorted({len(k.split()) for k in rare}, reverse=True)
    stats = Counter()
    with open(dedup_file, "r", encoding="utf-8") as f, open(args.output, "wb") as out:
        for line in f:
            stats["docs"] += 1
            try:
                doc = json.loads(line)
                pieces = free_ngram(doc[key], args, rare, sizes)
            except Exception as e:
                print("Error:", e)
                continue
            if len(pieces) == 1 and len(pieces[0]) == len(doc[key]):
                stats["untouched"] += 1
            elif len(pieces) > args.splits_count:
                stats["dropped_too_many_splits"] += 1
                continue
            else:
                stats["trimmed_or_split"] += 1
            for i, piece in enumerate(pieces):
                row = dict(doc)
                row[key] = piece
                if len(pieces) > 1:
                    row["split_id"] = i
                out.write(json.dumps(row, ensure_ascii=False).encode("utf-8") + b"\n")
                stats["written"] += 1
    print(" [FINAL] {:.2f}s {}".format(time.time() - t0, dict(stats)), flush=True)
    print("done :-)")

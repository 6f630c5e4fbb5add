"""Tokenize a JSON-lines corpus into indexed datasets (``<prefix>_<key>_<document|sentence>.{bin,idx}``).

Parity: tools/preprocess_data.py (same flags).  Sentence splitting uses NLTK punkt when it is installed and a
rule-based splitter otherwise (there is no network to download punkt data on an air-gapped cluster)."""
from __future__ import annotations

import argparse
import json
import multiprocessing
import os
import re
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))
from megatron_llm_b200.data import indexed_dataset  # noqa: E402
from megatron_llm_b200.tokenizer import build_tokenizer  # noqa: E402


class IdentitySplitter:
    def tokenize(self, *text):
        return text


class RegexSentenceSplitter:
    """Split after ., ! or ? followed by whitespace and an upper-case/digit/quote start."""
    _boundary = re.compile(r"(?<=[.!?])(\s+)(?=[\"'(\[]?[A-Z0-9])")

    def __init__(self, keep_newlines=False):
        self.keep_newlines = keep_newlines

    def tokenize(self, text):
        out, last = [], 0
        for m in self._boundary.finditer(text):
            end = m.end(1) if self.keep_newlines else m.start(1)
            out.append(text[last:end])
            last = m.end(1)
        out.append(text[last:])
        return [s for s in out if s]


def make_splitter(args):
    if not args.split_sentences:
        return IdentitySplitter()
    try:
        import nltk
        splitter = nltk.load(f"tokenizers/punkt/{args.lang}.pickle")
        if args.keep_newlines:
            class _Vars(nltk.tokenize.punkt.PunktLanguageVars):
                _period_context_fmt = r"""\S*%(SentEndChars)s\s*(?=(?P<after_tok>%(NonWord)s|(?P<next_tok>\S+)))"""
            splitter = nltk.tokenize.punkt.PunktSentenceTokenizer(train_text=splitter._params, lang_vars=_Vars())
        return splitter
    except Exception:
        print("NLTK punkt unavailable: using the rule-based sentence splitter", file=sys.stderr)
        return RegexSentenceSplitter(args.keep_newlines)


class Encoder:
    tokenizer = None
    splitter = None

    def __init__(self, args):
        self.args = args

    def initializer(self):
        Encoder.tokenizer = build_tokenizer(self.args)
        Encoder.splitter = make_splitter(self.args)

    def encode(self, json_line):
        data = json.loads(json_line)
        ids = {}
        for key in self.args.json_keys:
            doc_ids = [s for s in (Encoder.tokenizer.tokenize(sent) for sent in Encoder.splitter.tokenize(data[key]))
                       if len(s) > 0]
            if doc_ids and self.args.append_eod:
                doc_ids[-1].append(Encoder.tokenizer.eod)
            ids[key] = doc_ids
        return ids, len(json_line)


def add_tokenizer_args(parser):
    group = parser.add_argument_group(title="tokenizer")
    group.add_argument("--tokenizer_type", type=str, required=True,
                       choices=["BertWordPieceLowerCase", "BertWordPieceCase", "GPT2BPETokenizer",
                                "SentencePieceTokenizer", "FalconTokenizer", "NullTokenizer"],
                       help="What type of tokenizer to use.")
    group.add_argument("--vocab_file", type=str, default=None, help="Path to the vocab file")
    group.add_argument("--merge_file", type=str, default=None, help="Path to the BPE merge file (if necessary).")
    group.add_argument("--lang", type=str, default="english", help="Language to use for NLTK-powered sentence splitting.")
    group.add_argument("--vocab_extra_ids", type=int, default=0)
    group.add_argument("--vocab_extra_ids_list", type=str, default=None,
                       help="comma separated list of special vocab ids to add to the tokenizer")
    group.add_argument("--no_new_tokens", action="store_false", dest="new_tokens",
                       help="Whether to add special tokens (e.g. CLS, MASK, etc) in the sentenciepiece tokenizer or not")
    return group


def add_runtime_args(parser):
    group = parser.add_argument_group(title="runtime")
    group.add_argument("--workers", type=int, required=True, help="Number of worker processes to launch")
    group.add_argument("--chunk_size", type=int, required=True, help="Chunk size assigned to each worker process")
    group.add_argument("--log_interval", type=int, default=100, help="Interval between progress updates")
    return group


def tokenizer_defaults(args):
    """Dummy values ``build_tokenizer`` expects from the training arguments."""
    args.rank = 0
    args.make_vocab_size_divisible_by = 128
    args.tensor_model_parallel_size = 1
    return args


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    group = parser.add_argument_group(title="input data")
    group.add_argument("--input", type=str, required=True, help="Path to input JSON")
    group.add_argument("--json_keys", nargs="+", default=["text"], help="space separate listed of keys to extract from json")
    group.add_argument("--split_sentences", action="store_true", help="Split documents into sentences.")
    group.add_argument("--keep_newlines", action="store_true", help="Keep newlines between sentences when splitting.")
    tok = add_tokenizer_args(parser)
    tok.add_argument("--append_eod", action="store_true", help="Append an <eod> token to the end of a document.")
    group = parser.add_argument_group(title="output data")
    group.add_argument("--output_prefix", type=str, required=True, help="Path to binary output file without suffix")
    group.add_argument("--dataset_impl", type=str, default="mmap", choices=["lazy", "cached", "mmap"])
    add_runtime_args(parser)
    args = parser.parse_args(argv)
    args.keep_empty = False
    if args.tokenizer_type.lower().startswith("bert") and not args.split_sentences:
        print("Bert tokenizer detected, are you sure you don't want to split sentences?")
    return tokenizer_defaults(args)


def progress(i, t0, nbytes, interval):
    if i % interval == 0:
        dt = time.time() - t0
        print(f"Processed {i} documents ({i / dt} docs/s, {nbytes / dt / 1024 / 1024} MB/s).", file=sys.stderr)


def main(argv=None):
    args = get_args(argv)
    t_start = time.time()
    print("Opening", args.input)
    fin = open(args.input, "r", encoding="utf-8")
    encoder = Encoder(args)
    tokenizer = build_tokenizer(args)
    if args.workers > 1:
        pool = multiprocessing.Pool(args.workers, initializer=encoder.initializer)
        encoded_docs = pool.imap(encoder.encode, fin, args.chunk_size)
    else:
        encoder.initializer()
        encoded_docs = map(encoder.encode, fin)
    level = "sentence" if args.split_sentences else "document"
    print(f"Vocab size: {tokenizer.vocab_size}")
    print(f"Output prefix: {args.output_prefix}")
    builders, idx_files = {}, {}
    for key in args.json_keys:
        idx_files[key] = f"{args.output_prefix}_{key}_{level}.idx"
        builders[key] = indexed_dataset.make_builder(f"{args.output_prefix}_{key}_{level}.bin", impl=args.dataset_impl,
                                                     vocab_size=tokenizer.vocab_size)
    print("Time to startup:", time.time() - t_start)
    t0, nbytes = time.time(), 0
    for i, (doc, n) in enumerate(encoded_docs, start=1):
        nbytes += n
        for key, sentences in doc.items():
            if not sentences:
                continue
            for s in sentences:
                builders[key].add_item(torch.IntTensor(s))
            builders[key].end_document()
        progress(i, t0, nbytes, args.log_interval)
    print("Done! Now finalizing.")
    for key in args.json_keys:
        builders[key].finalize(idx_files[key])


if __name__ == "__main__":
    main()

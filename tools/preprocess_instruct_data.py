"""Tokenize instruction/chat JSON-lines into the ``<prefix>-text`` / ``<prefix>-role`` indexed dataset pair.

Parity: tools/preprocess_instruct_data.py.  Each line holds a question, an answer and optionally a system prompt;
every message is wrapped as ``<|im_start|>{role}\\n{message}<|im_end|>\\n`` and tokenised separately so each token
carries its author's :class:`Role` (the loss is masked to assistant tokens at training time)."""
from __future__ import annotations

import itertools
import json
import os
import sys
import time
from argparse import ArgumentParser
from multiprocessing import Pool
from pathlib import Path

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from megatron_llm_b200.data.indexed_dataset import make_builder  # noqa: E402
from megatron_llm_b200.data.instruction_dataset import Role  # noqa: E402
from megatron_llm_b200.tokenizer import build_tokenizer  # noqa: E402
from preprocess_data import add_runtime_args, add_tokenizer_args, progress, tokenizer_defaults  # noqa: E402


def format_message(message: str, role: str) -> str:
    return f"<|im_start|>{role}\n{message}<|im_end|>\n"


class Encoder:
    tokenizer = None

    def __init__(self, args):
        self.args = args

    def initializer(self):
        Encoder.tokenizer = build_tokenizer(self.args)

    def encode(self, line: str):
        data = json.loads(line)
        a = self.args
        turns = []
        if a.system_key is not None:
            turns.append((data[a.system_key], "system", Role.system))
        turns.append((data[a.question_key], "question", Role.prompter))
        turns.append((data[a.answer_key], "answer", Role.assistant))
        tokens, roles = [], []
        for message, tag, role in turns:
            ids = Encoder.tokenizer.tokenize(format_message(message, tag))
            tokens += ids
            roles += [role.value] * len(ids)
        return len(line), tokens, roles


class DatasetWriter:
    def __init__(self, prefix, vocab_size, dataset_impl="mmap", feature="text"):
        self.vocab_size, self.dataset_impl = vocab_size, dataset_impl
        self.bin_fname, self.idx_fname = f"{prefix}-{feature}.bin", f"{prefix}-{feature}.idx"
        self.builder = None

    def add_item(self, tokens):
        self.builder.add_item(torch.IntTensor(tokens))

    def __enter__(self):
        self.builder = make_builder(self.bin_fname, impl=self.dataset_impl, vocab_size=self.vocab_size)
        return self

    def __exit__(self, *_):
        self.builder.finalize(self.idx_fname)
        self.builder = None


def get_args(argv=None):
    parser = ArgumentParser()
    group = parser.add_argument_group(title="input data")
    group.add_argument("--input", type=str, nargs="+", help="Path(s) to input JSON file(s)")
    group.add_argument("--system_key", help="key to extract system info from json (optional)")
    group.add_argument("--question_key", default="input", help="key to extract questions from json")
    group.add_argument("--answer_key", default="output", help="key to extract answers from json")
    add_tokenizer_args(parser)
    group = parser.add_argument_group(title="output data")
    group.add_argument("--output_prefix", type=Path, required=True, help="Path to binary output file without suffix")
    group.add_argument("--dataset_impl", type=str, default="mmap", choices=["lazy", "cached", "mmap"])
    add_runtime_args(parser)
    args = parser.parse_args(argv)
    args.keep_empty = False
    return tokenizer_defaults(args)


def main(argv=None):
    args = get_args(argv)
    t_start = time.time()
    encoder = Encoder(args)
    vocab_size = build_tokenizer(args).vocab_size
    fs = [open(p) for p in args.input]
    lines = itertools.chain(*fs)
    if args.workers > 1:
        pool = Pool(args.workers, initializer=encoder.initializer)
        docs = pool.imap(encoder.encode, lines, args.chunk_size)
    else:
        encoder.initializer()
        docs = map(encoder.encode, lines)
    print("Time to startup:", time.time() - t_start)
    with DatasetWriter(args.output_prefix, vocab_size, args.dataset_impl, "text") as tw, \
            DatasetWriter(args.output_prefix, 16, args.dataset_impl, "role") as rw:
        t0, nbytes = time.time(), 0
        for i, (n, tokens, roles) in enumerate(docs, start=1):
            nbytes += n
            tw.add_item(tokens)
            rw.add_item(roles)
            progress(i, t0, nbytes, args.log_interval)
        print("Done! Now finalizing.")
    for f in fs:
        f.close()


if __name__ == "__main__":
    main()

"""Byte-level BPE (GPT-2) tokenizer (parity: megatron/tokenizer/gpt2_tokenization.py).

Own implementation of the published algorithm: bytes -> printable unicode alphabet, regex pre-tokenisation,
greedy lowest-rank pair merges with a per-word cache."""
from __future__ import annotations

import json
import os
from functools import lru_cache

import regex as re

_PAT = re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")


@lru_cache()
def bytes_to_unicode():
    """Reversible map byte -> printable unicode char (avoids whitespace/control chars the BPE code chokes on)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + \
        list(range(ord("®"), ord("ÿ") + 1))
    chars = keep[:]
    extra = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return dict(zip(keep, (chr(c) for c in chars)))


VOCAB_NAME, MERGES_NAME, SPECIAL_TOKENS_NAME = "vocab.json", "merges.txt", "special_tokens.txt"


def get_pairs(word):
    return {(a, b) for a, b in zip(word[:-1], word[1:])}


class GPT2Tokenizer:
    def __init__(self, vocab_file, merges_file, errors="replace", special_tokens=None, max_len=None):
        self.max_len = max_len if max_len is not None else int(1e12)
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder = json.load(f)
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.errors = errors
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().split("\n")[1:-1]
        merges = [tuple(l.split()) for l in lines if l and not l.startswith("#")]
        self.bpe_ranks = dict(zip(merges, range(len(merges))))
        self.cache = {}
        self.special_tokens, self.special_tokens_decoder = {}, {}
        self.set_special_tokens(special_tokens)

    def __len__(self):
        return len(self.encoder) + len(self.special_tokens)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, cache_dir=None, *inputs, **kwargs):
        """Load ``vocab.json`` / ``merges.txt`` (+ optional ``special_tokens.txt``) from a directory, the layout
        ``save_vocabulary`` writes (reference gpt2_tokenization.py:98-160; the reference can also download the
        published GPT-2 files by name -- there is no download path here, ``cache_dir`` is searched instead)."""
        base = pretrained_model_name_or_path
        if not os.path.isdir(base) and cache_dir is not None:
            base = os.path.join(cache_dir, pretrained_model_name_or_path)
        vocab_file, merges_file = os.path.join(base, VOCAB_NAME), os.path.join(base, MERGES_NAME)
        if not (os.path.exists(vocab_file) and os.path.exists(merges_file)):
            print("GPT2Tokenizer: no {} / {} under {!r}".format(VOCAB_NAME, MERGES_NAME, base), flush=True)
            return None
        special_file = os.path.join(base, SPECIAL_TOKENS_NAME)
        if os.path.exists(special_file) and "special_tokens" not in kwargs:
            with open(special_file, encoding="utf-8") as f:
                kwargs["special_tokens"] = f.read().split("\n")[:-1]
        return cls(vocab_file, merges_file, *inputs, **kwargs)

    def save_vocabulary(self, vocab_path):
        """Write the vocabulary, merges and special tokens into a directory; returns the three paths."""
        if not os.path.isdir(vocab_path):
            print("Vocabulary path ({}) should be a directory".format(vocab_path), flush=True)
            return None
        vocab_file, merge_file = os.path.join(vocab_path, VOCAB_NAME), os.path.join(vocab_path, MERGES_NAME)
        special_tokens_file = os.path.join(vocab_path, SPECIAL_TOKENS_NAME)
        with open(vocab_file, "w", encoding="utf-8") as f:
            f.write(json.dumps(self.encoder, ensure_ascii=False))
        with open(merge_file, "w", encoding="utf-8") as f:
            f.write("#version: 0.2\n")
            for pair, _ in sorted(self.bpe_ranks.items(), key=lambda kv: kv[1]):
                f.write(" ".join(pair) + "\n")
        with open(special_tokens_file, "w", encoding="utf-8") as f:
            for tok, _ in sorted(self.special_tokens.items(), key=lambda kv: kv[1]):
                f.write(tok + "\n")
        return vocab_file, merge_file, special_tokens_file

    def set_special_tokens(self, special_tokens):
        if not special_tokens:
            self.special_tokens, self.special_tokens_decoder = {}, {}
            return
        self.special_tokens = {tok: len(self.encoder) + i for i, tok in enumerate(special_tokens)}
        self.special_tokens_decoder = {v: k for k, v in self.special_tokens.items()}

    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token)
        pairs = get_pairs(word)
        if not pairs:
            return token
        while True:
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
            if len(word) == 1:
                break
            pairs = get_pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def tokenize(self, text):
        pieces = []
        for tok in re.findall(_PAT, text):
            mapped = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            pieces.extend(self.bpe(mapped).split(" "))
        return pieces

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self.special_tokens.get(tokens, self.encoder.get(tokens, 0))
        return [self.special_tokens[t] if t in self.special_tokens else self.encoder.get(t, 0) for t in tokens]

    def convert_ids_to_tokens(self, ids, skip_special_tokens=False):
        out = []
        for i in ids:
            if i in self.special_tokens_decoder:
                if not skip_special_tokens:
                    out.append(self.special_tokens_decoder[i])
            else:
                out.append(self.decoder[i])
        return out

    def encode(self, text):
        return self.convert_tokens_to_ids(self.tokenize(text))

    def decode(self, tokens):
        text = "".join(self.decoder[t] for t in tokens)
        return bytearray([self.byte_decoder[c] for c in text]).decode("utf-8", errors=self.errors)

"""BERT WordPiece tokenizer (parity: megatron/tokenizer/bert_tokenization.py): basic (whitespace / punctuation /
CJK splitting, optional lower-casing + accent stripping) followed by greedy longest-match-first WordPiece."""
from __future__ import annotations

import collections
import unicodedata


def convert_to_unicode(text):
    if isinstance(text, bytes):
        return text.decode("utf-8", "ignore")
    return str(text)


def load_vocab(vocab_file):
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            tok = convert_to_unicode(line).strip()
            if tok:
                vocab[tok] = i
            else:
                vocab.setdefault(line.rstrip("\n"), i)
    return vocab


def whitespace_tokenize(text):
    text = text.strip()
    return text.split() if text else []


def _is_whitespace(ch):
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch):
    if ch in "\t\n\r":
        return False
    return unicodedata.category(ch) in ("Cc", "Cf")


def _is_punctuation(ch):
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp):
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF
            or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer:
    def __init__(self, do_lower_case=True):
        self.do_lower_case = do_lower_case

    def tokenize(self, text):
        text = convert_to_unicode(text)
        cleaned = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_cjk(cp):
                cleaned.append(" " + ch + " ")
            else:
                cleaned.append(" " if _is_whitespace(ch) else ch)
        out = []
        for tok in whitespace_tokenize("".join(cleaned)):
            if self.do_lower_case:
                tok = "".join(c for c in unicodedata.normalize("NFD", tok.lower()) if unicodedata.category(c) != "Mn")
            cur = []
            for ch in tok:
                if _is_punctuation(ch):
                    if cur:
                        out.append("".join(cur))
                        cur = []
                    out.append(ch)
                else:
                    cur.append(ch)
            if cur:
                out.append("".join(cur))
        return out


class WordpieceTokenizer:
    def __init__(self, vocab, unk_token="[UNK]", max_input_chars_per_word=200):
        self.vocab, self.unk_token, self.max_input_chars_per_word = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text):
        out = []
        for token in whitespace_tokenize(convert_to_unicode(text)):
            if len(token) > self.max_input_chars_per_word:
                out.append(self.unk_token)
                continue
            pieces, start, bad = [], 0, False
            while start < len(token):
                end, found = len(token), None
                while start < end:
                    sub = token[start:end]
                    if start > 0:
                        sub = "##" + sub
                    if sub in self.vocab:
                        found = sub
                        break
                    end -= 1
                if found is None:
                    bad = True
                    break
                pieces.append(found)
                start = end
            out.extend([self.unk_token] if bad else pieces)
        return out


class FullTokenizer:
    def __init__(self, vocab_file, do_lower_case=True):
        self.vocab = load_vocab(vocab_file)
        self.inv_vocab = {v: k for k, v in self.vocab.items()}
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab)

    def tokenize(self, text):
        return [p for t in self.basic_tokenizer.tokenize(text) for p in self.wordpiece_tokenizer.tokenize(t)]

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab[t] for t in tokens]

    def convert_ids_to_tokens(self, ids):
        return [self.inv_vocab[i] for i in ids]

    @staticmethod
    def convert_tokens_to_string(tokens, clean_up_tokenization_spaces=True):
        text = " ".join(tokens).replace(" ##", "").strip()
        if clean_up_tokenization_spaces:
            for a, b in ((" .", "."), (" ?", "?"), (" !", "!"), (" ,", ","), (" ' ", "'"), (" n't", "n't"),
                         (" 'm", "'m"), (" 's", "'s"), (" 've", "'ve"), (" 're", "'re")):
                text = text.replace(a, b)
        return text

    def vocab_size(self):
        return len(self.vocab)

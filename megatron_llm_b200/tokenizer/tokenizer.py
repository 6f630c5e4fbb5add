"""Tokenizer factory and wrappers (parity: megatron/tokenizer/tokenizer.py:12-503).

``args.padded_vocab_size`` is padded to a multiple of ``make_vocab_size_divisible_by * tensor_parallel_size``.
``NullTokenizer`` (vocab size only) is an addition for synthetic-data runs without tokenizer files.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

from .bert_tokenization import FullTokenizer as FullBertTokenizer
from .gpt2_tokenization import GPT2Tokenizer


def build_tokenizer(args):
    if args.rank == 0:
        print("> building {} tokenizer ...".format(args.tokenizer_type), flush=True)
    t = args.tokenizer_type
    if t not in ("FalconTokenizer", "NullTokenizer"):
        assert args.vocab_file is not None
    if t == "BertWordPieceLowerCase":
        tokenizer = _BertWordPieceTokenizer(vocab_file=args.vocab_file, lower_case=True,
                                            vocab_extra_ids=args.vocab_extra_ids)
    elif t == "BertWordPieceCase":
        tokenizer = _BertWordPieceTokenizer(vocab_file=args.vocab_file, lower_case=False,
                                            vocab_extra_ids=args.vocab_extra_ids)
    elif t == "GPT2BPETokenizer":
        assert args.merge_file is not None
        tokenizer = _GPT2BPETokenizer(args.vocab_file, args.merge_file)
    elif t == "SentencePieceTokenizer":
        tokenizer = _SentencePieceTokenizer(args.vocab_file, vocab_extra_ids=args.vocab_extra_ids,
                                            vocab_extra_ids_list=args.vocab_extra_ids_list,
                                            new_tokens=args.new_tokens)
    elif t == "FalconTokenizer":
        tokenizer = _FalconTokenizer(vocab_extra_ids_list=args.vocab_extra_ids_list, new_tokens=args.new_tokens,
                                     path=getattr(args, "tokenizer_model", None))
    elif t == "NullTokenizer":
        tokenizer = _NullTokenizer(int(args.vocab_file) if args.vocab_file and str(args.vocab_file).isdigit()
                                   else getattr(args, "null_vocab_size", 32000))
    else:
        raise NotImplementedError("{} tokenizer is not implemented.".format(t))
    args.padded_vocab_size = _vocab_size_with_padding(tokenizer.vocab_size, args)
    return tokenizer


def _vocab_size_with_padding(orig_vocab_size, args):
    multiple = args.make_vocab_size_divisible_by * args.tensor_model_parallel_size
    after = ((orig_vocab_size + multiple - 1) // multiple) * multiple
    if args.rank == 0:
        print(" > padded vocab (size: {}) with {} dummy tokens (new size: {})".format(
            orig_vocab_size, after - orig_vocab_size, after), flush=True)
    return after


class AbstractTokenizer(ABC):
    def __init__(self, name):
        self.name = name
        super().__init__()

    @property
    @abstractmethod
    def vocab_size(self): ...

    @property
    @abstractmethod
    def vocab(self): ...

    @property
    @abstractmethod
    def inv_vocab(self): ...

    @abstractmethod
    def tokenize(self, text): ...

    def detokenize(self, token_ids):
        raise NotImplementedError("detokenizer is not implemented for {} tokenizer".format(self.name))

    def _missing(self, what):
        raise NotImplementedError("{} is not provided for {} tokenizer".format(what, self.name))

    @property
    def cls(self):
        self._missing("CLS")

    @property
    def sep(self):
        self._missing("SEP")

    @property
    def pad(self):
        self._missing("PAD")

    @property
    def eod(self):
        self._missing("EOD")

    @property
    def mask(self):
        self._missing("MASK")


class _NullTokenizer(AbstractTokenizer):
    """Whitespace-separated integer ids; ``eod`` is the last id of the vocabulary."""

    def __init__(self, vocab_size):
        super().__init__("NullTokenizer")
        self._vocab_size = int(vocab_size)

    @property
    def vocab_size(self):
        return self._vocab_size

    @property
    def vocab(self):
        return {str(i): i for i in range(self._vocab_size)}

    @property
    def inv_vocab(self):
        return {i: str(i) for i in range(self._vocab_size)}

    def tokenize(self, text):
        """Integer words map to themselves; anything else to a stable hash bucket (so free text can be smoke-tested)."""
        import zlib
        out = []
        for w in text.split():
            try:
                out.append(int(w) % self._vocab_size)
            except ValueError:
                out.append(zlib.crc32(w.encode()) % self._vocab_size)
        return out

    def detokenize(self, ids):
        return " ".join(str(int(i)) for i in ids)

    @property
    def eod(self):
        return self._vocab_size - 1

    @property
    def pad(self):
        return self._vocab_size - 1

    @property
    def cls(self):
        return -1

    @property
    def sep(self):
        return -1

    @property
    def mask(self):
        return -1


class _BertWordPieceTokenizer(AbstractTokenizer):
    def __init__(self, vocab_file, lower_case=True, vocab_extra_ids=0):
        super().__init__("BERT Lower Case" if lower_case else "BERT Upper Case")
        self.tokenizer = FullBertTokenizer(vocab_file, do_lower_case=lower_case)
        self.cls_id = self.tokenizer.vocab["[CLS]"]
        self.sep_id = self.tokenizer.vocab["[SEP]"]
        self.pad_id = self.tokenizer.vocab["[PAD]"]
        self.mask_id = self.tokenizer.vocab["[MASK]"]
        self._additional_special_tokens = []
        self._bos_token = "[BOS]"
        self.add_token(self._bos_token)
        self._bos_token_id = self.vocab.get(self._bos_token)
        self._eos_token = "[EOS]"
        self.add_token(self._eos_token)
        self._eos_token_id = self.vocab.get(self._eos_token)
        self.add_additional_special_tokens(["<extra_id_{}>".format(i) for i in range(vocab_extra_ids)])

    def add_token(self, token):
        if token not in self.vocab:
            self.inv_vocab[self.vocab_size] = token
            self.vocab[token] = self.vocab_size

    def add_additional_special_tokens(self, tokens_list):
        setattr(self, "additional_special_tokens", tokens_list)
        for value in tokens_list:
            self.add_token(value)

    @property
    def vocab_size(self):
        return len(self.tokenizer.vocab)

    @property
    def vocab(self):
        return self.tokenizer.vocab

    @property
    def inv_vocab(self):
        return self.tokenizer.inv_vocab

    def tokenize(self, text):
        return self.tokenizer.convert_tokens_to_ids(self.tokenizer.tokenize(text))

    def decode(self, ids):
        return self.tokenizer.convert_tokens_to_string(self.tokenizer.convert_ids_to_tokens(ids))

    def decode_token_ids(self, token_ids):
        tokens = self.tokenizer.convert_ids_to_tokens(token_ids)
        exclude = ["[PAD]", "[CLS]"]
        result = ""
        for s in (t for t in tokens if t not in exclude):
            result += s[2:] if s.startswith("##") else " " + s
        return result

    @property
    def cls(self):
        return self.cls_id

    @property
    def sep(self):
        return self.sep_id

    @property
    def pad(self):
        return self.pad_id

    @property
    def mask(self):
        return self.mask_id

    @property
    def bos_token(self):
        return self._bos_token

    @property
    def eos_token(self):
        return self._eos_token

    @property
    def additional_special_tokens(self):
        return self._additional_special_tokens

    @additional_special_tokens.setter
    def additional_special_tokens(self, value):
        self._additional_special_tokens = value

    @property
    def bos_token_id(self):
        return self._bos_token_id

    @property
    def eos_token_id(self):
        return self._eos_token_id

    @property
    def additional_special_tokens_ids(self):
        return [self.vocab.get(token) for token in self._additional_special_tokens]


class _GPT2BPETokenizer(AbstractTokenizer):
    def __init__(self, vocab_file, merge_file):
        super().__init__("GPT2 BPE")
        self.tokenizer = GPT2Tokenizer(vocab_file, merge_file, errors="replace", special_tokens=[], max_len=None)
        self.eod_id = self.tokenizer.encoder["<|endoftext|>"]

    @property
    def vocab_size(self):
        return len(self.tokenizer.encoder)

    @property
    def vocab(self):
        return self.tokenizer.encoder

    @property
    def inv_vocab(self):
        return self.tokenizer.decoder

    def tokenize(self, text):
        return self.tokenizer.encode(text)

    def detokenize(self, token_ids):
        return self.tokenizer.decode(token_ids)

    @property
    def eod(self):
        return self.eod_id


class _FalconTokenizer(AbstractTokenizer):
    """HF ``tiiuae/falcon-40b`` tokenizer; ``path`` may point at a local copy (no network needed)."""

    def __init__(self, vocab_extra_ids_list=None, new_tokens=True, path=None):
        super().__init__("FalconTokenizer")
        from transformers import AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(path or "tiiuae/falcon-40b")
        self._eod = self.tokenizer.vocab["<|endoftext|>"]
        if vocab_extra_ids_list and new_tokens:
            self.tokenizer.add_special_tokens({
                "additional_special_tokens": self.tokenizer.additional_special_tokens + vocab_extra_ids_list.split(",")})
        self._inv_vocab = {idx: token for token, idx in self.tokenizer.vocab.items()}

    @property
    def vocab_size(self):
        return len(self.tokenizer.vocab)

    @property
    def vocab(self):
        return self.tokenizer.vocab

    def tokenize(self, text):
        return self.tokenizer.encode(text)

    def detokenize(self, token_ids):
        return self.tokenizer.decode(token_ids)

    @property
    def inv_vocab(self):
        return self._inv_vocab

    @property
    def eod(self):
        return self._eod


class _SentencePieceTokenizer(AbstractTokenizer):
    """SentencePiece model + Megatron special tokens (<CLS>,<SEP>,<EOD>,<MASK>, pad/bos/eos, <extra_id_i>, user
    list).  ``new_tokens=False`` keeps the vocabulary untouched and maps pad -> eos."""

    def __init__(self, model_file, vocab_extra_ids=0, vocab_extra_ids_list=None, new_tokens=True):
        super().__init__("SentencePieceTokenizer")
        import sentencepiece
        self._tokenizer = sentencepiece.SentencePieceProcessor(model_file=model_file)
        self._initalize(vocab_extra_ids, vocab_extra_ids_list, new_tokens)

    def _initalize(self, vocab_extra_ids, vocab_extra_ids_list, new_tokens):
        self._vocab, self._inv_vocab = {}, {}
        self._special_tokens, self._inv_special_tokens = {}, {}
        self._t5_tokens = []
        for i in range(len(self._tokenizer)):
            piece = self._tokenizer.id_to_piece(i)
            self._inv_vocab[i], self._vocab[piece] = piece, i

        def add_special(tok, force=False):
            if tok not in self._vocab:
                if not (new_tokens or force):
                    return None
                nid = len(self._vocab)
                self._vocab[tok], self._inv_vocab[nid] = nid, tok
            self._special_tokens[tok] = self._vocab[tok]
            self._inv_special_tokens[self._vocab[tok]] = tok
            return self._vocab[tok]

        def piece_or(idx, default):
            try:
                return self._tokenizer.id_to_piece(idx)
            except IndexError:
                return default

        self._cls_id = add_special("<CLS>")
        self._sep_id = add_special("<SEP>")
        self._eod_id = add_special("<EOD>")
        self._mask_id = add_special("<MASK>")
        self._pad_id = add_special(piece_or(self._tokenizer.pad_id(), "<PAD>"))
        self._bos_id = add_special(piece_or(self._tokenizer.bos_id(), "<BOS>"))
        self._eos_id = add_special(piece_or(self._tokenizer.eos_id(), "<EOS>"))
        if not new_tokens:
            self._pad_id = self._eos_id
        for i in range(vocab_extra_ids):
            tok = "<extra_id_{}>".format(i)
            add_special(tok, force=True)
            self._t5_tokens.append(tok)
        if vocab_extra_ids_list:
            for tok in vocab_extra_ids_list.split(","):
                add_special(tok, force=True)
        print("Special tokens: {}".format(self._special_tokens))

    @property
    def vocab_size(self):
        return len(self._vocab)

    @property
    def vocab(self):
        return self._vocab

    @property
    def inv_vocab(self):
        return self._inv_vocab

    def tokenize(self, text):
        """Split on the earliest special token repeatedly; plain spans go through SentencePiece."""
        ids, pos = [], 0
        while True:
            best_tok, best_at = None, None
            for tok in self._special_tokens:
                at = text.find(tok, pos)
                if at != -1 and (best_at is None or at < best_at):
                    best_tok, best_at = tok, at
            if best_tok is None:
                break
            ids.extend(self._tokenizer.encode_as_ids(text[pos:best_at]))
            ids.append(self._special_tokens[best_tok])
            pos = best_at + len(best_tok)
        ids.extend(self._tokenizer.encode_as_ids(text[pos:]))
        return ids

    def detokenize(self, ids):
        text, last = "", 0
        for i, tid in enumerate(ids):
            if tid in self._inv_special_tokens:
                text += self._tokenizer.decode_ids(ids[last:i]) + " " + self._inv_special_tokens[tid] + " "
                last = i + 1
        text += self._tokenizer.decode_ids(ids[last:])
        return text.strip()

    @property
    def cls(self):
        return self._cls_id

    @property
    def sep(self):
        return self._sep_id

    @property
    def pad(self):
        return self._pad_id

    @property
    def bos_token_id(self):
        return self._bos_id

    @property
    def bos(self):
        return self._bos_id

    @property
    def eod(self):
        return self._eod_id if self._eod_id is not None else self._eos_id

    @property
    def eos_token_id(self):
        return self._eod_id if self._eod_id is not None else self._eos_id

    @property
    def eos(self):
        return self._eos_id

    @property
    def mask(self):
        return self._mask_id

    @property
    def additional_special_tokens_ids(self):
        return [self.vocab[k] for k in self._t5_tokens]

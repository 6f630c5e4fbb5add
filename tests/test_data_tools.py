"""Data tooling: preprocess -> indexed dataset -> GPT / instruction datasets -> merge."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _write_jsonl(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


def test_preprocess_and_gpt_dataset(tmp_path):
    import preprocess_data
    from megatron_llm_b200.data import indexed_dataset
    from megatron_llm_b200.data.gpt_dataset import build_train_valid_test_datasets
    rows = [{"text": " ".join(str((i * 7 + j) % 50) for j in range(5 + i % 11))} for i in range(64)]
    _write_jsonl(tmp_path / "c.jsonl", rows)
    prefix = str(tmp_path / "corpus")
    preprocess_data.main(["--input", str(tmp_path / "c.jsonl"), "--output_prefix", prefix, "--tokenizer_type",
                          "NullTokenizer", "--vocab_file", "100", "--workers", "1", "--chunk_size", "8",
                          "--append_eod"])
    ds = indexed_dataset.make_dataset(prefix + "_text_document", "mmap")
    assert len(ds) == 64
    assert ds[3].tolist()[:-1] == [int(t) for t in rows[3]["text"].split()]
    train, valid, test = build_train_valid_test_datasets([prefix + "_text_document"], "mmap", "8,1,1", [40, 4, 4],
                                                         16, 1234, True)
    assert len(train) >= 40
    s = train[0]["text"]
    assert s.shape == (17,) and s.dtype == np.int64


def test_instruction_pipeline(tmp_path):
    import preprocess_instruct_data
    from megatron_llm_b200.data.instruction_dataset import Role, build_train_valid_test_datasets, collate
    rows = [{"input": "1 2 3", "output": "4 5", "sys": "9"} for _ in range(10)]
    _write_jsonl(tmp_path / "i.jsonl", rows)
    prefix = str(tmp_path / "inst")
    preprocess_instruct_data.main(["--input", str(tmp_path / "i.jsonl"), "--output_prefix", prefix, "--system_key",
                                   "sys", "--tokenizer_type", "NullTokenizer", "--vocab_file", "100", "--workers", "1",
                                   "--chunk_size", "4"])
    train, valid, test = build_train_valid_test_datasets([prefix], "mmap", "8,2,0", [20, 2, 0], 32, 1, True)
    assert len(train) == 20 and len(valid) == 2 and test is None
    x = train[0]
    assert x["text"].shape == x["role"].shape
    assert set(x["role"].tolist()) == {Role.system.value, Role.prompter.value, Role.assistant.value}
    batch = collate([train[0], train[1]], 32, pad_id=0)
    assert batch["text"].shape == (2, 33)
    n = len(x["text"])
    assert batch["attention_mask"][0].sum().item() == n
    assert batch["assistant_mask"][0].sum().item() == int((x["role"] == Role.assistant.value).sum())
    short = collate([train[0]], 32, pad_id=0, variable_seq_lengths=True)
    assert short["text"].shape[1] == 17


def test_merge_datasets(tmp_path):
    import merge_datasets
    from megatron_llm_b200.data import indexed_dataset
    d = tmp_path / "parts"
    d.mkdir()
    expect = []
    for part in range(3):
        b = indexed_dataset.make_builder(str(d / f"p{part}.bin"), "mmap", vocab_size=1000)
        for i in range(4):
            item = list(range(part * 10 + i, part * 10 + i + 3))
            expect.append(item)
            b.add_item(torch.IntTensor(item))
            b.end_document()
        b.finalize(str(d / f"p{part}.idx"))
    merge_datasets.main(["--input", str(d), "--output_prefix", str(tmp_path / "merged")])
    ds = indexed_dataset.make_dataset(str(tmp_path / "merged"), "mmap")
    assert [x.tolist() for x in ds] == expect
    assert len(ds.doc_idx) == 13


class _FakeBertTok:
    """ids 0..4 special, 5..: words; odd ids in [21, 90) are '##' continuation pieces."""
    cls, sep, mask, pad = 1, 2, 3, 0
    bos_token_id, eos_token_id = 4, 2
    additional_special_tokens_ids = list(range(90, 100))

    @property
    def inv_vocab(self):
        v = {0: "[PAD]", 1: "[CLS]", 2: "[SEP]", 3: "[MASK]", 4: "[BOS]"}
        for i in range(5, 100):
            v[i] = ("##p%d" % i) if (21 <= i < 90 and i % 2 == 1) else "w%d" % i
        return v


def _sentence_corpus(tmp_path, name="sent", n_docs=30):
    from megatron_llm_b200.data import indexed_dataset
    rng = np.random.RandomState(0)
    b = indexed_dataset.make_builder(str(tmp_path / f"{name}.bin"), "mmap", vocab_size=100)
    for d in range(n_docs):
        for s in range(rng.randint(2, 6)):
            b.add_item(torch.IntTensor(rng.randint(5, 90, size=rng.randint(4, 12)).tolist()))
        b.end_document()
    b.finalize(str(tmp_path / f"{name}.idx"))
    return str(tmp_path / name)


def test_bert_and_t5_datasets(tmp_path):
    from megatron_llm_b200.data import indexed_dataset
    from megatron_llm_b200.data.bert_dataset import BertDataset
    from megatron_llm_b200.data.t5_dataset import T5Dataset
    prefix = _sentence_corpus(tmp_path)
    ds = indexed_dataset.make_dataset(prefix, "mmap", True)
    tok = _FakeBertTok()
    bert = BertDataset("train", ds, prefix, None, 50, 0.15, 48, 0.1, 1234, True, tokenizer=tok)
    assert len(bert) >= 50
    n_masked = 0
    for i in range(20):
        s = bert[i]
        assert s["text"].shape == (48,) and s["text"][0] == tok.cls
        real = int(s["padding_mask"].sum())
        assert (s["text"][real:] == tok.pad).all() and s["text"][real - 1] == tok.sep
        pos = np.nonzero(s["loss_mask"])[0]
        assert len(pos) >= 1 and (s["labels"][pos] >= 5).all() and (s["labels"][s["loss_mask"] == 0] == -1).all()
        assert set(np.unique(s["types"][:real])) <= {0, 1}
        n_masked += len(pos)
        assert (bert[i]["text"] == s["text"]).all()          # deterministic per index
    assert n_masked > 20
    t5 = T5Dataset("train", ds, prefix, None, 50, 0.15, 48, 32, 0.1, 1234, tokenizer=tok)
    for i in range(20):
        s = t5[i]
        assert s["text_enc"].shape == (48,) and s["text_dec"].shape == (32,) and s["labels"].shape == (32,)
        assert s["text_dec"][0] == tok.bos_token_id
        n_dec = int(s["loss_mask"].sum())
        assert s["labels"][n_dec - 1] == tok.eos_token_id
        sent_enc = [t for t in s["text_enc"] if t >= 90]
        sent_dec = [t for t in s["text_dec"] if t >= 90]
        assert sent_enc == sent_dec and len(sent_enc) >= 1
        assert s["enc_mask"].shape == (48, 48) and s["dec_mask"].shape == (32, 32)
        assert s["enc_dec_mask"].shape == (32, 48)
        assert np.triu(s["dec_mask"], 1).sum() == 0
        assert (s["text_dec"][1:n_dec] == s["labels"][:n_dec - 1]).all()


def test_ict_dataset(tmp_path):
    from megatron_llm_b200.data import indexed_dataset
    from megatron_llm_b200.data.ict_dataset import ICTDataset
    prefix = _sentence_corpus(tmp_path, "blocks", 20)
    b = indexed_dataset.make_builder(str(tmp_path / "titles.bin"), "mmap", vocab_size=100)
    for d in range(20):
        b.add_item(torch.IntTensor([5 + d % 10, 6]))
        b.end_document()
    b.finalize(str(tmp_path / "titles.idx"))
    blocks = indexed_dataset.make_dataset(prefix, "mmap", True)
    titles = indexed_dataset.make_dataset(str(tmp_path / "titles"), "mmap", True)
    tok = _FakeBertTok()
    ict = ICTDataset("full", blocks, titles, prefix, 1, None, 64, 0.1, 1, tokenizer=tok)
    assert len(ict) > 0
    s = ict[0]
    assert s["query_tokens"].shape == (64,) and s["context_tokens"].shape == (64,)
    assert s["query_tokens"][0] == tok.cls and s["context_tokens"][0] == tok.cls
    assert s["block_data"].shape == (4,)
    toks, pad = ict.get_block(*s["block_data"][:3])
    assert toks.shape == (64,) and pad.sum() > 3


def test_mips_index():
    from megatron_llm_b200.data.realm_index import FaissMIPSIndex, OpenRetreivalDataStore
    store = OpenRetreivalDataStore("/tmp/_unused_embeds.pkl", load_from_path=False, rank=0)
    rng = np.random.RandomState(0)
    emb = rng.randn(200, 16).astype(np.float32)
    store.add_block_data(np.arange(1000, 1200), emb)
    index = FaissMIPSIndex(16, store, chunk=64)
    q = rng.randn(5, 16).astype(np.float32)
    scores, ids = index.search_mips_index(q, 4, reconstruct=False)
    ref = np.argsort(-(q @ np.float32(np.float16(emb)).T), axis=1)[:, :4] + 1000
    assert (ids == ref).all()
    assert index.search_mips_index(q, 4, reconstruct=True).shape == (5, 4, 16)


def test_instruction_preprocessing_with_sentencepiece_chatml_tokens(tmp_path):
    """preprocess_instruct_data.py with a SentencePiece model + ``--vocab_extra_ids_list "<|im_start|>,<|im_end|>"``:
    every turn is wrapped in the two ChatML tokens (single ids appended after Megatron's special tokens), roles are
    aligned with the tokens, and the validation metrics' instruct mask drops the scaffolding around each turn."""
    import preprocess_instruct_data
    from megatron_llm_b200.data import indexed_dataset
    from megatron_llm_b200.data.instruction_dataset import Role
    from megatron_llm_b200.tokenizer.tokenizer import _SentencePieceTokenizer
    from tests.test_weights_conversion import _tiny_sentencepiece_model
    vocab_file = _tiny_sentencepiece_model(tmp_path)
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta"]
    rows = [{"input": " ".join(words[(i + j) % 8] for j in range(3 + i % 5)),
             "output": " ".join(words[(i * 3 + j) % 8] for j in range(2 + i % 4)), "sys": "alpha beta"} for i in range(12)]
    _write_jsonl(tmp_path / "i.jsonl", rows)
    extra = "<|im_start|>,<|im_end|>"
    preprocess_instruct_data.main(["--input", str(tmp_path / "i.jsonl"), "--output_prefix", str(tmp_path / "inst"),
                                   "--system_key", "sys", "--tokenizer_type", "SentencePieceTokenizer", "--vocab_file",
                                   str(vocab_file), "--vocab_extra_ids_list", extra, "--workers", "1", "--chunk_size", "4"])
    tok = _SentencePieceTokenizer(str(vocab_file), vocab_extra_ids_list=extra)
    start, end = tok.vocab["<|im_start|>"], tok.vocab["<|im_end|>"]
    assert (start, end) == (85, 86)                           # 80 pieces + <CLS> <SEP> <EOD> <MASK> <PAD>, then ours
    text = indexed_dataset.make_dataset(str(tmp_path / "inst-text"), "mmap")
    role = indexed_dataset.make_dataset(str(tmp_path / "inst-role"), "mmap")
    assert len(text) == len(role) == 12
    for t, r in zip(text, role):
        t, r = t.tolist(), r.tolist()
        assert len(t) == len(r) and t[0] == start and t[-1] == end and t.count(start) == t.count(end) == 3
        assert r == sorted(r) and set(r) == {Role.system.value, Role.prompter.value, Role.assistant.value}
        assert all(r[i] != r[i - 1] for i in range(1, len(t)) if t[i] == start)      # roles switch at <|im_start|>
    decoded = tok.detokenize(text[0].tolist())
    assert decoded.count("<|im_start|>") == 3 and "alpha beta" in decoded

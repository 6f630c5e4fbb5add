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

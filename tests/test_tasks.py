"""Downstream tasks through tasks/main.py on CPU: MNLI + RACE fine-tuning (1 epoch), WIKITEXT103 / LAMBADA zero-shot."""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
from tests.dist_utils import free_port  # noqa: E402
from tests.test_pretrain_entries import WORDS  # noqa: E402


def _run(argv, timeout=900):
    env = dict(os.environ, MLB200_FORCE_CPU="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), RANK="0",
               WORLD_SIZE="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tasks/main.py")] + argv, env=env, capture_output=True,
                       text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def _vocab(tmp):
    (tmp / "vocab.txt").write_text("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS) + "\n")
    return str(tmp / "vocab.txt")


def _sent(rnd, n=6):
    return " ".join(rnd.choice(WORDS[:21]) for _ in range(n))


MODEL = ["--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4", "--seq_length", "48",
         "--max_position_embeddings", "48", "--micro_batch_size", "4", "--global_batch_size", "4", "--lr", "1e-3",
         "--lr_decay_style", "constant", "--log_interval", "2", "--eval_interval", "1000", "--eval_iters", "1",
         "--num_workers", "0", "--make_vocab_size_divisible_by", "8", "--hidden_dropout", "0.0",
         "--attention_dropout", "0.0"]


def test_mnli(tmp_path):
    rnd = random.Random(0)
    labels = ["contradiction", "entailment", "neutral"]
    for name, n in (("MNLI_train.tsv", 24), ("MNLI_dev.tsv", 8)):
        with open(tmp_path / name, "w") as f:
            f.write("\t".join(["index"] + ["c%d" % i for i in range(1, 8)] + ["sentence1", "sentence2", "gold"]) + "\n")
            for i in range(n):
                f.write("\t".join([str(i)] + ["x"] * 7 + [_sent(rnd), _sent(rnd), rnd.choice(labels)]) + "\n")
    out = _run(MODEL + ["--task", "MNLI", "--epochs", "1", "--train_data", str(tmp_path / "MNLI_train.tsv"),
                        "--valid_data", str(tmp_path / "MNLI_dev.tsv"), "--tokenizer_type", "BertWordPieceLowerCase",
                        "--vocab_file", _vocab(tmp_path)])
    assert "overall: correct / total = " in out and "/ 8 =" in out


def test_race(tmp_path):
    rnd = random.Random(1)
    for split, n in (("train", 3), ("dev", 2)):
        d = tmp_path / "RACE" / split
        d.mkdir(parents=True)
        with open(d / "a.txt", "w") as f:
            for i in range(n):
                f.write(json.dumps({"article": _sent(rnd, 20), "questions": [_sent(rnd, 4) + " _", _sent(rnd, 5)],
                                    "options": [[_sent(rnd, 2) for _ in range(4)] for _ in range(2)],
                                    "answers": ["A", "C"]}) + "\n")
    args = [a if a != "4" else "2" for a in MODEL]
    out = _run(args + ["--task", "RACE", "--epochs", "1", "--train_data", str(tmp_path / "RACE" / "train"),
                       "--valid_data", str(tmp_path / "RACE" / "dev"), "--tokenizer_type", "BertWordPieceLowerCase",
                       "--vocab_file", _vocab(tmp_path)])
    assert "overall: correct / total = " in out and "/ 4 =" in out


def test_zeroshot_gpt(tmp_path):
    rnd = random.Random(2)
    (tmp_path / "wiki.test.tokens").write_text(" ".join(_sent(rnd, 8) + " ." for _ in range(60)))
    with open(tmp_path / "lambada.jsonl", "w") as f:
        for i in range(6):
            f.write(json.dumps({"text": _sent(rnd, 10)}) + "\n")
    common = [a for a in MODEL] + ["--tokenizer_type", "NullTokenizer", "--vocab_file", "64"]
    out = _run(common + ["--task", "WIKITEXT103", "--valid_data", str(tmp_path / "wiki.test.tokens"),
                         "--overlapping_eval", "16"])
    assert "validation results on WIKITEXT103" in out and "adjusted ppl" in out
    out = _run(common + ["--task", "LAMBADA", "--valid_data", str(tmp_path / "lambada.jsonl")])
    assert "validation results on LAMBADA" in out and "total examples: 6" in out

"""Downstream tasks through tasks/main.py on CPU: MNLI + RACE fine-tuning (1 epoch), WIKITEXT103 / LAMBADA zero-shot."""
import json
import os
import random
import subprocess
import sys

import pytest

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


def test_zeroshot_gpt_model_parallel(tmp_path):
    """WIKITEXT103 perplexity and LAMBADA accuracy under TP2 equal the single-rank numbers (same master-seeded CPU
    initialisation); PP2 seeds its stages differently, so only closeness is checked there.  PP2 covers the shape-less ``recv_forward()`` of the evaluation loop and the orderly exit
    of a first stage that finishes ahead of the last one."""
    import re
    from tests.test_pretrain_entries import _run_ranks
    rnd = random.Random(2)
    (tmp_path / "wiki.test.tokens").write_text(" ".join(_sent(rnd, 8) + " ." for _ in range(60)))
    with open(tmp_path / "lambada.jsonl", "w") as f:
        for i in range(6):
            f.write(json.dumps({"text": _sent(rnd, 10)}) + "\n")
    common = [a for a in MODEL] + ["--tokenizer_type", "NullTokenizer", "--vocab_file", "64", "--seed", "5",
                                   "--use_cpu_initialization"]
    wiki = ["--task", "WIKITEXT103", "--valid_data", str(tmp_path / "wiki.test.tokens"), "--overlapping_eval", "16"]
    lambada = ["--task", "LAMBADA", "--valid_data", str(tmp_path / "lambada.jsonl")]

    def numbers(out, key):
        return [float(x) for x in re.findall(key + r": ([0-9.E+-]+)", out)]
    ref_w = numbers(_run(common + wiki), "avg loss")
    ref_l = numbers(_run(common + lambada), "number correct")
    assert len(ref_w) == 1 and len(ref_l) == 1
    for layout in (["--tensor_model_parallel_size", "2"], ["--pipeline_model_parallel_size", "2"]):
        got_w = numbers(_run_ranks("tasks/main.py", common + wiki + layout, 2), "avg loss")
        got_l = numbers(_run_ranks("tasks/main.py", common + lambada + layout, 2), "number correct")
        if "--tensor_model_parallel_size" in layout:
            assert got_w == ref_w and got_l == ref_l, layout
        else:
            assert got_w == pytest.approx(ref_w, rel=5e-3) and len(got_l) == 1, layout


def test_qa_match_utils():
    from tasks.orqa.unsupervised.qa_utils import calculate_matches, exact_match_score, has_answer
    from tasks.orqa.unsupervised.tokenizers import SimpleTokenizer
    tok = SimpleTokenizer()
    assert tok.tokenize("Hello, World! 42").words(uncased=True) == ["hello", ",", "world", "!", "42"]
    assert has_answer(["new york"], "He moved to New York in 1999.", tok, "string")
    assert not has_answer(["york new"], "He moved to New York in 1999.", tok, "string")
    assert has_answer([r"19\d\d"], "He moved to New York in 1999.", tok, "regex")
    assert exact_match_score("The Eiffel Tower!", "eiffel tower")
    docs = {1: ("paris is the capital of france", "France"), 2: ("berlin is in germany", "Germany")}
    stats = calculate_matches(docs, [["paris"], ["rome"]], [([2, 1], [0.9, 0.8]), ([1, 2], [0.5, 0.4])], 1, "string")
    assert stats.top_k_hits == [0, 1] and stats.questions_doc_hits == [[False, True], [False, False]]


def _ctx(rnd):
    return {"title": _sent(rnd, 2), "text": _sent(rnd, 12)}


def test_retriever_finetune_and_eval(tmp_path):
    """RET-FINETUNE-NQ for one epoch (with hard negatives), then RETRIEVER-EVAL on the saved bi-encoder."""
    rnd = random.Random(3)
    rows = [{"question": _sent(rnd, 5) + "?", "answers": [rnd.choice(WORDS[:21])], "positive_ctxs": [_ctx(rnd)],
             "negative_ctxs": [_ctx(rnd) for _ in range(3)], "hard_negative_ctxs": [_ctx(rnd) for _ in range(2)]}
            for _ in range(8)]
    (tmp_path / "nq-train.json").write_text(json.dumps(rows))
    (tmp_path / "nq-dev.json").write_text(json.dumps(rows[:4]))
    ckpt = tmp_path / "ckpt"
    base = [a for a in MODEL] + ["--tokenizer_type", "BertWordPieceLowerCase", "--vocab_file", _vocab(tmp_path),
                                 "--retriever_seq_length", "48", "--biencoder_projection_dim", "16",
                                 "--retriever_report_topk_accuracies", "1", "2", "--bert_load", "none_given"]
    base = base[:-2]
    out = _run(base + ["--task", "RET-FINETUNE-NQ", "--epochs", "1", "--train_data", str(tmp_path / "nq-train.json"),
                       "--valid_data", str(tmp_path / "nq-dev.json"), "--train_with_neg", "--train_hard_neg", "2",
                       "--val_av_rank_hard_neg", "2", "--val_av_rank_other_neg", "2", "--eval_micro_batch_size", "2",
                       "--save", str(ckpt), "--save_interval", "1000"])
    assert "epoch:0|rank = " in out and "top1_acc" in out
    # evidence + questions for the zero-shot style evaluation of the fine-tuned retriever
    with open(tmp_path / "evidence.tsv", "w") as f:
        f.write("id\ttext\ttitle\n")
        for i in range(12):
            f.write("{}\t{}\t{}\n".format(i + 1, _sent(rnd, 14), _sent(rnd, 2)))
    with open(tmp_path / "nq-dev.tsv", "w") as f:
        for i in range(3):
            f.write("{}\t{}\n".format(_sent(rnd, 5), repr([rnd.choice(WORDS[:21])])))
    out = _run(base + ["--task", "RETRIEVER-EVAL", "--load", str(ckpt), "--evidence_data_path",
                       str(tmp_path / "evidence.tsv"), "--embedding_path", str(tmp_path / "evidence_embeds.pkl"),
                       "--qa_data_dev", str(tmp_path / "nq-dev.tsv"), "--faiss_topk_retrievals", "5",
                       "--indexer_batch_size", "4", "--indexer_log_interval", "1", "--no_load_optim", "--no_load_rng"])
    assert "DEV SET RESULTS" in out and "top-1:" in out
    assert os.path.exists(tmp_path / "evidence_embeds.pkl")


def test_msdp_metrics_and_preprocessing(tmp_path):
    from tasks.msdp.metrics import F1Metric, normalize_answer
    from tasks.msdp import preprocessing as pp
    assert normalize_answer("The cat, an animal!") == "cat animal"
    p, r, f = F1Metric.compute_all_pairs(["the cat sat", "anything", ""], ["a cat sat down", "", "dog"])
    assert abs(p - 0.5) < 1e-6 and abs(r - 1 / 3) < 1e-6 and f > 0
    wow = [{"chosen_topic": "Cats", "dialog": [
        {"speaker": "0_Apprentice", "text": "I love cats"},
        {"speaker": "1_Wizard", "text": "Cats are small felines", "checked_sentence": {"k": "The cat is a feline."},
         "checked_passage": {"p": "Cat"}},
        {"speaker": "0_Apprentice", "text": "Really?"},
        {"speaker": "1_Wizard", "text": "Yes!", "checked_sentence": {}, "checked_passage": {}}]}]
    (tmp_path / "wow.json").write_text(json.dumps(wow))
    pp.process_wow_dataset(str(tmp_path / "wow.json"), str(tmp_path / "proc.txt"), str(tmp_path / "k.txt"),
                           str(tmp_path / "r.txt"))
    rows = (tmp_path / "proc.txt").read_text().strip().split("\n")
    assert rows[0].split("\t") == ["Cat", "I love cats.", "The cat is a feline.", "Cats are small felines."]
    assert rows[1].split("\t") == ["Cats", "I love cats. [SEP] Cats are small felines. [SEP] Really?",
                                   "no_passages_used", "Yes!"]
    (tmp_path / "gen.txt").write_text("generated one<|endoftext|>\nsecond\n")
    pp.prepare_input_for_response_generation(str(tmp_path / "proc.txt"), str(tmp_path / "gen.txt"),
                                             str(tmp_path / "inp.txt"))
    assert (tmp_path / "inp.txt").read_text().split("\n")[0].split("\t")[2] == "generated one"
    by_topic, dialogs, examples = pp.get_database(str(tmp_path / "proc.txt"), str(tmp_path / "proc.txt"), "wow_seen")
    assert list(by_topic) == ["Cat"] and len(examples) == 1
    from tasks.msdp.evaluate import evaluate_f1
    import megatron_llm_b200.utils as u
    (tmp_path / "guess.txt").write_text("the cat is a feline\nwhatever\n")
    p, r, f = evaluate_f1(str(tmp_path / "guess.txt"), str(tmp_path / "k.txt"))
    assert f > 0.9


def test_msdp_prompting_end_to_end(tmp_path):
    """tasks/msdp/main.py MSDP-PROMPT (knowledge and response stages) with a byte-level GPT-2 tokenizer built here,
    single rank and TP2 (same padded vocabulary, master-seeded CPU init -> same generations: the model must be built
    with full-vocabulary logits), PP2, then MSDP-EVAL-F1 on files."""
    from megatron_llm_b200.tokenizer.gpt2_tokenization import bytes_to_unicode
    from tests.test_pretrain_entries import _run_ranks
    vocab = {c: i for i, c in enumerate(bytes_to_unicode().values())}
    merges = [("h", "e"), ("l", "l"), ("Ġ", "t"), ("Ġ", "a")]
    for a, b in merges:
        vocab[a + b] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    (tmp_path / "k_prompts.jsonl").write_text(json.dumps({"cats do cats purr": [
        "( do cats purr ) cats => cats purr when happy", "( what is a cat ) cats => a small animal"]}) + "\n")
    (tmp_path / "r_prompts.txt").write_text("Topic: cats. User says: hi. We know that: cats purr. System replies: yes\n")
    (tmp_path / "test.txt").write_text("cats\thello [SEP] do cats purr\tcats purr\tyes they do\n")
    model = ["--model_name", "gpt", "--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4",
             "--seq_length", "256", "--max_position_embeddings", "256", "--micro_batch_size", "1", "--tokenizer_type",
             "GPT2BPETokenizer", "--vocab_file", str(tmp_path / "vocab.json"), "--merge_file",
             str(tmp_path / "merges.txt"), "--out_seq_length", "8", "--seed", "3", "--use_cpu_initialization"]
    results = {}
    for name, world, extra in (("w1", 1, ["--make_vocab_size_divisible_by", "16"]),
                               ("tp2", 2, ["--make_vocab_size_divisible_by", "8", "--tensor_model_parallel_size", "2"]),
                               ("pp2", 2, ["--make_vocab_size_divisible_by", "16", "--pipeline_model_parallel_size",
                                           "2"])):
        for ptype, pfile in (("knowledge", "k_prompts.jsonl"), ("response", "r_prompts.txt")):
            out = tmp_path / f"{name}_{ptype}.out"
            _run_ranks("tasks/msdp/main.py", model + extra + [
                "--task", "MSDP-PROMPT", "--prompt_type", ptype, "--prompt_file", str(tmp_path / pfile),
                "--sample_input_file", str(tmp_path / "test.txt"), "--sample_output_file", str(out),
                "--num_prompt_examples", "2"], world)
            results[name, ptype] = out.read_bytes()
            assert results[name, ptype].endswith(b"\n") and results[name, ptype].count(b"\n") == 1
    assert results["tp2", "knowledge"] == results["w1", "knowledge"]
    assert results["tp2", "response"] == results["w1", "response"]
    (tmp_path / "guess.txt").write_text("cats purr when happy\n")
    (tmp_path / "answer.txt").write_text("cats purr\n")
    out = _run_ranks("tasks/msdp/main.py", model + ["--task", "MSDP-EVAL-F1", "--guess_file", str(tmp_path / "guess.txt"),
                                                    "--answer_file", str(tmp_path / "answer.txt")], 1)
    assert "Precision: 0.5000; recall: 1.0000; f1: 0.6667" in out

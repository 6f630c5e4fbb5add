"""pretrain_bert.py / pretrain_t5.py / pretrain_ict.py end to end on CPU: preprocess -> 2 training iterations."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
from tests.dist_utils import free_port  # noqa: E402

WORDS = ["the", "cat", "dog", "sat", "ran", "on", "mat", "road", "fast", "slow", "red", "blue", "big", "small", "and",
         "then", "over", "under", "house", "tree", "play", "##ing", "##ed", "##s", "."]


def _make_corpus(tmp, n_docs=60):
    import random
    rnd = random.Random(0)
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS
    (tmp / "vocab.txt").write_text("\n".join(vocab) + "\n")
    with open(tmp / "corpus.jsonl", "w") as f:
        for d in range(n_docs):
            sents = []
            for s in range(rnd.randint(3, 6)):
                ws = [rnd.choice(WORDS[:21]) for _ in range(rnd.randint(4, 9))]
                sents.append(" ".join(ws).capitalize() + " .")
            f.write(json.dumps({"text": " ".join(sents), "title": "The " + rnd.choice(WORDS[:21])}) + "\n")
    return tmp / "vocab.txt", tmp / "corpus.jsonl"


def _run(script, argv, timeout=900):
    env = dict(os.environ, MLB200_FORCE_CPU="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), RANK="0",
               WORLD_SIZE="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)] + argv, env=env, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def _run_ranks(script, argv, world, timeout=600):
    """``world`` CPU/gloo ranks of an entry point (what torchrun would start); returns rank 0's stdout."""
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, MLB200_FORCE_CPU="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r),
                   WORLD_SIZE=str(world), LOCAL_RANK=str(r), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, script)] + argv, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o[-2000:] + "\n" + e[-3000:]
    return "".join(o for o, _ in outs)


def _preprocess(tmp, keys=("text",), extra=()):
    vocab, corpus = _make_corpus(tmp)
    _run("tools/preprocess_data.py", ["--input", str(corpus), "--output_prefix", str(tmp / "data"), "--json_keys",
                                      *keys, "--tokenizer_type", "BertWordPieceLowerCase", "--vocab_file", str(vocab),
                                      "--split_sentences", "--workers", "1", "--chunk_size", "8", *extra])
    return vocab


COMMON = ["--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4", "--micro_batch_size", "2",
          "--global_batch_size", "2", "--train_iters", "2", "--lr", "1e-3", "--lr_decay_style", "constant",
          "--log_interval", "1", "--eval_iters", "1", "--eval_interval", "100", "--data_impl", "mmap", "--split",
          "90,5,5", "--num_workers", "0", "--hidden_dropout", "0.0", "--attention_dropout", "0.0"]


def test_pretrain_bert(tmp_path):
    vocab = _preprocess(tmp_path)
    # (also exercises --profile: iteration 1 of 2 goes through torch.profiler and must leave a trace + kernel table)
    out = _run("pretrain_bert.py", COMMON + ["--seq_length", "48", "--max_position_embeddings", "48", "--vocab_file",
                                             str(vocab), "--data_path", str(tmp_path / "data_text_sentence"),
                                             "--make_vocab_size_divisible_by", "8", "--profile",
                                             "--profile_step_start", "1", "--profile_step_end", "2", "--profile_dir",
                                             str(tmp_path / "trace")])
    assert "lm loss" in out and "sop loss" in out and "iteration        2/" in out
    assert (tmp_path / "trace" / "trace_rank0_it1-2.json").stat().st_size > 1000
    assert (tmp_path / "trace" / "kernels_rank0_it1-2.txt").exists()


def test_pretrain_t5(tmp_path):
    vocab = _preprocess(tmp_path)
    out = _run("pretrain_t5.py", COMMON + ["--encoder_seq_length", "48", "--decoder_seq_length", "32",
                                           "--max_position_embeddings", "48", "--vocab_file", str(vocab),
                                           "--vocab_extra_ids", "100", "--data_path",
                                           str(tmp_path / "data_text_sentence"), "--make_vocab_size_divisible_by", "8",
                                           "--kv_channels", "8", "--ffn_hidden_size", "64"])
    assert "lm loss" in out and "iteration        2/" in out


def test_pretrain_ict(tmp_path):
    vocab = _preprocess(tmp_path, keys=("text", "title"))
    out = _run("pretrain_ict.py", COMMON + ["--seq_length", "48", "--max_position_embeddings", "48", "--vocab_file",
                                            str(vocab), "--data_path", str(tmp_path / "data_text_sentence"),
                                            "--titles_data_path", str(tmp_path / "data_title_sentence"),
                                            "--make_vocab_size_divisible_by", "8", "--biencoder_projection_dim", "16",
                                            "--query_in_block_prob", "0.1", "--retriever_report_topk_accuracies", "1",
                                            "2", "--use_one_sent_docs"])
    assert "top1_acc" in out and "iteration        2/" in out


def test_pretrain_bert_tensor_parallel(tmp_path):
    """BERT with TP=2 + sequence parallelism: only the TP-rank-0 of every (DP, PP) coordinate builds datasets, so the
    wait for rank 0's index map must not be a world barrier (it once was: TP > 1 dead-locked in the dataset build)."""
    vocab = _preprocess(tmp_path)
    out = _run_ranks("pretrain_bert.py", COMMON + ["--seq_length", "48", "--max_position_embeddings", "48",
                                                   "--vocab_file", str(vocab), "--data_path",
                                                   str(tmp_path / "data_text_sentence"), "--make_vocab_size_divisible_by",
                                                   "8", "--tensor_model_parallel_size", "2", "--sequence_parallel"], 2)
    assert "lm loss" in out and "iteration        2/" in out


def test_pretrain_t5_pipeline_parallel(tmp_path):
    """T5 with the encoder on stage 0 and the decoder on stage 1 (``--pipeline_model_parallel_split_rank 1``);
    ``--decoder_num_layers`` is left to default to ``--num_layers`` (the reference fails on the unset value)."""
    vocab = _preprocess(tmp_path)
    out = _run_ranks("pretrain_t5.py", COMMON + ["--encoder_seq_length", "48", "--decoder_seq_length", "32",
                                                 "--max_position_embeddings", "48", "--vocab_file", str(vocab),
                                                 "--vocab_extra_ids", "100", "--data_path",
                                                 str(tmp_path / "data_text_sentence"), "--make_vocab_size_divisible_by",
                                                 "8", "--kv_channels", "8", "--ffn_hidden_size", "64",
                                                 "--pipeline_model_parallel_size", "2",
                                                 "--pipeline_model_parallel_split_rank", "1"], 2)
    assert "lm loss" in out and "iteration        2/" in out

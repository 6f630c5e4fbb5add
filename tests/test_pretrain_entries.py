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


LLAMA_TINY = ["--model_name", "llama2", "--num_layers", "4", "--hidden_size", "32", "--num_attention_heads", "4",
              "--num_attention_heads_kv", "2", "--ffn_hidden_size", "64", "--seq_length", "16",
              "--max_position_embeddings", "16", "--micro_batch_size", "2", "--global_batch_size", "8",
              "--tokenizer_type", "NullTokenizer", "--vocab_file", "100", "--make_vocab_size_divisible_by", "8",
              "--train_iters", "6", "--lr", "1e-2", "--min_lr", "1e-3", "--lr_decay_style", "cosine",
              "--lr_warmup_iters", "2", "--weight_decay", "0.01", "--clip_grad", "1.0", "--hidden_dropout", "0.0",
              "--attention_dropout", "0.0", "--use_rms_norm", "--glu_activation", "swiglu",
              "--position_embedding_type", "rotary", "--no_bias_gelu_fusion", "--no_tie_embed_logits", "--eval_iters",
              "2", "--eval_interval", "3", "--log_interval", "1", "--seed", "7", "--split", "8,1,1", "--num_workers",
              "0"]


def _loss_lines(out):
    import re
    train = [float(m.group(1)) for m in re.finditer(r"iteration\s+\d+/.*?lm loss: ([0-9.E+-]+)", out)]
    final = [float(m.group(1)) for m in re.finditer(r"for test data \| lm loss value: ([0-9.E+-]+)", out)]
    return train, final


def test_finetune_entry_on_preprocessed_gpt_and_instruction_data(tmp_path):
    """finetune.py end to end on data written by tools/preprocess_data.py / preprocess_instruct_data.py: a blended GPT
    corpus (train / valid / test split, cosine schedule with warm-up, periodic + final evaluation), the same run on two
    data-parallel ranks with loader workers (identical losses: the samplers shard one global batch), and instruction
    tuning with variable sequence lengths under TP2 x PP2."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import preprocess_data
    import preprocess_instruct_data
    rows = [{"text": " ".join(str((i * 7 + j) % 50) for j in range(5 + i % 23))} for i in range(200)]
    with open(tmp_path / "c.jsonl", "w") as f:
        f.writelines(json.dumps(r) + "\n" for r in rows)
    for name in ("a", "b"):
        preprocess_data.main(["--input", str(tmp_path / "c.jsonl"), "--output_prefix", str(tmp_path / name),
                              "--tokenizer_type", "NullTokenizer", "--vocab_file", "100", "--workers", "1",
                              "--chunk_size", "8", "--append_eod"])
    irows = [{"input": " ".join(str((i + j) % 40) for j in range(3 + i % 9)),
              "output": " ".join(str((i * 3 + j) % 40) for j in range(2 + i % 7)), "sys": "9 9"} for i in range(120)]
    with open(tmp_path / "i.jsonl", "w") as f:
        f.writelines(json.dumps(r) + "\n" for r in irows)
    preprocess_instruct_data.main(["--input", str(tmp_path / "i.jsonl"), "--output_prefix", str(tmp_path / "inst"),
                                   "--system_key", "sys", "--tokenizer_type", "NullTokenizer", "--vocab_file", "100",
                                   "--workers", "1", "--chunk_size", "4"])
    blend = ["--data_path", "0.3", str(tmp_path / "a_text_document"), "0.7", str(tmp_path / "b_text_document")]
    one = _run("finetune.py", LLAMA_TINY + blend)
    train1, final1 = _loss_lines(one)
    assert len(train1) == 6 and len(final1) == 1 and train1[-1] < train1[0]
    assert one.count("validation loss at iteration") == 2           # eval_interval 3 over 6 iterations
    two = _run_ranks("finetune.py", LLAMA_TINY + blend + ["--num_workers", "2"], 2)
    train2, final2 = _loss_lines(two)
    assert train2[:6] == pytest.approx(train1, rel=1e-4) and final2[0] == pytest.approx(final1[0], rel=1e-4)
    inst = _run_ranks("finetune.py", LLAMA_TINY + ["--data_path", str(tmp_path / "inst"), "--data_type", "instruction",
                                                   "--variable_seq_lengths", "--tensor_model_parallel_size", "2",
                                                   "--pipeline_model_parallel_size", "2", "--metrics", "all"], 4)
    train3, final3 = _loss_lines(inst)
    assert len(train3) >= 6 and len(final3) >= 1 and all(l == l and l > 0 for l in train3)
    assert "instruct accuracy" in inst and "count loss mask" in inst


def test_bert_checkpoint_reshards_to_tp2_and_pp2(tmp_path):
    """tools/checkpoint_util.py --model_type BERT: pooler, LM head (incl. its vocab-parallel output bias, which the
    reference's tool forgets) and binary head travel; continuing from the resharded checkpoint under TP2 and under PP2
    gives the single-rank LM / SOP losses."""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util
    vocab = _preprocess(tmp_path)
    base = COMMON + ["--seq_length", "48", "--max_position_embeddings", "48", "--vocab_file", str(vocab), "--data_path",
                     str(tmp_path / "data_text_sentence"), "--make_vocab_size_divisible_by", "8", "--seed", "11"]
    _run("pretrain_bert.py", base + ["--save", str(tmp_path / "ck"), "--save_interval", "2"])
    cont = list(base)
    cont[cont.index("--train_iters") + 1] = "4"
    cont += ["--finetune", "--no_load_optim", "--no_load_rng"]

    def losses(out):
        return re.findall(r"lm loss: ([0-9.E+-]+).*?sop loss: ([0-9.E+-]+)", out)
    ref = losses(_run("pretrain_bert.py", cont + ["--load", str(tmp_path / "ck")]))
    assert len(ref) == 4
    for tp, pp in ((2, 1), (1, 2)):
        dst = tmp_path / f"ck_tp{tp}_pp{pp}"
        checkpoint_util.main(["--model_type", "BERT", "--load_dir", str(tmp_path / "ck"), "--save_dir", str(dst),
                              "--target_tensor_parallel_size", str(tp), "--target_pipeline_parallel_size", str(pp)])
        out = _run_ranks("pretrain_bert.py", cont + ["--load", str(dst), "--tensor_model_parallel_size", str(tp),
                                                     "--pipeline_model_parallel_size", str(pp)], 2)
        got = losses(out)
        assert [tuple(map(float, g)) for g in got] == pytest.approx([tuple(map(float, r)) for r in ref], rel=2e-4), \
            (tp, pp, got, ref)

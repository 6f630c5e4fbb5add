"""Text generation on CPU: incremental KV-cache decoding must equal re-running the full prefix (greedy), for a GQA +
RoPE Llama and a learned-position GPT; scoring mode, beam search, the REST request validation and the server round trip."""
import json
import os
import sys
import threading
import urllib.request

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
from tests.dist_utils import run_distributed  # noqa: E402

LLAMA = ["--model_name", "llama2", "--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4",
         "--num_attention_heads_kv", "2", "--ffn_hidden_size", "64", "--use_rms_norm", "--glu_activation", "swiglu",
         "--position_embedding_type", "rotary", "--no_tie_embed_logits", "--no_bias_gelu_fusion"]
GPT = ["--model_name", "gpt", "--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4"]
COMMON = ["--seq_length", "32", "--max_position_embeddings", "32", "--micro_batch_size", "2", "--tokenizer_type",
          "NullTokenizer", "--vocab_file", "64", "--make_vocab_size_divisible_by", "8", "--train_iters", "1", "--lr",
          "1e-3", "--hidden_dropout", "0.0", "--attention_dropout", "0.0", "--seed", "3"]


def _setup(model_args, world):
    import finetune
    from megatron_llm_b200.initialize import initialize_megatron
    initialize_megatron(extra_args_provider=finetune.extra_args,
                        args_list=model_args + COMMON + ["--tensor_model_parallel_size", str(world)])
    model = finetune.model_provider(True, True)
    from megatron_llm_b200.utils import unwrap_model
    unwrap_model(model).parallel_output = False
    model.eval()
    return model


def _full_forward_greedy(model, prompt, n_new):
    """Oracle: no cache, recompute the whole prefix for every new token."""
    toks = list(prompt)
    for _ in range(n_new):
        t = torch.tensor([toks])
        pos = torch.arange(len(toks)).unsqueeze(0)
        mask = torch.tril(torch.ones(1, 1, len(toks), len(toks))) < 0.5
        with torch.no_grad():
            logits = model(t, pos, mask)
        toks.append(int(logits[0, -1].float().argmax()))
    return toks


def _generation_worker(rank, world, model_args):
    from megatron_llm_b200.text_generation import beam_search_and_post_process, generate_and_post_process
    model = _setup(model_args, world)
    prompts = ["5 9 13 2", "7 7 1 40 3 22"]
    texts, segments, logprobs, tokens = generate_and_post_process(
        model, prompts=prompts, tokens_to_generate=6, return_output_log_probs=True, top_k_sampling=1,
        use_eod_token_for_early_termination=False)
    for i, p in enumerate(prompts):
        prompt_ids = [int(x) for x in p.split()]
        expect = _full_forward_greedy(model, prompt_ids, 6)
        got = [int(x) for x in texts[i].split()]
        assert got[:len(expect)] == expect, (got, expect)
        assert len(logprobs[i]) == len(got) - 1 or len(logprobs[i]) >= len(expect) - 1
    # scoring mode: log-probs of a given text, no generation
    texts2, _, lp, _ = generate_and_post_process(model, prompts=prompts, tokens_to_generate=0,
                                                 return_output_log_probs=True)
    assert texts2 == prompts and all(v <= 0 for v in lp[0])
    t = torch.tensor([[5, 9, 13, 2]])
    with torch.no_grad():
        logits = model(t, torch.arange(4).unsqueeze(0), torch.tril(torch.ones(1, 1, 4, 4)) < 0.5)
    ref = torch.log_softmax(logits[0].float(), -1)[torch.arange(3), t[0, 1:]]
    assert torch.allclose(torch.tensor(lp[0][:3]), ref, atol=1e-4)
    # beam search returns beam_size hypotheses that extend the prompt
    out = beam_search_and_post_process(model, prompts=[prompts[0]], tokens_to_generate=5, beam_size=3, stop_token=63,
                                       num_return_gen=3, length_penalty=1.0)
    if out is not None:
        beams, _, scores = out
        assert len(beams) == 3 and all(b.startswith("5 9 13 2") for b in beams)
        assert scores[0] >= scores[-1]


@pytest.mark.parametrize("name,args,world", [("llama_gqa_rope", LLAMA, 1), ("gpt_learned_pos", GPT, 1),
                                             ("llama_tp2", LLAMA, 2)])
def test_kv_cache_decode_matches_full_forward(name, args, world):
    run_distributed(_generation_worker, world, args)


def _layout_generation_worker(rank, world, tp, pp, ckpt, out_path, save):
    """Greedy generation + beam search from a (resharded) checkpoint under tp x pp; the first-stage TP-rank-0 process
    holds the result (the reference broadcasts the tokens from the last stage back to it)."""
    import finetune
    from megatron_llm_b200 import get_args
    from megatron_llm_b200.checkpointing import load_checkpoint, save_checkpoint
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models.enums import ModelType
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.text_generation import beam_search_and_post_process, generate_and_post_process
    from megatron_llm_b200.training import get_model
    from megatron_llm_b200.utils import unwrap_model
    extra = ["--save", ckpt, "--save_interval", "1"] if save else ["--load", ckpt, "--no_load_optim", "--no_load_rng", "--finetune"]
    initialize_megatron(extra_args_provider=finetune.extra_args,
                        args_list=LLAMA + COMMON + ["--tensor_model_parallel_size", str(tp),
                                                    "--pipeline_model_parallel_size", str(pp)] + extra)

    def provider(pre_process=True, post_process=True):
        m = finetune.model_provider(pre_process, post_process)
        unwrap_model(m).parallel_output = False
        return m
    model = get_model(provider, ModelType.encoder_or_decoder, wrap_with_ddp=False, args=get_args())
    if save:
        save_checkpoint(1, model, None, None)
    else:
        load_checkpoint(model, None, None)
    model = model[0]
    model.eval()
    prompts = ["5 9 13 2", "7 7 1 40 3 22"]
    out = generate_and_post_process(model, prompts=prompts, tokens_to_generate=6, return_output_log_probs=True,
                                    top_k_sampling=1, use_eod_token_for_early_termination=False)
    beams = beam_search_and_post_process(model, prompts=[prompts[0]], tokens_to_generate=5, beam_size=3,
                                         stop_token=63, num_return_gen=3, length_penalty=1.0)
    if ps.is_pipeline_first_stage() and ps.get_tensor_model_parallel_rank() == 0:
        texts, _, logprobs, _ = out                           # (None on the other stages, as in the reference)
        with open(out_path, "w") as f:
            json.dump({"texts": texts, "logprobs": logprobs, "beams": beams[0],
                       "beam_scores": [float(x) for x in beams[2]]}, f)


@pytest.mark.parametrize("tp,pp", [(1, 2), (2, 2)], ids=["pp2", "tp2_pp2"])
def test_generation_under_pipeline_parallelism(tmp_path, tp, pp):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util
    ckpt = tmp_path / "ckpt"
    run_distributed(_layout_generation_worker, 1, 1, 1, str(ckpt), str(tmp_path / "ref.json"), True)
    dst = tmp_path / "resharded"
    checkpoint_util.main(["--model_type", "llama2", "--load_dir", str(ckpt), "--save_dir", str(dst),
                          "--target_tensor_parallel_size", str(tp), "--target_pipeline_parallel_size", str(pp),
                          "--true_vocab_size", "64"])
    run_distributed(_layout_generation_worker, tp * pp, tp, pp, str(dst), str(tmp_path / "got.json"), False)
    ref, got = json.load(open(tmp_path / "ref.json")), json.load(open(tmp_path / "got.json"))
    assert got["texts"] == ref["texts"] and got["beams"] == ref["beams"], (got, ref)
    for a, b in zip(got["logprobs"], ref["logprobs"]):
        assert a == pytest.approx(b, abs=1e-4)
    assert got["beam_scores"] == pytest.approx(ref["beam_scores"], abs=1e-4)


def test_request_validation():
    from megatron_llm_b200.text_generation_server import RequestError, parse_request
    ok = parse_request({"prompts": ["a"], "tokens_to_generate": 4, "top_k": 2, "temperature": 0.7})
    assert ok["top_k"] == 2 and ok["tokens_to_generate"] == 4
    for bad in ({}, {"prompts": "x"}, {"prompts": []}, {"prompts": ["a"] * 129}, {"prompts": ["a"], "max_len": 3},
                {"prompts": ["a"], "tokens_to_generate": -1}, {"prompts": ["a"], "tokens_to_generate": 0},
                {"prompts": ["a"], "top_k": 3, "top_p": 0.5}, {"prompts": ["a"], "top_p_decay": 0.5},
                {"prompts": ["a", "b"], "beam_width": 2}, {"prompts": [""]}, {"prompts": ["a"], "temperature": 0.0}):
        with pytest.raises(RequestError):
            parse_request(bad)


def _server_worker(rank, world, port):
    from megatron_llm_b200.text_generation_server import MegatronServer
    model = _setup(LLAMA, 1)
    server = MegatronServer(model)
    th = threading.Thread(target=server.run, kwargs={"url": "127.0.0.1", "port": port}, daemon=True)
    th.start()
    import time
    body = json.dumps({"prompts": ["5 9 13 2"], "tokens_to_generate": 4, "top_k": 1, "no_log": True}).encode()
    for _ in range(50):
        try:
            req = urllib.request.Request(f"http://127.0.0.1:{port}/api", data=body, method="PUT",
                                         headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=30) as r:
                out = json.loads(r.read())
            break
        except (ConnectionRefusedError, urllib.error.URLError):
            time.sleep(0.2)
    assert out["text"][0].startswith("5 9 13 2") and len(out["text"][0].split()) == 8
    bad = urllib.request.Request(f"http://127.0.0.1:{port}/api", data=b'{"prompts": []}', method="PUT")
    try:
        urllib.request.urlopen(bad, timeout=30)
        raise AssertionError("expected HTTP 400")
    except urllib.error.HTTPError as e:
        assert e.code == 400
    with urllib.request.urlopen(f"http://127.0.0.1:{port}/", timeout=30) as r:
        assert b"<html" in r.read().lower()
    server.httpd.shutdown()


def test_rest_server_roundtrip():
    from tests.dist_utils import free_port
    run_distributed(_server_worker, 1, free_port())


def test_server_tool_with_tensor_parallel_ranks():
    """tools/run_text_generation_server.py as two TP ranks (what torchrun starts): rank 0 serves HTTP and broadcasts the
    op code, rank 1 sits in the generate / beam-search loop.  Greedy, beam and top-p requests must all come back."""
    import subprocess
    import time
    from tests.dist_utils import free_port
    port, http = free_port(), free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, MLB200_FORCE_CPU="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r),
                   WORLD_SIZE="2", LOCAL_RANK=str(r), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "run_text_generation_server.py")]
                                      + LLAMA + COMMON + ["--tensor_model_parallel_size", "2", "--port", str(http)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT))

    def put(body):
        data = json.dumps(body).encode()
        for _ in range(300):
            try:
                req = urllib.request.Request(f"http://127.0.0.1:{http}/api", data=data, method="PUT",
                                             headers={"Content-Type": "application/json"})
                with urllib.request.urlopen(req, timeout=120) as resp:
                    return json.loads(resp.read())
            except (ConnectionRefusedError, urllib.error.URLError):
                assert all(p.poll() is None for p in procs), "a server rank exited"
                time.sleep(0.3)
        raise AssertionError("server did not come up")
    try:
        greedy = put({"prompts": ["5 9 13 2", "7 7 1"], "tokens_to_generate": 4, "top_k": 1, "logprobs": True})
        assert [len(t.split()) for t in greedy["text"]] == [8, 8] and len(greedy["logprobs"]) == 2
        again = put({"prompts": ["5 9 13 2", "7 7 1"], "tokens_to_generate": 4, "top_k": 1})
        assert again["text"] == greedy["text"]
        beam = put({"prompts": ["5 9 13 2"], "tokens_to_generate": 4, "beam_width": 2})
        assert len(beam["text"]) == 2 and beam["scores"][0] >= beam["scores"][1]
        assert all(t.startswith("5 9 13 2") for t in beam["text"])
        sampled = put({"prompts": ["1 2 3"], "tokens_to_generate": 3, "top_p": 0.9, "random_seed": 4})
        assert len(sampled["text"][0].split()) == 6
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        logs = [p.communicate()[0] for p in procs]
    assert not any("Traceback" in log for log in logs), logs

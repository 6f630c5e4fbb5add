"""HF <-> Megatron conversion round trip and logits parity (reference test strategy: verify_correctness.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)


def _tiny_hf_llama(n_kv):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=n_kv, max_position_embeddings=64,
                      rms_norm_eps=1e-5, tie_word_embeddings=False)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval()


@pytest.mark.parametrize("n_kv", [4, 2])
def test_qkv_roundtrip(n_kv):
    from weights_conversion.hf_to_megatron import hf_llama_state_to_meta_names, llama_like_to_megatron
    from weights_conversion.megatron_to_hf import llama_like_to_hf
    from argparse import Namespace
    hf = _tiny_hf_llama(n_kv)
    sd = {k: v.clone() for k, v in hf.state_dict().items()}
    mw = llama_like_to_megatron(hf_llama_state_to_meta_names(dict(sd)), 2, 64, 4, n_kv, "hf")
    args = Namespace(num_attention_heads=4, num_attention_heads_kv=n_kv, hidden_size=64, num_layers=2)
    enc = {k.replace(".attention.", ".self_attention."): v for k, v in mw["transformer"].items()}
    back = llama_like_to_hf(args, mw["embedding"]["word_embeddings.weight"], enc, mw["lm_head"])
    for k, v in back.items():
        assert torch.equal(v, sd[k]), k


@pytest.mark.parametrize("n_kv", [4, 2])
def test_converted_llama_matches_hf_logits(n_kv, tmp_path):
    """Our model loaded from the converted checkpoint reproduces the HF logits (fp32, CPU)."""
    from weights_conversion.hf_to_megatron import (architecture_args, hf_llama_state_to_meta_names,
                                                   llama_like_to_megatron, save_megatron)
    from tests.dist_utils import run_distributed
    hf = _tiny_hf_llama(n_kv)
    mw = llama_like_to_megatron(hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, n_kv, "hf")
    a = architecture_args("llama2", 7, 2, 64, 4, n_kv, 176, 96)
    a.update(max_position_embeddings=64, seq_length=32)
    save_megatron(tmp_path, mw, a, torch.float32)
    tokens = torch.randint(0, 96, (2, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = hf(tokens).logits
    run_distributed(_load_and_forward, 1, str(tmp_path), tokens, ref, backend="gloo")


def _load_and_forward(rank, world, path, tokens, ref):
    from megatron_llm_b200.checkpointing import load_checkpoint
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models import LlamaModel
    from megatron_llm_b200.models.enums import ModelType
    initialize_megatron(args_list=["--load", path, "--use_checkpoint_args", "--micro_batch_size", "2",
                                   "--tokenizer_type", "NullTokenizer", "--vocab_file", "96", "--no_load_optim",
                                   "--no_load_rng", "--finetune", "--train_iters", "1", "--lr", "1e-4"])
    model = LlamaModel(num_tokentypes=0, parallel_output=False, pre_process=True, post_process=True,
                       model_type=ModelType.encoder_or_decoder)
    load_checkpoint([model], None, None)
    model.eval()
    pos = torch.arange(tokens.size(1)).unsqueeze(0).expand_as(tokens).contiguous()
    mask = torch.tril(torch.ones(1, 1, tokens.size(1), tokens.size(1))) < 0.5     # True = masked
    with torch.no_grad():
        out = model(tokens, pos, mask).float()
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


def _flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, f"{prefix}{k}."))
        elif torch.is_tensor(v):
            out[f"{prefix}{k}"] = v
    return out


def test_reshard_roundtrip_and_tp2_logits(tmp_path):
    """tp1/pp1 -> tp2/pp2 -> tp1/pp1 is the identity; the tp2/pp1 checkpoint reproduces the HF logits on 2 ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from tools import checkpoint_util
    from weights_conversion.hf_to_megatron import (architecture_args, hf_llama_state_to_meta_names,
                                                   llama_like_to_megatron, save_megatron)
    from tests.dist_utils import run_distributed
    hf = _tiny_hf_llama(2)
    mw = llama_like_to_megatron(hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, 2, "hf")
    a = architecture_args("llama2", 7, 2, 64, 4, 2, 176, 96)
    a.update(max_position_embeddings=64, seq_length=32)
    src, mid, back, tp2 = (tmp_path / n for n in ("src", "mid", "back", "tp2"))
    save_megatron(src, mw, a, torch.float32)
    common = ["--model_type", "llama2", "--true_vocab_size", "96"]
    checkpoint_util.main(common + ["--load_dir", str(src), "--save_dir", str(mid),
                                   "--target_tensor_parallel_size", "2", "--target_pipeline_parallel_size", "2"])
    assert sorted(os.listdir(mid / "release")) == ["mp_rank_00_000", "mp_rank_00_001", "mp_rank_01_000",
                                                   "mp_rank_01_001"]
    checkpoint_util.main(common + ["--load_dir", str(mid), "--save_dir", str(back),
                                   "--target_tensor_parallel_size", "1", "--target_pipeline_parallel_size", "1"])
    x = _flatten(torch.load(src / "release/mp_rank_00/model_optim_rng.pt", weights_only=False)["model"])
    y = _flatten(torch.load(back / "release/mp_rank_00/model_optim_rng.pt", weights_only=False)["model"])
    x = {k.replace("transformer.", "encoder.").replace(".attention.", ".self_attention.")
         .replace("word_embeddings.weight", "word_embeddings.weight"): v for k, v in x.items()}
    y = {k.replace("word_embeddings.weight", "word_embeddings.weight"): v for k, v in y.items()}
    assert set(x) == set(y), set(x) ^ set(y)
    for k in x:
        assert torch.equal(x[k], y[k]), k
    checkpoint_util.main(common + ["--load_dir", str(src), "--save_dir", str(tp2),
                                   "--target_tensor_parallel_size", "2", "--target_pipeline_parallel_size", "1"])
    tokens = torch.randint(0, 96, (2, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = hf(tokens).logits
    run_distributed(_load_and_forward, 2, str(tp2), tokens, ref, backend="gloo")


def test_verify_correctness_script(tmp_path):
    """verify_correctness.py end to end on CPU: converted tiny Llama vs its HF original on synthetic batches."""
    import subprocess
    from weights_conversion.hf_to_megatron import (architecture_args, hf_llama_state_to_meta_names,
                                                   llama_like_to_megatron, save_megatron)
    from tests.dist_utils import free_port
    hf = _tiny_hf_llama(2)
    hf.save_pretrained(tmp_path / "hf")
    mw = llama_like_to_megatron(hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, 2, "hf")
    a = architecture_args("llama2", 7, 2, 64, 4, 2, 176, 96)
    a.update(max_position_embeddings=64, seq_length=32)
    save_megatron(tmp_path / "mega", mw, a, torch.float32)
    env = dict(os.environ, MLB200_FORCE_CPU="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), RANK="0",
               WORLD_SIZE="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.join(ROOT, "verify_correctness.py"), "--model_name", "llama2", "--load",
           str(tmp_path / "mega"), "--huggingface_cache", str(tmp_path / "hf"), "--huggingface_device", "cpu",
           "--data_type", "synthetic", "--tokenizer_type", "NullTokenizer", "--vocab_file", "96", "--no_load_optim",
           "--no_load_rng", "--finetune", "--train_iters", "10", "--global_batch_size", "1"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("Max absoulute error")]
    assert len(lines) == 10
    assert all(float(l.split("max=")[1].split(",")[0]) < 1e-3 for l in lines), lines


def _load_falcon_and_forward(rank, world, path, tokens, ref):
    from megatron_llm_b200.checkpointing import load_checkpoint
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models import FalconModel
    from megatron_llm_b200.models.enums import ModelType
    initialize_megatron(args_list=["--load", path, "--use_checkpoint_args", "--micro_batch_size", "2",
                                   "--tokenizer_type", "NullTokenizer", "--vocab_file", "128", "--no_load_optim",
                                   "--no_load_rng", "--finetune", "--train_iters", "1", "--lr", "1e-4"])
    model = FalconModel(num_tokentypes=0, parallel_output=False, pre_process=True, post_process=True,
                        model_type=ModelType.encoder_or_decoder)
    load_checkpoint([model], None, None)
    model.eval()
    pos = torch.arange(tokens.size(1)).unsqueeze(0).expand_as(tokens).contiguous()
    mask = torch.tril(torch.ones(1, 1, tokens.size(1), tokens.size(1))) < 0.5
    with torch.no_grad():
        out = model(tokens, pos, mask).float()
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


@pytest.mark.parametrize("size,new_arch,heads,kv", [(40, True, 16, 8), (7, False, 4, 1)], ids=["falcon40b_style", "falcon7b_style"])
def test_converted_falcon_matches_hf_logits(size, new_arch, heads, kv, tmp_path, monkeypatch):
    """Falcon: parallel attention + MLP, tied embeddings, LayerNorm; 40B style = GQA with separate attention / MLP norms,
    7B style = MQA with one norm.  A tiny random HF Falcon converted with ``falcon_to_megatron`` must give the same
    logits in this repo's FalconModel (fp32, CPU)."""
    import weights_conversion.hf_to_megatron as C
    from transformers import FalconConfig, FalconForCausalLM
    from tests.dist_utils import run_distributed
    torch.manual_seed(0)
    cfg = FalconConfig(vocab_size=128, hidden_size=128, num_hidden_layers=2, num_attention_heads=heads, num_kv_heads=kv,
                       new_decoder_architecture=new_arch, multi_query=not new_arch, parallel_attn=True, bias=False,
                       alibi=False, max_position_embeddings=64, tie_word_embeddings=True)
    hf = FalconForCausalLM(cfg).float().eval()
    monkeypatch.setitem(C.falcon_s2layer, size, 2)
    monkeypatch.setitem(C.falcon_s2heads, size, heads)
    monkeypatch.setitem(C.falcon_s2hidden, size, 128)
    sd = dict(hf.state_dict())
    sd.setdefault("lm_head.weight", sd["transformer.word_embeddings.weight"])
    mw = C.falcon_to_megatron(sd, size)
    a = C.architecture_args("falcon", size, vocab=128)
    a.update(max_position_embeddings=64, seq_length=32, make_vocab_size_divisible_by=1, tokenizer_type="NullTokenizer")
    C.save_megatron(tmp_path, mw, a, torch.float32)
    tokens = torch.randint(0, 128, (2, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = hf(tokens).logits
    run_distributed(_load_falcon_and_forward, 1, str(tmp_path), tokens, ref, backend="gloo")


def _load_mistral_and_forward(rank, world, path, tokens, ref):
    from megatron_llm_b200.checkpointing import load_checkpoint
    from megatron_llm_b200.initialize import initialize_megatron
    from megatron_llm_b200.models import MistralModel
    from megatron_llm_b200.models.enums import ModelType
    initialize_megatron(args_list=["--load", path, "--use_checkpoint_args", "--micro_batch_size", "2",
                                   "--tokenizer_type", "NullTokenizer", "--vocab_file", "96", "--no_load_optim",
                                   "--no_load_rng", "--finetune", "--train_iters", "1", "--lr", "1e-4",
                                   "--use_flash_attn"])
    model = MistralModel(num_tokentypes=0, parallel_output=False, pre_process=True, post_process=True,
                         model_type=ModelType.encoder_or_decoder)
    load_checkpoint([model], None, None)
    model.eval()
    pos = torch.arange(tokens.size(1)).unsqueeze(0).expand_as(tokens).contiguous()
    mask = torch.tril(torch.ones(1, 1, tokens.size(1), tokens.size(1))) < 0.5
    with torch.no_grad():
        out = model(tokens, pos, mask).float()
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()


def test_converted_mistral_matches_hf_logits(tmp_path):
    """Mistral (GQA, sliding window 4096) through the HF -> Megatron converter and this repo's MistralModel."""
    import weights_conversion.hf_to_megatron as C
    from transformers import MistralConfig, MistralForCausalLM
    from tests.dist_utils import run_distributed
    torch.manual_seed(0)
    cfg = MistralConfig(vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                        num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64, sliding_window=4096,
                        rms_norm_eps=1e-5, tie_word_embeddings=False, attn_implementation="eager")
    hf = MistralForCausalLM(cfg).float().eval()
    mw = C.llama_like_to_megatron(C.hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, 2, "hf")
    a = C.architecture_args("mistral", 7, 2, 64, 4, 2, 176, 96)
    a.update(max_position_embeddings=64, seq_length=32)
    C.save_megatron(tmp_path, mw, a, torch.float32)
    tokens = torch.randint(0, 96, (2, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = hf(tokens).logits
    run_distributed(_load_mistral_and_forward, 1, str(tmp_path), tokens, ref, backend="gloo")


@pytest.mark.parametrize("name", ["falcon40b_style", "falcon7b_style", "mistral"])
def test_hf_megatron_hf_round_trip(name, tmp_path, monkeypatch):
    """HF -> Megatron (hf_to_megatron) -> HF (megatron_to_hf.main, incl. the generated HF config) gives back a model with
    bit-identical logits."""
    import weights_conversion.hf_to_megatron as C
    import weights_conversion.megatron_to_hf as R
    from transformers import AutoModelForCausalLM, FalconConfig, FalconForCausalLM, MistralConfig, MistralForCausalLM
    torch.manual_seed(0)
    if name.startswith("falcon"):
        size, new_arch, heads, kv = (40, True, 16, 8) if name == "falcon40b_style" else (7, False, 4, 1)
        cfg = FalconConfig(vocab_size=128, hidden_size=128, num_hidden_layers=2, num_attention_heads=heads,
                           num_kv_heads=kv, new_decoder_architecture=new_arch, multi_query=not new_arch,
                           parallel_attn=True, bias=False, alibi=False, max_position_embeddings=64,
                           tie_word_embeddings=True)
        hf = FalconForCausalLM(cfg).float().eval()
        monkeypatch.setitem(C.falcon_s2layer, size, 2)
        monkeypatch.setitem(C.falcon_s2heads, size, heads)
        monkeypatch.setitem(C.falcon_s2hidden, size, 128)
        sd = dict(hf.state_dict())
        sd.setdefault("lm_head.weight", sd["transformer.word_embeddings.weight"])
        mw = C.falcon_to_megatron(sd, size)
        a = C.architecture_args("falcon", size, vocab=128)
        a.update(max_position_embeddings=64, seq_length=32, make_vocab_size_divisible_by=1)
        model, vocab = "falcon", 128
    else:
        cfg = MistralConfig(vocab_size=96, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                            num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                            sliding_window=4096, rms_norm_eps=1e-5, tie_word_embeddings=False)
        hf = MistralForCausalLM(cfg).float().eval()
        mw = C.llama_like_to_megatron(C.hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, 2, "hf")
        a = C.architecture_args("mistral", 7, 2, 64, 4, 2, 176, 96)
        a.update(max_position_embeddings=64, seq_length=32)
        model, vocab = "mistral", 96
    C.save_megatron(tmp_path / "meg", mw, a, torch.float32)
    R.main(model, tmp_path / "meg", tmp_path / "hf", dtype=torch.float32)
    back = AutoModelForCausalLM.from_pretrained(tmp_path / "hf").float().eval()
    tokens = torch.randint(0, vocab, (2, 32), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.equal(back(tokens).logits, hf(tokens).logits)


def _tiny_sentencepiece_model(d):
    import random
    import sentencepiece as spm
    rnd = random.Random(0)
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu"]
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 12))) for _ in range(500)))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "tokenizer"), vocab_size=80, model_type="bpe",
                                   pad_id=-1, unk_id=0, bos_id=1, eos_id=2, minloglevel=2)
    return d / "tokenizer.model"


def test_megatron_to_hf_writes_matching_tokenizer_and_shards(tmp_path):
    """megatron_to_hf with --vocab_file / --vocab_extra_ids_list / --override_special_tokens / --num_output_shards: the
    vocabulary is trimmed to the Megatron tokenizer's size, the saved HF tokenizer assigns the ids Megatron trained
    with (special tokens appended in the same order), the override lands, and the weights come in several shards."""
    import weights_conversion.hf_to_megatron as C
    import weights_conversion.megatron_to_hf as R
    from megatron_llm_b200.tokenizer.tokenizer import _SentencePieceTokenizer
    from transformers import AutoModelForCausalLM, AutoTokenizer, LlamaConfig, LlamaForCausalLM
    sp_dir = tmp_path / "sp"
    sp_dir.mkdir()
    vocab_file = _tiny_sentencepiece_model(sp_dir)
    extra = "<|im_start|>,<|im_end|>"
    mt = _SentencePieceTokenizer(str(vocab_file), vocab_extra_ids_list=extra, new_tokens=True)
    assert mt.vocab_size == 80 + 5 + 2                       # <CLS> <SEP> <EOD> <MASK> <PAD> + the two extra ids
    padded = 128
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=padded, hidden_size=64, intermediate_size=176, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64, rms_norm_eps=1e-5,
                      tie_word_embeddings=False)
    hf = LlamaForCausalLM(cfg).float().eval()
    mw = C.llama_like_to_megatron(C.hf_llama_state_to_meta_names(dict(hf.state_dict())), 2, 64, 4, 2, "hf")
    a = C.architecture_args("llama2", 7, 2, 64, 4, 2, 176, padded)
    a.update(max_position_embeddings=64, seq_length=32)
    C.save_megatron(tmp_path / "meg", mw, a, torch.float32)
    R.main("llama2", tmp_path / "meg", tmp_path / "hf", vocab_file=vocab_file, no_new_tokens=False, dtype=torch.float32,
           override_special_tokens=["eos=<|im_end|>", "bogus", "xyz=<PAD>", "pad=<nope>"], num_output_shards=3,
           vocab_extra_ids_list=extra)
    shards = [f for f in os.listdir(tmp_path / "hf") if f.endswith(".safetensors")]
    assert len(shards) >= 3 and (tmp_path / "hf" / "model.safetensors.index.json").exists()
    back = AutoModelForCausalLM.from_pretrained(tmp_path / "hf").float().eval()
    assert back.config.vocab_size == mt.vocab_size
    tokens = torch.randint(0, mt.vocab_size, (2, 16), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.equal(back(tokens).logits, hf(tokens).logits[..., :mt.vocab_size])
    tok = AutoTokenizer.from_pretrained(tmp_path / "hf")
    vocab = tok.get_vocab()
    for name in ("<CLS>", "<SEP>", "<EOD>", "<MASK>", "<PAD>", "<|im_start|>", "<|im_end|>", "<s>", "</s>"):
        assert vocab[name] == mt.vocab[name], name
    assert tok.eos_token == "<|im_end|>" and tok.eos_token_id == mt.vocab["<|im_end|>"]
    assert tok.pad_token_id == mt.pad and tok.cls_token_id == mt.cls and tok.mask_token_id == mt.mask
    text = "alpha beta<|im_end|> gamma"
    assert tok.encode(text, add_special_tokens=False) == mt.tokenize(text)

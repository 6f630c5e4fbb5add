"""BERT / T5 / classification / multiple-choice / bi-encoder: forward+backward on CPU (gloo), TP=1 and TP=2."""
import os
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
sys.path.insert(0, ROOT)
from tests.dist_utils import run_distributed  # noqa: E402

BASE = ["--num_layers", "2", "--hidden_size", "32", "--num_attention_heads", "4", "--seq_length", "16",
        "--max_position_embeddings", "16", "--micro_batch_size", "2", "--tokenizer_type", "NullTokenizer",
        "--vocab_file", "64", "--make_vocab_size_divisible_by", "8", "--train_iters", "1", "--lr", "1e-3",
        "--hidden_dropout", "0.0", "--attention_dropout", "0.0"]


def _init(world, extra=()):
    from megatron_llm_b200.initialize import initialize_megatron
    initialize_megatron(args_list=BASE + ["--tensor_model_parallel_size", str(world)] + list(extra))
    from megatron_llm_b200.global_vars import get_args
    return get_args()


def _bert_family(rank, world):
    args = _init(world)
    from megatron_llm_b200.models import BertModel, Classification, MultipleChoice
    from megatron_llm_b200.models.enums import ModelType
    torch.manual_seed(0)
    tokens = torch.randint(0, 60, (2, 16))
    pad = torch.ones(2, 16, dtype=torch.long)
    pad[1, 12:] = 0
    types = torch.zeros(2, 16, dtype=torch.long)
    labels = torch.randint(0, 60, (2, 16))
    bert = BertModel(num_tokentypes=2, add_binary_head=True, model_type=ModelType.encoder_or_decoder)
    loss, binary = bert(tokens, pad, tokentype_ids=types, lm_labels=labels)
    assert loss.shape == (2, 16) and binary.shape == (2, 2)
    (loss.mean() + binary.float().mean()).backward()
    assert bert.lm_head.dense.weight.grad is not None
    logits, _ = bert(tokens, pad, tokentype_ids=types)
    assert logits.shape == (2, 16, args.padded_vocab_size // world)
    sd = bert.state_dict_for_save_checkpoint()
    assert set(sd) == {"language_model", "lm_head", "binary_head"}
    bert.load_state_dict(sd)
    cls = Classification(num_classes=3, num_tokentypes=2, model_type=ModelType.encoder_or_decoder)
    out = cls(tokens, pad, tokentype_ids=types)
    assert out.shape == (2, 3)
    out.sum().backward()
    mc = MultipleChoice(num_tokentypes=2, model_type=ModelType.encoder_or_decoder)
    t4 = tokens.view(1, 2, 16).repeat(2, 1, 1)
    out = mc(t4, pad.view(1, 2, 16).repeat(2, 1, 1), tokentype_ids=types.view(1, 2, 16).repeat(2, 1, 1))
    assert out.shape == (2, 2)
    assert "multichoice_head" in mc.state_dict_for_save_checkpoint()


def _t5(rank, world):
    args = _init(world, ["--decoder_seq_length", "8"])
    from megatron_llm_b200.models import T5Model
    from megatron_llm_b200.models.enums import ModelType
    args.model_type = ModelType.encoder_and_decoder
    torch.manual_seed(0)
    enc = torch.randint(0, 60, (2, 16))
    dec = torch.randint(0, 60, (2, 8))
    enc_mask = torch.zeros(2, 16, 16, dtype=torch.bool)
    dec_mask = torch.triu(torch.ones(8, 8, dtype=torch.bool), 1).unsqueeze(0).expand(2, 8, 8)
    cross = torch.zeros(2, 8, 16, dtype=torch.bool)
    model = T5Model(num_tokentypes=0, model_type=ModelType.encoder_and_decoder)
    loss = model(enc, dec, enc_mask, dec_mask, cross, lm_labels=torch.randint(0, 60, (2, 8)))
    assert loss.shape == (2, 8)
    loss.mean().backward()
    logits = model(enc, dec, enc_mask, dec_mask, cross)
    assert logits.shape == (2, 8, args.padded_vocab_size // world)
    # causality of the decoder: changing a later decoder token must not change earlier logits
    dec2 = dec.clone()
    dec2[:, -1] = (dec2[:, -1] + 1) % 60
    model.eval()
    with torch.no_grad():
        a, b = model(enc, dec, enc_mask, dec_mask, cross), model(enc, dec2, enc_mask, dec_mask, cross)
    assert torch.allclose(a[:, :-1], b[:, :-1], atol=1e-5) and not torch.allclose(a[:, -1], b[:, -1])
    sd = model.state_dict_for_save_checkpoint()
    assert set(sd) == {"language_model", "lm_head"} and "decoder" in sd["language_model"]
    model.load_state_dict(sd)


def _biencoder(rank, world):
    _init(world, ["--biencoder_projection_dim", "8"])
    from megatron_llm_b200.models.biencoder_model import biencoder_model_provider
    from megatron_llm_b200.models.enums import ModelType
    model = biencoder_model_provider(model_type=ModelType.encoder_or_decoder)
    tokens = torch.randint(0, 60, (2, 16))
    mask = torch.zeros(2, 16, 16, dtype=torch.bool)
    types = torch.zeros(2, 16, dtype=torch.long)
    q, c = model(tokens, mask, types, tokens, mask, types)
    assert q.shape == (2, 8) and c.shape == (2, 8)
    (q @ c.t()).sum().backward()
    sd = model.state_dict_for_save_checkpoint()
    assert set(sd) == {"query_model", "context_model"} and "projection_enc" in sd["query_model"]
    model.load_state_dict(sd)


@pytest.mark.parametrize("world", [1, 2])
def test_bert_family(world):
    run_distributed(_bert_family, world)


@pytest.mark.parametrize("world", [1, 2])
def test_t5(world):
    run_distributed(_t5, world)


def test_biencoder():
    run_distributed(_biencoder, 1)

"""Run our model and a Hugging Face / Meta baseline on the same batches and report max / mean absolute logit error.

Parity: verify_correctness.py (reference :1-217).  Same flags (``--huggingface_cache``, ``--huggingface_device``,
``--model_size`` on top of finetune.py's), same two modes:
  * ``--load`` is a Megatron checkpoint  -> our model is built by ``finetune.model_provider`` and loaded from it;
  * ``--load`` is a HF directory         -> a converted-back checkpoint (megatron_to_hf) is verified against the baseline.
The baseline always comes from ``--huggingface_cache`` (a local HF directory or Meta's raw ``*.pth`` directory; there is
no network here so hub names are only tried as a last resort).  ``--data_type synthetic`` verifies on random tokens."""
from __future__ import annotations

import json
import os
import sys
import warnings
from pathlib import Path
from typing import Optional

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from finetune import data_provider, extra_args, get_batch, loss_func, model_provider  # noqa: E402
from megatron_llm_b200 import get_args, update_num_microbatches  # noqa: E402
from megatron_llm_b200.initialize import initialize_megatron, set_jit_fusion_options  # noqa: E402
from megatron_llm_b200.training import _setup_model_and_optimizer, build_train_valid_test_data_iterators  # noqa: E402

HUB_NAMES = {"falcon": "tiiuae/falcon-{size}b", "llama": "decapoda-research/llama-{size}b-hf",
             "llama2": "meta-llama/Llama-2-{size}b-hf", "codellama": "codellama/CodeLlama-{size}b-hf",
             "mistral": "mistralai/Mistral-{size}B-v0.1"}


class MetaLlamaWrapper(nn.Module):
    """Meta's raw ``consolidated.00.pth`` + ``params.json`` run through the HF architecture (the reference imports
    Meta's ``llama`` package and fairscale; neither is needed: the weights are renamed and un-permuted on the fly)."""

    def __init__(self, cache_dir: Path, dtype=None):
        super().__init__()
        from transformers import LlamaConfig, LlamaForCausalLM
        from weights_conversion.utils.permute_qkv import _permute_head
        ckpts = sorted(cache_dir.glob("*.pth"))
        assert len(ckpts) == 1, "Currently, only llama2 unsharded models implemented"
        params = json.loads((cache_dir / "params.json").read_text())
        sd = torch.load(ckpts[0], map_location="cpu")
        h, n = params["dim"], params["n_heads"]
        nkv = params.get("n_kv_heads", n)
        ffn = sd["layers.0.feed_forward.w1.weight"].size(0)
        cfg = LlamaConfig(vocab_size=sd["tok_embeddings.weight"].size(0), hidden_size=h, intermediate_size=ffn,
                          num_hidden_layers=params["n_layers"], num_attention_heads=n, num_key_value_heads=nkv,
                          rms_norm_eps=params.get("norm_eps", 1e-5), max_position_embeddings=4096)
        ren = {"attention.wq": "self_attn.q_proj", "attention.wk": "self_attn.k_proj", "attention.wv": "self_attn.v_proj",
               "attention.wo": "self_attn.o_proj", "feed_forward.w1": "mlp.gate_proj",
               "feed_forward.w2": "mlp.down_proj", "feed_forward.w3": "mlp.up_proj",
               "attention_norm": "input_layernorm", "ffn_norm": "post_attention_layernorm"}
        out = {"model.embed_tokens.weight": sd["tok_embeddings.weight"], "model.norm.weight": sd["norm.weight"],
               "lm_head.weight": sd["output.weight"]}
        hn = h // n
        for k, w in sd.items():
            parts = k.split(".")
            if parts[0] != "layers":
                continue
            name = ".".join(parts[2:-1])
            if name not in ren:
                continue
            if name in ("attention.wq", "attention.wk"):       # interleaved (Meta) -> halves (HF)
                w = torch.cat([_permute_head(x, revert=True) for x in w.split(hn, dim=0)], dim=0)
            out[f"model.layers.{parts[1]}.{ren[name]}.weight"] = w
        self.model = LlamaForCausalLM(cfg)
        self.model.load_state_dict(out, strict=False)
        if dtype is not None:
            self.model.to(dtype)

    def forward(self, input_ids, position_ids=None, attention_mask=None, labels=None):
        return self.model(input_ids=input_ids, position_ids=position_ids, labels=labels)


Llama2Wrapper = MetaLlamaWrapper          # the reference's name (verify_correctness.py:21)


def is_meta_llama2_path(path: Optional[Path]) -> bool:
    return path is not None and len(list(Path(path).glob("*.pth"))) > 0


def hf_provider(name: str, cache_dir: Optional[Path], device: str, size: int = 7, bf16: bool = False):
    from transformers import AutoModelForCausalLM
    print("Getting huggingface model...")
    kw = {"torch_dtype": torch.bfloat16} if bf16 else {}
    if name in ("llama2", "llama", "codellama") and is_meta_llama2_path(cache_dir):
        print(f"baseline path {cache_dir} does not look like a huggingface, assuming it's raw llama weights instead")
        model = MetaLlamaWrapper(Path(cache_dir), torch.bfloat16 if bf16 else None)
    elif name in HUB_NAMES:
        try:
            model = AutoModelForCausalLM.from_pretrained(cache_dir, trust_remote_code=(name == "falcon"), **kw)
        except (OSError, TypeError, ValueError):
            print(f"Cache dir {cache_dir} does not look like a huggingface checkpoint, assuming cache_dir instead")
            model = AutoModelForCausalLM.from_pretrained(HUB_NAMES[name].format(size=size), cache_dir=cache_dir,
                                                         trust_remote_code=(name == "falcon"), **kw)
    else:
        raise KeyError(f"Model {name} not implemented")
    return model.eval().requires_grad_(False).to(device)


def hf_our_provider(name: str, data_dir: Path, device: str, size: int = 7, bf16: bool = False):
    from transformers import AutoModelForCausalLM
    model = AutoModelForCausalLM.from_pretrained(data_dir, **({"torch_dtype": torch.bfloat16} if bf16 else {}))
    return model.eval().requires_grad_(False).to(device)


def hf_forward(model, batch):
    device = next(p.device for p in model.parameters())
    tokens, labels, loss_mask, attention_mask, position_ids = [t.to(device) if t is not None else None for t in batch]
    out = model(input_ids=tokens, position_ids=position_ids, labels=tokens)
    return out["logits"], out["loss"]


def mega_provider(name: str):
    print("Getting megatron model...")
    from megatron_llm_b200.models.enums import ModelType
    model, _, _ = _setup_model_and_optimizer(model_provider, ModelType.encoder_or_decoder, args=get_args())
    assert len(model) == 1, "correctness verification only supported with unsharded models"
    return model[0].eval().requires_grad_(False)


def mega_forward(model, batch):
    tokens, labels, loss_mask, attention_mask, position_ids = batch
    assert torch.all(loss_mask)
    out = model(tokens, position_ids, attention_mask, labels=labels)
    _, logits = out
    loss, _ = loss_func(model.training, batch, out)
    return logits, loss


def verify_step(our_forward, our_model, base_forward, base_model, batch):
    our_logits, our_loss = our_forward(our_model, batch)
    base_logits, base_loss = base_forward(base_model, batch)
    v = min(our_logits.size(-1), base_logits.size(-1))          # ours may carry vocab padding
    our_logits, base_logits = our_logits[..., :v].float().cpu(), base_logits[..., :v].float().cpu()
    assert our_logits.size() == base_logits.size(), f"ours={our_logits.size()}, true={base_logits.size()}"
    err = (our_logits - base_logits).abs()
    print(f"Max absoulute error in the logits: max={err.max():.6f}, avg={err.mean():.6f}")
    our_loss, base_loss = our_loss.float().cpu(), base_loss.float().cpu()
    print(f"Abs loss error: {(our_loss - base_loss).abs():.6f} Our loss: {our_loss:.3f}, theirs: {base_loss:.3f}")
    return err.max().item(), err.mean().item()


def is_megatron_path(path) -> bool:
    return (Path(path) / "latest_checkpointed_iteration.txt").exists()


def main(n_iters: int = 10):
    print("Starting megatron vs huggingface verification")
    args = get_args()
    set_jit_fusion_options(args)
    dev0 = "cuda:0" if torch.cuda.is_available() else "cpu"
    print("Loading our model!")
    if is_megatron_path(args.load):
        our_model, our_forward = mega_provider(args.model_name), mega_forward
    else:
        print("NOTE: The given path does not look like a megatron checkpoint, assuming it's a huggingface checkpoint "
              f"instead (path={args.load})")
        our_model, our_forward = hf_our_provider(args.model_name, args.load, dev0, bf16=args.bf16), hf_forward
        args.iteration = 0
    print("Loading baseline model!")
    base_dev = args.baseline_device
    if base_dev.startswith("cuda") and (not torch.cuda.is_available()
                                        or int(base_dev.split(":")[-1] or 0) >= torch.cuda.device_count()):
        base_dev = dev0
    base_model = hf_provider(args.model_name, args.cache_dir, base_dev, size=args.model_size, bf16=args.bf16)
    print("Loading dataset!")
    data_iterator, _, _ = build_train_valid_test_data_iterators(data_provider, args)
    worst = 0.0
    for iteration in range(n_iters):
        print(f"Iteration {iteration}...")
        update_num_microbatches(args.consumed_train_samples)
        args.curr_iteration = iteration
        mx, _ = verify_step(our_forward, our_model, hf_forward, base_model, get_batch(data_iterator))
        worst = max(worst, mx)
    return worst


def extra_extra_args(parser):
    parser = extra_args(parser)
    group = parser.add_argument_group(title="huggingface")
    group.add_argument("--huggingface_cache", type=Path, default=None, dest="cache_dir",
                       help="local HF checkpoint directory / HF cache dir / Meta raw-weight directory of the baseline")
    group.add_argument("--huggingface_device", default="cuda:1", dest="baseline_device",
                       help="Device to use for the baseline model")
    group.add_argument("--model_size", type=int, default=7)
    return parser


if __name__ == "__main__":
    defaults = {"micro_batch_size": 1, "use_checkpoint_args": True, "train_iters": 10, "lr": 1.0}
    initialize_megatron(extra_extra_args, args_defaults=defaults)
    main()

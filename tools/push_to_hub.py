"""Export a converted HF-format checkpoint to a local folder (dtype cast / resharding) and/or push it to the Hub.

Parity: tools/push_to_hub.py (same flags)."""
import argparse
import sys

import torch

DTYPES = {"float16": torch.float16, "fp16": torch.float16, "float32": torch.float32, "fp32": torch.float32,
          "bfloat16": torch.bfloat16, "bf16": torch.bfloat16, "auto": None}


def parse_args(argv=None):
    p = argparse.ArgumentParser(
        description="Push checkpoints in HF transformers format to the Huggingface Hub.",
        epilog="Example: python push_to_hub.py /path/to/checkpoint --hf_repo_name org/model --dtype bf16 --auth_token hf_...")
    p.add_argument("model_name", type=str, help="Path to checkpoint or model name")
    p.add_argument("--dtype", type=str, default="auto", help="auto (default), bf16, fp16 or fp32")
    p.add_argument("--hf_repo_name", type=str, help="HuggingFace repository name")
    p.add_argument("--auth_token", type=str, help="User access token (HuggingFace) used for model upload")
    p.add_argument("--output_folder", type=str, help="Output folder path (e.g. for dtype conversion)")
    p.add_argument("--max_shard_size", type=str, default="10GB",
                   help="Maximum size for a checkpoint before being sharded (default: 10GB)")
    p.add_argument("--unsafe", action="store_true", default=False, help="Disable safetensor serialization")
    p.add_argument("--rope_scaling_type", type=str, default="linear", help="Overwrite rope scaling type (linear, dynamic)")
    p.add_argument("--rope_scaling_factor", type=float, help="Overwrite rope scaling factor (float >1.0)")
    p.add_argument("--trust_remote_code", action="store_true", default=False, help="Allow custom model code")
    return p.parse_args(argv)


def main(argv=None):
    from transformers import AutoModelForCausalLM, AutoTokenizer
    args = parse_args(argv)
    print(args)
    if args.dtype not in DTYPES:
        print(f"Unsupported dtype: {args.dtype}")
        sys.exit(1)
    if not args.hf_repo_name and not args.output_folder:
        print("Please specify either `--hf_repo_name` to push to HF or `--output_folder` to export the model to a "
              "local folder.")
        sys.exit(1)
    tokenizer = None
    try:
        tokenizer = AutoTokenizer.from_pretrained(args.model_name)
        print(f"Tokenizer: {type(tokenizer).__name__} (vocab_size: {len(tokenizer):,})")
        for token in tokenizer.all_special_tokens:
            print(f"{token}: {tokenizer.convert_tokens_to_ids(token)}")
    except Exception as e:
        print(f"No tokenizer found next to the model ({e}); exporting weights only")
    model = AutoModelForCausalLM.from_pretrained(args.model_name, torch_dtype=DTYPES[args.dtype],
                                                 trust_remote_code=args.trust_remote_code)
    print(f"Model: {type(model).__name__} (num_parameters={model.num_parameters():,})")
    if args.rope_scaling_type is not None and args.rope_scaling_factor is not None:
        assert args.rope_scaling_type in ("linear", "dynamic") and args.rope_scaling_factor >= 1.0
        model.config.rope_scaling = {"type": args.rope_scaling_type, "factor": args.rope_scaling_factor}
    print(model.config)
    safe = not args.unsafe
    if args.output_folder:
        model.save_pretrained(args.output_folder, max_shard_size=args.max_shard_size, safe_serialization=safe)
        if tokenizer is not None:
            tokenizer.save_pretrained(args.output_folder)
    if args.hf_repo_name:
        model.push_to_hub(args.hf_repo_name, token=args.auth_token, max_shard_size=args.max_shard_size,
                          safe_serialization=safe)
        if tokenizer is not None:
            tokenizer.push_to_hub(args.hf_repo_name, token=args.auth_token)


if __name__ == "__main__":
    main()

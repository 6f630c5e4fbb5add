#!/bin/bash
# Convert HF / Meta weights to the Megatron layout (parity: examples/hf_to_megatron.sh).
#   examples/hf_to_megatron.sh llama2 --size 7 --model-path /models/Llama-2-7b-hf --out /checkpoints/llama2-7b
source "$(dirname "$0")/_common.sh"
MODEL=$1; shift
python $REPO/weights_conversion/hf_to_megatron.py $MODEL "$@"

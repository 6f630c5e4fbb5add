#!/bin/bash
# Compare our logits with the HF / Meta baseline on real batches (parity: examples/verify.sh).
#   examples/verify.sh llama2 7 /checkpoints/llama2-7b /models/Llama-2-7b-hf /data/corpus_text_document
source "$(dirname "$0")/_common.sh"
MODEL=$1; SIZE=$2; CKPT=$3; HF=$4; DATA=$5
case $MODEL in falcon) TOK="--tokenizer_type FalconTokenizer";; *) TOK="--tokenizer_type SentencePieceTokenizer --vocab_file ${TOKENIZER_MODEL:-$HF/tokenizer.model} --no_new_tokens";; esac
GPUS_PER_NODE=1 launch $REPO/verify_correctness.py --model_name $MODEL --model_size $SIZE --load $CKPT --huggingface_cache $HF \
  --huggingface_device cuda:1 --data_path $DATA $TOK --bf16 --use_flash_attn --micro_batch_size 1 --global_batch_size 1 \
  --no_bias_gelu_fusion --no_bias_dropout_fusion --hidden_dropout 0.0 --attention_dropout 0.0 --split 100,0,0

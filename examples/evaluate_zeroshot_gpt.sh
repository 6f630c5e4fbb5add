#!/bin/bash
# Zero-shot WikiText-103 perplexity / LAMBADA accuracy (parity: examples/evaluate_zeroshot_gpt.sh).
source "$(dirname "$0")/_common.sh"
TASK=${TASK:-LAMBADA}; VALID_DATA=${VALID_DATA:-lambada_test.jsonl}
launch $REPO/tasks/main.py --task $TASK --valid_data $VALID_DATA --tokenizer_type GPT2BPETokenizer --strict_lambada \
  --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} --merge_file ${MERGE_FILE:-gpt2-merges.txt} --load ${CHECKPOINT:-checkpoints/gpt2_345m} \
  --tensor_model_parallel_size 1 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --micro_batch_size 8 --seq_length 1024 \
  --max_position_embeddings 1024 --log_interval 10 --bf16 --no_load_optim --no_load_rng

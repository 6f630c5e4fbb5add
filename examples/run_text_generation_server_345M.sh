#!/bin/bash
# REST text-generation server for a 345M GPT (parity: examples/run_text_generation_server_345M.sh); query with tools/text_generation_cli.py host:5000
source "$(dirname "$0")/_common.sh"
GPUS_PER_NODE=1 launch $REPO/tools/run_text_generation_server.py --tensor_model_parallel_size 1 --pipeline_model_parallel_size 1 \
  --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --max_position_embeddings 1024 --seq_length 1024 \
  --load ${CHECKPOINT:-checkpoints/gpt2_345m} --tokenizer_type GPT2BPETokenizer --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} \
  --merge_file ${MERGE_FILE:-gpt2-merges.txt} --bf16 --micro_batch_size 1 --out_seq_length 1024 --temperature 1.0 --top_p 0.9 --seed 42

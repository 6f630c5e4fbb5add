#!/bin/bash
# TP=2 x PP=2 x DP=2 GPT-2 345M (parity: examples/pretrain_gpt_distributed_with_mp.sh).
source "$(dirname "$0")/_common.sh"
CHECKPOINT_PATH=${CHECKPOINT_PATH:-checkpoints/gpt2_345m}; DATA_PATH=${DATA_PATH:-my-gpt2_text_document}
launch $REPO/finetune.py --model_name gpt --tensor_model_parallel_size 2 --pipeline_model_parallel_size 2 --sequence_parallel \
  --micro_batch_size 4 --global_batch_size 16 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --seq_length 1024 --max_position_embeddings 1024
  --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} --merge_file ${MERGE_FILE:-gpt2-merges.txt} --tokenizer_type GPT2BPETokenizer \
  --train_iters 500000 --lr_decay_iters 320000 --lr 0.00015 --min_lr 1.0e-5 --lr_decay_style cosine --lr_warmup_fraction .01
  --weight_decay 1e-2 --clip_grad 1.0 --log_interval 100 --save_interval 10000 --eval_interval 1000 --eval_iters 10 --split 949,50,1 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH --data_path $DATA_PATH --bf16

#!/bin/bash
# BERT-large (345M) pre-training (parity: examples/pretrain_bert.sh).
source "$(dirname "$0")/_common.sh"
CHECKPOINT_PATH=${CHECKPOINT_PATH:-checkpoints/bert_345m}; DATA_PATH=${DATA_PATH:-my-bert_text_sentence}
GPUS_PER_NODE=1 launch $REPO/pretrain_bert.py --micro_batch_size 4 --global_batch_size 8 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --seq_length 512 --max_position_embeddings 512
  --vocab_file ${VOCAB_FILE:-bert-vocab.txt} --tokenizer_type BertWordPieceLowerCase \
  --train_iters 1000000 --lr 0.0001 --min_lr 1.0e-5 --lr_decay_style linear --lr_decay_iters 990000 --lr_warmup_fraction .01
  --weight_decay 1e-2 --clip_grad 1.0 --log_interval 100 --save_interval 10000 --eval_interval 1000 --eval_iters 10 --split 949,50,1 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH --data_path $DATA_PATH --bf16

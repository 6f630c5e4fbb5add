#!/bin/bash
# RACE fine-tuning of BERT-large (parity: examples/finetune_race_distributed.sh).
source "$(dirname "$0")/_common.sh"
TRAIN_DATA="${TRAIN_DATA:-data/RACE/train/middle}"
VALID_DATA="${VALID_DATA:-data/RACE/dev/middle data/RACE/dev/high}"
launch $REPO/tasks/main.py --task RACE --seed 1234 --train_data $TRAIN_DATA --valid_data $VALID_DATA --epochs 3 \
  --pretrained_checkpoint ${PRETRAINED_CHECKPOINT:-checkpoints/bert_345m} --save ${CHECKPOINT_PATH:-checkpoints/bert_345m_race} \
  --tensor_model_parallel_size 1 --micro_batch_size 4 --lr 1.0e-5 --clip_grad 1.0 --hidden_dropout 0.1 --attention_dropout 0.1 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --seq_length 512 --max_position_embeddings 512
  --vocab_file ${VOCAB_FILE:-bert-vocab.txt} --tokenizer_type BertWordPieceLowerCase --lr_decay_style linear --lr_warmup_fraction 0.065
  --save_interval 500000 --log_interval 10 --eval_interval 100 --eval_iters 50 --weight_decay 1.0e-1 --bf16

#!/bin/bash
# T5-base (220M) span-corruption pre-training (parity: examples/pretrain_t5_distributed_with_mp.sh).
source "$(dirname "$0")/_common.sh"
CHECKPOINT_PATH=${CHECKPOINT_PATH:-checkpoints/t5_base}; DATA_PATH=${DATA_PATH:-my-t5_text_sentence}
launch $REPO/pretrain_t5.py --tensor_model_parallel_size 2 --micro_batch_size 16 --global_batch_size 128 --num_layers 12 --hidden_size 768 --num_attention_heads 12 --kv_channels 64 --ffn_hidden_size 3072 --encoder_seq_length 512
  --decoder_seq_length 128 --max_position_embeddings 512 --vocab_file ${VOCAB_FILE:-bert-vocab.txt} --vocab_extra_ids 100
  --tokenizer_type BertWordPieceLowerCase \
  --train_iters 1000000 --lr 0.0001 --min_lr 1.0e-5 --lr_decay_style linear --lr_decay_iters 990000 --lr_warmup_fraction .01
  --weight_decay 1e-2 --clip_grad 1.0 --log_interval 100 --save_interval 10000 --eval_interval 1000 --eval_iters 10 --split 949,50,1 --save $CHECKPOINT_PATH --load $CHECKPOINT_PATH --data_path $DATA_PATH --bf16

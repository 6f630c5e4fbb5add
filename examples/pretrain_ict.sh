#!/bin/bash
# Inverse-cloze-task pre-training of the bi-encoder retriever from a BERT checkpoint (parity: examples/pretrain_ict.sh).
source "$(dirname "$0")/_common.sh"
CHECKPOINT_PATH=${CHECKPOINT_PATH:-checkpoints/ict}; BERT_LOAD_PATH=${BERT_LOAD_PATH:-checkpoints/bert_345m}
TEXT_DATA_PATH=${TEXT_DATA_PATH:-wiki_text_sentence}; TITLE_DATA_PATH=${TITLE_DATA_PATH:-wiki_title_sentence}
launch $REPO/pretrain_ict.py --num_layers 12 --hidden_size 768 --num_attention_heads 12 --tensor_model_parallel_size 1 \
  --micro_batch_size 32 --seq_length 256 --max_position_embeddings 512 --train_iters 100000 --vocab_file ${VOCAB_FILE:-bert-vocab.txt} \
  --tokenizer_type BertWordPieceLowerCase --DDP_impl local --bert_load $BERT_LOAD_PATH --log_interval 100 --eval_interval 1000 \
  --eval_iters 10 --retriever_report_topk_accuracies 1 5 10 20 100 --retriever_score_scaling --load $CHECKPOINT_PATH --save $CHECKPOINT_PATH \
  --data_path $TEXT_DATA_PATH --titles_data_path $TITLE_DATA_PATH --lr 0.0001 --lr_decay_style linear --weight_decay 1e-2 --clip_grad 1.0 \
  --lr_warmup_fraction 0.01 --save_interval 4000 --exit_interval 8000 --query_in_block_prob 0.1 --bf16

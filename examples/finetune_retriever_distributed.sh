#!/bin/bash
# Supervised fine-tuning of the ICT-pretrained retriever on Natural Questions (parity: examples/finetune_retriever_distributed.sh).
source "$(dirname "$0")/_common.sh"
DATA_DIR=${DATA_DIR:-data/nq}
launch $REPO/tasks/main.py --task RET-FINETUNE-NQ --train_with_neg --train_hard_neg 1 --pretrained_checkpoint ${CHECKPOINT_PATH:-checkpoints/ict} \
  --save ${SAVE_PATH:-checkpoints/ret_nq} --num_layers 12 --hidden_size 768 --num_attention_heads 12 --tensor_model_parallel_size 1 \
  --tokenizer_type BertWordPieceLowerCase --vocab_file ${VOCAB_FILE:-bert-vocab.txt} --train_data $DATA_DIR/biencoder-nq-train.json \
  --valid_data $DATA_DIR/biencoder-nq-dev.json --evidence_data_path $DATA_DIR/psgs_w100.tsv --epochs 80 --micro_batch_size 8 \
  --eval_micro_batch_size 16 --indexer_batch_size 128 --lr 2e-5 --lr_warmup_fraction 0.01 --weight_decay 1e-1 --clip_grad 2.0 \
  --seq_length 512 --retriever_seq_length 256 --max_position_embeddings 512 --retriever_score_scaling --log_interval 10 \
  --eval_interval 500 --eval_iters 10 --save_interval 500 --retriever_report_topk_accuracies 1 5 20 100 --bf16 --DDP_impl local

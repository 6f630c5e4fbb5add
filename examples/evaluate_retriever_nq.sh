#!/bin/bash
# Index the evidence with the context encoder and evaluate retrieval on NQ (parity: examples/evaluate_retriever_nq.sh).
source "$(dirname "$0")/_common.sh"
DATA_DIR=${DATA_DIR:-data/nq}
launch $REPO/tasks/main.py --task RETRIEVER-EVAL --tokenizer_type BertWordPieceLowerCase --num_layers 12 --hidden_size 768 \
  --num_attention_heads 12 --tensor_model_parallel_size 1 --micro_batch_size 128 --seq_length 512 --max_position_embeddings 512 \
  --load ${CHECKPOINT_PATH:-checkpoints/ret_nq} --evidence_data_path $DATA_DIR/psgs_w100.tsv --embedding_path $DATA_DIR/evidence_embeds.pkl \
  --retriever_seq_length 256 --vocab_file ${VOCAB_FILE:-bert-vocab.txt} --qa_data_test $DATA_DIR/nq-test.csv --faiss_use_gpu \
  --retriever_report_topk_accuracies 1 5 20 100 --bf16 --indexer_log_interval 1000 --indexer_batch_size 128

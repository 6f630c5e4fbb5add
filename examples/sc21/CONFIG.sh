#!/bin/bash
# Cluster settings + argument assembly (sourced by the run_*.sh scripts after they set TP PP MBS GBS NLS HS NAH DDP NNODES).
export SLURM_PARTITION=${SLURM_PARTITION:-batch}
export SLURM_ACCOUNT=${SLURM_ACCOUNT:-account}
export MEGATRON_CODE_DIR=${MEGATRON_CODE_DIR:-$(cd "$(dirname "${BASH_SOURCE[0]}")/../.." && pwd)}
MEGATRON_DATA=${MEGATRON_DATA:-/data/gpt2_text_document}
BPE_VOCAB_FILE=${BPE_VOCAB_FILE:-/data/gpt2-vocab.json}
BPE_MERGE_FILE=${BPE_MERGE_FILE:-/data/gpt2-merges.txt}
export MEGATRON_PARAMS=" ${MEGATRON_EXTRA_PARAMS} --model_name gpt \
  --tensor_model_parallel_size ${TP} --pipeline_model_parallel_size ${PP} --micro_batch_size ${MBS} --global_batch_size ${GBS} \
  --num_layers ${NLS} --hidden_size ${HS} --num_attention_heads ${NAH} --DDP_impl ${DDP} --data_path ${MEGATRON_DATA} \
  --vocab_file ${BPE_VOCAB_FILE} --merge_file ${BPE_MERGE_FILE} --tokenizer_type GPT2BPETokenizer --log_interval 5 \
  --seq_length 2048 --max_position_embeddings 2048 --train_iters 500 --lr_decay_iters 320 --lr 0.0001 --min_lr 0.00001 \
  --lr_decay_style cosine --lr_warmup_fraction 0.01 --split 969,30,1 --eval_iters 100 --eval_interval 1000 --clip_grad 1.0 --bf16 "

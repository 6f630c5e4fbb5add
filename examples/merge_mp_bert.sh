#!/bin/bash
# Merge a TP=2 BERT checkpoint into a single-rank one (parity: examples/merge_mp_bert.sh; the reference's
# tools/merge_mp_partitions.py no longer exists upstream -- the checkpoint resharder does the same job).
source "$(dirname "$0")/_common.sh"
python $REPO/tools/checkpoint_util.py --model_type BERT --load_dir ${CHECKPOINT_PATH:-checkpoints/bert_345m} \
  --save_dir ${CHECKPOINT_PATH:-checkpoints/bert_345m}-merged --target_tensor_parallel_size 1 --target_pipeline_parallel_size 1

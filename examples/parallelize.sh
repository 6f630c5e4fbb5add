#!/bin/bash
# Reshard a checkpoint to a new TP x PP layout (parity: examples/parallelize.sh).
#   examples/parallelize.sh llama2 7 8 1   ->  /checkpoints/llama2-7b-tp8-pp1
source "$(dirname "$0")/_common.sh"
MODEL=$1; SIZE=$2; TP=$3; PP=$4
ROOT=${CKPT_ROOT:-/checkpoints}
case $MODEL in llama|llama2|codellama|mistral) EXTRA="--true_vocab_size 32000";; *) EXTRA="";; esac
python $REPO/tools/checkpoint_util.py --model_type $MODEL --load_dir $ROOT/${MODEL}-${SIZE}b \
  --save_dir $ROOT/${MODEL}-${SIZE}b-tp$TP-pp$PP --target_tensor_parallel_size $TP --target_pipeline_parallel_size $PP --bf16 $EXTRA

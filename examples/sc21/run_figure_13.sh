#!/bin/bash
# Tensor vs pipeline parallelism on 64 GPUs: PP in {2,4,8,16,32}.
cd "$(dirname "$0")"
PP=${PP:-2}; GBS=${GBS:-32}; TP=$((64/PP)); MBS=1; NLS=32; HS=20480; NAH=128; DDP=local; NNODES=8
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "
export JOB_NAME=results_figure_13_pipeline_parallel_size_${PP}_tensor_parallel_size_${TP}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

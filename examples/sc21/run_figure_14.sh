#!/bin/bash
# Pipeline vs data parallelism on 64 GPUs.
cd "$(dirname "$0")"
PP=${PP:-2}; GBS=${GBS:-32}; DP=$((64/PP)); TP=1; MBS=1; NLS=32; HS=3840; NAH=32; DDP=local; NNODES=8
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "
export JOB_NAME=results_figure_14_pipeline_parallel_size_${PP}_data_parallel_size_${DP}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

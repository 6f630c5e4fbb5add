#!/bin/bash
# Pipeline-parallel weak scaling: PP in {1,2,4,8}, GBS in {8,128}.
cd "$(dirname "$0")"
PP=${PP:-1}; GBS=${GBS:-8}; NLS=$((3*PP)); NNODES=$PP; TP=8; MBS=1; HS=20480; NAH=128; DDP=local
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "
export JOB_NAME=results_figure_11_pipeline_parallel_size_${PP}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

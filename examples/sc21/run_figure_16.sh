#!/bin/bash
# Micro-batch size: MBS in {1,2,4,8}, GBS in {128,512}.
cd "$(dirname "$0")"
MBS=${MBS:-1}; GBS=${GBS:-128}; TP=8; PP=8; NLS=32; HS=15360; NAH=128; DDP=local; NNODES=8
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "
export JOB_NAME=results_figure_16_microbatch_size_${MBS}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

#!/bin/bash
# Activation recomputation on/off.
cd "$(dirname "$0")"
ACTIVATION_RECOMPUTATION=${ACTIVATION_RECOMPUTATION:-YES}; GBS=${GBS:-1}; TP=8; PP=16; MBS=1; NLS=80; HS=12288; NAH=96; DDP=local; NNODES=16
MEGATRON_EXTRA_PARAMS=""; [ $ACTIVATION_RECOMPUTATION = YES ] && MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "
export JOB_NAME=results_figure_17_activation_recomputation_${ACTIVATION_RECOMPUTATION}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

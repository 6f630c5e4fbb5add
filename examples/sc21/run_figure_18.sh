#!/bin/bash
# Scatter/gather of pipeline activations over the TP group on/off.
cd "$(dirname "$0")"
SCATTER_GATHER=${SCATTER_GATHER:-YES}; GBS=${GBS:-12}; TP=8; PP=12; MBS=1; NLS=96; HS=12288; NAH=96; DDP=local; NNODES=12
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform --num_layers_per_virtual_pipeline_stage 2 "
[ $SCATTER_GATHER = NO ] && MEGATRON_EXTRA_PARAMS+="--no_scatter_gather_tensors_in_pipeline "
export JOB_NAME=results_figure_18_scatter_gather_${SCATTER_GATHER}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

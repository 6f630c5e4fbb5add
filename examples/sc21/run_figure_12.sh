#!/bin/bash
# Interleaved schedule on/off: GBS in {12,24,36,48,60}.
cd "$(dirname "$0")"
INTERLEAVED=${INTERLEAVED:-YES}; GBS=${GBS:-12}; TP=8; PP=12; MBS=1; NLS=96; HS=12288; NAH=96; DDP=local; NNODES=12
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "; [ $INTERLEAVED = YES ] && MEGATRON_EXTRA_PARAMS+="--num_layers_per_virtual_pipeline_stage 2 "
export JOB_NAME=results_figure_12_interleaved_${INTERLEAVED}_batch_size_${GBS}
. ./CONFIG.sh
. ./SBATCH.sh

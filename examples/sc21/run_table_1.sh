#!/bin/bash
# Weak scaling table: MODEL_SIZE in {1.7B,3.6B,7.5B,18B,39B,76B,145B,310B,530B,1T}.
cd "$(dirname "$0")"
MODEL_SIZE=${MODEL_SIZE:-1.7B}
#            TP PP MBS GBS  NLS HS    NAH NNODES layers/virtual-stage
declare -A T=( [1.7B]="1 1 16 512 24 2304 24 4 0"    [3.6B]="2 1 16 512 30 3072 32 8 0"     [7.5B]="4 1 16 512 36 4096 32 16 0"
               [18B]="8 1 8 1024 40 6144 48 32 0"    [39B]="8 2 4 1536 48 8192 64 64 0"     [76B]="8 4 2 1792 60 10240 80 128 5"
               [145B]="8 8 2 2304 80 12288 96 192 5" [310B]="8 16 1 2160 96 16384 128 240 3" [530B]="8 35 1 2520 105 20480 128 315 1"
               [1T]="8 64 1 3072 128 25600 160 384 0" )
[ -z "${T[$MODEL_SIZE]}" ] && { echo "Invalid configuration"; exit 1; }
read TP PP MBS GBS NLS HS NAH NNODES VPP <<< "${T[$MODEL_SIZE]}"; DDP=local
MEGATRON_EXTRA_PARAMS="--recompute_granularity full --recompute_method uniform "; [ $VPP != 0 ] && MEGATRON_EXTRA_PARAMS+="--num_layers_per_virtual_pipeline_stage $VPP "
export JOB_NAME=results_table_1_model_size_${MODEL_SIZE}
. ./CONFIG.sh
. ./SBATCH.sh

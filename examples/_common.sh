#!/bin/bash
# Shared launcher pieces for the example scripts (sourced).  One process per GPU over NCCL/NVLink; single node by default.
REPO=$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)
GPUS_PER_NODE=${GPUS_PER_NODE:-8}
NNODES=${NNODES:-1}
NODE_RANK=${NODE_RANK:-0}
MASTER_ADDR=${MASTER_ADDR:-127.0.0.1}
MASTER_PORT=${MASTER_PORT:-6000}
DISTRIBUTED_ARGS="--nproc_per_node $GPUS_PER_NODE --nnodes $NNODES --node_rank $NODE_RANK --master_addr $MASTER_ADDR --master_port $MASTER_PORT"
launch () { python -m torch.distributed.run $DISTRIBUTED_ARGS "$@"; }
# B200 defaults: bf16, hand-written tcgen05 attention + fused TP collectives are on by default
COMMON_ARGS="--bf16 --use_flash_attn --no_bias_gelu_fusion --no_bias_dropout_fusion"

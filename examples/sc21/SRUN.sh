#!/bin/bash
#SBATCH -t 0:30:00 --exclusive --mem=0 --overcommit
# One process per GPU; rank / world size come from SLURM.
export MASTER_ADDR=$(scontrol show hostnames $SLURM_JOB_NODELIST | head -n 1); export MASTER_PORT=6000
srun --ntasks-per-node=8 bash -c 'RANK=$SLURM_PROCID WORLD_SIZE=$SLURM_NTASKS LOCAL_RANK=$SLURM_LOCALID \
  python -u ${MEGATRON_CODE_DIR}/finetune.py ${MEGATRON_PARAMS}'

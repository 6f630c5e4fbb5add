#!/bin/bash
# Submit the job assembled by CONFIG.sh.
mkdir -p logs
sbatch -p ${SLURM_PARTITION} -A ${SLURM_ACCOUNT} --job-name=${JOB_NAME} --nodes=${NNODES} --ntasks-per-node=8 --exclusive \
  --export=ALL,MEGATRON_CODE_DIR,MEGATRON_PARAMS -o logs/${JOB_NAME}.log "$(dirname "${BASH_SOURCE[0]}")/SRUN.sh"

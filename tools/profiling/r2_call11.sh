#!/bin/bash
# round-2 call 11 (2 GPUs): the default policy (selector + async Column backward) against the all-fused policy at N=2
mkdir -p gpurun_out/r2c11
O=gpurun_out/r2c11
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29600 bench.py --gpus 2 --steps 4 --warmup 3 --no_e2e > $O/bench_n2_default.json 2> $O/bench_n2_default.err; echo "default rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n2_default.json | head -1) $(grep -o '"exposed_tp_collective_ms_per_step": {[^}]*}' $O/bench_n2_default.json)"
MLB200_FUSED_TP_FORCE=1 timeout 200 $TR --master-port 29601 bench.py --gpus 2 --steps 4 --warmup 3 --no_e2e > $O/bench_n2_force.json 2> $O/bench_n2_force.err; echo "force rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n2_force.json | head -1) $(grep -o '"exposed_tp_collective_ms_per_step": {[^}]*}' $O/bench_n2_force.json)"

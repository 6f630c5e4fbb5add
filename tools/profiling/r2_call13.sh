#!/bin/bash
# round-2 call 13 (4 GPUs): headline config at N=4, both arms (the scaling points N=1/2/8 are measured already)
mkdir -p gpurun_out/r2c13
O=gpurun_out/r2c13
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29610 bench.py --gpus 4 --steps 4 --warmup 3 --no_e2e > $O/bench_n4_ours.json 2> $O/bench_n4_ours.err; echo "ours rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n4_ours.json | head -1) $(grep -o '"exposed_tp_collective_ms_per_step": {[^}]*}' $O/bench_n4_ours.json)"
timeout 240 $TR --master-port 29611 bench.py --impl reference --gpus 4 --steps 4 --warmup 3 > $O/bench_n4_ref.json 2> $O/bench_n4_ref.err; echo "ref rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n4_ref.json | head -1)"

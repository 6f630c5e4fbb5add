#!/bin/bash
mkdir -p gpurun_out
run () { # name, env...
  name=$1; shift
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_n8_$name.json 2> gpurun_out/bench_7b_n8_$name.err
  echo "== $name rc=$?"; tail -1 gpurun_out/bench_7b_n8_$name.json | cut -c1-330; grep -iE "error|Traceback|WARNING: symm" gpurun_out/bench_7b_n8_$name.err | head -5
}
run fused MLB200_FUSED_TP=1
run nccl MLB200_FUSED_TP=0
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; head -12 gpurun_out/topo_n8.txt

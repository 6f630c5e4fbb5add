#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_n8_$name.json 2> gpurun_out/bench_7b_n8_$name.err
  echo "== $name rc=$?"; tail -1 gpurun_out/bench_7b_n8_$name.json | cut -c1-330; grep -v "^\s*$" gpurun_out/bench_7b_n8_$name.err | grep -iE "error" | head -5 | cut -c1-300
}
run graph MLB200_BENCH_GRAPH=1

#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_n2_$name.json 2> gpurun_out/bench_7b_n2_$name.err
  echo "== $name rc=$?"; tail -1 gpurun_out/bench_7b_n2_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')}, d['config'].get('cuda_graph_microbatch'))" || (grep -v "^\s*$" gpurun_out/bench_7b_n2_$name.err | grep "rank0\]" | tail -12 | cut -c1-250)
}
run graph MLB200_BENCH_GRAPH=1
run nccl MLB200_BENCH_GRAPH=0 MLB200_FUSED_TP=0
run fused MLB200_BENCH_GRAPH=0 MLB200_FUSED_TP=1

#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_n8_graph.json 2> gpurun_out/bench_7b_n8_graph.err
echo "== graph rc=$?"; tail -1 gpurun_out/bench_7b_n8_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')}, d['config'].get('cuda_graph_microbatch'))" || (grep -v "^\s*$" gpurun_out/bench_7b_n8_graph.err | grep "rank0\]" | tail -12 | cut -c1-250)

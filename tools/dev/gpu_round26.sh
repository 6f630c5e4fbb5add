#!/bin/bash
# N=8: fused TP microbench + flagship bench with fused kernels inside the micro-batch graph
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/dev/fused_bench.py > gpurun_out/r26_fused_bench_n8.jsonl 2> gpurun_out/r26_fused_bench_n8.err
echo "microbench exit $?"; tail -3 gpurun_out/r26_fused_bench_n8.err
MLB200_FUSED_TP=1 MLB200_BENCH_GRAPH=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r26_bench_n8_fused_graph.json 2> gpurun_out/r26_bench_n8_fused_graph.err
echo "bench fused+graph exit $?"; tail -3 gpurun_out/r26_bench_n8_fused_graph.err; cat gpurun_out/r26_bench_n8_fused_graph.json

#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
MLB200_FUSED_TP=1 timeout 400 python -m pytest tests/test_fused_comm_gpu.py -m gpu -x -q > gpurun_out/r25_fused_test.log 2>&1
echo "fused test exit $?"; tail -5 gpurun_out/r25_fused_test.log
for c in 32; do
MLB200_COMM_CTAS=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dev/fused_bench.py > gpurun_out/r25_fused_bench_c$c.jsonl 2> gpurun_out/r25_fused_bench_c$c.err
echo "exit $?"; tail -3 gpurun_out/r25_fused_bench_c$c.err
done
MLB200_FUSED_TP=1 MLB200_BENCH_GRAPH=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r25_bench_n2_fused_graph.json 2> gpurun_out/r25_bench_n2_fused_graph.err
echo "bench fused+graph exit $?"; tail -3 gpurun_out/r25_bench_n2_fused_graph.err; cat gpurun_out/r25_bench_n2_fused_graph.json

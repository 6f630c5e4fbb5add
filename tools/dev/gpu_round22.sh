#!/bin/bash
# N=2: fused 2-CTA TP kernels (correctness + bench) and the graph-vs-eager test
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
MLB200_FUSED_TP=1 timeout 400 python -m pytest tests/test_fused_comm_gpu.py -m gpu -x -q > gpurun_out/r22_fused_test.log 2>&1
echo "fused test exit $?" | tee -a gpurun_out/r22_fused_test.log
tail -15 gpurun_out/r22_fused_test.log
MLB200_FUSED_TP=1 MLB200_BENCH_GRAPH=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r22_bench_n2_fused2.json 2> gpurun_out/r22_bench_n2_fused2.err
echo "bench fused exit $?"; tail -3 gpurun_out/r22_bench_n2_fused2.err; cat gpurun_out/r22_bench_n2_fused2.json
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "cuda_graph" > gpurun_out/r22_graph_test.log 2>&1
echo "graph test exit $?"; tail -8 gpurun_out/r22_graph_test.log

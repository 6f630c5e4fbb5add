#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
for c in 8 16; do
MLB200_COMM_CTAS=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dev/fused_bench.py > gpurun_out/r24_fused_bench_c$c.jsonl 2> gpurun_out/r24_fused_bench_c$c.err
echo "exit $?"; tail -3 gpurun_out/r24_fused_bench_c$c.err
done

#!/bin/bash
# (record of the first round-2 GPU session: the stream / push all-gather variants it measured were deleted afterwards)
# round-2 call 1 (2 GPUs): NVLink instruction-path rates + the opt-in all-gather transports
mkdir -p gpurun_out/r2c1
export MASTER_ADDR=127.0.0.1
O=gpurun_out/r2c1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt 2>&1
echo "== nvlink"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
  tools/profiling/nvlink_bench.py 16 > $O/nvlink_n2.jsonl 2> $O/nvlink.err
tail -3 $O/nvlink.err; wc -l $O/nvlink_n2.jsonl
echo "== fused tests (experimental variants)"
MLB200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_fused_comm_gpu.py -m gpu -q -x > $O/fused_tests.log 2>&1; tail -5 $O/fused_tests.log
echo "== fused bench"
for v in default stream push; do
  env=""; [ $v = stream ] && env="MLB200_AG_STREAM=1"; [ $v = push ] && env="MLB200_AG_PUSH=1"
  for seq in 4096 1024; do
  env $env timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29562 tools/profiling/fused_bench.py 4096 11008 $seq > $O/fused_bench_n2_${v}_s$seq.jsonl 2> $O/fused_bench_${v}_s$seq.err
  echo "$v $seq: exit $? ($(grep -c '^{' $O/fused_bench_n2_${v}_s$seq.jsonl) shapes)"
  done
done

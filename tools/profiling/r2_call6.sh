#!/bin/bash
# round-2 call 6 (1 GPU): GPU suite after the latest changes, single-rank fused-kernel overhead, elementwise kernel timings
mkdir -p gpurun_out/r2c6
O=gpurun_out/r2c6
export MASTER_ADDR=127.0.0.1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.log
echo "== fused kernels, one virtual rank, vs plain GEMM"
for g in auto 2 8; do
  if [ $g = auto ]; then timeout 200 python tools/profiling/fused_single_bench.py 8192 > $O/fused_single_$g.jsonl 2> $O/fused_single_$g.err
  else MLB200_FUSED_GROUP=$g timeout 200 python tools/profiling/fused_single_bench.py 8192 > $O/fused_single_$g.jsonl 2> $O/fused_single_$g.err; fi
  echo "group $g:"; cat $O/fused_single_$g.jsonl
done
timeout 200 python tools/profiling/fused_single_bench.py 4096 > $O/fused_single_m4096.jsonl 2>> $O/fused_single_auto.err; cat $O/fused_single_m4096.jsonl
echo "== elementwise kernels"
timeout 300 python tools/profiling/ew_drive.py --time > $O/ew_times.jsonl 2> $O/ew_times.err; cat $O/ew_times.jsonl; tail -3 $O/ew_times.err

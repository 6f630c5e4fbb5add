#!/bin/bash
# round-2 call 7 (1 GPU): evidence -- ncu --set full of every hot kernel, final-build step breakdown, sanitizer
mkdir -p gpurun_out/r2c7
O=gpurun_out/r2c7
export MASTER_ADDR=127.0.0.1
echo "== bench N=1 (micro-batch graph on, pipelined norm backward)"
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; tail -c 1600 $O/bench_n1.json
echo "== elementwise kernels"
timeout 200 python tools/profiling/ew_drive.py --time > $O/ew_times.jsonl 2> $O/ew_times.err; cat $O/ew_times.jsonl
echo "== step breakdown (8 layers, torch profiler)"
timeout 300 python tools/profiling/profile_step.py 8 $O/step_breakdown_8layers_r2.txt > /dev/null 2> $O/step_breakdown.err; head -30 $O/step_breakdown_8layers_r2.txt
echo "== ncu captures"
bash tools/profiling/ncu_kernels.sh $O 2>&1 | tail -25
echo "== compute-sanitizer memcheck + synccheck (small cases)"
out=$O/sanitize; mkdir -p $out
for t in memcheck; do
  timeout 700 compute-sanitizer --tool $t --error-exitcode 3 --launch-timeout 120 \
    python -m pytest tests/test_ops_gpu.py tests/test_fused_loopback_gpu.py -m gpu -q -x \
      -k "not 4096 and not 2048 and not 8192 and not world8 and not 8-256 and not fp16_operands and not norm_fwd_bwd" \
    > $out/$t.log 2>&1
  echo "$t: exit $? ; $(grep -h 'ERROR SUMMARY' $out/$t.log | tail -1) ; $(tail -1 $out/$t.log)"
done
ls -la $O | head -40

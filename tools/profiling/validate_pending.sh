#!/bin/bash
# One GPU-box call that validates everything that was written without hardware access (see docs/NEXT_STEPS.md):
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/profiling/validate_pending.sh'
# Each step is bounded by its own timeout; results land in gpurun_out/pending_*.
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
N=$(nvidia-smi -L | wc -l)
echo "== GPUs: $N"

echo "== 1. head_dim 64 attention (1 GPU)"
bash tools/profiling/attn_hd64_checks.sh > gpurun_out/pending_hd64.log 2>&1; tail -12 gpurun_out/pending_hd64.log

if [ "$N" -ge 2 ]; then
  echo "== 2. fused TP variants: streaming pullers, push all-gather, rank skew (2 GPUs)"
  MLB200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_fused_comm_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q \
    > gpurun_out/pending_fused_tests.log 2>&1; tail -6 gpurun_out/pending_fused_tests.log

  echo "== 3. NVLink transfer rates per instruction path"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29561 \
    tools/profiling/nvlink_bench.py $([ "$N" -ge 8 ] && echo 4 || echo 16) > gpurun_out/pending_nvlink_n$N.jsonl 2> gpurun_out/pending_nvlink.err
  tail -3 gpurun_out/pending_nvlink.err; wc -l gpurun_out/pending_nvlink_n$N.jsonl

  echo "== 4. per-pair fused times: default pullers / streaming / push"
  for v in default stream push; do
    env=""; [ $v = stream ] && env="MLB200_AG_STREAM=1"; [ $v = push ] && env="MLB200_AG_PUSH=1"
    env $env timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29562 tools/profiling/fused_bench.py > gpurun_out/pending_fused_bench_n${N}_$v.jsonl 2> gpurun_out/pending_fused_bench_$v.err
    echo "$v: exit $? ($(grep -c '^{' gpurun_out/pending_fused_bench_n${N}_$v.jsonl) shapes)"
  done
  python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/pending_fused_bench_n*_*.jsonl")):
    rows = [json.loads(l) for l in open(f) if l.startswith("{")]
    if rows:
        print(f.split("/")[-1], "sum fused %.0f us, sum nccl+gemm %.0f us, sum gemm %.0f us" % (
            1e3 * sum(r["fused_ms"] for r in rows), 1e3 * sum(r["nccl_plus_gemm_ms"] for r in rows),
            1e3 * sum(r["gemm_ms"] for r in rows)))
PY
fi

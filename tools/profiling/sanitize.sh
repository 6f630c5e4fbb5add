#!/bin/bash
# compute-sanitizer over the kernel numerics tests (one GPU).  memcheck = out-of-bounds / misaligned accesses,
# racecheck = shared-memory hazards, synccheck = illegal barrier use, initcheck = reads of uninitialised global memory.
#   tools/profiling/sanitize.sh [tool ...]        default: memcheck synccheck
# The tcgen05 / TMA kernels synchronise through mbarriers and the async proxy, which racecheck does not model: expect
# false positives there and read its report per kernel.  Sizes are the small cases of tests/ (seconds each natively,
# minutes under the sanitizer).  Without a GPU: the barrier protocols of exactly those kernels (mbarrier phases included)
# are checked by ThreadSanitizer on the functional model -- python -m pytest tests -q -k "thread_sanitizer or races".
out=gpurun_out/sanitize; mkdir -p $out
tools=${@:-memcheck synccheck}
for t in $tools; do
  timeout 1500 compute-sanitizer --tool $t --error-exitcode 3 --launch-timeout 120 \
    python -m pytest tests/test_ops_gpu.py tests/test_fused_loopback_gpu.py -m gpu -q -k "not 4096 and not 2048 and not world8 and not 8-256" \
    > $out/$t.log 2>&1
  echo "$t: exit $? ($(grep -c 'ERROR SUMMARY' $out/$t.log) summaries; $(grep -h 'ERROR SUMMARY' $out/$t.log | tail -1))"
done

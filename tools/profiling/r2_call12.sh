#!/bin/bash
# round-2 call 12 (1 GPU): what the driver runs at round end -- smoke() and the whole GPU suite -- on the final code
mkdir -p gpurun_out/r2c12
O=gpurun_out/r2c12
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log

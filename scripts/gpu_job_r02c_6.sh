#!/bin/bash
# session 3, call 6: deferred run factor of the T-side gather: A/B + parity suite
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/exp_defer.py 1300000 > gpurun_out/exp_defer.log 2>&1; grep -E "gather_defer|Error|error" gpurun_out/exp_defer.log | head
timeout 600 python scripts/exp_defer.py 162500 > gpurun_out/exp_defer_162k.log 2>&1; grep -E "gather_defer|Error|error" gpurun_out/exp_defer_162k.log | head
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c6.log

#!/bin/bash
# round 2, session 2, job 2: GPU test suite with the new kernel variants on by default, A/B of the variants at C3 size
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/exp_variants.py 1300000 > gpurun_out/exp_variants.log 2>&1
cat gpurun_out/exp_variants.log
SCANB200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -k "experimental" > gpurun_out/pytest_experimental.log 2>&1
tail -5 gpurun_out/pytest_experimental.log
timeout 600 python scripts/exp_gather_split.py 1300000 > gpurun_out/exp_gather_split.log 2>&1
cat gpurun_out/exp_gather_split.log

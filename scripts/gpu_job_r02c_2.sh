#!/bin/bash
# session 3, call 2: calibration indexed by the share a CTA drew (was: by block index), per-CTA {SM, cycles} dump
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/gather_dump.txt
SCANB200_GATHER_DUMP=gpurun_out/gather_dump.txt SCANB200_TRACE=1 timeout 600 python scripts/exp_calibrate.py 1300000 > gpurun_out/exp_calibrate2.log 2>&1; grep -E "gather_calibrate|recalibrated" gpurun_out/exp_calibrate2.log
timeout 600 python scripts/exp_calibrate.py 162500 > gpurun_out/exp_calibrate2_162k.log 2>&1; grep -E "gather_calibrate" gpurun_out/exp_calibrate2_162k.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c2.log

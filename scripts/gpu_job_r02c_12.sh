#!/bin/bash
# session 3, call 12: skinny tall GEMM A/B, full GPU suite
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/exp_skinny.py 1300000 > gpurun_out/exp_skinny.log 2>&1; grep -E "gemm_skinny|rror" gpurun_out/exp_skinny.log | head
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c12.log

#!/bin/bash
# session 3, call 11: host top-k eigensolver A/B, full GPU suite, C++ host test
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/exp_eig.py 1300000 > gpurun_out/exp_eig.log 2>&1; grep -E "eig_host|rror" gpurun_out/exp_eig.log | head
timeout 600 python scripts/exp_eig.py 162500 > gpurun_out/exp_eig_162k.log 2>&1; grep -E "eig_host|rror" gpurun_out/exp_eig_162k.log | head
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c11.log

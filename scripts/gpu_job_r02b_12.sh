#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k_e.json 2> gpurun_out/bench_n1_162k_e.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_v5.json 2> gpurun_out/bench_n1_v5.err; tail -c 300 gpurun_out/bench_n1_v5.err
timeout 300 python -m pytest tests -m gpu -q -k "plane or products or golden or bksvd" > gpurun_out/pytest_ws.log 2>&1; tail -3 gpurun_out/pytest_ws.log

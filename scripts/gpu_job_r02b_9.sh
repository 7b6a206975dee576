#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bksvd or randsvd or config1 or shortcut or k30 or k50 or k100 or drift or cpp_host or irlba or variance or properties" > gpurun_out/pytest_dense2.log 2>&1; tail -4 gpurun_out/pytest_dense2.log
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k_b.json 2> gpurun_out/bench_n1_162k_b.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err; tail -c 300 gpurun_out/bench_n1_v2.err

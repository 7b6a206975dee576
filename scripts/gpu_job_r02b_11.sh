#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k_d.json 2> gpurun_out/bench_n1_162k_d.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_v4.json 2> gpurun_out/bench_n1_v4.err; tail -c 300 gpurun_out/bench_n1_v4.err
timeout 600 python scripts/exp_segcost.py 1300000 > gpurun_out/exp_segcost.log 2>&1; cat gpurun_out/exp_segcost.log
timeout 300 python -m pytest tests -m gpu -q -k "bksvd or randsvd or config1 or shortcut or k30 or k100" > gpurun_out/pytest_dense3.log 2>&1; tail -3 gpurun_out/pytest_dense3.log

#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_162k.csv python bench.py --cells 162500 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_162k.json 2> gpurun_out/bench_ncu_162k.err
tail -c 300 gpurun_out/bench_ncu_162k.err
timeout 300 python -m pytest tests -m gpu -q -k "bksvd or randsvd or config1 or shortcut" > gpurun_out/pytest_dense.log 2>&1; tail -3 gpurun_out/pytest_dense.log

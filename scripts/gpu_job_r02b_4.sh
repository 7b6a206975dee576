#!/bin/bash
# round 2, session 2, job 4: re-run of the fixed tests, decomposition of k_planes_t, bench N=1 with the new defaults, ncu of the V1 kernels
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "moment or cpp_host or irlba or h5 or multi_gpu" > gpurun_out/pytest_fix.log 2>&1; tail -5 gpurun_out/pytest_fix.log
timeout 600 python scripts/exp_pl_debug2.py 1300000 > gpurun_out/exp_pl_debug2.log 2>&1; cat gpurun_out/exp_pl_debug2.log
timeout 900 python bench.py > gpurun_out/bench_n1_v1.json 2> gpurun_out/bench_n1_v1.err; tail -c 400 gpurun_out/bench_n1_v1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_planes_t|k_planes_n' --launch-skip 2 -c 4 \
  -o gpurun_out/planes_v1_r02 -f python scripts/prof_passes.py 400000 2 2 > gpurun_out/ncu_planes_v1.log 2>&1
tail -3 gpurun_out/ncu_planes_v1.log

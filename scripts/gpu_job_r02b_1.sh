#!/bin/bash
# round 2, session 2, job 1: fixed-cost trace at one rank's share of C3 on 8 GPUs, ncu --set full of the four SpMM kernels, bench N=1
set -x
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1
SCANB200_TRACE=1 timeout 300 python scripts/trace_step.py 162500 > gpurun_out/trace_162k.out 2> gpurun_out/trace_162k.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_planes_t|k_planes_n|k_gather|k_pl_digits_n|k_pl_reduce_t' --launch-skip 8 -c 8 \
  -o gpurun_out/spmm_r02 -f python scripts/prof_passes.py 400000 2 2 > gpurun_out/ncu_spmm.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.err
ls -la gpurun_out

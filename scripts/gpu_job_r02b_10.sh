#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log; tail -6 gpurun_out/pytest_gpu2.log | cut -c1-250
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k_c.json 2> gpurun_out/bench_n1_162k_c.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_v3.json 2> gpurun_out/bench_n1_v3.err; tail -c 300 gpurun_out/bench_n1_v3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_v3.json 2> gpurun_out/bench_n2_v3.err; tail -c 300 gpurun_out/bench_n2_v3.err

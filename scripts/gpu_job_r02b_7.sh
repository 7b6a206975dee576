#!/bin/bash
# 2 GPUs: one rank's share of the 8-GPU C3 run (162.5k cells per rank) with and without a second rank -- isolates what the collectives
# and the rank skew cost per step; then the multi-GPU parity worker (new sharded paths) at world 2
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/exp_items.py 1300000 > gpurun_out/exp_items2.log 2>&1; cat gpurun_out/exp_items2.log
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k.json 2> gpurun_out/bench_n1_162k.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --cells 325000 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n2_325k.json 2> gpurun_out/bench_n2_325k.err
tail -c 300 gpurun_out/bench_n2_325k.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/mgpu_worker.py > gpurun_out/mgpu_parity_w2.log 2>&1
grep MGPU_PARITY_OK gpurun_out/mgpu_parity_w2.log; tail -5 gpurun_out/mgpu_parity_w2.log
timeout 600 python -m pytest tests -m gpu -q -k "single_process_multi_gpu or two_gpu" > gpurun_out/pytest_mgpu.log 2>&1; tail -3 gpurun_out/pytest_mgpu.log

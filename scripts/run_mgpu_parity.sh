#!/bin/bash
# Multi-GPU parity against the CPU oracle at world sizes 2, 4, 8 (as many as the box has), then bench.py at each size.
# usage (under gpurun --gpus N): bash scripts/run_mgpu_parity.sh <outdir> [bench steps]
OUT=${1:-gpurun_out/mgpu}; STEPS=${2:-5}
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
for W in 2 4 8; do
  [ $W -le $NG ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500+W)) tests/mgpu_worker.py > $OUT/parity_w$W.log 2>&1
  echo "world $W parity rc=$?" | tee -a $OUT/summary.log
  grep MGPU_PARITY_OK $OUT/parity_w$W.log | tee -a $OUT/summary.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29600+W)) bench.py --gpus $W --steps $STEPS --warmup 3 > $OUT/bench_n$W.json 2> $OUT/bench_n$W.err
  echo "world $W bench rc=$?" | tee -a $OUT/summary.log
done

#!/bin/bash
# 8-GPU session: oracle parity at world 4 and 8, then bench.py on 8 GPUs for the headline config and the large / skewed ones.
OUT=${1:-gpurun_out/mgpu8}
mkdir -p $OUT
for W in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29500+W)) tests/mgpu_worker.py > $OUT/parity_w$W.log 2>&1
  echo "world $W parity rc=$?" | tee -a $OUT/summary.log
  grep MGPU_PARITY_OK $OUT/parity_w$W.log | tee -a $OUT/summary.log
done
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name rc=$?" | tee -a $OUT/summary.log
}
run c3_n8 --steps 10 --warmup 3
run c4_n8 --config c4 --steps 3 --warmup 2 --e2e-steps 1
run c5_n8 --config c5 --steps 5 --warmup 2 --e2e-steps 1
nvidia-smi topo -m > $OUT/topo.txt 2>&1

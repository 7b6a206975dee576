#!/bin/bash
set -x
cd /root/repo/profiles/microbench
for b in umma_i8_rate ring_rate sync_latency; do nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o $b $b.cu; done
mkdir -p /root/repo/gpurun_out
(echo "== umma_i8_rate"; timeout 120 ./umma_i8_rate; echo "== ring_rate"; timeout 120 ./ring_rate; echo "== sync_latency"; timeout 120 ./sync_latency) > /root/repo/gpurun_out/microbench_r02.txt 2>&1
cat /root/repo/gpurun_out/microbench_r02.txt

#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
SCANB200_TRACE=1 timeout 300 python scripts/trace_step.py 1300000 > gpurun_out/trace_1300k.out 2> gpurun_out/trace_1300k.log
tail -45 gpurun_out/trace_1300k.log
timeout 300 python scripts/trace_step.py 1300000 2>&1 | tail -5

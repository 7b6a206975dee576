#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
(for n in 1300000 162500; do timeout 600 python scripts/exp_variants.py $n; done) > gpurun_out/exp_variants2.log 2>&1; cat gpurun_out/exp_variants2.log
(for n in 1300000 162500; do timeout 600 python scripts/exp_items.py $n; done) > gpurun_out/exp_items.log 2>&1; cat gpurun_out/exp_items.log
timeout 300 python -m pytest tests -m gpu -q -k "plane or products or golden" > gpurun_out/pytest_planes.log 2>&1; tail -3 gpurun_out/pytest_planes.log

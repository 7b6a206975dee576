#!/bin/bash
# round 2, session 2, job 3: whole GPU suite (new tests: IRLBA, moment consumers, HDF5 loader, sb_multi; shutdown fix), A/B of the
# second-generation T-side kernel at three sizes, plane-density sweep, fixed-cost trace at one rank's share of the 8-GPU run
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for n in 1300000 400000 162500; do timeout 600 python scripts/exp_variants.py $n; done > gpurun_out/exp_variants.log 2>&1
cat gpurun_out/exp_variants.log
timeout 600 python scripts/exp_plane_density.py 1300000 > gpurun_out/exp_plane_density.log 2>&1
cat gpurun_out/exp_plane_density.log

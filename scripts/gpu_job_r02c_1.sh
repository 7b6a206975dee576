#!/bin/bash
# session 3, call 1: full GPU suite on the current tree, A/B of the calibrated T-side gather shares, C3 bench line
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c1.log
SCANB200_TRACE=1 timeout 600 python scripts/exp_calibrate.py 1300000 > gpurun_out/exp_calibrate.log 2>&1; grep -E "gather_calibrate|recalibrated" gpurun_out/exp_calibrate.log
timeout 600 python scripts/exp_calibrate.py 162500 > gpurun_out/exp_calibrate_162k.log 2>&1; grep -E "gather_calibrate" gpurun_out/exp_calibrate_162k.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c1_c3.json 2> gpurun_out/bench_c1_c3.err; tail -c 300 gpurun_out/bench_c1_c3.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_c1_c3.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', d['e2e']['ms_per_step'], d['parity']['ok'], d['roofline']['frac'])
PY

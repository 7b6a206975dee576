#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log; tail -4 gpurun_out/pytest_gpu3.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_n1_v7.json 2> gpurun_out/bench_n1_v7.err; tail -c 300 gpurun_out/bench_n1_v7.err
timeout 600 python bench.py --cells 162500 --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/bench_n1_162k_f.json 2> gpurun_out/bench_n1_162k_f.err
python - <<'PY'
import json
for f in ('bench_n1_v7','bench_n1_162k_f'):
    d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline']['phase_ms_per_step'], (d.get('e2e') or {}).get('ms_per_step'))
PY

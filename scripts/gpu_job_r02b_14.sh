#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_v6a.json 2> gpurun_out/bench_n1_v6a.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_v6b.json 2> gpurun_out/bench_n1_v6b.err; tail -c 300 gpurun_out/bench_n1_v6b.err
python - <<'PY'
import json
for f in ('bench_n1_v6a','bench_n1_v6b'):
    d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['roofline']['phase_ms_per_step'], d['wall_s_timed_region'])
PY

#!/bin/bash
# session 3, call 14: SYRK / GEMM tile skipping: full GPU suite + C3 bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c14.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c14.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c14_c3.json 2> gpurun_out/bench_c14_c3.err; tail -c 300 gpurun_out/bench_c14_c3.err
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c14_c2.json 2> gpurun_out/bench_c14_c2.err
python - <<'PY'
import json
for f in ('bench_c14_c3','bench_c14_c2'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), e['calls_ms_host_clock[upload,normalize+pca,free]'], e.get('host_form_build_s_outside_timed_region'), d['parity']['ok'], round(d['roofline']['frac'],3))
    except Exception as ex:
        print(f, 'ERR', ex)
PY

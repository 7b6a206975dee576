#!/bin/bash
# session 3, call 16: cp.async-staged SYRK / GEMM with tile skipping: tests, per-launch times, bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c16.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:'k_syrk_tall|k_gemm_tall|k_gemm_skinny' --csv --log-file gpurun_out/dense_launches_c16.csv python scripts/trace_step.py 1300000 > gpurun_out/trace_c16.out 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/root/repo/gpurun_out/dense_launches_c16.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
byid={}
for r in rows:
    byid.setdefault(r['ID'],{'name':r['Kernel Name'][:14]})[r['Metric Name']]=r['Metric Value']
big=[v for v in byid.values() if float(v.get('gpu__time_duration.sum','0').replace(',',''))>200000]
for v in big[-3:]: print(v)
PY
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c16_c3.json 2> gpurun_out/bench_c16_c3.err; tail -c 300 gpurun_out/bench_c16_c3.err
python - <<'PY'
import json
for f in ('bench_c16_c3',):
    d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
    e=d['e2e']
    print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), d['parity']['ok'], round(d['roofline']['frac'],3))
PY

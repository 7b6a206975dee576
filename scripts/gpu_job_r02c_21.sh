#!/bin/bash
# session 3, call 21: row-tile GEMM with compile-time triangular skipping (and the live-tile SYRK): per-launch times, full GPU suite, bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_syrk|k_gemm_tall|k_gemm_rows' --csv --log-file gpurun_out/dense_launches_c21.csv python scripts/trace_step.py 1300000 > gpurun_out/trace_c21.out 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/root/repo/gpurun_out/dense_launches_c21.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
byid={}
for r in rows:
    byid.setdefault(r['ID'],{'name':r['Kernel Name'][:22]})[r['Metric Name']]=r['Metric Value']
vals=list(byid.values())
big=[v for v in vals if float(v.get('gpu__time_duration.sum','0').replace(',',''))>200000]
for v in big[-2:]: print(v)
small=[v for v in vals if 'k_gemm_rows' in v['name']]
print('gemm_rows launches (ns):', [v['gpu__time_duration.sum'] for v in small[-19:-1]])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c21.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c21_c3.json 2> gpurun_out/bench_c21_c3.err; tail -c 300 gpurun_out/bench_c21_c3.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_c21_c3.json').read().strip().splitlines()[-1])
e=d['e2e']
print(round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), d['parity']['ok'], round(d['roofline']['frac'],3))
PY

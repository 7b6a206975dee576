#!/bin/bash
# session 3, call 15: per-launch times of the dense kernels after the tile skipping
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:'k_syrk_tall|k_gemm_tall|k_gemm_skinny' --csv --log-file gpurun_out/dense_launches_c15.csv python scripts/trace_step.py 1300000 > gpurun_out/trace_c15.out 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/root/repo/gpurun_out/dense_launches_c15.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
byid={}
for r in rows:
    byid.setdefault(r['ID'],{'name':r['Kernel Name'][:14]})[r['Metric Name']]=r['Metric Value']
big=[v for v in byid.values() if float(v.get('gpu__time_duration.sum','0').replace(',',''))>200000]
for v in big[-9:]: print(v)
PY

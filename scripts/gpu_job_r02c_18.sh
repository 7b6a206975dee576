#!/bin/bash
# session 3, call 18: SYRK with 32-row stages and a conflict-free stride: per-launch times + dense parity tests
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'k_syrk_tall|k_gemm_tall' --csv --log-file gpurun_out/dense_launches_c18.csv python scripts/trace_step.py 1300000 > gpurun_out/trace_c18.out 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/root/repo/gpurun_out/dense_launches_c18.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
byid={}
for r in rows:
    byid.setdefault(r['ID'],{'name':r['Kernel Name'][:14]})[r['Metric Name']]=r['Metric Value']
big=[v for v in byid.values() if float(v.get('gpu__time_duration.sum','0').replace(',',''))>200000]
for v in big[-2:]: print(v)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bksvd or randsvd or irlba or pca or k30 or k50 or k100 or dense" > gpurun_out/pytest_gpu_c18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c18.log

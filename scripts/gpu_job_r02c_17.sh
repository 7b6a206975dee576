#!/bin/bash
# session 3, call 17: ncu --set full of the tall SYRK / GEMM (what are they bound by?)
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_syrk_tall|k_gemm_tall' --launch-skip 60 -c 12 -o gpurun_out/dense_r02c -f python scripts/trace_step.py 1300000 > gpurun_out/ncu_dense_r02c.log 2>&1
ncu -i gpurun_out/dense_r02c.ncu-rep --page raw --csv > gpurun_out/dense_r02c_raw.csv 2>/dev/null
python profiles/ncu_extract.py gpurun_out/dense_r02c_raw.csv > gpurun_out/dense_r02c_metrics.txt 2>&1
python - <<'PY'
import re
txt=open('/root/repo/gpurun_out/dense_r02c_metrics.txt').read()
for b in txt.split('-----')[1:]:
    m=re.search(r'gpu__time_duration.sum\s+(\S+)',b)
    if m and float(m.group(1).replace(',',''))>300: print(b[:3000])
PY

#!/bin/bash
# session 3, call 3: packed host form (sb_upload_packed): parity tests, untraced and traced upload timings, bench with both host forms
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or compact" > gpurun_out/pytest_packed.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_packed.log
timeout 600 python scripts/trace_upload.py packed > gpurun_out/upload_packed.log 2>&1; grep -E "pack_csc|total" gpurun_out/upload_packed.log
SCANB200_TRACE=2 timeout 600 python scripts/trace_upload.py packed > gpurun_out/trace_upload_packed.log 2>&1; tail -45 gpurun_out/trace_upload_packed.log | cut -c1-150
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c3_packed.json 2> gpurun_out/bench_c3_packed.err; tail -c 300 gpurun_out/bench_c3_packed.err
timeout 900 python bench.py --no-cpu-baseline --host-form compact > gpurun_out/bench_c3_compact.json 2> gpurun_out/bench_c3_compact.err; tail -c 300 gpurun_out/bench_c3_compact.err
python - <<'PY'
import json
for f in ('bench_c3_packed','bench_c3_compact'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), e['h2d_bytes_per_step'], e['calls_ms_host_clock[upload,normalize+pca,free]'], d['parity']['ok'], round(d['roofline']['frac'],3))
    except Exception as ex:
        print(f, 'ERR', ex)
PY

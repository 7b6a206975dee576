#!/bin/bash
# session 3, call 4: exact-size block cache A/B on the end-to-end arm (packed and compact host forms), full GPU suite
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c4.log
timeout 900 python bench.py --no-cpu-baseline --e2e-steps 4 > gpurun_out/bench_c4_packed.json 2> gpurun_out/bench_c4_packed.err; tail -c 300 gpurun_out/bench_c4_packed.err
SCANB200_BLOCK_CACHE_GB=0 timeout 900 python bench.py --no-cpu-baseline --e2e-steps 4 > gpurun_out/bench_c4_packed_nocache.json 2> gpurun_out/bench_c4_packed_nocache.err; tail -c 300 gpurun_out/bench_c4_packed_nocache.err
timeout 900 python bench.py --no-cpu-baseline --host-form compact --e2e-steps 4 > gpurun_out/bench_c4_compact.json 2> gpurun_out/bench_c4_compact.err; tail -c 300 gpurun_out/bench_c4_compact.err
python - <<'PY'
import json
for f in ('bench_c4_packed','bench_c4_packed_nocache','bench_c4_compact'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), e['h2d_bytes_per_step'], e['calls_ms_host_clock[upload,normalize+pca,free]'], d['parity']['ok'], round(d['roofline']['frac'],3))
    except Exception as ex:
        print(f, 'ERR', ex)
PY

#!/bin/bash
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_pool_private.json 2> gpurun_out/bench_pool_private.err
SCANB200_DEFAULT_POOL=1 timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_pool_default.json 2> gpurun_out/bench_pool_default.err
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_pool_private2.json 2> gpurun_out/bench_pool_private2.err
python - <<'PY'
import json
for f in ('bench_pool_private','bench_pool_default','bench_pool_private2'):
    d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
    e=d['e2e']
    print(f, round(d['ms_per_step'],1), 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), e['calls_ms_host_clock[upload,normalize+pca,free]'])
PY

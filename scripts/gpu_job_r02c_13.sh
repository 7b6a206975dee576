#!/bin/bash
# session 3, call 13 (8 GPUs): oracle parity worker at world 8, C3 and C4 bench lines with the session-3 code
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_worker.py > gpurun_out/mgpu_parity_c13_w8.log 2>&1
echo "parity rc=$?"; grep -c MGPU_PARITY_OK gpurun_out/mgpu_parity_c13_w8.log; grep MGPU_PARITY_OK gpurun_out/mgpu_parity_c13_w8.log | tail -3 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_r02c_n8_c3.json 2> gpurun_out/bench_r02c_n8_c3.err; tail -c 200 gpurun_out/bench_r02c_n8_c3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 8 --config c4 --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/bench_r02c_n8_c4.json 2> gpurun_out/bench_r02c_n8_c4.err; tail -c 200 gpurun_out/bench_r02c_n8_c4.err
python - <<'PY'
import json
for f in ('bench_r02c_n8_c3','bench_r02c_n8_c4'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d.get('e2e') or {}
        print(f, round(d['ms_per_step'],1), round(d['value']/1e6,1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', e.get('ms_per_step'), 'upload', e.get('upload_ms'), d['parity']['ok'])
    except Exception as ex:
        print(f, 'ERR', ex)
PY

#!/bin/bash
# 8 GPUs: parity worker at world 8 (new sharded paths included), C3 / C4 / C5 bench lines
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_worker.py > gpurun_out/mgpu_parity_w8.log 2>&1
grep -c MGPU_PARITY_OK gpurun_out/mgpu_parity_w8.log; tail -3 gpurun_out/mgpu_parity_w8.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_final_n8_c3.json 2> gpurun_out/bench_final_n8_c3.err; tail -c 200 gpurun_out/bench_final_n8_c3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 8 --config c4 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/bench_final_n8_c4.json 2> gpurun_out/bench_final_n8_c4.err; tail -c 200 gpurun_out/bench_final_n8_c4.err
python - <<'PY'
import json
for f in ('bench_final_n8_c3','bench_final_n8_c4'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), round(d['value']/1e6,1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), d['parity']['ok'])
    except Exception as e:
        print(f, 'ERR', e)
PY

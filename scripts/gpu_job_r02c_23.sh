#!/bin/bash
# session 3, call 23 (2 GPUs, final code): oracle parity worker at world 2 (packed + compact sharded uploads), C3 bench at N=2
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_worker.py > gpurun_out/mgpu_parity_final_w2.log 2>&1
echo "parity rc=$?"; grep -c MGPU_PARITY_OK gpurun_out/mgpu_parity_final_w2.log; tail -4 gpurun_out/mgpu_parity_final_w2.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; tail -c 300 gpurun_out/bench_final_n2.err
python - <<'PY'
import json
for f in ('bench_final_n2',):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), e['calls_ms_host_clock[upload,normalize+pca,free]'], d['parity']['ok'])
    except Exception as ex:
        print(f, 'ERR', ex)
PY

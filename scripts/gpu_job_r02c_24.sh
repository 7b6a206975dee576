#!/bin/bash
# session 3, call 24 (4 GPUs, final code): C3 bench line
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_final_n4.json 2> gpurun_out/bench_final_n4.err; tail -c 200 gpurun_out/bench_final_n4.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_final_n4.json').read().strip().splitlines()[-1])
e=d['e2e']
print(round(d['ms_per_step'],1), round(d['value']/1e6,1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), d['parity']['ok'])
PY

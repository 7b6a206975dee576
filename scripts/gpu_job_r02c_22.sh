#!/bin/bash
# session 3, call 22 (final, one GPU): full GPU suite, smoke(), the bench line of record with its CPU baseline leg, launch list of the same command, C2
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final_n1_c3.json 2> gpurun_out/bench_final_n1_c3.err; tail -c 300 gpurun_out/bench_final_n1_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final_bench_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_final.json 2> gpurun_out/bench_under_ncu_final.err
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_final_n1_c2.json 2> gpurun_out/bench_final_n1_c2.err
python - <<'PY'
import json
for f in ('bench_final_n1_c3','bench_final_n1_c2'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), d['parity']['ok'], round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY

#!/bin/bash
# final single-GPU evidence: bench line, launch list of the same command under ncu, the other configs
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_final_n1_c3.json 2> gpurun_out/bench_final_n1_c3.err; tail -c 300 gpurun_out/bench_final_n1_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02b_bench_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
for c in c2 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_final_n1_$c.json 2> gpurun_out/bench_final_n1_$c.err; done
timeout 600 python bench.py --config c4 --cells 500000 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench_final_n1_c4_500k.json 2> gpurun_out/bench_final_n1_c4_500k.err
python - <<'PY'
import json
for f in ('bench_final_n1_c3','bench_final_n1_c2','bench_final_n1_c5','bench_final_n1_c4_500k'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', (d.get('e2e') or {}).get('ms_per_step'), d['parity']['ok'])
    except Exception as e:
        print(f, 'ERR', e)
PY

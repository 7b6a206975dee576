#!/bin/bash
# session 3, call 10: ncu --set full of the SpMM kernels as they are now (deferred T-side gather, calibrated shares), launch list of the bench command, bench lines
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_planes_t|k_planes_n|k_gather|k_pl_digits_n|k_pl_reduce_t' --launch-skip 8 -c 12 \
  -o gpurun_out/spmm_r02c -f python scripts/prof_passes.py 400000 2 3 > gpurun_out/ncu_spmm_r02c.log 2>&1
ncu -i gpurun_out/spmm_r02c.ncu-rep --page raw --csv > gpurun_out/spmm_r02c_raw.csv 2>/dev/null
python profiles/ncu_extract.py gpurun_out/spmm_r02c_raw.csv > gpurun_out/spmm_r02c_metrics.txt 2>&1; grep -E "Kernel Name|gpu__time_duration" gpurun_out/spmm_r02c_metrics.txt | paste - - | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02c_bench_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_c.json 2> gpurun_out/bench_under_ncu_c.err
timeout 900 python bench.py > gpurun_out/bench_r02c_n1_c3.json 2> gpurun_out/bench_r02c_n1_c3.err; tail -c 300 gpurun_out/bench_r02c_n1_c3.err
for c in c2 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_r02c_n1_$c.json 2> gpurun_out/bench_r02c_n1_$c.err; done
timeout 600 python bench.py --config c4 --cells 500000 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench_r02c_n1_c4_500k.json 2> gpurun_out/bench_r02c_n1_c4_500k.err
python - <<'PY'
import json
for f in ('bench_r02c_n1_c3','bench_r02c_n1_c2','bench_r02c_n1_c5','bench_r02c_n1_c4_500k'):
    try:
        d=json.loads(open(f'/root/repo/gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'e2e', round(e['ms_per_step'],1), 'upload', round(e['upload_ms'],1), d['parity']['ok'], round(d['roofline']['frac'],3))
    except Exception as ex:
        print(f, 'ERR', ex)
PY

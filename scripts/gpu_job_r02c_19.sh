#!/bin/bash
# session 3, call 19: ncu --set full of the large k_syrk_tall launch (n x 100 block)
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_syrk_tall' --launch-skip 18 -c 1 -o gpurun_out/syrk_big -f python scripts/trace_step.py 1300000 > gpurun_out/ncu_syrk_big.log 2>&1
ncu -i gpurun_out/syrk_big.ncu-rep --page raw --csv > gpurun_out/syrk_big_raw.csv 2>/dev/null
python profiles/ncu_extract.py gpurun_out/syrk_big_raw.csv > gpurun_out/syrk_big_metrics.txt 2>&1; head -60 gpurun_out/syrk_big_metrics.txt | cut -c1-150
ncu -i gpurun_out/syrk_big.ncu-rep --page details 2>/dev/null | grep -E "Duration|Theoretical Occupancy|Achieved Occupancy|Registers|Shared Memory Config|Dynamic Shared|Block Limit|Waves|One or More|Stall|stall|Est. Speedup|DRAM Throughput|L2 Cache Throughput|Mem Busy|Max Bandwidth|Issue Slots Busy|Eligible|No Eligible" | head -40 | cut -c1-200

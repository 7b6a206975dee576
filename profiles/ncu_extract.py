"""Pull the metrics we care about out of an `ncu --page raw --csv` export.
usage: python profiles/ncu_extract.py raw.csv"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','sm__cycles_active.avg','sm__cycles_active.max','sm__cycles_active.min','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','launch__grid_size','launch__block_size','sm__inst_executed_pipe_lsu.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_red.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum']
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in idx: print(f"{w:75s} {r[idx[w]]:>20s} {units[idx[w]]}")
    stall=[h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
    vals=sorted([(float(r[idx[h]].replace(',','') or 0),h) for h in stall], reverse=True)[:7]
    for v,h in vals: print(f"   stall {v:8.2f} {h.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','')}")

"""Print the key ncu metrics of a raw-page CSV (ncu -i X.ncu-rep --page raw --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'launch__shared_mem_per_block_dynamic']
idx = {h: i for i, h in enumerate(hdr)}
for w in want:
    if w in idx:
        print(f"{w[:66]:66s} {units[idx[w]][:9]:9s}", [r[idx[w]][:34] for r in rows[2:]])
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        vals = [r[i] for r in rows[2:]]
        try:
            if max(float(v) for v in vals) < 4: continue
        except Exception: continue
        print(f"{h.replace('smsp__average_','').replace('smsp__warp_issue_stalled_','stall:')[:66]:66s}", [v[:6] for v in vals])

"""Per-kernel stall / pipe digest of an `ncu --page raw --csv` export:  python profiles/stalls.py <raw.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
ki = H.index("Kernel Name")
names = [r[ki].split('(')[0][:18] for r in rows[2:]]
print(names)
KEEP = ('gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__maximum_warps_per_active_cycle_pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active')
for i, h in enumerate(H):
    vals = [r[i] for r in rows[2:]]
    if 'issue_stalled' in h and 'not_issued' not in h and h.endswith('.ratio'):
        try:
            if max(float(v.replace(',', '')) for v in vals) < 0.7:
                continue
        except Exception:
            pass
        print('stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', 'L:').replace('_per_issue_active.ratio', '').replace('.ratio', ''), [v[:6] for v in vals])
    elif h in KEEP:
        print(h, [v[:10] for v in vals])

"""Print the metrics we read from an `ncu --page raw --csv` export (one row per profiled launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__cycles_active.avg']
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
stall = [c for c in h if 'warp_issue_stalled' in c and c.endswith('_per_warp_active.pct')]
for r in rows[2:]:
    print("==", r[h.index('Kernel Name')][:90] if 'Kernel Name' in h else '')
    for w in WANT:
        if w in h:
            i = h.index(w)
            print(f"  {w:75s} {r[i]:>18s} {units[i]}")
    vals = sorted(((float(r[h.index(c)].replace(',', '')), c) for c in stall if r[h.index(c)]), reverse=True)
    for v, c in vals[:8]:
        print(f"  stall {c.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v:8.2f} %")

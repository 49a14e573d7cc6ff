"""Print selected metrics of every kernel in an ncu report:  python tools/ncu_raw_pick.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__cluster_max_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__lsu_writeback_active', 'launch__occupancy_limit',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if any(h == w or h.startswith(w + '.') or h.startswith(w) and w.startswith('launch__occ') for w in want) or ('issue_stalled' in h and h.endswith('.ratio')):
        vals = [r[i][:14] for r in rows[2:]]
        if 'issue_stalled' in h and all(float(v or 0) < 0.15 for v in vals):
            continue
        print(h[:88], units[i], vals)

"""Concise view of an .ncu-rep: python tools/ncu_brief.py file.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    for k in want:
        if k in hdr:
            i = hdr.index(k); print(f"{k:78s} {r[i]:>44s} {units[i]}")
    for i, k in enumerate(hdr):
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
            try:
                if float(r[i]) > 0.05: print(f"  stall {k[34:-23]:40s} {float(r[i]):8.3f}")
            except ValueError:
                pass
    print()

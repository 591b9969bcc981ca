"""Print the handful of ncu metrics used to steer kernel work from a .ncu-rep (one block per captured launch)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:60])
    for k in KEYS:
        if k in hdr:
            print('  ', k, r[hdr.index(k)], units[hdr.index(k)])
    out = []
    for i, k in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in k and k.endswith('_not_issued'):
            try:
                out.append((float(r[i]), k))
            except ValueError:
                pass
    tot = sum(v for v, k in out) or 1
    print('   stalls (not-issued samples):', ', '.join(f'{k[33:-11]} {v / tot * 100:.0f}%' for v, k in sorted(out, reverse=True)[:9]))

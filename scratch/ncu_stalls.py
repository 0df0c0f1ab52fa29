#!/usr/bin/env python3
"""Key utilisation + stall metrics of .ncu-rep files."""
import csv, subprocess, sys
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    print('##', r[hdr.index('Kernel Name')][:40])
    def g(n):
        return r[hdr.index(n)] if n in hdr else 'n/a'
    for n in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
              'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
              'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
              'smsp__thread_inst_executed_per_inst_executed.ratio']:
        print(f'  {n:66s} {g(n):>14s}')
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    st.sort(reverse=True)
    for v, h in st[:7]:
        print(f'  stall {h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""):40s} {v:8.2f}')

#!/usr/bin/env python
"""Text digest of a small ncu report (one or two launches, --set full --import-source on): key raw metrics per launch
and the source lines ranked by stall samples.   python tools/ncu_report_md.py report.ncu-rep "title" > out.md"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
print('# %s\n' % title)
print('Digest of `%s` (`ncu --set full --import-source on --clock-control none`; the report itself is not in the '
      'repository, `*.ncu-rep` is git-ignored).\n' % rep.split('/')[-1])
for v in rows[2:]:
    print('| metric | value |\n|---|---|')
    for w in want:
        if w in h:
            i = h.index(w)
            print('| `%s` | %s %s |' % (w, v[i][:90].replace('|', '/'), u[i]))
    stalls = [(float(v[i] or 0), x) for i, x in enumerate(h) if 'warp_issue_stalled' in x and 'per_warp_active' in x]
    for val, x in sorted(stalls, reverse=True)[:5]:
        print('| `%s` | %.1f %% |' % (x.replace('smsp__warp_issue_stalled_', 'stall: ').replace('_per_warp_active.pct', ''), val))
    print()
lines = subprocess.run([sys.executable, 'tools/ncu_lines.py', rep, '30'], capture_output=True, text=True).stdout
print('## Source lines by stall samples (all launches of the report)\n\n```')
print('\n'.join(l[:170] for l in lines.splitlines()))
print('```')

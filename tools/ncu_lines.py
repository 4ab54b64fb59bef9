#!/usr/bin/env python
"""Per-source-line totals of an ncu --import-source report:  python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
h = rows[hdr]
ci, si = h.index('Instructions Executed'), h.index('# Samples')
data = []
for r in rows[hdr + 1:]:
    if len(r) > ci and r[0].strip().isdigit():
        try:
            data.append((int(r[0]), r[1], int(r[ci] or 0), int(r[si] or 0)))
        except ValueError:
            pass
ti, ts = sum(d[2] for d in data) or 1, sum(d[3] for d in data) or 1
print('warp instructions %d, samples %d' % (ti, ts))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print('%5d  instr %5.1f%%  samples %5.1f%% | %s' % (d[0], 100.0 * d[2] / ti, 100.0 * d[3] / ts, d[1][:120]))

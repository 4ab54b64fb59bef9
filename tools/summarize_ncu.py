#!/usr/bin/env python
"""Turn gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/prof_<tag>.ncu-rep (ncu --set full) into a committed summary under profiles/."""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else 'r01'
out = open('profiles/%s_%s_summary.md' % (rnd, tag), 'w')


def p(*a):
    print(*a, file=out)


lines = [l for l in open('gpurun_out/launches_%s.csv' % tag) if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'])
    v = {'ns': v / 1000, 'us': v, 'usecond': v, 'ms': v * 1000, 'msecond': v * 1000, 'nsecond': v / 1000}[row['Metric Unit']]
    seq.append((row['Kernel Name'].split('(')[0][:60], v, row['Grid Size'], row['Block Size']))
tot, cnt = collections.Counter(), collections.Counter()
for n, v, _, _ in seq:
    tot[n] += v
    cnt[n] += 1
total = sum(tot.values())
p('# ncu launch list, `bench.py --precision %s` (%s)\n' % (tag, rnd))
p('`ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200` over the timed loop; per-launch times are')
p('cold-cache and serialised, so compare SHARES, not absolutes.  %d launches captured.\n' % len(seq))
p('| kernel | launches | total us | share | avg us |')
p('|---|---:|---:|---:|---:|')
for n, v in tot.most_common():
    p('| `%s` | %d | %.1f | %.1f%% | %.2f |' % (n, cnt[n], v, 100 * v / total, v / cnt[n]))
p('\n## one step in launch order (first 48 launches)\n')
p('| # | kernel | grid | block | us |')
p('|---:|---|---|---|---:|')
for i, (n, v, g, b) in enumerate(seq[:48]):
    p('| %d | `%s` | %s | %s | %.2f |' % (i, n, g, b, v))

try:
    raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s.ncu-rep' % tag, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, timeout=300).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum',
            'smsp__cycles_active.avg', 'l1tex__t_bytes.sum']
    idx = {h: hdr.index(h) for h in want if h in hdr}
    p('\n## `ncu --set full` capture of the dominant kernel (%d launches)\n' % (len(rows) - 2))
    p('units row: ' + ', '.join('%s=%s' % (h, rows[1][i]) for h, i in idx.items() if rows[1][i]))
    p('')
    p('| ' + ' | '.join(idx) + ' |')
    p('|' + '---|' * len(idx))
    for r in rows[2:]:
        p('| ' + ' | '.join(r[i][:48] for i in idx.values()) + ' |')
except Exception as e:  # noqa
    p('\n(full capture not summarised: %s)' % e)
out.close()
print(open('profiles/%s_%s_summary.md' % (rnd, tag)).read()[:6000])

#!/usr/bin/env python
"""profiles/r02_final_summary.md + profiles/r02_traffic.json from the evidence run of tools/gpu_final2.sh:

    gpurun_out/launches_r02.csv        ncu --metrics gpu__time_duration.sum over bench.py (one batch in flight)
    gpurun_out/r02_ncu_full_raw.csv    `ncu -i prof --page raw --csv` of an `ncu --set full` capture of one step's
                                       tcgen05 launches (the 68 MB report itself stays on the GPU box)
"""
import collections
import csv
import json
import re
import sys

sys.path.insert(0, '.')
out = open('profiles/r02_final_summary.md', 'w')


def p(*a):
    print(*a, file=out)


def short(n):
    return re.sub(r'<.*', '', re.sub(r'\(.*', '', n).replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', ''))


lines = [l for l in open('gpurun_out/launches_r02.csv') if l.startswith('"')]
rows = list(csv.DictReader(lines))
seq = []
for r in rows:
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'])
    v = {'ns': v / 1000, 'us': v, 'usecond': v, 'ms': v * 1000, 'msecond': v * 1000, 'nsecond': v / 1000}[r['Metric Unit']]
    seq.append((short(r['Kernel Name']), v, r['Grid Size'], r['Block Size']))
starts = [i for i, s in enumerate(seq) if s[0].startswith('k_stem')]
step = seq[starts[1]:starts[2]]
p('# r02 - ncu evidence for the default bench line (BASELINE configs[2]: SSD-MobileNet-v2, 90 classes, 8 cameras, tf32x3)\n')
p('## Launch list of one step (`ncu --metrics gpu__time_duration.sum --clock-control none`, one batch in flight)\n')
p('Per-launch times under ncu are cold-cache and serialised: compare SHARES.  SM time = duration x min(1, CTAs / (148 x CTAs')
p('per SM)), the quantity that bounds the step with six batches in flight (sum ~ measured ms per step).\n')
fam = collections.OrderedDict()
tot = smt = 0.0
for n, t, g, b in step:
    ctas = 1
    for x in re.findall(r'\d+', g):
        ctas *= int(x)
    thr = int(re.findall(r'\d+', b)[0])
    per_sm = 1 if ('gemm_tc' in n or 'dwpw' in n or 'irb' in n) else max(1, min(2048 // thr, 8))
    occ = min(1.0, ctas / (148.0 * per_sm))
    f = fam.setdefault(n, [0, 0.0, 0.0])
    f[0] += 1
    f[1] += t
    f[2] += t * occ
    tot += t
    smt += t * occ
p('| kernel | launches | serial us | share | SM-time us | share |')
p('|---|---:|---:|---:|---:|---:|')
for n, (c, t, sm) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    p('| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f%% |' % (n, c, t, 100 * t / tot, sm, 100 * sm / smt))
p('| **one step** | **%d** | **%.1f** | | **%.1f** | |' % (len(step), tot, smt))
p('\n### in launch order\n')
p('| # | kernel | grid | block | us |')
p('|---:|---|---|---|---:|')
for i, (n, t, g, b) in enumerate(step):
    p('| %d | `%s` | %s | %s | %.2f |' % (i, n, g, b, t))

traffic = {}
try:
    raw = list(csv.reader(open('gpurun_out/r02_ncu_full_raw.csv')))
    hdr = raw[0]
    want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
    idx = collections.OrderedDict((h, hdr.index(h)) for h in want if h in hdr)
    p('\n## `ncu --set full` of %d consecutive tcgen05 launches of one step\n' % (len(raw) - 2))
    p('units: ' + ', '.join('%s=%s' % (h, raw[1][i]) for h, i in idx.items() if raw[1][i]))
    p('')
    p('| ' + ' | '.join(h.replace('.avg.pct_of_peak_sustained_', ' %') for h in idx) + ' |')
    p('|' + '---|' * len(idx))
    dram = []
    for r in raw[2:]:
        cells = [short(r[i]) if h == 'Kernel Name' else r[i][:40] for h, i in idx.items()]
        p('| ' + ' | '.join(cells) + ' |')
        try:
            def num(h):
                return float(r[idx[h]].replace(',', ''))
            ur, uw = raw[1][idx['dram__bytes_read.sum']], raw[1][idx['dram__bytes_write.sum']]
            mul = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            dram.append(num('dram__bytes_read.sum') * mul.get(ur, 1.0) + num('dram__bytes_write.sum') * mul.get(uw, 1.0))
        except Exception:
            pass
    if dram:
        from tests.workload import v2_coco_model
        name = v2_coco_model().name
        traffic = {name: {'tf32x3': {'gemm': {
            'launches': len(dram), 'dram_bytes_per_launch': sum(dram) / len(dram),
            'source': 'profiles/r02_final_summary.md (ncu --set full, cold caches per replay, %d consecutive tcgen05 '
                      'launches of one step of the default bench workload)' % len(dram)}}}}
        json.dump(traffic, open('profiles/r02_traffic.json', 'w'), indent=1)
        p('\nDRAM read + write per launch, mean over the capture: %.2f MB (`profiles/r02_traffic.json`, copied into '
          '`roofline.traffic` by bench.py).' % (sum(dram) / len(dram) / 1e6))
except FileNotFoundError:
    p('\n(no --set full capture found)')
out.close()
print(open('profiles/r02_final_summary.md').read()[:3000])

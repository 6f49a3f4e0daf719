"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/xxx_launches.md"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        rows.append((re.sub(r'\(.*', '', r['Kernel Name']), r['Grid Size'], r['Block Size'], us))
rows = rows[skip:]
agg = OrderedDict()
for name, grid, block, us in rows:
    a = agg.setdefault(name, [0, 0.0, grid, block])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f'launches: {len(rows)} (first {skip} skipped), total device time {tot / 1e3:.2f} ms (cold-cache, serialised under ncu: compare shares)\n')
print('| kernel | launches | total us | share | avg us | grid (first) | block |')
print('|---|---|---|---|---|---|---|')
for name, (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| {name} | {n} | {us:.1f} | {100 * us / tot:.1f}% | {us / n:.1f} | {grid} | {block} |')

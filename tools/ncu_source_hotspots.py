"""Aggregate the warp-stall samples of an ncu report per CUDA source line.
    ncu -i report.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_source_hotspots.py src.csv [top_n] > profiles/<name>.md
(the kernel must have been built with -lineinfo and captured with --import-source on)."""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    agg, cur = collections.Counter(), None
    for r in csv.reader(open(path)):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1]
        elif len(r) > 6 and r[0].strip().isdigit() and r[2] == '-':      # a source line row (its SASS rows follow)
            try:
                agg[(cur, int(r[0]))] += int(r[4] or 0)
            except ValueError:
                pass
    tot = sum(agg.values())
    cache = {}

    def text(f, ln):
        if f not in cache:
            p = f if os.path.exists(f) else os.path.join(ROOT, 'xframe_b200', 'csrc', os.path.basename(f))
            cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        return cache[f][ln - 1].strip()[:110] if 0 < ln <= len(cache[f]) else ''
    print(f'warp-stall samples: {tot}\n\n| share | file:line | source |\n|---|---|---|')
    for (f, ln), s in agg.most_common(top_n):
        print(f'| {100 * s / tot:.1f} % | {os.path.basename(f)}:{ln} | `{text(f, ln)}` |')


if __name__ == '__main__':
    main()

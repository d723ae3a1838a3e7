"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv [--top 30] > profiles/<name>.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 30
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((r['Kernel Name'], float(r['Metric Value'].replace(',', ''))))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r'\(.*', '', name).replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
        short = re.sub(r'^void ', '', short)
        short = re.sub(r'<.*', '', short)[:90]
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f'# {path}: {len(rows)} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)\n')
    print('| kernel | launches | total ms | share | mean us |')
    print('|---|---:|---:|---:|---:|')
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |')
    ours = sum(v[1] for k, v in agg.items() if 'abx::' in k)
    print(f'\nabx:: kernels: {100 * ours / total:.1f}% of kernel time')


if __name__ == '__main__':
    main()

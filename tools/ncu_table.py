"""Markdown table of an ncu --set full report: one row per captured kernel launch.
    python tools/ncu_table.py gpurun_out/x.ncu-rep [--json out.json --B 8 --N 350]"""
import csv
import io
import json
import subprocess
import sys

COLS = [('time us', 'gpu__time_duration.sum', 1.0), ('DRAM read MB', 'dram__bytes_read.sum', 1.0), ('DRAM write MB', 'dram__bytes_write.sum', 1.0),
        ('DRAM active %', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 1.0), ('SM %', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
        ('issue active %', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1.0),
        ('warps active %', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1.0), ('regs', 'launch__registers_per_thread', 1.0),
        ('tensor pipe %', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 1.0),
        ('L1 %', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 1.0), ('L2 %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 1.0)]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    print('| kernel | ' + ' | '.join(c[0] for c in COLS) + ' |')
    print('|---|' + '---:|' * len(COLS))
    tot = {'time': 0.0, 'rd': 0.0, 'wr': 0.0}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get('Kernel Name', '').split('(')[0].replace('abx::', '').replace('void ', '')[:60]
        vals = []
        for label, key, _ in COLS:
            v = d.get(key, '')
            try:
                f = float(v.replace(',', ''))
            except ValueError:
                vals.append(v)
                continue
            u = unit.get(key, '')
            if label.startswith('DRAM') and 'MB' in label:
                f = f * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
                tot['rd' if 'read' in label else 'wr'] += f
            if label == 'time us':
                f = f * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1.0)
                tot['time'] += f
            vals.append(f'{f:.1f}' if label != 'regs' else f'{int(f)}')
        print(f'| `{name}` | ' + ' | '.join(vals) + ' |')
    print(f'| **sum** | {tot["time"]:.1f} | {tot["rd"]:.1f} | {tot["wr"]:.1f} |' + ' |' * (len(COLS) - 3))
    if '--json' in sys.argv:
        path = sys.argv[sys.argv.index('--json') + 1]
        B = int(sys.argv[sys.argv.index('--B') + 1]) if '--B' in sys.argv else None
        N = int(sys.argv[sys.argv.index('--N') + 1]) if '--N' in sys.argv else None
        json.dump({'B': B, 'N': N, 'dram_bytes_read': tot['rd'] * 1e6, 'dram_bytes_write': tot['wr'] * 1e6,
                   'traffic_bytes': (tot['rd'] + tot['wr']) * 1e6, 'time_us_cold_serialised': tot['time'], 'source': rep}, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()

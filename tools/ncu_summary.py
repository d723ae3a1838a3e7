"""Key metrics of an .ncu-rep (raw page) and, with --source, the hottest source lines by stall samples.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source] [--top 40]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, 'raw')
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:90])
        for k in KEYS:
            if k in d:
                print(f'  {k:90s} {d[k]}')
        for k, v in d.items():
            if 'warp_issue_stalled' in k and k.endswith('_per_warp_active.pct') and float(v or 0) > 2.0:
                print(f'  {k:90s} {v}')
    if '--source' in sys.argv:
        top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
        rows = page(rep, 'source')
        hdr = rows[0]
        ci = {h: i for i, h in enumerate(hdr)}
        samp = next((h for h in hdr if h.startswith('# Samples') or h == 'Warp Stall Sampling (All Samples)'), None) or \
            next(h for h in hdr if 'Sampl' in h)
        print('sampling column:', samp, '| columns:', [h for h in hdr][:12])
        body = [r for r in rows[1:] if len(r) == len(hdr)]
        def val(r):
            try:
                return float(r[ci[samp]])
            except ValueError:
                return 0.0
        tot = sum(val(r) for r in body) or 1.0
        for r in sorted(body, key=val, reverse=True)[:top]:
            print(f'{100 * val(r) / tot:5.1f}%  {r[ci.get("Source", 1)][:150]}')


if __name__ == '__main__':
    main()

"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU):
   python scripts/ncu_summary.py launches <launches.csv>      -> per-kernel totals / shares
   python scripts/ncu_summary.py full <prof.ncu-rep>          -> key metrics per captured launch"""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
        k = row['Kernel Name'].split('(')[0]
        agg[k][0] += 1; agg[k][1] += v; tot += v
    print(f'| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|')
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'| {k[:70]} | {n} | {t / 1e3:.2f} | {t / n:.1f} | {100 * t / tot:.1f}% |')
    print(f'\ntotal device time {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"\n### {r[idx['Kernel Name']].split('(')[0]}  (grid {r[idx['Grid Size']]}, block {r[idx['Block Size']]})")
        for w in WANT:
            if w in idx:
                print(f'- {w} = {r[idx[w]]} {units[idx[w]]}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])

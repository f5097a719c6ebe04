"""Summarise the warp-stall samples of an .ncu-rep per SASS instruction (CPU only; needs `ncu`).

    python tools/ncu_stalls.py gpurun_out/k1fused_full_r1.ncu-rep [--top 12] [--csv out.csv]

Reads `ncu -i REP --page source --csv`, prints the stall-reason totals, the hottest instructions with
the few instructions that precede them (the load a stalled instruction waits on is normally one of
those), and the share of samples per 40-instruction window.
"""
import argparse
import csv
import io
import subprocess
import sys


def load(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = next(n for n, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
    return rows[head], rows[head + 1:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--top', type=int, default=12)
    ap.add_argument('--context', type=int, default=6)
    ap.add_argument('--csv', default=None, help='write the instructions holding >= 0.5 %% of the samples')
    a = ap.parse_args()
    h, data = load(a.rep)
    ix = {k: n for n, k in enumerate(h)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    tot = sum(f(r, '# Samples') for r in data) or 1.0
    stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    print('samples %d over %d instructions' % (tot, len(data)))
    for k, v in sorted(((k, sum(f(r, k) for r in data)) for k in stalls), key=lambda x: -x[1])[:8]:
        print('  %-24s %5.1f %%' % (k, 100 * v / tot))
    order = sorted(range(len(data)), key=lambda n: -f(data[n], '# Samples'))
    for n in order[:a.top]:
        share = 100 * f(data[n], '# Samples') / tot
        if share < 0.5:
            break
        why = max(stalls, key=lambda k: f(data[n], k))
        print('--- #%d  %.1f %%  %s' % (n, share, why))
        for m in range(max(0, n - a.context), n + 1):
            print('   %5.1f %%  %s' % (100 * f(data[m], '# Samples') / tot, data[m][ix['Source']][:100]))
    print('--- share of samples per 40-instruction window')
    for lo in range(0, len(data), 40):
        s = sum(f(r, '# Samples') for r in data[lo:lo + 40])
        if s / tot >= 0.005:
            tags = sorted({t for r in data[lo:lo + 40] for t in ('LDG', 'REDG', 'MUFU', 'SHFL', 'STG', 'ATOM', 'BAR',
                                                                   'LDS', 'STS', 'UTCHMMA', 'LDTM', 'SYNCS')
                           if t in r[ix['Source']]})
            print('   %4d-%4d  %5.1f %%  %s' % (lo, lo + 40, 100 * s / tot, ' '.join(tags)))
    if a.csv:
        with open(a.csv, 'w', newline='') as fh:
            w = csv.writer(fh)
            w.writerow(['index', 'share_pct', 'top_stall', 'sass'])
            for n in sorted(order[:200]):
                share = 100 * f(data[n], '# Samples') / tot
                if share >= 0.5:
                    w.writerow([n, '%.2f' % share, max(stalls, key=lambda k: f(data[n], k)), data[n][ix['Source']]])
    return 0


if __name__ == '__main__':
    sys.exit(main())

"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel:
    python tools/launch_summary.py profiles/r01_sparse_launches_v4.csv [name-filter]
Prints launches, total and average duration per kernel name (template arguments kept, call arguments dropped)."""
import collections
import csv
import re
import sys


def main(path, flt=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    H = rows[hi]
    kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r'\(.*', '', r[kn]).replace('void ', '')
        if flt and flt not in name:
            continue
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print('%-72s %8s %12s %10s %7s' % ('kernel', 'launches', 'total us', 'avg us', 'share'))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-72s %8d %12.1f %10.1f %6.1f%%' % (k[:72], c, t, t / c, 100.0 * t / tot))
    print('%-72s %8d %12.1f' % ('total', sum(c for c, _ in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
    python tools/launch_summary.py gpurun_out/launches.csv
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hdr]
    idx = {k: j for j, k in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        name = r[idx['Kernel Name']][:110]
        val = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        val *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.3f ms of kernel time" % (path, sum(a[0] for a in agg.values()), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6d launches %10.3f ms total %8.3f ms avg %5.1f%%  %s" % (a[0], a[1], a[1] / a[0], 100 * a[1] / tot, k))


if __name__ == '__main__':
    main(sys.argv[1])

#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.   python tools/launch_summary.py file.csv "header" """
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki][:86], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    for h in sys.argv[2:]:
        print("# " + h)
    print("# per-launch times under ncu are serialised and cold-cache: use the SHARES")
    print("launches %d   total kernel time %.2f ms" % (sum(a[0] for a in agg.values()), tot))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot >= 0.001:
            print("%-88s n=%5d %10.3f ms %6.1f%%  avg %.4f ms" % (n, c, t, 100 * t / tot, t / c))


if __name__ == "__main__":
    main()

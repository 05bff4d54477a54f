#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page source --csv` (SASS view) into instruction windows: share of executed warp instructions,
share of stall samples, average active lanes, and the opcode mix.   python tools/ncu_source_windows.py file.source.csv [window]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    rows = list(csv.reader(open(path)))
    start = 0
    while start < len(rows):
        if rows[start] and rows[start][0] == "Kernel Name":
            name = rows[start][1]
            hdr = rows[start + 1]
            end = start + 2
            while end < len(rows) and not (rows[end] and rows[end][0] == "Kernel Name"):
                end += 1
            body = [dict(zip(hdr, r)) for r in rows[start + 2:end] if len(r) >= len(hdr) - 1]
            report(name, body, win)
            start = end
        else:
            start += 1


def report(name, body, win):
    def num(d, k):
        try:
            return float(d.get(k, "0") or 0)
        except ValueError:
            return 0.0
    inst = [num(d, "Instructions Executed") for d in body]
    thr = [num(d, "Thread Instructions Executed") for d in body]
    smp = [num(d, "# Samples") for d in body]
    ti, tt, ts = sum(inst), sum(thr), sum(smp)
    print("kernel %s" % name[:100])
    print("warp instructions %d   thread instructions %d   average active lanes %.2f of 32   stall samples %d" % (ti, tt, tt / max(ti, 1), ts))
    ops = collections.defaultdict(lambda: [0.0, 0.0])
    for d, i, s in zip(body, inst, smp):
        op = d["Source"].split()
        op = [x for x in op if not x.startswith("@")]
        o = op[0].split(".")[0] if op else "?"
        ops[o][0] += i; ops[o][1] += s
    print("by opcode: share of executed warp instructions / share of stall samples")
    for o, (i, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
        print("  %-10s %5.1f%%  %5.1f%%" % (o, 100 * i / max(ti, 1), 100 * s / max(ts, 1)))
    print("SASS windows (%d instructions each) above 2%%: index range, share of warp instructions, share of samples, average lanes, first instruction" % win)
    for a in range(0, len(body), win):
        i = sum(inst[a:a + win]); t = sum(thr[a:a + win]); s = sum(smp[a:a + win])
        if i / max(ti, 1) >= 0.02 or s / max(ts, 1) >= 0.03:
            print("  [%4d,%4d)  %5.1f%%  %5.1f%%  lanes %4.1f   %s" % (a, a + win, 100 * i / max(ti, 1), 100 * s / max(ts, 1), t / max(i, 1), body[a]["Source"].strip()[:60]))
    print()


if __name__ == "__main__":
    main()

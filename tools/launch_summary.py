"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches,
total us and share of the profiled step.  usage: python tools/launch_summary.py file.csv [--all]"""
import collections
import csv
import re
import sys


def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    out = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        out.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("o3dml::", ""), v, r[gi], r[bi]))
    return out


def main():
    rows = load(sys.argv[1])
    tot = sum(r[1] for r in rows)
    if "--all" in sys.argv:
        for n, v, g, b in rows:
            print("%9.1f us  grid %-16s block %-14s %s" % (v, g, b, n))
    agg = collections.OrderedDict()
    for n, v, _, _ in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (n, c, v, 100 * v / tot))
    print("| **sum** | %d | %.1f | |" % (len(rows), tot))


if __name__ == "__main__":
    main()

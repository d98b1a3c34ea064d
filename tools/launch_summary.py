#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def summarise(path, skip=0):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1 + skip:]:
        if len(r) > vi:
            agg.setdefault(r[ki][:80], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| total ms | launches | avg us | share | kernel |", "|---:|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {sum(v) / 1e6:.3f} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% | `{k}` |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0))

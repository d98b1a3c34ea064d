#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
per kernel: launches, average duration, average DRAM bytes read / written."""
import collections
import csv
import sys


def summarise(path, min_ms=0.05):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki][:64], {"t": [], "rd": [], "wr": []})
        v = float(r[vi].replace(",", ""))
        if "time_duration" in r[mi]:
            a["t"].append(v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui][:2], 1e-6))
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
            a["rd" if "read" in r[mi] else "wr"].append(v / 1e9)
    out = ["| kernel | launches | avg ms | DRAM read GB | DRAM write GB |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]["t"]) / max(len(kv[1]["t"]), 1)):
        n = len(a["t"])
        if n == 0 or sum(a["t"]) / n < min_ms:
            continue
        rd = sum(a["rd"]) / max(len(a["rd"]), 1)
        wr = sum(a["wr"]) / max(len(a["wr"]), 1)
        out.append(f"| `{k}` | {n} | {sum(a['t']) / n:.3f} | {rd:.2f} | {wr:.2f} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.05))

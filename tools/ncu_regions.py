#!/usr/bin/env python
"""Group the SASS of one kernel in an `ncu --page source --csv` dump into regions of equal execution count."""
import csv
import sys


def main(path, pattern, thr=0.01):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for si, st in enumerate(starts):
        name = rows[st][1]
        if pattern not in name:
            continue
        end = starts[si + 1] if si + 1 < len(starts) else len(rows)
        h = rows[st + 1]
        isrc, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        data = [(int(r[ie]), int(r[iss]), r[isrc].strip()) for r in rows[st + 2:end] if len(r) > ie and r[ie].isdigit()]
        tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
        print(name, "instr", tot, "samples", tots)
        start = 0
        for i in range(1, len(data) + 1):
            if i == len(data) or abs(data[i][0] - data[start][0]) > 0.05 * max(data[start][0], 1):
                e, s = sum(d[0] for d in data[start:i]), sum(d[1] for d in data[start:i])
                if e > thr * tot or s > thr * tots:
                    ops = {}
                    for d in data[start:i]:
                        t = d[2].split()
                        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                        ops[op] = ops.get(op, 0) + 1
                    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
                    print(f"instr {start:4d}-{i - 1:4d} n={i - start:3d} exec/instr={data[start][0]:11d} instr%={100 * e / tot:5.1f} samples%={100 * s / tots:5.1f}  {top}")
                start = i


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.01)

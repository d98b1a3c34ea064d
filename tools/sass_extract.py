#!/usr/bin/env python
"""SASS of the step kernels out of libflexigal_gpu.so (cuobjdump -sass), one file per kernel under profiles/, plus an opcode
histogram: the evidence that the hot path is FP64 DFMA / DMMA.8x8x4 code with cp.async (LDGSTS) staging and no tensor-core GEMM."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flexigal_b200", "lib", "libflexigal_gpu.so")
WANT = {"k_nb_cluster<3,0,false>": r"k_nb_clusterILi3ELi0ELb0E", "k_imls_c<3,0>": r"k_imls_cILi3ELi0E", "k_imls_p<3,0>": r"k_imls_pILi3ELi0E", "k_pattern_clusters<3>": r"k_pattern_clustersILi3E",
        "k_cluster_build_masks": r"k_cluster_build_masks", "k_linear_pts<3>": r"k_linear_ptsILi3E", "k_bilinear_pts<3,GRADGRAD>": r"k_bilinear_ptsILi3ELi1E",
        "k_elast_diag<3>": r"k_elast_diagILi3E", "k_elast_off<3>": r"k_elast_offILi3E", "k_generic_pts<2>": r"k_generic_ptsILi2E"}


def main(tag):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    chunks = re.split(r"\n\s*Function : ", out)
    rows = []
    for name, pat in WANT.items():
        hit = [c for c in chunks if re.match(r"_Z\d+" + pat, c)]
        if not hit:
            continue
        body = hit[0]
        fn = os.path.join(ROOT, "profiles", f"{tag}_sass_{re.sub(r'[^a-z0-9_]', '', name.split('<')[0])}.txt")
        lines = [ln for ln in body.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        with open(fn, "w") as f:
            f.write(f"// {name}  (cuobjdump -sass flexigal_b200/lib/libflexigal_gpu.so; sm_100a)\n")
            f.write("\n".join(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip() for ln in lines) + "\n")
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if m:
                ops[m.group(1)] += 1
        rows.append((name, len(lines), ops))
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.md"), "w") as f:
        f.write(f"# SASS opcode histogram of the step kernels ({tag}, static counts, `tools/sass_extract.py`)\n\n")
        f.write("| kernel | instructions | DFMA | DMUL | DADD | DMMA | LDG | LDS | STS | STG | RED/ATOM | LDGSTS (cp.async) | SHFL | HMMA / UTC*MMA |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for name, n, ops in rows:
            f.write(f"| `{name}` | {n} | {ops['DFMA']} | {ops['DMUL']} | {ops['DADD']} | {ops['DMMA']} | {ops['LDG']} | {ops['LDS']} | {ops['STS']} | {ops['STG']} | "
                    f"{ops['RED'] + ops['REDG'] + ops['ATOMG'] + ops['ATOM']} | {ops['LDGSTS']} | {ops['SHFL']} | {ops['HMMA'] + sum(v for k, v in ops.items() if k.startswith('UTC'))} |\n")
        f.write("\nFP64 arithmetic only (DFMA/DMUL/DADD and the FP64 tensor instruction DMMA.8x8x4); no HMMA/UTC*MMA: nothing on this path is a dense "
                "low-precision contraction (north star).  LDGSTS = the cp.async staging of the flush tables.\n")
    print("wrote", [r[0] for r in rows])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r03")

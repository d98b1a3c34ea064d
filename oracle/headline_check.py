"""Checker for the headline workload (Divisions = (N, N, N), 3D Poisson, trilinear IMLS): TEST INFRASTRUCTURE, like
everything under oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's untimed parity leg import it.

The oracle cannot run the 10^6-node problem whole (its domain() is the reference's brute force, O(nnodes) per Gauss point),
but it can run a stratified SAMPLE of it against the full node cloud: blocks of B x B x B background cells at the corner
(0,0,0), on an edge, on a face, in the centre and at the far corner.  For every sampled Gauss point the reference algorithm
(brute-force domain, src/Shape_Functions.jl:3-27; IMLS :88-162 in FP64 and in __float128) gives DOM, phi, grad phi; and
for every node whose support only holds sampled Gauss points ("complete" nodes, the interior nodes of a block) the
reference's column of K (src/Assembling_Operators.jl:2-26) follows from the sample alone.
"""
from __future__ import annotations

import numpy as np

import efg_oracle as oracle


def block_origins(n, B, which=("corner", "edge", "face", "centre", "far_corner", "far_edge")):
    mid = (n - B) // 2
    table = {"corner": (0, 0, 0), "edge": (mid, 0, 0), "face": (mid, mid, 0), "centre": (mid, mid, mid),
             "far_corner": (n - B, n - B, n - B), "far_edge": (n - B, mid, n - B)}
    return [(name, table[name]) for name in which]


def block_cells(n, origin, B):
    """0-based cell ids (x fastest, src/Meshing.jl:44) of the B^3 block at `origin`, ascending."""
    ox, oy, oz = origin
    k, j, i = np.meshgrid(np.arange(oz, oz + B), np.arange(oy, oy + B), np.arange(ox, ox + B), indexing="ij")
    return np.sort((i + n * j + n * n * k).reshape(-1))


def gauss_runs(cells, per_cell=27):
    """Contiguous runs [g_lo, g_hi) of Gauss-point ids covered by the (ascending) cell ids."""
    runs, start, prev = [], cells[0], cells[0]
    for c in cells[1:]:
        if c != prev + 1:
            runs.append((int(start) * per_cell, int(prev + 1) * per_cell))
            start = c
        prev = c
    runs.append((int(start) * per_cell, int(prev + 1) * per_cell))
    return runs


def complete_nodes(n, origin, B, reach=2):
    """0-based ids of the nodes whose whole support lies inside the block: the Gauss points within reach of node index m
    sit in cells m-reach .. m+reach-1 (clipped to the grid), all of which must belong to the block."""
    per_dim = []
    for o in origin:
        ok = [m for m in range(n + 1) if max(m - reach, 0) >= o and min(m + reach - 1, n - 1) < o + B]
        per_dim.append(np.array(ok, dtype=np.int64))
    kk, jj, ii = np.meshgrid(per_dim[2], per_dim[1], per_dim[0], indexing="ij")
    return np.sort((ii + (n + 1) * jj + (n + 1) ** 2 * kk).reshape(-1))


def reference_block(x, dm, gs_block, nthreads=8, want_truth=True):
    """The reference algorithm on the block's Gauss points against the FULL cloud: brute-force DOM, FP64 phi/dphi and the
    __float128 'truth' on the same lists."""
    off, dom, phi64, dphi64 = oracle.shape_fun(gs_block, x, dm, nthreads=nthreads)
    if want_truth:
        phiq, dphiq = oracle.shape_fun_given_dom(gs_block, x, dm, off, dom, nthreads=nthreads, quad=True)
    else:
        phiq, dphiq = phi64, dphi64
    return off, dom, phi64, dphi64, phiq, dphiq


def reference_columns(nnodes, gs_block, off, dom, phi, dphi, nodes, quad=True):
    """Columns `nodes` (0-based) of K = int grad(dT).grad(T) from the block's Gauss points alone: {node: (rows 1-based, values)}."""
    cp, rv = oracle.pattern(nnodes, off, dom)
    K = oracle.assemble_bilinear(gs_block, off, dom, phi, dphi, oracle.coef_gradgrad(3), 1, cp, rv, quad=quad)
    return {int(j): (rv[cp[j] - 1:cp[j + 1] - 1].copy(), K[cp[j] - 1:cp[j + 1] - 1].copy()) for j in nodes}


def rel(a, b):
    den = float(np.abs(b).max()) if b.size else 1.0
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0)) if a.size else 0.0


def check_against_device(ctx, sid, n, x, dm, gs, blocks, B=5, nthreads=8, with_K=True, tol=1e-12):
    """Compare the device results of set `sid` (the whole N^3 workload, already searched, IMLS done, K = gradgrad assembled
    when with_K) with the reference on the sampled blocks.  Returns a report dict; raises AssertionError on a mismatch."""
    nnodes = x.shape[0]
    rep = {"blocks": [], "points": 0, "dom_bit_exact": True, "phi_err_vs_truth": 0.0, "dphi_err_vs_truth": 0.0,
           "oracle64_phi_err_vs_truth": 0.0, "oracle64_dphi_err_vs_truth": 0.0, "K_columns": 0, "K_err_vs_truth": 0.0,
           "K_pattern_bit_exact": True}
    for name, origin in blocks:
        cells = block_cells(n, origin, B)
        runs = gauss_runs(cells)
        gidx = np.concatenate([np.arange(a, b) for a, b in runs])
        gsb = np.asfortranarray(gs[gidx, :])
        off, dom, phi64, dphi64, phiq, dphiq = reference_block(x, dm, gsb, nthreads)
        # device results for the same Gauss ranges
        d_off, d_dom, d_phi, d_dphi = [np.zeros(1, dtype=np.int64)], [], [], []
        for a, b in runs:
            o, d, p, dp = ctx.get_shapes_range(sid, a, b)
            d_off.append(o[1:] + d_off[-1][-1]); d_dom.append(d); d_phi.append(p); d_dphi.append(dp)
        d_off = np.concatenate(d_off); d_dom = np.concatenate(d_dom); d_phi = np.concatenate(d_phi); d_dphi = np.vstack(d_dphi)
        ok = np.array_equal(d_off, off) and np.array_equal(d_dom, dom)
        rep["dom_bit_exact"] &= bool(ok)
        assert ok, f"block {name}: neighbour lists / offsets differ from the brute-force domain()"
        e_phi, e_dphi = rel(d_phi, phiq), rel(d_dphi, dphiq)
        rep["phi_err_vs_truth"] = max(rep["phi_err_vs_truth"], e_phi); rep["dphi_err_vs_truth"] = max(rep["dphi_err_vs_truth"], e_dphi)
        rep["oracle64_phi_err_vs_truth"] = max(rep["oracle64_phi_err_vs_truth"], rel(phi64, phiq))
        rep["oracle64_dphi_err_vs_truth"] = max(rep["oracle64_dphi_err_vs_truth"], rel(dphi64, dphiq))
        assert e_phi <= tol and e_dphi <= tol, (name, e_phi, e_dphi)
        entry = {"block": name, "origin": list(origin), "points": int(gsb.shape[0]), "phi_err": e_phi, "dphi_err": e_dphi}
        if with_K:
            nodes = complete_nodes(n, origin, B)
            cols = reference_columns(nnodes, gsb, off, dom, phiq, dphiq, nodes, quad=True)
            kmax, kerr = 0.0, 0.0
            for j in nodes:
                rows_ref, vals_ref = cols[int(j)]
                cp, rows = ctx.get_pattern_cols(sid, 1, int(j), int(j) + 1)
                vals = np.empty(rows.size)
                ctx.get_values_cols(sid, int(j), int(j) + 1, vals)
                same = np.array_equal(rows, rows_ref)
                rep["K_pattern_bit_exact"] &= bool(same)
                assert same, f"block {name}: rows of column {j} differ"
                kmax = max(kmax, float(np.abs(vals_ref).max())); kerr = max(kerr, float(np.abs(vals - vals_ref).max()))
            e_k = kerr / kmax
            rep["K_err_vs_truth"] = max(rep["K_err_vs_truth"], e_k); rep["K_columns"] += int(nodes.size)
            entry["K_columns"] = int(nodes.size); entry["K_err"] = e_k
            assert e_k <= tol, (name, e_k)
        rep["points"] += int(gsb.shape[0])
        rep["blocks"].append(entry)
    return rep

#!/usr/bin/env python
"""bench.py -- Gauss points/second of the EFG hot path (neighbour search + MLS + K assembly, FP64).

Workload (BASELINE.json configs[1] scaled as the metric is quoted): examples/Poisson3D_Source.jl on a
synthetic 10^6-node grid -- Domain (1,1,1), Divisions (100,100,100) => 1 030 301 nodes, 27 Gauss points per
cell => 2.7e7 Gauss points, dmax 1.35, trilinear IMLS, K = int grad(dT).grad(T).  Deterministic, no RNG.

A "step" = one pass of the hot path over the whole Gauss table: fg_neighbours (Gauss-point binning included) ->
fg_shape_functions -> fg_assemble_bilinear with inputs resident in HBM and the CSC pattern prebuilt (it depends on the
node cloud and Gauss table only and is amortised over Newton steps; its build time is reported in `config`).
`e2e` = the same pass through the reference-facing API with HOST buffers, what build_space + the assemblers do for a
model: upload x, dm and the cell connectivity from pinned memory, Gauss table (fg_generate_gauss, bit-identical to egauss),
neighbours, shape functions, pattern, K and f, download nzval and F.

N > 1 (torchrun): the SAME 10^6-node problem split into z-slabs of background cells, one per GPU (strong scaling, the
north star's "1/2/4/8 B200" figure; --scaling weak gives every rank its own 100^3-cell slab instead).  Halo columns are
exchanged inside the library (fg_exchange_columns: NCCL send/recv + add on the context's stream).  value = all ranks'
Gauss points / max-over-ranks device time.  After the timed region every rank cross-checks the columns around its slab
interfaces against an independent single-GPU assembly of a window of the global grid (`parity_check`).

--config selects the other BASELINE.json workloads (single GPU): elasticity3d (C3: 3 DOF/node), newton (C4: re-assembly
with per-Gauss-point tensors + field evaluation per iteration), mk (Moving Kriging shape functions), solve (f3: PCG with K
resident).  Each prints its own JSON line with the same keys.

--impl reference: the reference's own algorithm (brute-force domain() + Gram-Schmidt IMLS + COO/sparse())
restated in C (oracle/, parity unpinned: Julia is not installed) timed on the host cores on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gauss points/sec (MLS + K assembly, FP64)"
UNIT = "Gauss points/s"
DMAX = 1.35
DEGREE = 3
FP64_PEAK_TFLOPS = 36.4          # measured DFMA peak of this pool's B200 (profiles/r01_microbench_b200.json, tools/microbench.cu)
SURVEY_FLOP_PER_POINT = 13.3e3   # algorithmic FLOPs per Gauss point at C2'


def algorithmic_bytes_per_point(dim, nnodes, ngauss, total_L, nnz, D=1):
    """B = 8(dim+2) + 48 nnodes/ngauss + Lbar (8 + 8 dim + 4) + 4 + 8 nnz D^2 / ngauss."""
    lbar = total_L / ngauss
    return 8 * (dim + 2) + 48.0 * nnodes / ngauss + lbar * (8 + 8 * dim + 4) + 4 + 8.0 * nnz * D * D / ngauss


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the GPU is under load (B200_PROFILING.md recipe): one streaming
    nvidia-smi process (-lms 100) started before the warm-up and stopped after the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            pass

    def start(self):
        return self

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[t.strip() for t in ln.split(",")] for ln in out.splitlines() if ln.count(",") >= 6]
            except Exception:
                pass
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[3 + k].lower().startswith("active")})
        # samples under load = those in the upper half by power draw
        load = [s for s, p in zip(sm, pw) if p >= (sorted(pw)[len(pw) // 2] if pw else 0)] if len(pw) == len(sm) else sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def pinned_array(shape, torch, dtype=None):
    dtype = dtype or torch.float64
    t = torch.empty(int(np.prod(shape)), dtype=dtype, pin_memory=True)
    return t, t.numpy().reshape(shape, order="F")


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    return (peaks["hbm_gbs"], "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline (the oracle port of the reference algorithm; the one place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import efg_oracle as oracle
    oracle.build()
    return oracle


def cpu_leg(x, dm, gs, seconds, threads):
    """Brute-force domain() + Gram-Schmidt IMLS + COO/sparse() on a bounded sample from the middle of the grid."""
    oracle = _oracle()
    ng = gs.shape[0]
    mid = (ng // 54) * 27
    probe = np.asfortranarray(gs[mid:mid + 27 * 2 * threads])
    t0 = time.perf_counter()
    oracle.shape_fun(probe, x, dm, nthreads=threads)
    rate = probe.shape[0] / (time.perf_counter() - t0)
    cells = int(max(2, min(ng // 27, seconds * rate / 27)))
    sample = np.asfortranarray(gs[mid:mid + 27 * cells])
    t0 = time.perf_counter()
    off, dom, phi, dphi = oracle.shape_fun(sample, x, dm, nthreads=threads)          # domain() + IMLS
    oracle.assemble_gradgrad_coo(sample, off, dom, dphi, x.shape[0])                  # COO push + sparse()
    dt = time.perf_counter() - t0
    return sample.shape[0] / dt, dt, sample.shape[0], cells


def cpu_baseline_leg(x, dm, gs, seconds=12.0):
    """1 core (the reference is single-threaded: BASELINE.md's headline baseline) and all host cores (courtesy bound)."""
    cores = os.cpu_count() or 1
    v1, dt1, n1, c1 = cpu_leg(x, dm, gs, seconds, 1)
    va, dta, na, ca = cpu_leg(x, dm, gs, seconds, cores)
    what = "of the same 10^6-node workload, brute-force domain() + Gram-Schmidt IMLS + COO/sparse() assembly"
    return {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n1} Gauss points ({c1} cells from the middle of the grid) {what}, {dt1:.1f} s on 1 core",
            "all_cores": {"value": va, "cores": cores, "sample": f"{na} Gauss points ({ca} cells), {dta:.1f} s on {cores} threads"}}


def run_reference(args):
    """The reference arm: the reference's CPU algorithm (oracle port; Julia is absent) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import flexigal_b200.geometry as G
    n = args.n
    model = G.create_model((1.0, 1.0, 1.0), (n, n, n), boundary=False)
    dm = G.Influence_Domains(model, DMAX)
    gs = G.egauss(model.x, model.conn, G.pgauss(DEGREE))
    cores = os.cpu_count() or 1
    times, pts, last = [], 0, None
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for it in range(args.warmup + args.steps):
        v, dt, npts, cells = cpu_leg(model.x, dm, gs, per_step, cores)
        last = (npts, cells, dt)
        if it >= args.warmup:
            times.append(dt); pts += npts
    value = pts / sum(times)
    sample = (f"{last[0]} Gauss points ({last[1]} cells from the middle of the grid) of the same 10^6-node workload, brute-force domain() + "
              f"Gram-Schmidt IMLS + COO/sparse() assembly, {last[2]:.1f} s per step on {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, n), "sampled": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def global_counts(n, nz):
    """(nodes, nnz) of the GLOBAL (n, n, nz)-cell problem: on the tensor grid the sparsity count factorises into 1-D counts
    (pairs of nodes that share a Gauss point), computed here by brute force with the reference's predicate."""
    import flexigal_b200.geometry as G
    def pairs(m):
        h = 1.0 / m
        nodes = h * np.arange(m + 1)
        g = (np.arange(m)[:, None] + 0.5 * (1.0 + np.unique(G.pgauss(DEGREE)[0]))[None, :]).reshape(-1) * h
        inside = np.abs(nodes[None, :] - g[:, None]) <= DMAX * h
        return int(((inside.T.astype(np.int64) @ inside.astype(np.int64)) > 0).sum())
    return (n + 1) * (n + 1) * (nz + 1), pairs(n) ** 2 * pairs(nz)


def workload_name(n, nz):
    return f"Poisson3D_Source scaled: Divisions=({n},{n},{nz}), dmax={DMAX}, degree {DEGREE}, IMLS, K=gradgrad"


# ------------------------------------------------------------------------------------------------------------------
# parity legs (untimed): the oracle as the checker
# ------------------------------------------------------------------------------------------------------------------
def headline_parity(ctx, sid, n, x, dm, gs):
    """N = 1: two blocks of 5^3 cells (corner + centre, 6 750 Gauss points) of the workload that was just timed, against the
    reference's brute-force domain() / __float128 IMLS / K columns."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    _oracle()
    import headline_check as hc
    t0 = time.perf_counter()
    try:
        rep = hc.check_against_device(ctx, sid, n, x, dm, gs, hc.block_origins(n, 5, ("corner", "centre")), B=5,
                                      nthreads=os.cpu_count() or 1, with_K=True)
        rep["ok"] = True
    except AssertionError as e:
        rep = {"ok": False, "error": str(e)}
    rep["seconds"] = round(time.perf_counter() - t0, 1)
    rep["what"] = "sampled blocks of the timed workload vs oracle: DOM/offsets bit-exact, phi/dphi/K columns <= 1e-12 vs __float128"
    return rep


def interface_parity(fg, part, ctx, sid, op, device, stream):
    """N > 1: the columns within two node layers of each of this rank's slab interfaces (after the NCCL exchange) against an
    independent single-context assembly of a 10-layer window of the global grid: rows bit-exact, values <= 1e-12."""
    from flexigal_b200.distributed import SlabPartition
    rep = {"interfaces": 0, "columns": 0, "pattern_bit_exact": True, "max_rel_err": 0.0}
    nz = part.divisions[2]
    own_lo, own_hi = part.owned_layers()
    for c in ([part.c0] if part.rank > 0 else []) + ([part.c1] if part.rank < part.world - 1 else []):
        lo, hi = max(0, c - 5), min(nz, c + 5)
        win = SlabPartition.window(part.domain, part.divisions, DMAX, lo, hi, DEGREE)
        cloud = fg.NodeCloud(win.local_nodes(), win.local_dm(), "rectangular", device=device, stream=stream)
        w = cloud.ctx
        ws = w.new_set_id()
        w.set_gauss(ws, win.local_gauss())
        w.neighbours(ws); w.shape_functions(ws, "IMLS"); w.pattern(ws); w.assemble_bilinear(ws, op, None)
        a_lay = max(c - 2, own_lo, lo + 2 if lo > 0 else 0)
        b_lay = min(c + 2, own_hi, (hi - 2 if hi < nz else nz) + 1)
        if a_lay < b_lay:
            a, b = part.local_node_range(a_lay, b_lay)
            wa, wb = win.local_node_range(a_lay, b_lay)
            cp, rv = ctx.get_pattern_cols(sid, 1, a, b, part.node_offset)
            vals = np.empty(rv.size); ctx.get_values_cols(sid, a, b, vals)
            cpw, rvw = w.get_pattern_cols(ws, 1, wa, wb, win.node_offset)
            valsw = np.empty(rvw.size); w.get_values_cols(ws, wa, wb, valsw)
            same = bool(np.array_equal(cp, cpw) and np.array_equal(rv, rvw))
            rep["pattern_bit_exact"] &= same
            if same:
                rep["max_rel_err"] = max(rep["max_rel_err"], float(np.abs(vals - valsw).max() / np.abs(valsw).max()))
                # K 1 = 0 on the checked columns (complete columns of a symmetric operator)
                sums = np.add.reduceat(vals, cp[:-1] - 1)
                rep["max_col_sum"] = max(rep.get("max_col_sum", 0.0), float(np.abs(sums).max() / np.abs(valsw).max()))
            rep["columns"] += b - a
        rep["interfaces"] += 1
        cloud.close()
    return rep


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--cells", dest="n", type=int, default=100, help="cells per direction (use --cells under torchrun: its parser claims --n)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: split the n^3 grid (strong) or one n^3 slab per rank (weak)")
    ap.add_argument("--config", default="headline", choices=["headline", "elasticity3d", "newton", "mk", "solve", "irregular"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="fg_set_option switch for A/B runs of kernel variants (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "headline":
        import bench_configs
        return bench_configs.run(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner to fd 1 when the communicator is created (NCCL_DEBUG=VERSION/WARN in the
        # environment): stdout must carry the JSON line only, so fd 1 points at stderr until the communicators exist
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        import flexigal_b200 as fg
        from flexigal_b200.distributed import SlabPartition, HaloExchange

        n = args.n
        weak = args.scaling == "weak" and world > 1
        nz = n * world if weak else n
        t_setup = time.perf_counter()
        part = SlabPartition((1.0, 1.0, 1.0 * (world if weak else 1)), (n, n, nz), DMAX, world, rank, degree=DEGREE)
        x, dm, gs = part.local_nodes(), part.local_dm(), part.local_gauss()
        ngauss = gs.shape[0]
        keep = []
        def pin(a, dtype=None):
            t, v = pinned_array(a.shape, torch, dtype); v[...] = a; keep.append(t); return v
        x_p, dm_p, gs_p = pin(x), pin(dm), pin(gs)
        pconn_p = pin(part.pattern_conn(), torch.int64) if world > 1 else None   # own + halo cells: the witness points are generated on the device
        conn_p = pin(part.local_conn(), torch.int64)      # e2e: build_space's inputs are the mesh (x, conn), the Gauss table is made on the device (f4)
        t_setup = time.perf_counter() - t_setup

        # a dedicated (non-legacy) stream shared by torch and the library: CUDA events recorded by torch
        # bracket exactly the kernels the C ABI launches
        tstream = torch.cuda.Stream()
        torch.cuda.set_stream(tstream)
        stream = tstream.cuda_stream
        cloud = fg.NodeCloud(x_p, dm_p, "rectangular", device=local_rank, stream=stream)
        ctx = cloud.ctx
        ctx.set_option("nb_rebin", 1)              # every timed search bins the Gauss points again, as for new inputs
        for kv in args.opt:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, gs_p)
        op = fg.GradGrad(1.0).c_struct(3, ngauss)
        t0 = time.perf_counter()
        if world > 1:
            ps = ctx.new_set_id()
            ctx.generate_gauss(ps, pconn_p, fg.pgauss(DEGREE))
            nnz = ctx.pattern_from(sid, ps)        # sparsity from own + halo Gauss positions: shared columns agree on both ranks
        else:
            nnz = ctx.pattern(sid)
        ctx.sync()
        t_pattern = time.perf_counter() - t0
        halo = HaloExchange(part, ctx, sid, torch, dist) if world > 1 else None     # creates the library's NCCL communicator
    finally:
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def step():
        ctx.neighbours(sid)
        ctx.shape_functions(sid, "IMLS")
        ctx.assemble_bilinear(sid, op, None)
        if halo is not None:
            halo.exchange_values()

    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2; inputs are > L2 anyway

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank).start()
    for _ in range(args.warmup):
        step()
    barrier()
    total_L, max_L = ctx.neighbours(sid)
    ctx.shape_functions(sid, "IMLS")
    barrier()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = {k: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)] for k in ("nb", "mls", "asm", "xchg")}
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        kev["nb"][i][0].record(); ctx.neighbours(sid); kev["nb"][i][1].record()
        kev["mls"][i][0].record(); ctx.shape_functions(sid, "IMLS"); kev["mls"][i][1].record()
        kev["asm"][i][0].record(); ctx.assemble_bilinear(sid, op, None); kev["asm"][i][1].record()
        kev["xchg"][i][0].record()
        if halo is not None:
            halo.exchange_values()
        kev["xchg"][i][1].record()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches0) // args.steps
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    phase_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in kev.items()}
    if halo is None:
        phase_ms.pop("xchg")
    phase_ms_ranks = None
    if world > 1:       # where the step goes on every rank (the exchange time of a rank includes waiting for its slower neighbour)
        mine = [phase_ms.get(k, 0.0) for k in ("nb", "mls", "asm", "xchg")] + [1e3 * t_dev / args.steps]
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        phase_ms_ranks = {k: [round(r[i], 3) for r in allp] for i, k in enumerate(("nb", "mls", "asm", "xchg", "step"))}
    tmax = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(ngauss), float(total_L), float(nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev_max, ngauss_all = float(tmax.item()), float(tot[0].item())
    value = ngauss_all * args.steps / t_dev_max

    # ---- parity of what was just timed (untimed; the oracle as the checker) -----------------------------
    parity = None
    if not args.no_parity:
        if world == 1:
            parity = headline_parity(ctx, sid, n, x, dm, gs) if n >= 10 else None
        else:
            mine = interface_parity(fg, part, ctx, sid, op, local_rank, stream)
            allrep = [None] * world
            dist.all_gather_object(allrep, mine)
            parity = {"what": "columns within 2 node layers of every slab interface, after the NCCL exchange, vs an independent single-GPU "
                              "assembly of a 10-layer window of the global grid (rows bit-exact, values relative to max |K|)",
                      "interfaces_checked": sum(r["interfaces"] for r in allrep), "columns": sum(r["columns"] for r in allrep),
                      "pattern_bit_exact": all(r["pattern_bit_exact"] for r in allrep),
                      "max_rel_err": max(r["max_rel_err"] for r in allrep), "max_col_sum": max(r.get("max_col_sum", 0.0) for r in allrep)}
            parity["ok"] = bool(parity["pattern_bit_exact"] and parity["max_rel_err"] <= 1e-12 and parity["columns"] > 0)

    # ---- e2e: host buffers in, host buffers out, through the public API -------------------------------
    e2e = None
    if not args.no_e2e:
        a_own, b_own = part.owned_columns()
        nnz_own = ctx.get_pattern_cols(sid, 1, a_own, b_own)[1].size if world > 1 else nnz
        nz_t, nz_host = pinned_array((nnz_own,), torch)
        F_t, F_host = pinned_array((x.shape[0],), torch)
        src = fg.Source(5000.0).c_struct(3, ngauss)
        # one long-lived context, like a Julia session holding the library handle: every step re-uploads the
        # inputs and redoes ALL the work (binning, neighbours, shape functions, sparsity, cluster index, K, f)
        c2 = fg.NodeCloud(x_p, dm_p, "rectangular", device=local_rank, stream=stream)
        s2 = c2.ctx.new_set_id()
        ps2 = c2.ctx.new_set_id() if world > 1 else None
        h2 = None
        if world > 1:
            sys.stdout.flush(); fd = os.dup(1); os.dup2(2, 1)
            try:
                h2 = HaloExchange(part, c2.ctx, s2, torch, dist)
            finally:
                sys.stdout.flush(); os.dup2(fd, 1); os.close(fd)
        pg = fg.pgauss(DEGREE)
        if world == 1:
            # build_space never reads the shape functions on the host: with "lazy_shapes" the IMLS kernels run inside the streamed
            # assembly, range by range, and the download of finished columns of K starts after the first range
            c2.ctx.set_option("lazy_shapes", 1)
        def e2e_step():
            c2.ctx.set_nodes(x_p, dm_p, "rectangular")                                       # uploads x, dm; bins nodes
            c2.ctx.generate_gauss(s2, conn_p, pg)                                            # uploads conn; egauss on the device (bit-identical table)
            c2.ctx.neighbours(s2)
            c2.ctx.shape_functions(s2, "IMLS")
            if world == 1:
                c2.ctx.pattern(s2)
            else:
                c2.ctx.generate_gauss(ps2, pconn_p, fg.pgauss(DEGREE))                       # uploads the connectivity, not a second Gauss table
                c2.ctx.pattern_from(s2, ps2)
            if world > 1:
                c2.ctx.assemble_bilinear(s2, op, None)
                h2.exchange_values()
                c2.ctx.get_values_cols(s2, a_own, b_own, nz_host)                            # downloads the owned column block
            else:
                c2.ctx.assemble_bilinear_async(s2, op, nz_host)                              # finished column blocks of nzval leave on the copy stream ...
            c2.ctx.assemble_linear(s2, src, None)                                            # ... while the load vector is assembled
            if world > 1:
                h2.exchange_vector()
            c2.ctx.get_vector(s2, F_host)                                                    # f, downloaded
            c2.ctx.sync()                                                                    # both downloads complete
        e2e_step()
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            e2e_step()                      # ends with a blocking D2H: the step is complete when it returns
            step_ms.append(round(1e3 * (time.perf_counter() - ts), 2))
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        parts = {}
        if world == 1:
            # untimed diagnostic pass: the same calls with a sync after each one (where the e2e time goes)
            c2.ctx.set_option("lazy_shapes", 0)           # the breakdown times every call on its own
            def timed(name, fn):
                c2.ctx.sync(); t = time.perf_counter(); fn(); c2.ctx.sync(); parts[name] = round(1e3 * (time.perf_counter() - t), 2)
            timed("set_nodes_h2d", lambda: c2.ctx.set_nodes(x_p, dm_p, "rectangular"))
            timed("generate_gauss_conn_h2d", lambda: c2.ctx.generate_gauss(s2, conn_p, pg))
            timed("neighbours", lambda: c2.ctx.neighbours(s2))
            timed("shape_functions", lambda: c2.ctx.shape_functions(s2, "IMLS"))
            timed("pattern", lambda: c2.ctx.pattern(s2))
            timed("assemble_bilinear_incl_cluster_index", lambda: c2.ctx.assemble_bilinear(s2, op, None))
            timed("get_values_d2h", lambda: c2.ctx.get_values(s2, nz_host))
            timed("assemble_linear_d2h", lambda: c2.ctx.assemble_linear(s2, src, F_host))
        e2e = {"value": ngauss_all / float(t_e2e.item()), "unit": UNIT, "breakdown_ms": parts, "step_ms": step_ms,
               "h2d_bytes_per_step": int(8 * (x_p.size + dm_p.size + conn_p.size + (pconn_p.size if pconn_p is not None else 0))),
               "d2h_bytes_per_step": int(8 * (nnz_own + x.shape[0])),
               "ms_per_step": 1e3 * float(t_e2e.item()), "steps": e2e_steps,
               "includes": "H2D of the mesh (x, dm, connectivity) from pinned host memory, node binning, Gauss table (egauss on the device), neighbours, IMLS, sparsity + cluster index build, K and f assembly"
                           + (", halo exchange, D2H of the owned column block of nzval and of F" if world > 1 else ", D2H of nzval/F")
                           + "; one long-lived context (device buffers reused); bytes are per rank"}

    if rank == 0:
        peak, peak_src = load_peaks()
        total_L_all, nnz_all = float(tot[1].item()), float(tot[2].item())
        # whole path, algorithmic bytes of the GLOBAL problem (halo duplicates are not algorithmic)
        nn_global, nnz_global = global_counts(n, nz)       # N>1: the local patterns also hold halo columns, which are not algorithmic
        bpp = algorithmic_bytes_per_point(3, nn_global, ngauss_all, total_L_all, nnz_global)
        lbar = total_L / ngauss
        # per-phase algorithmic bytes of THIS rank:  nb: read gs row, write offsets + hit masks ; mls: read gs row + masks + node
        # records, write DOM/phi/dphi ; asm: read DPHI + the 32-byte gather map + weights, write K
        # (the DOM write, 4 Lbar bytes per point, moved from the search to the shape-function kernel with the deferred fill pass)
        phase_bytes = {"nb": ngauss * (8 * 5 + 4 + 16) + 48.0 * x.shape[0],
                       "mls": ngauss * (8 * 5 + 16 + lbar * (4 + 8 + 24)) + 48.0 * x.shape[0],
                       "asm": ngauss * (16 + 32 + lbar * 24) + 8.0 * nnz}
        phases = {k: {"ms": phase_ms[k], "algorithmic_GB": phase_bytes[k] / 1e9, "GBps": phase_bytes[k] / (phase_ms[k] * 1e-3) / 1e9,
                      "hbm_frac": phase_bytes[k] / (phase_ms[k] * 1e-3) / 1e9 / peak} for k in ("nb", "mls", "asm")}
        dom_phase = max(phases, key=lambda k: phases[k]["ms"])
        achieved = phases[dom_phase]["GBps"]
        names = {"nb": "k_nb_cluster", "mls": "k_imls_p", "asm": "k_bilinear_pts"}
        traffic = None
        try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at this workload (ncu, profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(names[dom_phase]) if (world == 1 and n == 100) else None
        except Exception:
            pass
        fp64_tflops = SURVEY_FLOP_PER_POINT * value / 1e12 / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, nz),
                       "nodes_per_rank": int(x.shape[0]), "gauss_points_per_rank": int(ngauss), "sum_L_per_rank": int(total_L), "max_L": int(max_L),
                       "nnz_per_rank": int(nnz), "gauss_points_total": int(ngauss_all),
                       "partition": (f"z-slabs of background cells, one per GPU ({'one n^3 slab each' if weak else 'the n^3 grid split'}), halo columns "
                                     "exchanged inside the library (NCCL send/recv + add)") if world > 1 else "single GPU",
                       "l2": "inputs (1.08 GB gs + 18 GB shape functions) exceed the 126 MB L2; a 192 MB buffer is also rewritten between steps",
                       "pattern_build_ms": 1e3 * t_pattern, "host_setup_s": t_setup,
                       "phase_ms": phase_ms, "phase_ms_per_rank": phase_ms_ranks, "wall_ms_per_step": 1e3 * wall / args.steps,
                       "step": "fg_neighbours (Gauss-point binning included) + fg_shape_functions(IMLS) + fg_assemble_bilinear(gradgrad)"
                               + (" + fg_exchange_columns" if world > 1 else "")},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": names[dom_phase], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "phases": phases,
                         "algorithmic_bytes_per_gauss_point_whole_path": bpp,
                         "whole_path_achieved_GBps": bpp * value / 1e9 / world, "whole_path_frac": bpp * value / 1e9 / world / peak,
                         "fp64": {"flop_per_gauss_point_survey": SURVEY_FLOP_PER_POINT, "achieved_tflops_per_gpu": fp64_tflops,
                                  "peak_tflops": FP64_PEAK_TFLOPS, "peak_source": "measured DFMA microbenchmark (profiles/r01_microbench_b200.json)",
                                  "frac": fp64_tflops / FP64_PEAK_TFLOPS},
                         "fp64_frac": fp64_tflops / FP64_PEAK_TFLOPS,
                         "note": "the path is FP64-pipe bound (DESIGN.md): fp64_frac is the binding roofline; frac = dominant kernel's algorithmic bytes / its CUDA-event time / HBM copy peak"},
        }
        if parity is not None:
            line["parity_check"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(x, dm, gs)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- Gauss points/second of the EFG hot path (neighbour search + MLS + K assembly, FP64).

Workload (BASELINE.json configs[1] scaled as the metric is quoted): examples/Poisson3D_Source.jl on a
synthetic 10^6-node grid -- Domain (1,1,1), Divisions (100,100,100) => 1 030 301 nodes, 27 Gauss points per
cell => 2.7e7 Gauss points, dmax 1.35, trilinear IMLS, K = int grad(dT).grad(T).  Deterministic, no RNG.

A "step" = one pass of the hot path over the whole Gauss table: fg_neighbours -> fg_shape_functions ->
fg_assemble_bilinear with inputs resident in HBM and the CSC pattern prebuilt (it depends on the node
cloud and Gauss table only and is amortised over Newton steps; its build time is reported in `config`).
`e2e` = the same pass through the reference-facing API with HOST buffers: upload x, dm, gs from pinned
memory, neighbours, shape functions, pattern, K and f, download nzval and F.

N > 1 (torchrun): weak scaling, one z-slab of 100^3 cells per rank, halo rows exchanged over NCCL
(flexigal_b200/distributed.py); value = all ranks' Gauss points / max-over-ranks device time.

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


def pinned_array(shape, torch):
    t = torch.empty(int(np.prod(shape)), dtype=torch.float64, pin_memory=True)
    return t, t.numpy().reshape(shape, order="F")


def build_workload(n, rank=0, world=1):
    """Rank's z-slab of the (n, n, n*world) cell grid: local node cloud (halo layers included) and Gauss table."""
    import flexigal_b200 as fg
    from flexigal_b200.distributed import SlabPartition
    part = SlabPartition((1.0, 1.0, 1.0 * world), (n, n, n * world), DMAX, world, rank)
    return part, fg


def cpu_baseline_leg(part, x, dm, gs, seconds=15.0, threads=None):
    """Oracle port of the reference algorithm on the host cores, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import efg_oracle as oracle
    oracle.build()
    threads = threads or (os.cpu_count() or 1)
    ng = gs.shape[0]
    # calibrate on 27*8 points, then size the sample for ~`seconds`
    mid = (ng // 54) * 27
    probe = np.asfortranarray(gs[mid:mid + 27 * 8])
    t0 = time.perf_counter()
    oracle.shape_fun(probe, x, dm, nthreads=threads)
    rate = probe.shape[0] / (time.perf_counter() - t0)
    cells = int(max(8, min(ng // 27, seconds * rate / 27)))
    sample = np.asfortranarray(gs[mid:mid + 27 * cells])
    t0 = time.perf_counter()
    off, dom, phi, dphi = oracle.shape_fun(sample, x, dm, nthreads=threads)          # domain() + IMLS
    oracle.assemble_gradgrad_coo(sample, off, dom, dphi, x.shape[0])                  # COO push + sparse()
    dt = time.perf_counter() - t0
    return {"value": sample.shape[0] / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample.shape[0]} Gauss points ({cells} cells from the middle of the grid) of the same 10^6-node workload, "
                      f"brute-force domain() + Gram-Schmidt IMLS + COO/sparse() assembly, {dt:.1f} s"}, dt, sample.shape[0]


def run_reference(args):
    """The reference arm: the reference's CPU algorithm (oracle port; Julia is absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import flexigal_b200.geometry as G
    n = args.n
    model = G.create_model((1.0, 1.0, 1.0), (n, n, n), boundary=False)
    dm = G.Influence_Domains(model, DMAX)
    gs = G.egauss(model.x, model.conn, G.pgauss(DEGREE))
    times, pts = [], 0
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for it in range(args.warmup + args.steps):
        cb, dt, npts = cpu_baseline_leg(None, model.x, dm, gs, seconds=per_step)
        if it >= args.warmup:
            times.append(dt); pts += npts
    value = pts / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Poisson3D_Source scaled: Divisions=({n},{n},{n}), dmax={DMAX}, degree {DEGREE}, IMLS, K=gradgrad",
                       "sampled": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100, help="cells per direction (per rank along z)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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
        # environment): stdout must carry the JSON line only, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    import flexigal_b200 as fg
    from flexigal_b200 import _lib
    from flexigal_b200.distributed import SlabPartition, HaloExchange

    n = args.n
    t_setup = time.perf_counter()
    part = SlabPartition((1.0, 1.0, 1.0 * world), (n, n, n * world), DMAX, world, rank, degree=DEGREE)
    x, dm, gs = part.local_nodes(), part.local_dm(), part.local_gauss()
    ngauss = gs.shape[0]
    # pinned host copies (the e2e leg uploads from these every step)
    keep = []
    def pin(a):
        t, v = pinned_array(a.shape, torch); v[...] = a; keep.append(t); return v
    x_p, dm_p, gs_p = pin(x), pin(dm), pin(gs)
    pgs_p = pin(part.pattern_gauss()) if world > 1 else None      # own + halo Gauss positions (sparsity of shared columns)
    t_setup = time.perf_counter() - t_setup

    # a dedicated (non-legacy) stream shared by torch and the library: CUDA events recorded by torch
    # bracket exactly the kernels the C ABI launches, and NCCL ops are ordered with them
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    cloud = fg.NodeCloud(x_p, dm_p, "rectangular", device=local_rank, stream=stream)
    ctx = cloud.ctx
    sid = ctx.new_set_id()
    ctx.set_gauss(sid, gs_p)
    op = fg.GradGrad(1.0).c_struct(3, ngauss)
    t0 = time.perf_counter()
    pat_sid = part.pattern_set(ctx, sid, pgs_p)  # N>1: pattern from own + halo Gauss points so shared columns agree
    nnz = ctx.pattern(sid) if pat_sid is None else part.build_pattern(ctx, sid, pat_sid)
    ctx.sync()
    t_pattern = time.perf_counter() - t0
    halo = HaloExchange(part, ctx, sid, torch, dist) if world > 1 else None

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
    kev = {k: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)] for k in ("nb", "mls", "asm")}
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        kev["nb"][i][0].record(); ctx.neighbours(sid); kev["nb"][i][1].record()
        kev["mls"][i][0].record(); ctx.shape_functions(sid, "IMLS"); kev["mls"][i][1].record()
        kev["asm"][i][0].record(); ctx.assemble_bilinear(sid, op, None)
        if halo is not None:
            halo.exchange_values()
        kev["asm"][i][1].record()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches0) // args.steps
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    phase_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in kev.items()}
    tmax = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(ngauss)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev_max, ngauss_all = float(tmax.item()), float(tot.item())
    value = ngauss_all * args.steps / t_dev_max

    # ---- e2e: host buffers in, host buffers out, through the public API -------------------------------
    e2e = None
    if not args.no_e2e:
        nz_t, nz_host = pinned_array((nnz,), torch)
        F_t, F_host = pinned_array((x.shape[0],), torch)
        src = fg.Source(5000.0).c_struct(3, ngauss)
        # one long-lived context, like a Julia session holding the library handle: every step re-uploads the
        # inputs and redoes ALL the work (binning, neighbours, shape functions, sparsity, cluster index, K, f)
        c2 = fg.NodeCloud(x_p, dm_p, "rectangular", device=local_rank, stream=stream)
        s2 = c2.ctx.new_set_id()
        ps2 = c2.ctx.new_set_id() if world > 1 else None
        def e2e_step():
            c2.ctx.set_nodes(x_p, dm_p, "rectangular")                                       # uploads x, dm; bins nodes
            c2.ctx.set_gauss(s2, gs_p)                                                       # uploads gs
            c2.ctx.neighbours(s2)
            c2.ctx.shape_functions(s2, "IMLS")
            if world == 1:
                c2.ctx.pattern(s2)
            else:
                c2.ctx.set_gauss(ps2, pgs_p)
                c2.ctx.pattern_from(s2, ps2)
            c2.ctx.assemble_bilinear(s2, op, None)
            if world > 1:
                HaloExchange(part, c2.ctx, s2, torch, dist).exchange_values()
            c2.ctx.get_values(s2, nz_host)                                                   # downloads nzval
            c2.ctx.assemble_linear(s2, src, F_host)                                          # f, downloaded
        # free the device-resident arm first: the e2e arm owns its own context like a user would
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
        # untimed diagnostic pass: the same calls with a sync after each one (where the e2e time goes)
        parts = {}
        def timed(name, fn):
            c2.ctx.sync(); t = time.perf_counter(); fn(); c2.ctx.sync(); parts[name] = round(1e3 * (time.perf_counter() - t), 2)
        timed("set_nodes_h2d", lambda: c2.ctx.set_nodes(x_p, dm_p, "rectangular"))
        timed("set_gauss_h2d", lambda: c2.ctx.set_gauss(s2, gs_p))
        timed("neighbours", lambda: c2.ctx.neighbours(s2))
        timed("shape_functions", lambda: c2.ctx.shape_functions(s2, "IMLS"))
        if world == 1:
            timed("pattern", lambda: c2.ctx.pattern(s2))
        else:
            timed("pattern", lambda: (c2.ctx.set_gauss(ps2, pgs_p), c2.ctx.pattern_from(s2, ps2)))
        timed("assemble_bilinear_incl_cluster_index", lambda: c2.ctx.assemble_bilinear(s2, op, None))
        timed("get_values_d2h", lambda: c2.ctx.get_values(s2, nz_host))
        timed("assemble_linear_d2h", lambda: c2.ctx.assemble_linear(s2, src, F_host))
        e2e = {"value": ngauss_all / float(t_e2e.item()), "unit": UNIT, "breakdown_ms": parts, "step_ms": step_ms,
               "h2d_bytes_per_step": int(8 * (x_p.size + dm_p.size + gs_p.size + (pgs_p.size if pgs_p is not None else 0))),
               "d2h_bytes_per_step": int(8 * (nnz + x.shape[0])),
               "ms_per_step": 1e3 * float(t_e2e.item()), "steps": e2e_steps,
               "includes": "H2D of x/dm/gs from pinned host memory, node binning, neighbours, IMLS, sparsity + cluster index build, K and f assembly, D2H of nzval/F; one long-lived context (device buffers reused)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        bpp = algorithmic_bytes_per_point(3, x.shape[0], ngauss, total_L, nnz)
        # dominant kernel = the longest phase; its algorithmic bytes per launch:
        #   mls: read gs row + node data + DOM, write phi/dphi ; asm: read PHI/DPHI/DOM + gs, write K
        lbar = total_L / ngauss
        phase_bytes = {"nb": ngauss * (8 * 5 + 4 + 4 * lbar) + 48.0 * x.shape[0],
                       "mls": ngauss * (8 * 5 + lbar * (4 + 8 + 24)) + 48.0 * x.shape[0],
                       "asm": ngauss * (16 + lbar * (4 + 8 + 24)) + 8.0 * nnz}
        dom_phase = max(phase_ms, key=phase_ms.get)
        achieved = phase_bytes[dom_phase] / (phase_ms[dom_phase] * 1e-3) / 1e9
        names = {"nb": "k_nb_cluster", "mls": "k_imls", "asm": "k_bilinear_reg"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Poisson3D_Source scaled: Divisions=({n},{n},{n * world}), dmax={DMAX}, degree {DEGREE}, IMLS, K=gradgrad",
                       "nodes_per_rank": int(x.shape[0]), "gauss_points_per_rank": int(ngauss), "sum_L_per_rank": int(total_L), "max_L": int(max_L),
                       "nnz_per_rank": int(nnz), "partition": "z-slabs, one per GPU" if world > 1 else "single GPU",
                       "l2": "inputs (1.08 GB gs + 18 GB shape functions) exceed the 126 MB L2; a 192 MB buffer is also rewritten between steps",
                       "pattern_build_ms": 1e3 * t_pattern, "host_setup_s": t_setup,
                       "phase_ms": phase_ms, "wall_ms_per_step": 1e3 * wall / args.steps},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": names[dom_phase], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of that kernel per launch at this workload (ncu, profiles/r01_launches_final.md)
                         "traffic": {"k_bilinear_reg": 18.33e9, "k_imls": 21.50e9, "k_nb_cluster": 5.38e9}.get(names[dom_phase]) if (world == 1 and n == 100) else None,
                         "algorithmic_bytes_per_gauss_point_whole_path": bpp,
                         "whole_path_achieved_GBps": bpp * value / 1e9 / world, "whole_path_frac": bpp * value / 1e9 / world / peak,
                         "note": "the path is FP64-pipe bound (DESIGN.md); frac is the dominant kernel's algorithmic bytes / its CUDA-event time"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_baseline_leg(part, x, dm, gs)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

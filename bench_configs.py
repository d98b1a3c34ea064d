"""bench.py --config <name>: the other BASELINE.json workloads, one JSON line each (single GPU, same keys as the headline line).

  elasticity3d  C3 as SURVEY 8(d) scales it: Domain (5,1,1), Divisions (200,40,40), 3 DOF per node, E = 1000, nu = 0.3
                (examples/Elasticity_Test.jl:7-10,25-26), trilinear IMLS.  step = search + IMLS + grad(du) (.) sigma(u) assembly.
  newton        C4 flavour (examples/SaintVenant_Kirchoff_Test.jl:28-34, src/Solvers.jl:184-189): every iteration evaluates
                grad u_h at the Gauss points (fg_evaluate_field) and re-assembles the tangent with a per-Gauss-point coefficient
                tensor on the SAME sparsity; 2D, VectorField{2}, Divisions (900,400).  step = one iteration.
  mk            Moving Kriging shape functions (technique = :MK, what C3/C4/C5 select as shipped): 2D, Divisions (1500,300),
                dmax 1.75.  step = search + MK.
  irregular     the headline workload with every interior node moved by 0.2 h U(-1,1) (no structured-grid shortcuts).
  solve         SURVEY 8(f) f3: the 10^6-node Poisson system with the penalty Dirichlet step, solved by Jacobi-PCG with K RESIDENT
                on the device.  metric = unknowns x iterations / s; iteration count and residual reported.

All inputs are synthetic and deterministic.  `value` is device-timed with the inputs resident in HBM, `e2e` goes through the
public API with host buffers (uploads and downloads inside the timed region)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _events(torch, n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def _line(metric, unit, value, args, ms, workload, extra_cfg, e2e, launches, clocks, roofline, dtype="f64"):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": dict({"workload": workload}, **extra_cfg), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}


def run(args):
    import torch
    import bench
    import flexigal_b200 as fg
    torch.cuda.set_device(0)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    peak, peak_src = bench.load_peaks()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sampler = bench.ClockSampler(0).start()
    cfg = args.config

    def timed_steps(step_fns, names):
        """W warm-up steps, then K timed ones with per-phase CUDA events; returns (ms per step, {phase: ms})."""
        for _ in range(args.warmup):
            for f in step_fns:
                f()
        torch.cuda.synchronize()
        ev = _events(torch, args.steps)
        kev = {k: _events(torch, args.steps) for k in names}
        for i in range(args.steps):
            flush.zero_()
            ev[i][0].record()
            for k, f in zip(names, step_fns):
                kev[k][i][0].record(); f(); kev[k][i][1].record()
            ev[i][1].record()
        torch.cuda.synchronize()
        return (sum(a.elapsed_time(b) for a, b in ev) / args.steps,
                {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in kev.items()})

    if cfg == "elasticity3d":
        div = (200, 40, 40) if args.n == 100 else (2 * args.n, args.n // 2 + 1, args.n // 2 + 1)
        model = fg.create_model((5.0, 1.0, 1.0), div, boundary=False)
        dm = fg.Influence_Domains(model, bench.DMAX)
        gs = fg.egauss(model.x, model.conn, fg.pgauss(3))
        ng, nn, D = gs.shape[0], model.nnodes, 3
        E, nu = 1000.0, 0.3
        G, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        cloud = fg.NodeCloud(model.x, dm, "rectangular", device=0, stream=stream)
        for kv in getattr(args, "opt", []):             # --opt KEY=VALUE: kernel-variant switches for A/B runs
            cloud.ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ctx = cloud.ctx
        ctx.set_option("nb_rebin", 1)
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, gs)
        op = fg.Elasticity(G, lam, D).c_struct(3, ng)
        total_L, max_L = ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        t0 = time.perf_counter(); nnz = ctx.pattern(sid); ctx.sync(); t_pat = time.perf_counter() - t0
        l0 = ctx.launch_count()
        ms, ph = timed_steps([lambda: ctx.neighbours(sid), lambda: ctx.shape_functions(sid, "IMLS"), lambda: ctx.assemble_bilinear(sid, op, None)],
                             ["nb", "mls", "asm"])
        launches = (ctx.launch_count() - l0) // (args.steps + args.warmup)
        # invariants of the assembled operator (rigid translations in the null space), on the device-resident values
        nzd = np.empty(nnz * D * D); ctx.get_values(sid, nzd)
        colptr, rowval = ctx.get_pattern(sid, D, nnz)
        import scipy.sparse as sp
        K = sp.csc_matrix((nzd, rowval - 1, colptr - 1), shape=(nn * D, nn * D))
        t = np.zeros(nn * D); t[0::D] = 1.0
        null_err = float(np.abs(K @ t).max() / np.abs(nzd).max())
        sym_err = float(abs(K - K.T).max() / np.abs(nzd).max())
        del K, colptr, rowval
        # e2e
        x_t, x_p = bench.pinned_array(model.x.shape, torch); x_p[...] = model.x
        d_t, d_p = bench.pinned_array(dm.shape, torch); d_p[...] = dm
        g_t, g_p = bench.pinned_array(gs.shape, torch); g_p[...] = gs
        nz_t, nz_h = bench.pinned_array((nnz * D * D,), torch)
        def e2e_step():
            ctx.set_nodes(x_p, d_p, "rectangular"); ctx.set_gauss(sid, g_p)
            ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS"); ctx.pattern(sid)
            ctx.assemble_bilinear(sid, op, nz_h)
        e2e_step(); ctx.sync()
        t0 = time.perf_counter(); e2e_step(); e2e_step(); t_e2e = (time.perf_counter() - t0) / 2
        lbar = total_L / ng
        bpp = bench.algorithmic_bytes_per_point(3, nn, ng, total_L, nnz, D)
        asm_bytes = ng * (16 + lbar * (4 + 24)) + 8.0 * nnz * D * D
        line = _line("Gauss points/sec (MLS + K assembly, FP64), 3 DOF per node", bench.UNIT, ng / (ms * 1e-3), args, ms,
                     f"Elasticity_Test 3D variant: Domain=(5,1,1), Divisions={div}, D=3, E=1000, nu=0.3, dmax={bench.DMAX}, degree 3, IMLS, K=grad(du)(.)sigma(u)",
                     {"nodes": nn, "gauss_points": ng, "sum_L": int(total_L), "nnz_dof": int(nnz * D * D), "phase_ms": ph, "pattern_build_ms": 1e3 * t_pat,
                      "checks": {"rigid_translation_residual": null_err, "symmetry": sym_err},
                      "l2": "inputs exceed the 126 MB L2; a 192 MB buffer is rewritten between steps"},
                     {"value": ng / t_e2e, "unit": bench.UNIT, "ms_per_step": 1e3 * t_e2e,
                      "h2d_bytes_per_step": int(8 * (x_p.size + d_p.size + g_p.size)), "d2h_bytes_per_step": int(8 * nnz * D * D)},
                     launches, sampler.stop(),
                     {"bound": "hbm", "kernel": "k_elast_diag + k_elast_off", "achieved": asm_bytes / (ph["asm"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": asm_bytes / (ph["asm"] * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                      "algorithmic_bytes_per_gauss_point_whole_path": bpp, "whole_path_frac": bpp * ng / (ms * 1e-3) / 1e9 / peak,
                      "asm_ms_per_1e6_points": ph["asm"] / (ng / 1e6)})
        print(json.dumps(line))
        return 0

    if cfg == "irregular":
        # SURVEY 8(d)'s guard against structured-grid shortcuts: the headline grid with every interior node moved by 0.2 h U(-1, 1)
        n = args.n
        model = fg.create_model((1.0, 1.0, 1.0), (n, n, n), boundary=False)
        rng = np.random.default_rng(0)
        h = 1.0 / n
        x = model.x.copy(order="F")
        interior = np.all((model.x0 > 1e-12) & (model.x0 < 1.0 - 1e-12), axis=1)
        x[interior] += 0.2 * h * rng.uniform(-1, 1, size=(int(interior.sum()), 3))
        x = np.asfortranarray(x)
        dm = fg.Influence_Domains(model, bench.DMAX)
        gs = fg.egauss(x, model.conn, fg.pgauss(3))
        ng, nn = gs.shape[0], model.nnodes
        cloud = fg.NodeCloud(x, dm, "rectangular", device=0, stream=stream)
        for kv in getattr(args, "opt", []):             # --opt KEY=VALUE: kernel-variant switches for A/B runs
            cloud.ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ctx = cloud.ctx
        ctx.set_option("nb_rebin", 1)
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, gs)
        op = fg.GradGrad(1.0).c_struct(3, ng)
        total_L, max_L = ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        t0 = time.perf_counter(); nnz = ctx.pattern(sid); ctx.sync(); t_pat = time.perf_counter() - t0
        ctx.assemble_bilinear(sid, op, None)
        stats = ctx.assembly_stats(sid)
        l0 = ctx.launch_count()
        ms, ph = timed_steps([lambda: ctx.neighbours(sid), lambda: ctx.shape_functions(sid, "IMLS"), lambda: ctx.assemble_bilinear(sid, op, None)],
                             ["nb", "mls", "asm"])
        launches = (ctx.launch_count() - l0) // (args.steps + args.warmup)
        # invariants on the device result: K annihilates constants; IMLS reproduces the trilinear monomial on any cloud
        val, grad = ctx.evaluate_field(sid, ng, x[:, 0] * x[:, 1] * x[:, 2], 1)
        rep = float(np.abs(val[:, 0] - gs[:, 0] * gs[:, 1] * gs[:, 2]).max())
        del val, grad
        nz = np.empty(nnz); ctx.get_values(sid, nz)
        colptr, rowval = ctx.get_pattern(sid, 1, nnz)
        import scipy.sparse as sp
        K = sp.csc_matrix((nz, (rowval - 1).astype(np.int32), colptr - 1), shape=(nn, nn))
        k1 = float(np.abs(K @ np.ones(nn)).max() / np.abs(nz).max())
        del K, colptr, rowval, nz
        lbar = total_L / ng
        bpp = bench.algorithmic_bytes_per_point(3, nn, ng, total_L, nnz)
        mls_bytes = ng * (8 * 5 + 16 + lbar * 36) + 48.0 * nn
        line = _line(bench.METRIC + ", irregular cloud", bench.UNIT, ng / (ms * 1e-3), args, ms,
                     f"Poisson3D_Source scaled, nodes jittered by 0.2 h (numpy default_rng(0)): Divisions=({n},{n},{n}), dmax={bench.DMAX}, degree 3, IMLS, K=gradgrad",
                     {"nodes": nn, "gauss_points": ng, "sum_L": int(total_L), "max_L": int(max_L), "nnz": int(nnz), "phase_ms": ph, "pattern_build_ms": 1e3 * t_pat,
                      "assembly": stats, "checks": {"trilinear_reproduction_max_err": rep, "K_times_ones_rel": k1},
                      "l2": "inputs exceed the 126 MB L2; a 192 MB buffer is rewritten between steps"},
                     None, launches, sampler.stop(),
                     {"bound": "hbm", "kernel": "k_imls_p", "achieved": mls_bytes / (ph["mls"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": mls_bytes / (ph["mls"] * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                      "algorithmic_bytes_per_gauss_point_whole_path": bpp, "whole_path_frac": bpp * ng / (ms * 1e-3) / 1e9 / peak})
        print(json.dumps(line))
        return 0

    if cfg == "newton":
        div = (900, 400) if args.n == 100 else (9 * args.n, 4 * args.n)
        model = fg.create_model((2.25, 1.0), div, boundary=False)
        dm = fg.Influence_Domains(model, 1.5)
        gs = fg.egauss(model.x, model.conn, fg.pgauss(3))
        ng, nn, D, nb = gs.shape[0], model.nnodes, 2, 3
        cloud = fg.NodeCloud(model.x, dm, "rectangular", device=0, stream=stream)
        for kv in getattr(args, "opt", []):             # --opt KEY=VALUE: kernel-variant switches for A/B runs
            cloud.ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ctx = cloud.ctx
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, gs)
        total_L, _ = ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        nnz = ctx.pattern(sid)
        X = model.x
        u = np.stack([0.01 * np.sin(np.pi * X[:, 0]) * X[:, 1], 0.02 * X[:, 0] * X[:, 1] ** 2], axis=1).reshape(-1)
        n2 = (D * nb) ** 2
        # device-resident arm: tensors live on the GPU (a Newton loop that keeps u_h and the constitutive update on the device)
        Cdev = torch.zeros(ng * n2, dtype=torch.float64, device="cuda")
        import ctypes
        from flexigal_b200 import _lib
        op_dev = _lib.fg_operator(_lib.OP_GENERIC, D, (ctypes.c_double * 4)(0, 0, 0, 0), ctypes.c_void_p(Cdev.data_ptr()), n2)

        def tangent_from(grad):          # a St-Venant-Kirchhoff-like tangent: C = lam tr + 2 mu sym + geometric term ~ grad u
            C = np.zeros((ng, D * nb, D * nb))
            F11 = 1.0 + grad[:, 0, 0]; F22 = 1.0 + grad[:, 1, 1]
            for i in range(D):
                for j in range(D):
                    for a in range(2):
                        for b in range(2):
                            C[:, i * nb + 1 + a, j * nb + 1 + b] = (384.6 * (i == j) * (a == b) + 576.9 * (i == a) * (j == b) + 384.6 * (i == b) * (j == a))
                    C[:, i * nb + 1 + i, j * nb + 1 + j] *= F11 * F22
            return C

        grad0 = ctx.evaluate_field(sid, ng, u, D, values=False, gradients=True)[1]
        C_host = tangent_from(grad0)
        Cdev.copy_(torch.from_numpy(C_host.reshape(-1)))
        ctx.set_option("coef_on_device", 1)
        ctx.assemble_bilinear(sid, op_dev, None); ctx.sync()
        l0 = ctx.launch_count()
        ms, ph = timed_steps([lambda: ctx.assemble_bilinear(sid, op_dev, None)], ["asm"])
        launches = (ctx.launch_count() - l0) // (args.steps + args.warmup)
        ctx.set_option("coef_on_device", 0)
        # e2e iteration: upload u, field gradients down, tensors up (host-computed, as the Julia closure would), K down
        nz_t, nz_h = bench.pinned_array((nnz * D * D,), torch)
        c_t, c_p = bench.pinned_array((ng * n2,), torch); c_p[...] = C_host.reshape(-1)
        op_h = _lib.fg_operator(_lib.OP_GENERIC, D, (ctypes.c_double * 4)(0, 0, 0, 0), ctypes.c_void_p(c_p.ctypes.data), n2)
        def e2e_iter():
            ctx.evaluate_field(sid, ng, u, D, values=False, gradients=True)
            ctx.assemble_bilinear(sid, op_h, nz_h)
        e2e_iter(); ctx.sync()
        t0 = time.perf_counter(); e2e_iter(); e2e_iter(); t_e2e = (time.perf_counter() - t0) / 2
        lbar = total_L / ng
        asm_bytes = ng * (16 + 8 * n2 + lbar * (4 + 8 + 16)) + 8.0 * nnz * D * D
        line = _line("Gauss points/sec (re-assembly of the tangent K with per-Gauss-point tensors, FP64)", bench.UNIT, ng / (ms * 1e-3), args, ms,
                     f"SaintVenant_Kirchoff-like Newton iteration: 2D, Divisions={div}, VectorField{{2}}, dmax 1.5, degree 3, IMLS, generic 6x6 tensor per Gauss point, sparsity reused",
                     {"nodes": nn, "gauss_points": ng, "sum_L": int(total_L), "nnz_dof": int(nnz * D * D), "phase_ms": ph,
                      "l2": "a 192 MB buffer is rewritten between steps"},
                     {"value": ng / t_e2e, "unit": bench.UNIT, "ms_per_step": 1e3 * t_e2e,
                      "includes": "fg_evaluate_field (u up, gradients down) + fg_assemble_bilinear with host tensors (up) and nzval (down)",
                      "h2d_bytes_per_step": int(8 * (u.size + ng * n2)), "d2h_bytes_per_step": int(8 * (ng * D * 2 + nnz * D * D))},
                     launches, sampler.stop(),
                     {"bound": "hbm", "kernel": "k_bilinear_warp<GENERIC>", "achieved": asm_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": asm_bytes / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None})
        print(json.dumps(line))
        return 0

    if cfg == "mk":
        div = (1500, 300) if args.n == 100 else (15 * args.n, 3 * args.n)
        model = fg.create_model((5.0, 1.0), div, boundary=False)
        dm = fg.Influence_Domains(model, 1.75)
        gs = fg.egauss(model.x, model.conn, fg.pgauss(3))
        ng, nn = gs.shape[0], model.nnodes
        cloud = fg.NodeCloud(model.x, dm, "rectangular", device=0, stream=stream)
        for kv in getattr(args, "opt", []):             # --opt KEY=VALUE: kernel-variant switches for A/B runs
            cloud.ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ctx = cloud.ctx
        ctx.set_option("nb_rebin", 1)
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, gs)
        total_L, max_L = ctx.neighbours(sid)
        l0 = ctx.launch_count()
        ms, ph = timed_steps([lambda: ctx.neighbours(sid), lambda: ctx.shape_functions(sid, "MK")], ["nb", "mk"])
        launches = (ctx.launch_count() - l0) // (args.steps + args.warmup)
        sing = ctx.singular_count(sid)
        val, _ = ctx.evaluate_field(sid, ng, np.ones(nn), 1, values=True, gradients=False)
        pu = float(np.abs(val[:, 0] - 1.0).max())
        g_t, g_p = bench.pinned_array(gs.shape, torch); g_p[...] = gs
        phi_h = np.empty(total_L); dphi_h = np.empty((total_L, 2)); off_h = np.empty(ng + 1, dtype=np.int64); dom_h = np.empty(total_L, dtype=np.int64)
        def e2e_step():
            ctx.set_gauss(sid, g_p); ctx.neighbours(sid); ctx.shape_functions(sid, "MK")
            ctx.lib.fg_get_shapes(ctx.h, sid, off_h.ctypes.data, dom_h.ctypes.data, phi_h.ctypes.data, dphi_h.ctypes.data)
        e2e_step()
        t0 = time.perf_counter(); e2e_step(); t_e2e = time.perf_counter() - t0
        lbar = total_L / ng
        mk_bytes = ng * (8 * 4 + lbar * (4 + 8 + 16)) + 32.0 * nn
        line = _line("Gauss points/sec (neighbour search + Moving Kriging shape functions, FP64)", bench.UNIT, ng / (ms * 1e-3), args, ms,
                     f"Elasticity_Test-shaped cloud: 2D, Domain=(5,1), Divisions={div}, dmax 1.75, degree 3, technique MK",
                     {"nodes": nn, "gauss_points": ng, "sum_L": int(total_L), "max_L": int(max_L), "phase_ms": ph,
                      "checks": {"singular_points": int(sing), "partition_of_unity_max_dev": pu}},
                     {"value": ng / t_e2e, "unit": bench.UNIT, "ms_per_step": 1e3 * t_e2e, "includes": "gs up, search, MK, (offsets, DOM, PHI, DPHI) down",
                      "h2d_bytes_per_step": int(8 * g_p.size), "d2h_bytes_per_step": int(8 * (ng + 1) + total_L * (8 + 8 + 16))},
                     launches, sampler.stop(),
                     {"bound": "hbm", "kernel": "k_mk", "achieved": mk_bytes / (ph["mk"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": mk_bytes / (ph["mk"] * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                      "note": "per point an L x L Gaussian correlation matrix is factorised (L = 9..16): FP64/latency bound, not HBM bound"})
        print(json.dumps(line))
        return 0

    if cfg == "solve":
        n = args.n
        model = fg.create_model((1.0, 1.0, 1.0), (n, n, n))
        tags = ["Left", "Right", "Top", "Bottom", "Back", "Front"]
        tri = fg.Triangulation(model, "Domain")
        recipe = fg.ApproxSpace(model, tri, D=1, dmax=bench.DMAX)
        cloud = recipe.cloud()
        ctx = cloud.ctx
        ctx.set_stream(stream)
        dO = fg.IntegrationSet(tri, 3)
        sp_dom = fg.build_space(recipe, [dO])
        sA = sp_dom.merged().set_id
        ngauss = sp_dom.merged().ngauss
        ctx.pattern(sA)
        ctx.assemble_bilinear(sA, fg.GradGrad(1.0).c_struct(3, ngauss), None)
        alpha = ctx.values_max(sA) * 100.0
        F = np.empty(model.nnodes); ctx.assemble_linear(sA, fg.Source(5000.0).c_struct(3, ngauss), F)
        sp_bnd = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, t), 3) for t in tags])
        sB = sp_bnd.merged().set_id
        ctx.pattern(sB)
        ctx.assemble_bilinear(sB, fg.Mass(alpha).c_struct(3, sp_bnd.merged().ngauss), None)
        nnzA, nnzB = ctx.pattern(sA), ctx.pattern(sB)
        rtol = 1e-10
        x, it, res = ctx.solve_spd_sets([sA, sB], F, 1, None, rtol, 20000)          # warm-up + iteration count
        ev = _events(torch, args.steps)
        for i in range(args.steps):
            ev[i][0].record(); ctx.solve_spd_sets([sA, sB], F, 1, None, rtol, 20000); ev[i][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
        per_it = ms / max(it, 1)
        spmv_bytes = (nnzA + nnzB) * 12.0 + model.nnodes * 8.0 * 14
        line = _line("unknowns x PCG iterations / s (Jacobi-PCG, K resident on the device, FP64)", "unknown-iterations/s", model.nnodes * it / (ms * 1e-3), args, ms,
                     f"Poisson3D_Source scaled: Divisions=({n},{n},{n}), penalty Dirichlet on six faces (alpha = 100 max A), u = (A + Ap) \\\\ F to ||r|| <= 1e-10 ||F||",
                     {"unknowns": model.nnodes, "nnz_A": int(nnzA), "nnz_Ap": int(nnzB), "iterations": int(it), "relative_residual": res, "ms_per_iteration": per_it,
                      "alpha": alpha, "u_max": float(x.max())},
                     {"value": model.nnodes * it / (ms * 1e-3), "unit": "unknown-iterations/s", "ms_per_step": ms,
                      "includes": "rhs up, solution down (the operators stay on the device)", "h2d_bytes_per_step": int(8 * model.nnodes), "d2h_bytes_per_step": int(8 * model.nnodes)},
                     0, sampler.stop(),
                     {"bound": "hbm", "kernel": "k_spmv_set", "achieved": spmv_bytes / (per_it * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": spmv_bytes / (per_it * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                      "note": "per iteration: both operators streamed once (8 B value + 4 B row id per entry) + 14 vector passes"})
        print(json.dumps(line))
        return 0
    raise SystemExit(f"unknown config {cfg}")

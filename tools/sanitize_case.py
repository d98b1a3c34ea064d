#!/usr/bin/env python
"""A tiny pass through every kernel of the library (for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py`):
3D and 2D clouds, cluster and point-wise search, the three IMLS kernels, Moving Kriging, grad.grad / mass (D = 1, 2), elasticity,
advection, generic tensors (constant and per point), linear forms, node subsets + concatenation, field evaluation, resident PCG,
device Gauss tables.  Prints a checksum per stage; no oracle involved."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flexigal_b200 as fg  # noqa: E402


def run(dim):
    model = fg.create_model((1.0,) * dim, (5,) * dim if dim == 3 else (9, 7))
    tri = fg.Triangulation(model, "Domain")
    for D, tech in ((1, "IMLS"), (dim, "IMLS"), (1, "MK")):
        recipe = fg.ApproxSpace(model, tri, D=D, dmax=1.35 if dim == 3 else 1.6, technique=tech)
        space = fg.build_space(recipe, [fg.IntegrationSet(tri, 3)])
        S = space.merged()
        ctx = space.cloud.ctx
        ng = S.ngauss
        K = fg.Bilinear_Assembler(fg.GradGrad(1.0, D=D), space)
        M = fg.Bilinear_Assembler(fg.Mass(2.0, D=D), space)
        F = fg.Linear_Assembler(fg.Source(3.0, D=D), space)
        out = [K.data.sum(), M.data.sum(), F.sum()]
        if D == dim:
            out.append(fg.Bilinear_Assembler(fg.Elasticity(1.0, 2.0, D), space).data.sum())
        nb = 1 + dim
        C = np.random.default_rng(0).standard_normal((ng, D * nb, D * nb))
        out.append(fg.Bilinear_Assembler(fg.Generic(C, D=D), space).data.sum())
        out.append(fg.Bilinear_Assembler(fg.Generic(C[0], D=D), space).data.sum())
        if D == 1:
            out.append(fg.Bilinear_Assembler(fg.Advection(np.arange(1.0, dim + 1)), space).data.sum())
        u = np.linspace(0, 1, model.nnodes * D)
        out.append(fg.Get_Point_Values(u, space).sum())
        out.append(fg.Get_Point_Values(u, space, gradient=True).sum())
        for opt, val in (("imls_cluster", 0), ("imls_warp", 1), ("nb_pointwise", 1), ("assembly_pts", 0), ("assembly_direct", 1), ("cluster_from_masks", 0)):
            ctx.set_option(opt, val)
            sp2 = fg.build_space(recipe, [fg.IntegrationSet(tri, 2)])
            out.append(fg.Bilinear_Assembler(fg.GradGrad(1.0, D=D), sp2).data.sum())
        for opt, val in (("imls_cluster", 1), ("imls_warp", 0), ("nb_pointwise", 0), ("assembly_pts", 1), ("assembly_direct", 0), ("cluster_from_masks", 1)):
            ctx.set_option(opt, val)
        print(dim, D, tech, [float(f"{v:.10g}") for v in out])
    # node subsets + concatenation + boundary sets + resident solve
    fg.set_tag(model, lambda c: c[0] < 0.5, "A"); fg.set_tag(model, lambda c: c[0] >= 0.5, "B")
    tags = ["Left", "Right"]
    recipe = fg.ApproxSpace(model, tri, D=1, dmax=1.5, Dirichlet_Boundaries=[fg.Triangulation(model, t) for t in tags], Dirichlet_Values=[0.0, 1.0])
    both = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, "A"), 2), fg.IntegrationSet(fg.Triangulation(model, "B"), 2)])
    print("subsets", float(fg.Bilinear_Assembler(fg.Mass(1.0), both).data.sum()))
    dO = fg.IntegrationSet(tri, 2)
    A, b, Spaces = fg.Linear_Problem([(fg.GradGrad(1.0), dO)], [(fg.Source(1.0), dO)], recipe)
    x, it, res = fg.Solve_On_Device(Spaces, b, recipe, rtol=1e-10)
    print("solve", it, float(res), float(x.sum()))
    conn, _ = fg.geometry.get_entity_info(model, "Domain")
    c2 = recipe.cloud().ctx
    print("gauss", float(c2.generate_gauss(c2.new_set_id(), conn, fg.pgauss(2), fetch=True).sum()))


if __name__ == "__main__":
    run(3)
    run(2)
    print("sanitize case done")

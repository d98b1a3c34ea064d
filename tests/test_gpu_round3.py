"""GPU parity, part 3: the switches added with the persistent IMLS blocks, the index-build shortcuts and the fused sparsity / flush-table
pass.  Every variant must give the numbers of the default path (and the default path the oracle's: test_gpu_parity.py, test_gpu_round2.py)."""
import numpy as np
import pytest

from cases import jitter, poisson2d, poisson3d, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(fg, dim, jit):
    model, dmax, deg = (poisson2d(fg, 22, 1.75) if dim == 2 else poisson3d(fg, 9, 1.35))
    if jit:
        jitter(model, amount=0.08)
    dm = fg.Influence_Domains(model, dmax)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    return model, dm, gs


@pytest.mark.parametrize("dim", [2, 3])
def test_imls_work_distribution_does_not_change_the_numbers(fg, oracle, dim):
    """k_imls_p takes its (cluster, group) items from a device counter on persistent single-warp blocks ("imls_grid" -1, default), from a
    block stride (w > 0 blocks per resident slot) or one item per block (0).  The arithmetic of an item is the same in all three: PHI,
    DPHI and the DOM written by the kernel must agree bit for bit, and with the oracle."""
    model, dm, gs = _setup(fg, dim, True)
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    got = {}
    for grid in (-1, 0, 1, 3):
        cloud = fg.NodeCloud(model.x, dm)
        cloud.ctx.set_option("imls_grid", grid)
        S = fg.SHAPE_FUN(gs, cloud=cloud)
        assert np.array_equal(S.offsets, off) and np.array_equal(S.DOM, dom), grid
        assert rel(S.PHI, phiq) <= TOL and rel(S.DPHI, dphiq) <= TOL, grid
        got[grid] = (np.array(S.PHI), np.array(S.DPHI))
        cloud.close()
    for grid in (0, 1, 3):
        assert np.array_equal(got[grid][0], got[-1][0]) and np.array_equal(got[grid][1], got[-1][1]), grid


@pytest.mark.parametrize("jit", [False, True])
def test_flush_tables_from_the_sparsity_pass_equal_the_table_kernel(fg, oracle, jit):
    """After the set's own search fg_pattern derives the sparsity from the cluster index and, in the same pass, writes the flush tables
    of the symmetric assembly kernels ("pattern_tables" 1, default); with 0 k_cluster_tables builds them from the finished pattern.
    Same pattern (the oracle's), same K for the symmetric forms, and an unsymmetric form afterwards (advection: full tables, lidx on
    demand) still matches the oracle."""
    model, dm, gs = _setup(fg, 3, jit)
    ng = gs.shape[0]
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_gradgrad(3), 1, cp, rv, quad=True)
    adv = fg.Advection((1.0, -0.5, 0.25))
    Aq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_advection(3, np.array(adv.v)), 1, cp, rv, quad=True)
    got = {}
    for tables in (1, 0):
        cloud = fg.NodeCloud(model.x, dm)
        ctx, sid = cloud.ctx, cloud.ctx.new_set_id()
        ctx.set_option("pattern_tables", tables)
        ctx.set_gauss(sid, gs)
        ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        nnz = ctx.pattern(sid)
        colptr, rowval = ctx.get_pattern(sid, 1, nnz)
        assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv), tables
        K = np.empty(nnz); ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), K)
        M = np.empty(nnz); ctx.assemble_bilinear(sid, fg.Mass(2.0).c_struct(3, ng), M)
        A = np.empty(nnz); ctx.assemble_bilinear(sid, adv.c_struct(3, ng), A)
        K2 = np.empty(nnz); ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), K2)      # symmetric again, on the full tables
        assert rel(K, Kq) <= TOL and rel(A, Aq) <= TOL, tables
        assert np.abs(K2 - K).max() <= 1e-13 * np.abs(K).max()
        got[tables] = (K, M, A)
        cloud.close()
    for a, b in zip(got[1], got[0]):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


def test_local_indices_on_demand_equal_the_eager_ones(fg, oracle):
    """32-node blocks built from the search's hit masks leave lidx (local index of every DOM entry) to the first kernel that reads it
    ("cluster_lidx_eager" 0, default); the point-batched kernels never do.  Kernels that do -- the advection block kernel, the
    per-point linear form -- must give the same numbers as with the eager build, and the oracle's."""
    model, dm, gs = _setup(fg, 3, True)
    ng = gs.shape[0]
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    adv = fg.Advection((0.3, 1.0, -2.0))
    Aq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_advection(3, np.array(adv.v)), 1, cp, rv, quad=True)
    got = {}
    for eager in (0, 1):
        cloud = fg.NodeCloud(model.x, dm)
        ctx, sid = cloud.ctx, cloud.ctx.new_set_id()
        ctx.set_option("cluster_lidx_eager", eager)
        ctx.set_gauss(sid, gs)
        ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        nnz = ctx.pattern(sid)
        K = np.empty(nnz); ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), K)       # point-batched: no lidx
        A = np.empty(nnz); ctx.assemble_bilinear(sid, adv.c_struct(3, ng), A)                    # block kernel: lidx
        ctx.set_option("linear_pts", 0)
        F = np.empty(model.nnodes); ctx.assemble_linear(sid, fg.Source(2.0).c_struct(3, ng), F)   # k_linear_cluster: lidx
        assert rel(A, Aq) <= TOL, eager
        got[eager] = (K, A, F)
        cloud.close()
    for a, b in zip(got[0], got[1]):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()

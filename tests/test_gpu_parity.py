"""GPU parity: every result of the CUDA path, fetched through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  * neighbour lists, offsets, CSC sparsity: bit-exact against the oracle's literal restatement;
  * phi, dphi, K, M, f: <= 1e-12 relative (max-norm) against the __float128 evaluation of the
    reference formulas ("truth"), and inside 10x the reference-like FP64 restatement's own error
    ball |oracle64 - truth| (the reference's unshifted Gram-Schmidt is itself ~1e-11 off);
  * solution fields: <= 1e-10 against the oracle's own CPU path on the same inputs.
"""
import numpy as np
import pytest

from cases import jitter, poisson2d, poisson3d, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _shapes(fg, oracle, model, dmax, degree, shape="rectangular", tag="Domain", dm_map=None):
    recipe = fg.ApproxSpace(model, fg.Triangulation(model, "Domain"), D=1, dmax=dmax, shape=shape, dm_map=dm_map)
    iset = fg.IntegrationSet(fg.Triangulation(model, tag), degree)
    space = fg.build_space(recipe, [iset])
    S = space.merged()
    gs = space.Measures[0].gs
    dm = recipe.dm()
    off, dom, phi64, dphi64 = oracle.shape_fun(gs, model.x, dm, shape=shape, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, shape=shape, nthreads=8, quad=True)
    return recipe, space, S, gs, dm, (off, dom, phi64, dphi64, phiq, dphiq)


def _check_shapes(S, ref):
    off, dom, phi64, dphi64, phiq, dphiq = ref
    assert np.array_equal(S.offsets, off), "offsets differ"
    assert np.array_equal(S.DOM, dom), "neighbour lists differ"
    assert S.singular_count() == 0
    e_phi, e_dphi = rel(S.PHI, phiq), rel(S.DPHI, dphiq)
    assert e_phi <= TOL and e_dphi <= TOL, (e_phi, e_dphi)
    # inside the reference-like FP64 restatement's own error ball
    assert rel(S.PHI, phi64) <= 10 * rel(phi64, phiq) + TOL
    assert rel(S.DPHI, dphi64) <= 10 * rel(dphi64, dphiq) + TOL
    return e_phi, e_dphi


@pytest.mark.parametrize("case", ["poisson2d", "poisson3d"])
def test_shape_functions_as_shipped(fg, oracle, case):
    model, dmax, deg = (poisson2d if case == "poisson2d" else poisson3d)(fg)
    _, _, S, _, _, ref = _shapes(fg, oracle, model, dmax, deg)
    _check_shapes(S, ref)
    # ragged view = the reference's Vector{Vector} containers
    PHI, DPHI, DOM = S.ragged()
    assert len(PHI) == S.ngauss and all(len(p) == len(d) for p, d in zip(PHI[:50], DOM[:50]))


@pytest.mark.parametrize("dim", [2, 3])
def test_shape_functions_irregular_nodes(fg, oracle, dim):
    model, dmax, deg = (poisson2d(fg, 24, 1.75) if dim == 2 else poisson3d(fg, 8, 1.6))
    jitter(model)
    _, _, S, _, _, ref = _shapes(fg, oracle, model, dmax, deg)
    _check_shapes(S, ref)


@pytest.mark.parametrize("dim", [2, 3])
def test_shape_functions_circular(fg, oracle, dim):
    model, _, deg = (poisson2d(fg, 20) if dim == 2 else poisson3d(fg, 7))
    _, _, S, _, _, ref = _shapes(fg, oracle, model, 2.1, deg, shape="circular")
    _check_shapes(S, ref)


def test_shape_functions_dm_map_and_anisotropic(fg, oracle):
    """Non-uniform supports (dm_map, FlexiGal.jl:173-187) and a dmax vector (Poisson3D_No_Source.jl:4)."""
    model, _, deg = poisson2d(fg, 20)
    _, _, S, _, _, ref = _shapes(fg, oracle, model, [1.5, 2.2], deg, dm_map=lambda p: 1.0 + 0.5 * p[0])
    _check_shapes(S, ref)


def test_boundary_set_searches_all_nodes(fg, oracle):
    model, dmax, deg = poisson3d(fg, 6)
    _, _, S, _, _, ref = _shapes(fg, oracle, model, dmax, deg, tag="Left")
    _check_shapes(S, ref)


def test_mls_invariants_partition_of_unity_and_reproduction(fg, oracle):
    """Any correct MLS with this basis reproduces every basis monomial."""
    model, dmax, deg = poisson3d(fg, 6)
    jitter(model)
    _, _, S, gs, _, _ = _shapes(fg, oracle, model, dmax, deg)
    off, dom = S.offsets, S.DOM - 1
    seg = off[:-1]
    X = model.x[dom]
    g = gs[:, :3]
    assert np.abs(np.add.reduceat(S.PHI, seg) - 1).max() < 1e-13               # partition of unity
    assert np.abs(np.add.reduceat(S.DPHI, seg, axis=0)).max() < 1e-10           # sum grad phi = 0
    monos = [(X[:, 0], g[:, 0]), (X[:, 1] * X[:, 2], g[:, 1] * g[:, 2]), (X[:, 0] * X[:, 1] * X[:, 2], g[:, 0] * g[:, 1] * g[:, 2])]
    for at_nodes, at_gauss in monos:
        assert np.abs(np.add.reduceat(S.PHI * at_nodes, seg) - at_gauss).max() < 1e-13
    dx = np.add.reduceat(S.DPHI * X[:, :1], seg, axis=0)                         # grad of the monomial x
    assert np.abs(dx - np.array([1.0, 0.0, 0.0])).max() < 1e-10


def _assembled(fg, oracle, model, dmax, deg, D, op_gpu, coef, shape="rectangular"):
    recipe = fg.ApproxSpace(model, fg.Triangulation(model, "Domain"), D=D, dmax=dmax, shape=shape)
    space = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)])
    S = space.merged()
    gs = space.Measures[0].gs
    off, dom = S.offsets, S.DOM
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, recipe.dm(), off, dom, shape=shape, nthreads=8, quad=True)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    K = fg.Bilinear_Assembler(op_gpu, space)
    cpd, rvd = oracle.dof_csc(cp, rv, D)
    assert np.array_equal(K.indptr + 1, cpd) and np.array_equal(K.indices + 1, rvd), "sparsity differs"
    Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, coef(gs), D, cp, rv, quad=True)
    return K, Kq, space, (gs, off, dom, phiq, dphiq, cp, rv)


@pytest.mark.parametrize("case", ["poisson2d", "poisson3d"])
def test_poisson_operators_and_solution(fg, oracle, case):
    """C1 / C2 end to end: K, M, f, penalty Dirichlet, solution field (Poisson*_Source.jl)."""
    model, dmax, deg = (poisson2d if case == "poisson2d" else poisson3d)(fg)
    dim = model.dim
    K, Kq, space, (gs, off, dom, phiq, dphiq, cp, rv) = _assembled(fg, oracle, model, dmax, deg, 1, fg.GradGrad(1.0),
                                                                   lambda gs: oracle.coef_gradgrad(dim))
    assert rel(K.data, Kq) <= TOL
    M = fg.Bilinear_Assembler(fg.Mass(2.5), space)
    Mq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_mass(dim, 2.5), 1, cp, rv, quad=True)
    assert rel(M.data, Mq) <= TOL
    src = 50.0 if dim == 2 else 5000.0
    F = fg.Linear_Assembler(fg.Source(src), space)
    Fq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, oracle.coef_source(dim, src), 1, model.nnodes, quad=True)
    assert rel(F, Fq) <= TOL
    # invariants: sum_ab K_ab = 0, sum_ab M_ab = c * volume, K symmetric
    assert abs(K.sum()) <= 1e-9 * np.abs(K.data).max() * K.shape[0]
    assert abs(M.sum() - 2.5) <= 1e-12
    assert abs(K - K.T).max() <= 1e-12 * np.abs(K.data).max()
    # full Linear_Problem with penalty Dirichlet + solve, against the oracle's own CPU path
    tags = ["Left", "Right", "Top", "Bottom"] if dim == 2 else ["Left", "Right", "Top", "Bottom", "Back", "Front"]
    tri = fg.Triangulation(model, "Domain")
    recipe = fg.ApproxSpace(model, tri, D=1, dmax=dmax, Dirichlet_Boundaries=[fg.Triangulation(model, t) for t in tags],
                            Dirichlet_Values=[0.0] * len(tags))
    dO = fg.IntegrationSet(tri, deg)
    A, b, _ = fg.Linear_Problem([(fg.GradGrad(1.0), dO)], [(fg.Source(src), dO)], recipe)
    u = fg.Solve(A, b)
    omodel = oracle.create_model(model.domain, model.divisions)
    Ao, bo, _ = oracle.linear_problem_scalar(omodel, dmax, deg, oracle.coef_gradgrad(dim), oracle.coef_source(dim, src), tags, [0.0] * len(tags))
    uo = fg.Solve(Ao, bo)
    # (A's stored pattern after the sparse `+` depends on which tie-node entries are EXACTLY zero, i.e. on
    # last-bit rounding of the weights; the structural pattern is compared bit-exactly in _assembled above)
    assert A.shape == Ao.shape and abs(A - Ao).max() <= 1e-10 * abs(Ao).max()   # the FP64 restatement own error ball
    assert rel(u, uo) <= 1e-10


@pytest.mark.parametrize("dim", [2, 3])
def test_elasticity_blocks(fg, oracle, dim):
    """C3: grad(du) (.) sigma(u) with E = 1000, nu = 0.3 (Elasticity_Test.jl:7-10,25-26), D = dim DOF per node."""
    E, nu = 1000.0, 0.3
    G, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    model = fg.create_model((5.0, 1.0), (20, 4)) if dim == 2 else fg.create_model((5.0, 1.0, 1.0), (10, 3, 3))
    K, Kq, *_ = _assembled(fg, oracle, model, 1.75 if dim == 2 else 1.5, 3, dim, fg.Elasticity(G, lam, dim),
                           lambda gs: oracle.coef_elasticity(dim, G, lam))
    assert rel(K.data, Kq) <= TOL
    assert abs(K - K.T).max() <= 1e-12 * np.abs(K.data).max()
    # rigid translations are in the null space
    for i in range(dim):
        t = np.zeros(K.shape[0]); t[i::dim] = 1.0
        assert np.abs(K @ t).max() <= 1e-9 * np.abs(K.data).max()


def test_vector_gradgrad_and_mass_keep_structural_zeros(fg, oracle):
    model = fg.create_model((1.0, 1.0), (10, 10))
    K, Kq, space, (gs, off, dom, phiq, dphiq, cp, rv) = _assembled(fg, oracle, model, 1.5, 2, 2, fg.GradGrad(3.0, D=2),
                                                                   lambda gs: oracle.coef_gradgrad(2, 3.0, 2))
    assert rel(K.data, Kq) <= TOL
    assert (K.data == 0).sum() >= K.nnz // 2        # off-diagonal block entries are explicit zeros, like sparse()
    M = fg.Bilinear_Assembler(fg.Mass(7.0, D=2), space)
    Mq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_mass(2, 7.0, 2), 2, cp, rv, quad=True)
    assert rel(M.data, Mq) <= TOL


def test_advection_and_generic_tensors(fg, oracle):
    """Non-symmetric closure dT*(v.grad T) (Poisson2D_No_Source.jl:30), as a closed form, as a constant
    tensor, as a per-Gauss-point tensor (the SVK / Navier-Stokes re-assembly path) and as a probed closure."""
    model = fg.create_model((1.0, 1.0), (12, 12))
    v = np.array([0.7, -1.3])
    K, Kq, space, (gs, off, dom, phiq, dphiq, cp, rv) = _assembled(fg, oracle, model, 1.5, 3, 1, fg.Advection(v),
                                                                   lambda gs: oracle.coef_advection(2, v))
    assert rel(K.data, Kq) <= TOL
    K2 = fg.Bilinear_Assembler(fg.Generic(oracle.coef_advection(2, v)), space)
    assert rel(K2.data, Kq) <= TOL
    K3 = fg.Bilinear_Assembler(lambda sa, sb: sa.phi * float(v @ sb.dphi), space)
    assert rel(K3.data, Kq) <= TOL
    vg = np.stack([1.0 + gs[:, 0], gs[:, 1] ** 2], axis=1)
    Cg = oracle.coef_advection(2, vg)
    K4 = fg.Bilinear_Assembler(fg.Generic(Cg), space)
    K4q = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, Cg, 1, cp, rv, quad=True)
    assert rel(K4.data, K4q) <= TOL
    # per-point linear form
    c = np.sin(gs[:, 0]) + 2.0
    F = fg.Linear_Assembler(fg.Source(c), space)
    Fq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, oracle.coef_source(2, c), 1, model.nnodes, quad=True)
    assert rel(F, Fq) <= TOL


def test_cluster_and_direct_assembly_paths_agree(fg, oracle):
    """The shared-memory cluster path and the direct global-atomics path (fg_set_option "assembly_direct") give the same K;
    supports too wide for a cluster (dmax 2.6 in 3D: 125+ neighbours) fall back to the direct path on their own."""
    model, dmax, deg = poisson3d(fg, 6)
    K, Kq, space, _ = _assembled(fg, oracle, model, dmax, deg, 1, fg.GradGrad(1.0), lambda gs: oracle.coef_gradgrad(3))
    st = space.cloud.ctx.assembly_stats(space.merged().set_id)
    assert st["clusters"] > 0 and st["direct_points"] == 0 and st["max_union"] <= 128
    space.cloud.ctx.set_option("assembly_direct", 1)
    Kd = fg.Bilinear_Assembler(fg.GradGrad(1.0), space)
    space.cloud.ctx.set_option("assembly_direct", 0)
    assert rel(Kd.data, Kq) <= TOL and rel(K.data, Kq) <= TOL
    model = fg.create_model((1.0, 1.0, 1.0), (6, 6, 6))
    Kw, Kwq, spacew, _ = _assembled(fg, oracle, model, 2.6, 2, 1, fg.Mass(1.0), lambda gs: oracle.coef_mass(3))
    stw = spacew.cloud.ctx.assembly_stats(spacew.merged().set_id)
    assert stw["direct_points"] > 0
    assert rel(Kw.data, Kwq) <= TOL


def test_empty_and_tiny_sets(fg, oracle):
    model = fg.create_model((1.0, 1.0), (4, 4))
    dm = fg.Influence_Domains(model, 1.5)
    cloud = fg.NodeCloud(model.x, dm)
    S0 = fg.SHAPE_FUN(np.zeros((0, 4)), cloud=cloud)
    assert S0.ngauss == 0 and S0.total_L == 0 and S0.offsets.tolist() == [0]
    # one point far outside every support: L = 0, like findall on an all-false mask
    S1 = fg.SHAPE_FUN(np.array([[10.0, 10.0, 1.0, 1.0], [0.5, 0.5, 1.0, 1.0]]), cloud=cloud)
    assert S1.offsets[1] == 0 and S1.offsets[2] > 0
    ids = oracle.domain(np.array([0.5, 0.5]), model.x, dm)
    assert np.array_equal(S1.DOM, ids)


def test_error_behaviour_matches_reference(fg):
    model = fg.create_model((1.0, 1.0), (4, 4))
    dm = fg.Influence_Domains(model, 1.5)
    with pytest.raises(ValueError, match="Unknown shape of influence domain"):
        fg.NodeCloud(model.x, dm, shape="triangular")
    cloud = fg.NodeCloud(model.x, dm)
    with pytest.raises(ValueError, match="Unknown technique"):
        fg.SHAPE_FUN(np.array([[0.5, 0.5, 1.0, 1.0]]), cloud=cloud, technique="MLS")


def test_slab_sharding_matches_single_domain_on_the_gpu(fg, oracle):
    """two z-slabs assembled by two contexts (own Gauss points, sparsity from own + halo Gauss
    positions via fg_pattern_from), halo column slices added to their owner -- equals the single-domain K.
    (Both 'ranks' run on cuda:0 here and the exchange is a host copy; the NCCL path is bench.py --gpus N.)"""
    from flexigal_b200.distributed import SlabPartition
    from flexigal_b200 import _lib
    domain, div, dmax, deg, world = (1.0, 1.0, 2.0), (5, 5, 10), 1.35, 3, 2
    locals_ = []
    for r in range(world):
        part = SlabPartition(domain, div, dmax, world, r, degree=deg)
        cloud = fg.NodeCloud(part.local_nodes(), part.local_dm())
        ctx = cloud.ctx
        sid = ctx.new_set_id()
        ctx.set_gauss(sid, part.local_gauss())
        ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        nnz = part.build_pattern(ctx, sid, part.pattern_set(ctx, sid))
        colptr, rowval = ctx.get_pattern(sid, 1, nnz)
        vals = np.empty(nnz)
        ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, 0), vals)
        locals_.append((part, colptr - 1, rowval - 1, vals))
    # the exchange plan, executed on the host
    merged = [v.copy() for (_, _, _, v) in locals_]
    for r, (part, cp, rv, vals) in enumerate(locals_):
        for peer, send_l, recv_l in part.plan():
            ppart, pcp, prv, _ = locals_[peer]
            a, b = part.local_node_range(*send_l)
            pa, pb = ppart.local_node_range(*send_l)
            assert np.array_equal(rv[cp[a]:cp[b]] + part.node_offset, prv[pcp[pa]:pcp[pb]] + ppart.node_offset), "shared columns differ"
            merged[peer][pcp[pa]:pcp[pb]] += vals[cp[a]:cp[b]]
    model = fg.create_model(domain, div)
    K, Kq, *_ = _assembled(fg, oracle, model, dmax, deg, 1, fg.GradGrad(1.0), lambda gs: oracle.coef_gradgrad(3))
    Kc = K.tocsc()
    for r, (part, cp, rv, _) in enumerate(locals_):
        lo, hi = part.owned_layers()
        a, b = part.local_node_range(lo, hi)
        for jl in range(a, b):
            jg = jl + part.node_offset
            assert np.array_equal(rv[cp[jl]:cp[jl + 1]] + part.node_offset, Kc.indices[Kc.indptr[jg]:Kc.indptr[jg + 1]])
        ga, gb = Kc.indptr[a + part.node_offset], Kc.indptr[b + part.node_offset]
        assert rel(merged[r][cp[a]:cp[b]], Kq[ga:gb]) <= TOL


@pytest.mark.parametrize("case", ["elasticity2d", "poisson2d", "poisson3d", "circular2d"])
def test_moving_kriging_shape_functions(fg, oracle, case):
    """technique=:MK (Shape_Functions.jl:164-239), the technique C3/C4/C5 select as shipped.  The Gaussian
    correlation matrix is ill-conditioned (entries 0.7..1), so the bar is the FP64 restatement's own error
    ball around the __float128 truth, plus the interpolation invariants."""
    shape = "rectangular"
    if case == "elasticity2d":
        model, dmax, deg = fg.create_model((5.0, 1.0), (40, 8)), 1.75, 3         # Elasticity_Test.jl:4-5,22 (coarser)
    elif case == "poisson2d":
        model, dmax, deg = fg.create_model((1.0, 1.0), (16, 16)), 1.5, 3
    elif case == "poisson3d":
        model, dmax, deg = fg.create_model((1.0, 1.0, 1.0), (5, 5, 5)), 1.35, 2
    else:
        model, dmax, deg, shape = fg.create_model((1.0, 1.0), (12, 12)), 2.1, 2, "circular"
    dm = fg.Influence_Domains(model, dmax, shape)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    S = fg.SHAPE_FUN(gs, model.x, dm, shape=shape, technique="MK")
    off, dom, phi64, dphi64 = oracle.shape_fun(gs, model.x, dm, shape=shape, technique="MK", nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, shape=shape, technique="MK", nthreads=8, quad=True)
    assert np.array_equal(S.offsets, off) and np.array_equal(S.DOM, dom) and S.singular_count() == 0
    assert rel(S.PHI, phiq) <= 10 * rel(phi64, phiq) + 1e-12, (rel(S.PHI, phiq), rel(phi64, phiq))
    assert rel(S.DPHI, dphiq) <= 10 * rel(dphi64, dphiq) + 1e-12, (rel(S.DPHI, dphiq), rel(dphi64, dphiq))
    seg = off[:-1]
    assert np.abs(np.add.reduceat(S.PHI, seg) - 1).max() < 1e-7
    X = model.x[dom - 1]
    assert np.abs(np.add.reduceat(S.PHI * X[:, 0] * X[:, 1], seg) - gs[:, 0] * gs[:, 1]).max() < 1e-7


def test_recomputed_neighbours_use_the_single_pass_kernel_and_agree(fg, oracle):
    """With the option "nb_fused", second and later fg_neighbours calls on a set take the single-pass (decoupled look-back)
    kernel; a new Gauss table with longer lists than the capacity hint falls back to the two-pass kernel.  Same DOM bits."""
    model, dmax, deg = poisson3d(fg, 7)
    jitter(model)
    dm = fg.Influence_Domains(model, dmax)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    cloud = fg.NodeCloud(model.x, dm)
    ctx = cloud.ctx
    ctx.set_option("nb_fused", 1)
    ctx.set_option("nb_pointwise", 1)         # the single-pass kernel belongs to the point-wise search
    sid = ctx.new_set_id()
    ng = ctx.set_gauss(sid, gs)
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    for _ in range(3):
        tot, mx = ctx.neighbours(sid)
        o, d, _, _ = ctx.get_shapes(sid, ng, tot, want_values=False)
        assert tot == off[-1] and mx == np.diff(off).max() and np.array_equal(o, off) and np.array_equal(d, dom)
    # same set id, a different table: first a sparse one (short lists), then the dense one again (longer than the hint)
    few = np.asfortranarray(gs[::50])
    ng2 = ctx.set_gauss(sid, few)
    tot2, _ = ctx.neighbours(sid)
    o2, d2, _, _ = ctx.get_shapes(sid, ng2, tot2, want_values=False)
    sel = np.concatenate([np.arange(off[g], off[g + 1]) for g in range(0, gs.shape[0], 50)])
    assert np.array_equal(d2, dom[sel])
    ctx.set_gauss(sid, gs)
    tot3, _ = ctx.neighbours(sid)
    o3, d3, _, _ = ctx.get_shapes(sid, ng, tot3, want_values=False)
    assert np.array_equal(o3, off) and np.array_equal(d3, dom)


@pytest.mark.parametrize("D", [1, 2])
def test_field_evaluation_at_gauss_points(fg, oracle, D):
    """Get_Point_Values of a nodal field and of its gradient (Fields_Operations.jl:40-112)."""
    model = fg.create_model((1.0, 1.0), (10, 8))
    recipe = fg.ApproxSpace(model, fg.Triangulation(model, "Domain"), D=D, dmax=1.6)
    space = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, "Domain"), 2)])
    S = space.merged()
    gs = space.Measures[0].gs
    rng = np.random.default_rng(3)
    u = rng.standard_normal(model.nnodes * D)
    val = fg.Get_Point_Values(u, space)
    grad = fg.Get_Point_Values(u, space, gradient=True)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, recipe.dm(), S.offsets, S.DOM, quad=True)
    assert rel(val, oracle.point_values(u, S.offsets, S.DOM, phiq, D)) <= 1e-12
    assert rel(grad, oracle.point_gradients(u, S.offsets, S.DOM, dphiq, D)) <= 1e-12
    # a field in the span of the basis is reproduced exactly: u = 2 + 3x - y + 0.5xy
    X = model.x
    ulin = np.repeat((2 + 3 * X[:, 0] - X[:, 1] + 0.5 * X[:, 0] * X[:, 1])[:, None], D, axis=1).reshape(-1)
    v2 = fg.Get_Point_Values(ulin, space)
    g2 = fg.Get_Point_Values(ulin, space, gradient=True)
    assert np.abs(v2[:, 0] - (2 + 3 * gs[:, 0] - gs[:, 1] + 0.5 * gs[:, 0] * gs[:, 1])).max() < 1e-12
    assert np.abs(g2[:, 0, 0] - (3 + 0.5 * gs[:, 1])).max() < 1e-10 and np.abs(g2[:, 1, 0] - (-1 + 0.5 * gs[:, 0])).max() < 1e-10


def test_device_cg_solve_matches_direct_solve(fg, oracle):
    """u = A \\ F on the device for the penalty-Dirichlet Poisson system (Solvers.jl:24-31)."""
    model, dmax, deg = poisson2d(fg, 24)
    tags = ["Left", "Right", "Top", "Bottom"]
    tri = fg.Triangulation(model, "Domain")
    recipe = fg.ApproxSpace(model, tri, D=1, dmax=dmax, Dirichlet_Boundaries=[fg.Triangulation(model, t) for t in tags],
                            Dirichlet_Values=[0.0] * 4)
    dO = fg.IntegrationSet(tri, deg)
    A, b, _ = fg.Linear_Problem([(fg.GradGrad(1.0), dO)], [(fg.Source(50.0), dO)], recipe)
    u_direct = fg.Solve(A, b)
    u_cg = fg.Solve(A, b, cloud=recipe.cloud())
    assert rel(u_cg, u_direct) <= 1e-10
    x, it, res = recipe.cloud().ctx.solve_spd(A, b, rtol=1e-13)
    assert res <= 1e-12 and it < 5000


def test_gauss_tables_generated_on_the_device_are_bit_identical(fg):
    """egauss / egauss_bound (Integration.jl:148-218) on the device == the host table, bit for bit,
    for cells and boundary entities in 2D and 3D, on a mapped (non-rectangular) mesh too."""
    for dom, div, mp in [((2.0, 1.0), (7, 5), None), ((1.0, 2.0, 1.5), (5, 4, 3), None),
                         ((1.0, 1.0), (6, 6), lambda p: (p[0] + 0.1 * p[1] ** 2, p[1] + 0.05 * np.sin(3 * p[0]))),
                         ((1.0, 1.0, 1.0), (4, 3, 3), lambda p: (p[0] * (1 + 0.2 * p[1]), p[1] + 0.1 * p[2] * p[0], p[2] + 0.05 * p[0] ** 2))]:
        model = fg.create_model(dom, div, map=mp)
        cloud = fg.NodeCloud(model.x, fg.Influence_Domains(model, 1.5))
        for deg in (1, 2, 3, 4):
            for tag in ["Domain", "Left", "Top", "Boundary"]:
                conn, _ = fg.geometry.get_entity_info(model, tag)
                host = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, tag), deg)).gs
                dev = cloud.ctx.generate_gauss(cloud.ctx.new_set_id(), conn, fg.pgauss(deg), fetch=True)
                assert np.array_equal(host, dev), (dom, deg, tag, np.abs(host - dev).max())
    # the generated set feeds the rest of the path like an uploaded one
    sid = cloud.ctx.new_set_id()
    conn, _ = fg.geometry.get_entity_info(model, "Domain")
    ng = cloud.ctx.generate_gauss(sid, conn, fg.pgauss(2))
    tot, _ = cloud.ctx.neighbours(sid)
    S2 = fg.SHAPE_FUN(fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), 2)).gs, cloud=cloud)
    assert ng == S2.ngauss and tot == S2.total_L


def test_newton_like_reassembly_with_per_point_tensors(fg, oracle):
    """C4 flavour (SaintVenant_Kirchoff_Test.jl:28-34, Solvers.jl:184-189): every iteration evaluates grad u_h at the
    Gauss points (f2), builds a per-Gauss-point coefficient tensor from it and re-assembles K and the residual on the
    SAME sparsity (pattern and cluster index are built once).  Model problem: -div((1 + |grad u|^2) grad u) = 1."""
    model = fg.create_model((1.0, 1.0), (10, 10))
    recipe = fg.ApproxSpace(model, fg.Triangulation(model, "Domain"), D=1, dmax=1.5)
    space = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, "Domain"), 3)])
    S = space.merged()
    gs = space.Measures[0].gs
    off, dom = S.offsets, S.DOM
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, recipe.dm(), off, dom, quad=True)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    colptr0, rowval0 = S.pattern(1)
    u = 0.1 * np.sin(np.pi * model.x[:, 0]) * model.x[:, 1]
    ctx = space.cloud.ctx
    launches = []
    for it in range(3):
        g = fg.Get_Point_Values(u, space, gradient=True)[:, :, 0]            # (ngauss, 2)
        k = 1.0 + (g ** 2).sum(axis=1)
        # tangent: (1+|g|^2) grad dT . grad T + 2 (g . grad dT)(g . grad T)
        C = np.zeros((gs.shape[0], 3, 3))
        C[:, 1:, 1:] = k[:, None, None] * np.eye(2)[None] + 2.0 * g[:, :, None] * g[:, None, :]
        before = ctx.launch_count()
        K = fg.Bilinear_Assembler(fg.Generic(C), space)
        R = fg.Linear_Assembler(fg.GenericLinear(np.concatenate([np.ones((gs.shape[0], 1)), -(k[:, None] * g)], axis=1)), space)
        launches.append(ctx.launch_count() - before)
        Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, C.reshape(gs.shape[0], -1), 1, cp, rv, quad=True)
        gq = oracle.point_gradients(u, off, dom, dphiq, 1)[:, :, 0]
        assert rel(g, gq) <= 1e-12
        assert np.array_equal(K.indptr + 1, cp) and np.array_equal(K.indices + 1, rv)
        assert rel(K.data, Kq) <= 1e-12
        Rq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, np.concatenate([np.ones((gs.shape[0], 1)), -(k[:, None] * g)], axis=1), 1, model.nnodes, quad=True)
        assert rel(R, Rq) <= 1e-12
        u = u + 0.05 * R / np.abs(R).max()                                    # any update; the point is the re-assembly
    assert launches[1] == launches[2] and launches[1] < launches[0]           # iterations 2+ reuse pattern, clusters and tables
    assert np.array_equal(S.pattern(1)[0], colptr0)


def test_repeated_assembly_with_neighbour_lists_longer_than_a_warp(fg, oracle):
    """Supports of 3 spacings in 2D give 36 neighbours per interior Gauss point: the cluster kernels hand such points
    to the direct kernel through a per-call list.  Re-assembling on the same set (Newton iterations, Solvers.jl:184-189)
    must give the same matrices every time (the list starts from the index build's own entries on every call), and
    the linear forms -- accumulated per cluster once the cluster index exists -- the same vector."""
    model = fg.create_model((1.0, 1.0), (12, 12))
    recipe, space, S, gs, dm, ref = _shapes(fg, oracle, model, 3.0, 3)
    off, dom, phi64, dphi64, phiq, dphiq = ref
    assert int(np.diff(off).max()) > 32
    _check_shapes(S, ref)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_gradgrad(2), 1, cp, rv, quad=True)
    Mq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_mass(2, 1.5), 1, cp, rv, quad=True)
    Fq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, oracle.coef_source(2, 7.0), 1, model.nnodes, quad=True)
    F0 = fg.Linear_Assembler(fg.Source(7.0), space)               # before any bilinear form: point-wise kernel
    assert rel(F0, Fq) <= TOL
    direct = []
    for _ in range(3):
        K = fg.Bilinear_Assembler(fg.GradGrad(1.0), space)
        direct.append(space.cloud.ctx.assembly_stats(S.set_id)["direct_points"])
        assert np.array_equal(K.indptr + 1, cp) and np.array_equal(K.indices + 1, rv)
        assert rel(K.data, Kq) <= TOL
        M = fg.Bilinear_Assembler(fg.Mass(1.5), space)
        assert rel(M.data, Mq) <= TOL
        F = fg.Linear_Assembler(fg.Source(7.0), space)            # cluster index exists now: per-cluster kernel
        assert rel(F, Fq) <= TOL
    # the long-list points go to the direct kernel: the same ones, once, on every call
    assert 0 < direct[0] <= gs.shape[0] and direct[1] == direct[0] and direct[2] == direct[0]


def test_benchmark_workload_properties_at_full_size(fg):
    """BASELINE.json's configuration (Divisions 100^3, dmax 1.35, degree 3: 1 030 301 nodes, 2.7e7 Gauss points), far
    beyond what the oracle finishes in seconds: size-independent properties instead.  On the tensor grid the counts
    factorise (sum L = s^3, nnz = p^3 with 1-D counts s, p computed here by brute force); IMLS reproduces the trilinear
    basis exactly, so fields interpolate exactly, K annihilates constants and u^T K u = integral |grad u|^2 in closed form."""
    import scipy.sparse as sp
    from flexigal_b200.distributed import SlabPartition
    n, dmax = 100, 1.35
    part = SlabPartition((1.0, 1.0, 1.0), (n, n, n), dmax, 1, 0, degree=3)
    x, dm, gs = part.local_nodes(), part.local_dm(), part.local_gauss()
    nn, ng = x.shape[0], gs.shape[0]
    assert (nn, ng) == ((n + 1) ** 3, 27 * n ** 3)
    # 1-D brute force with the reference's predicate |x_a - g| <= dm
    # (membership has a margin of 0.13 spacings here, so the clean 1-D coordinates decide it like the table's own)
    from flexigal_b200 import geometry as G
    h = 1.0 / n
    nodes1 = h * np.arange(n + 1)
    g1 = (np.arange(n)[:, None] + 0.5 * (1.0 + np.unique(G.pgauss(3)[0]))[None, :]).reshape(-1) * h
    assert g1.size == 3 * n
    inside = np.abs(nodes1[None, :] - g1[:, None]) <= dm[0, 0]
    s1 = int(inside.sum())
    p1 = int(((inside.T.astype(np.int64) @ inside.astype(np.int64)) > 0).sum())
    cloud = fg.NodeCloud(x, dm, "rectangular", device=0)
    ctx = cloud.ctx
    sid = ctx.new_set_id()
    ctx.set_gauss(sid, gs)
    total_L, max_L = ctx.neighbours(sid)
    assert total_L == s1 ** 3 and max_L == 27
    ctx.shape_functions(sid, "IMLS")
    assert ctx.singular_count(sid) == 0
    # reproduction: constants and the trilinear monomial, values and gradients, at every Gauss point
    val, grad = ctx.evaluate_field(sid, ng, np.ones(nn))
    assert np.abs(val[:, 0] - 1.0).max() < 1e-13 and np.abs(grad).max() < 1e-10
    u3 = x[:, 0] * x[:, 1] * x[:, 2]
    val, grad = ctx.evaluate_field(sid, ng, u3)
    gx, gy, gz = gs[:, 0], gs[:, 1], gs[:, 2]
    assert np.abs(val[:, 0] - gx * gy * gz).max() < 1e-13
    assert np.abs(grad[:, :, 0] - np.stack([gy * gz, gx * gz, gx * gy], axis=1)).max() < 1e-10
    del val, grad
    # sparsity and operators
    nnz = ctx.pattern(sid)
    assert nnz == p1 ** 3
    colptr, rowval = ctx.get_pattern(sid, 1, nnz)
    assert colptr[0] == 1 and colptr[-1] == nnz + 1 and rowval.min() == 1 and rowval.max() == nn
    indptr, indices = (colptr - 1), (rowval - 1).astype(np.int32)
    del rowval
    vals = np.empty(nnz)
    ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), vals)
    K = sp.csc_matrix((vals.copy(), indices, indptr), shape=(nn, nn))
    scale = np.abs(K.data).max()
    assert np.abs(K @ np.ones(nn)).max() <= 1e-10 * scale                      # constants are in the kernel
    ux = x[:, 0].copy()
    assert abs(ux @ (K @ ux) - 1.0) <= 1e-10                                   # integral |grad x|^2 = volume
    assert abs(u3 @ (K @ u3) - 1.0 / 3.0) <= 1e-10                             # integral |grad xyz|^2 = 3 (1/3)^2
    r = np.random.default_rng(0).standard_normal((2, nn))
    assert abs(r[0] @ (K @ r[1]) - r[1] @ (K @ r[0])) <= 1e-9 * scale * nn ** 0.5   # symmetric
    ctx.assemble_bilinear(sid, fg.Mass(2.5).c_struct(3, ng), vals)
    M = sp.csc_matrix((vals, indices, indptr), shape=(nn, nn))
    assert abs(M.sum() - 2.5) <= 1e-11                                         # sum_ab M_ab = c * volume
    assert abs(ux @ (M @ ux) - 2.5 / 3.0) <= 1e-11                             # c * integral x^2
    F = np.empty(nn)
    ctx.assemble_linear(sid, fg.Source(5000.0).c_struct(3, ng), F)
    assert abs(F.sum() - 5000.0) <= 1e-8 and abs(F @ ux - 2500.0) <= 1e-8      # integral f, integral f x
    stats = ctx.assembly_stats(sid)
    assert stats["clusters"] > 0 and stats["direct_points"] == 0 and stats["max_union"] <= 32
    cloud.close()


def test_irregular_cloud_assembly_falls_back_to_larger_blocks(fg, oracle):
    """Jittered 3D cloud: the node-centred 32-node blocks of the structured case overflow for most clusters;
    the index is rebuilt with 64-node blocks instead of sending those points through the global-atomics kernel.
    Parity is the same either way."""
    model = jitter(fg.create_model((1.0, 1.0, 1.0), (7, 7, 7)), amount=0.3)
    K, Kq, space, (gs, off, dom, phiq, dphiq, cp, rv) = _assembled(fg, oracle, model, 1.35, 3, 1, fg.GradGrad(1.0),
                                                                   lambda gs: oracle.coef_gradgrad(3))
    assert rel(K.data, Kq) <= TOL
    stats = space.cloud.ctx.assembly_stats(space.merged().set_id)
    assert stats["clusters"] > 0 and stats["direct_points"] * 20 <= gs.shape[0] or stats["max_union"] > 32
    F = fg.Linear_Assembler(fg.Source(3.0), space)
    Fq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, oracle.coef_source(3, 3.0), 1, model.nnodes, quad=True)
    assert rel(F, Fq) <= TOL

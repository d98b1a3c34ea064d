"""GPU parity, part 2: the headline workload pinned against the reference algorithm on a stratified sample, sub-domain
tags, the Lid_Cavity-shaped combination (MK + map + dm_map + reduced set + fixed penalty), the call sequence of
Linear_Problem replayed symbol by symbol through ctypes, and the solve with K resident on the device.

Same bars as test_gpu_parity.py: lists / sparsity bit-exact, values <= 1e-12 relative against the __float128 evaluation of the
reference formulas, solution fields <= 1e-10 against the oracle's own CPU path."""
import ctypes

import numpy as np
import pytest

from cases import jitter, poisson2d, poisson3d, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_headline_workload_sample_against_the_reference_algorithm(fg, oracle):
    """BASELINE.json's configuration (Divisions 100^3, dmax 1.35, degree 3): the kernels run the WHOLE workload (2.7e7 Gauss
    points, 1 030 301 nodes); 6 blocks of 5^3 cells (corner, edge, face, centre, far corner, far edge: 20 250 Gauss points)
    are compared with the reference's brute-force domain() over all 10^6 nodes (bit-exact lists), with the __float128 IMLS
    (phi, grad phi <= 1e-12) and, for the 216 nodes whose supports lie inside a block, with the reference's columns of K."""
    import headline_check as hc
    from flexigal_b200.distributed import SlabPartition
    n = 100
    part = SlabPartition((1.0, 1.0, 1.0), (n, n, n), 1.35, 1, 0, degree=3)
    x, dm, gs = part.local_nodes(), part.local_dm(), part.local_gauss()
    cloud = fg.NodeCloud(x, dm, "rectangular", device=0)
    ctx = cloud.ctx
    sid = ctx.new_set_id()
    ctx.set_gauss(sid, gs)
    ctx.neighbours(sid)
    ctx.shape_functions(sid, "IMLS")
    ctx.pattern(sid)
    ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, gs.shape[0]), None)
    rep = hc.check_against_device(ctx, sid, n, x, dm, gs, hc.block_origins(n, 5), B=5, nthreads=8, with_K=True, tol=TOL)
    print("headline sample:", {k: v for k, v in rep.items() if k != "blocks"})
    assert rep["points"] == 6 * 125 * 27 and rep["dom_bit_exact"] and rep["K_pattern_bit_exact"]
    assert rep["K_columns"] == 64 + 32 + 16 + 8 + 64 + 32
    assert rep["phi_err_vs_truth"] <= TOL and rep["dphi_err_vs_truth"] <= TOL and rep["K_err_vs_truth"] <= TOL
    # the reference's own FP64 arithmetic is 1e-7 .. 1e-6 away from the truth at this size (unshifted Gram-Schmidt)
    assert rep["oracle64_dphi_err_vs_truth"] > 100 * rep["dphi_err_vs_truth"]
    cloud.close()


def _sub_reference(oracle, model, dm, gs, ids, technique="IMLS"):
    """build_space's domain branch (src/FlexiGal.jl:250-252): SHAPE_FUN on x[ids,:], dm[ids,:], DOM mapped back through ids."""
    xs, dms = np.asfortranarray(model.x[ids - 1, :]), np.asfortranarray(dm[ids - 1])
    off, dom_l, phi64, dphi64 = oracle.shape_fun(gs, xs, dms, technique=technique, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, xs, dms, off, dom_l, technique=technique, nthreads=8, quad=True)
    return off, ids[dom_l - 1], phi64, dphi64, phiq, dphiq


def test_subdomain_tags_search_the_subset_in_global_numbering(fg, oracle):
    """set_tag! sub-domains (src/Geometry.jl:303): a domain set only sees the nodes of its own cells; DOM, the sparsity and the
    operators come back in the numbering of the full cloud -- one context, one set per tag, no host-side re-indexing.  A space
    built from BOTH tags (Flexi_Measure concatenation, FlexiGal.jl:207-224) keeps each point's own subset."""
    model = fg.create_model((2.0, 1.0), (16, 8))
    assert fg.set_tag(model, lambda c: c[0] < 0.8, "Soft") > 0 and fg.set_tag(model, lambda c: c[0] >= 0.8, "Stiff") > 0
    tris = {t: fg.Triangulation(model, t) for t in ("Soft", "Stiff", "Domain")}
    recipe = fg.ApproxSpace(model, list(tris.values()), D=1, dmax=1.6)
    dm = recipe.dm()
    refs, spaces = {}, {}
    for tag, k in (("Soft", 1.0), ("Stiff", 50.0)):
        iset = fg.IntegrationSet(tris[tag], 3)
        space = fg.build_space(recipe, [iset])
        S = space.merged()
        gs = space.Measures[0].gs
        conn, _ = fg.geometry.get_entity_info(model, tag)
        ids = np.unique(conn.reshape(-1))
        assert ids.size < model.nnodes
        off, dom, phi64, dphi64, phiq, dphiq = _sub_reference(oracle, model, dm, gs, ids)
        assert np.array_equal(S.offsets, off) and np.array_equal(S.DOM, dom), tag
        assert rel(S.PHI, phiq) <= TOL and rel(S.DPHI, dphiq) <= TOL
        cp, rv = oracle.pattern(model.nnodes, off, dom)
        K = fg.Bilinear_Assembler(fg.GradGrad(k), space)
        assert np.array_equal(K.indptr + 1, cp) and np.array_equal(K.indices + 1, rv), "sparsity differs"
        Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_gradgrad(2, k), 1, cp, rv, quad=True)
        assert rel(K.data, Kq) <= TOL
        F = fg.Linear_Assembler(fg.Source(3.0), space)
        Fq = oracle.assemble_linear(gs, off, dom, phiq, dphiq, oracle.coef_source(2, 3.0), 1, model.nnodes, quad=True)
        assert rel(F, Fq) <= TOL
        outside = np.setdiff1d(np.arange(1, model.nnodes + 1), ids)
        assert np.all(F[outside - 1] == 0.0) and np.all(np.diff(K.indptr)[outside - 1] == 0)      # nodes of the other material are never touched
        refs[tag], spaces[tag] = (gs, off, dom, phiq, dphiq), space
    # the restriction matters: an unrestricted search over the same Gauss points finds more neighbours near the interface
    free = fg.SHAPE_FUN(refs["Soft"][0], cloud=recipe.cloud())
    assert free.total_L > spaces["Soft"].merged().total_L
    # same context for every tag (one handle per model, set ids per tag)
    assert spaces["Soft"].cloud is spaces["Stiff"].cloud
    # both tags in one space: merged set with two different node subsets
    both = fg.build_space(recipe, [fg.IntegrationSet(tris["Soft"], 3), fg.IntegrationSet(tris["Stiff"], 3)])
    gs = np.asfortranarray(np.vstack([refs["Soft"][0], refs["Stiff"][0]]))
    off = np.concatenate([refs["Soft"][1], refs["Stiff"][1][1:] + refs["Soft"][1][-1]])
    dom = np.concatenate([refs["Soft"][2], refs["Stiff"][2]])
    phiq = np.concatenate([refs["Soft"][3], refs["Stiff"][3]]); dphiq = np.vstack([refs["Soft"][4], refs["Stiff"][4]])
    M = both.merged()
    assert np.array_equal(M.offsets, off) and np.array_equal(M.DOM, dom)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    K = fg.Bilinear_Assembler(fg.Mass(2.0), both)
    assert np.array_equal(K.indptr + 1, cp) and np.array_equal(K.indices + 1, rv), "sparsity of the merged space differs"
    Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_mass(2, 2.0), 1, cp, rv, quad=True)
    assert rel(K.data, Kq) <= TOL
    assert abs(K.sum() - 2.0 * 2.0) <= 1e-12            # c * area


def test_merged_space_reports_singular_points_of_its_sources(fg):
    """fg_concat_sets carries the singular-moment-matrix counters of its sources (a Gauss point with too few neighbours
    for the bilinear basis: its shape functions are NaN, like the reference's 1/0)."""
    model = fg.create_model((1.0, 1.0), (6, 6))
    dm = fg.Influence_Domains(model, 1.5)
    cloud = fg.NodeCloud(model.x, dm)
    # a point that only sees the corner node (1 neighbour, 4 basis functions) and regular points
    bad = fg.SHAPE_FUN(np.array([[-0.2, -0.2, 1.0, 1.0]]), cloud=cloud)
    good = fg.SHAPE_FUN(np.array([[0.5, 0.5, 1.0, 1.0], [0.3, 0.6, 1.0, 1.0]]), cloud=cloud)
    assert bad.total_L == 1 and bad.singular_count() == 1 and good.singular_count() == 0
    ctx = cloud.ctx
    dst = ctx.new_set_id()
    ctx.concat_sets([good.set_id, bad.set_id], dst)
    assert ctx.singular_count(dst) == 1
    dst2 = ctx.new_set_id()
    ctx.concat_sets([good.set_id], dst2)
    assert ctx.singular_count(dst2) == 0


@pytest.mark.parametrize("variant", ["block", "warp", "cluster"])
@pytest.mark.parametrize("dim", [2, 3])
def test_imls_kernel_variants_agree_with_the_oracle(fg, oracle, variant, dim):
    """The default IMLS kernel after a cluster search is the software-pipelined cluster-centric one (k_imls_p: candidate records in
    shared memory, entry descriptors from the hit masks, next record in flight); k_imls_c is its unpipelined predecessor
    ("imls_pipe" 0), the block-sorted kernel (k_imls) serves point-wise searches and very long lists, k_imls_w is the
    warp-autonomous variant.  All four must give the oracle's numbers, on a jittered cloud with non-uniform supports."""
    model, dmax, deg = (poisson2d(fg, 18, 1.75) if dim == 2 else poisson3d(fg, 7, 1.5))
    jitter(model)
    dm_map = (lambda p: 1.0 + 0.4 * p[0]) if dim == 2 else None
    dm = fg.Influence_Domains(model, dmax, "rectangular", dm_map)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    off, dom, phi64, dphi64 = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    cloud = fg.NodeCloud(model.x, dm)
    ref = fg.SHAPE_FUN(gs, cloud=cloud)                                  # default: k_imls_p
    assert np.array_equal(ref.offsets, off) and np.array_equal(ref.DOM, dom)
    assert rel(ref.PHI, phiq) <= TOL and rel(ref.DPHI, dphiq) <= TOL
    if variant == "cluster":
        cloud.ctx.set_option("imls_pipe", 0)
    else:
        cloud.ctx.set_option("imls_cluster", 0)
        cloud.ctx.set_option("imls_warp", 1 if variant == "warp" else 0)
    alt = fg.SHAPE_FUN(gs, cloud=cloud)
    assert np.array_equal(alt.DOM, dom)
    assert rel(alt.PHI, phiq) <= TOL and rel(alt.DPHI, dphiq) <= TOL
    assert rel(alt.PHI, ref.PHI) <= 1e-13 and rel(alt.DPHI, ref.DPHI) <= 1e-13
    # point-wise search (no candidate lists): the default falls back to the block-sorted kernel on its own
    cloud.ctx.set_option("imls_cluster", 1)
    cloud.ctx.set_option("nb_pointwise", 1)
    pw = fg.SHAPE_FUN(gs, cloud=cloud)
    assert np.array_equal(pw.DOM, dom) and rel(pw.PHI, phiq) <= TOL and rel(pw.DPHI, dphiq) <= TOL


@pytest.mark.parametrize("jittered", [False, True])
def test_cluster_index_from_search_masks_equals_the_hashed_index(fg, oracle, jittered):
    """The assembly's cluster index (unions, local indices, gather maps) is derived from the search's candidate lists and hit
    masks when the cluster grid nests in the search grid, else hashed from DOM: both must give the oracle's K, with the neighbour
    lists still pending when the mask path runs (DOM is only materialised for the getters)."""
    model, dmax, deg = poisson3d(fg, 8)
    if jittered:
        jitter(model, amount=0.15)
    dm = fg.Influence_Domains(model, dmax)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    Kq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_gradgrad(3), 1, cp, rv, quad=True)
    Mq = oracle.assemble_bilinear(gs, off, dom, phiq, dphiq, oracle.coef_mass(3, 2.0), 1, cp, rv, quad=True)
    stats = []
    for from_masks in (1, 0):
        cloud = fg.NodeCloud(model.x, dm)
        ctx = cloud.ctx
        ctx.set_option("cluster_from_masks", from_masks)
        sid = ctx.new_set_id()
        ng = ctx.set_gauss(sid, gs)
        ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
        nnz = ctx.pattern(sid)
        vals = np.empty(nnz)
        ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), vals)
        assert rel(vals, Kq) <= TOL, from_masks
        ctx.assemble_bilinear(sid, fg.Mass(2.0).c_struct(3, ng), vals)
        assert rel(vals, Mq) <= TOL, from_masks
        stats.append(ctx.assembly_stats(sid))
        o, d, _, _ = ctx.get_shapes(sid, ng, off[-1])
        assert np.array_equal(o, off) and np.array_equal(d, dom)
        cloud.close()
    assert stats[0] == stats[1], stats


def _clustering(x):
    e = 2.0
    return [0.5 + 0.5 * np.tanh((2.0 * x[0] - 1.0) * e) / np.tanh(e), 0.5 + 0.5 * np.tanh((2.0 * x[1] - 1.0) * e) / np.tanh(e)]


def _dm_clustered(x0):
    e = 2.0
    return [e / (np.tanh(e) * np.cosh((2.0 * x0[0] - 1.0) * e) ** 2), e / (np.tanh(e) * np.cosh((2.0 * x0[1] - 1.0) * e) ** 2)]


def test_lid_cavity_shaped_combination(fg, oracle):
    """C5 (examples/Lid_Cavity.jl:5-35) at a parity-friendly size: Moving Kriging + `map` (tanh clustering) + `dm_map`
    (non-uniform supports), VectorField{2}, Navier-Stokes Jacobian and residual with u_h-dependent per-Gauss-point tensors on
    the degree-3 set, the grad-div term on the REDUCED (degree-1) set, and NL_Solver's fixed penalty alpha = 1e6
    (src/Solvers.jl:139) on the four Dirichlet sides merged into one space."""
    Re, D, dim, nb = 500.0, 2, 2, 3
    model = fg.create_model((1.0, 1.0), (20, 20), map=_clustering)
    tri = fg.Triangulation(model, "Domain")
    sides = ["Left", "Bottom", "Right", "Top"]
    recipe = fg.ApproxSpace(model, tri, D=D, dmax=1.5, technique="MK", dm_map=_dm_clustered,
                            Dirichlet_Boundaries=[fg.Triangulation(model, s) for s in sides])
    dm = recipe.dm()
    assert dm.std(axis=0).min() > 0.0                                  # really non-uniform
    dO, dOr = fg.IntegrationSet(tri, 3), fg.IntegrationSet(tri, 1)
    sp, spr = fg.build_space(recipe, [dO]), fg.build_space(recipe, [dOr])
    checks = {}
    for name, space in (("full", sp), ("reduced", spr)):
        S = space.merged(); gs = space.Measures[0].gs
        off, dom, phi64, dphi64 = oracle.shape_fun(gs, model.x, dm, technique="MK", nthreads=8)
        phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, technique="MK", nthreads=8, quad=True)
        assert np.array_equal(S.offsets, off) and np.array_equal(S.DOM, dom) and S.singular_count() == 0
        e_gpu, e_ref = (rel(S.PHI, phiq), rel(S.DPHI, dphiq)), (rel(phi64, phiq), rel(dphi64, dphiq))
        print(f"MK {name}: gpu vs truth {e_gpu}, reference-like FP64 vs truth {e_ref}")
        assert e_gpu[0] <= 10 * e_ref[0] + TOL and e_gpu[1] <= 10 * e_ref[1] + TOL
        checks[name] = (S, gs, off, dom)
    # a velocity field and its values / gradients at the Gauss points (f2) -- the operands of the Jacobian
    X = model.x
    u = np.stack([np.sin(np.pi * X[:, 0]) * X[:, 1], 0.3 * np.cos(2.0 * X[:, 1]) * X[:, 0]], axis=1).reshape(-1)
    S, gs, off, dom = checks["full"]
    ng = gs.shape[0]
    val = fg.Get_Point_Values(u, sp)                                    # (ng, D)
    grad = fg.Get_Point_Values(u, sp, gradient=True)                    # (ng, dim, D): [g, d, i] = d_d u_i
    assert rel(val, oracle.point_values(u, off, dom, S.PHI, D)) <= TOL
    assert rel(grad, oracle.point_gradients(u, off, dom, S.DPHI, D)) <= TOL
    # Jacobian tensor C[(i,al),(j,be)], al/be = 0 (phi), 1+d (d_d phi):
    C = np.zeros((ng, D * nb, D * nb))
    for i in range(D):
        for j in range(D):
            if i == j:
                for d in range(dim):
                    C[:, i * nb + 1 + d, j * nb + 1 + d] += 1.0 / Re                   # grad(du_i).grad(u_i)
                    C[:, i * nb, i * nb + 1 + d] += val[:, d]                           # du_i (u . grad) u_i
            C[:, i * nb + 1 + j, j * nb + 1 + i] += 1.0 / Re                            # the transposed-gradient part
            C[:, i * nb, j * nb] += grad[:, j, i]                                       # du_i (d_j u_i) u_j
    J = fg.Bilinear_Assembler(fg.Generic(C, D=D), sp)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    cpd, rvd = oracle.dof_csc(cp, rv, D)
    assert np.array_equal(J.indptr + 1, cpd) and np.array_equal(J.indices + 1, rvd)
    Jq = oracle.assemble_bilinear(gs, off, dom, S.PHI, S.DPHI, C.reshape(ng, -1), D, cp, rv, quad=True)
    assert rel(J.data, Jq) <= TOL
    # residual: du . ((u . grad) u) + (1/Re) grad(du) (.) (grad u + grad u^T)
    c = np.zeros((ng, D * nb))
    for i in range(D):
        c[:, i * nb] = sum(val[:, d] * grad[:, d, i] for d in range(dim))
        for d in range(dim):
            c[:, i * nb + 1 + d] = (grad[:, d, i] + grad[:, i, d]) / Re
    R = fg.Linear_Assembler(fg.GenericLinear(c, D=D), sp)
    Rq = oracle.assemble_linear(gs, off, dom, S.PHI, S.DPHI, c, D, model.nnodes, quad=True)
    assert rel(R, Rq) <= TOL
    # grad-div stabilisation on the reduced set: (div du) 10 (div u), a constant tensor
    Sr, gsr, offr, domr = checks["reduced"]
    Cd = np.zeros((D * nb, D * nb))
    for i in range(D):
        for j in range(D):
            Cd[i * nb + 1 + i, j * nb + 1 + j] = 10.0
    Gd = fg.Bilinear_Assembler(fg.Generic(Cd, D=D), spr)
    cpr, rvr = oracle.pattern(model.nnodes, offr, domr)
    Gq = oracle.assemble_bilinear(gsr, offr, domr, Sr.PHI, Sr.DPHI, Cd, D, cpr, rvr, quad=True)
    assert rel(Gd.data, Gq) <= TOL
    # NL_Solver's penalty: alpha = 1e6 on the four sides as ONE space (Space_D), lid velocity U0 on "Top"
    alpha = 1.0e6
    spD = fg.build_space(recipe, [fg.IntegrationSet(fg.Triangulation(model, s), 3) for s in sides])
    M = spD.merged()
    Ap = fg.Bilinear_Assembler(fg.Mass(alpha, D=D), spD)
    offb, domb = M.offsets, M.DOM
    cpb, rvb = oracle.pattern(model.nnodes, offb, domb)
    cpbd, rvbd = oracle.dof_csc(cpb, rvb, D)
    assert np.array_equal(Ap.indptr + 1, cpbd) and np.array_equal(Ap.indices + 1, rvbd)
    Apq = oracle.assemble_bilinear(M.gs, offb, domb, M.PHI, M.DPHI, oracle.coef_mass(dim, alpha, D), D, cpb, rvb, quad=True)
    assert rel(Ap.data, Apq) <= TOL
    top = fg.FlexiSpace(spD.sets, D, [spD.Measures[3]], spD.nnodes, spD.cloud)
    gt = spD.Measures[3].gs
    U0 = np.stack([4.0 * gt[:, 0] * (1.0 - gt[:, 0]), np.zeros(gt.shape[0])], axis=1)
    Fp = fg.Linear_Assembler(fg.Source(alpha * U0, D=D), top)
    St = spD.sets["Top"]
    ct = np.zeros((gt.shape[0], D * nb)); ct[:, 0] = alpha * U0[:, 0]; ct[:, nb] = alpha * U0[:, 1]
    Fpq = oracle.assemble_linear(gt, St.offsets, St.DOM, St.PHI, St.DPHI, ct, D, model.nnodes, quad=True)
    assert rel(Fp, Fpq) <= TOL


def test_linear_problem_call_sequence_replayed_through_ctypes(fg, oracle):
    """INTEGRATION.md's shim, call for call: what `Linear_Problem` (src/Assembling_Operators.jl:167-267) makes of the shipped
    3D Poisson example -- one context for the model, one set per tag, the SIX Dirichlet faces as one Space_D whose measures are
    assembled one by one for Fp (temp_space, :258-259) and all together for Ap (:263) -- through the raw C symbols only."""
    from flexigal_b200 import _lib
    lib = _lib.load()
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    model, dmax, deg = poisson3d(fg, 8)
    x = np.asfortranarray(model.x); dm = fg.Influence_Domains(model, dmax)
    nn, dim = x.shape
    h = ctypes.c_void_p()
    assert lib.fg_create(ctypes.byref(h), 0) == 0                                            # gpu_context(model): once per model

    def ck(rc):
        if rc != 0:
            raise RuntimeError(lib.fg_last_error(h).decode())

    ck(lib.fg_set_nodes(h, dim, nn, P(x), P(dm), 0))
    sets, shapes = {}, {}

    def build_space(tags, degree):                                                           # build_space(recipe, isets)
        out = []
        for tag in tags:
            gs = np.asfortranarray(fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, tag), degree)).gs)
            sid = len(sets)
            ck(lib.fg_set_gauss(h, sid, gs.shape[0], P(gs)))
            conn, cat = fg.geometry.get_entity_info(model, tag)
            if cat == "volumes":                                                             # domain set: nodes of the tagged cells
                ids = np.unique(conn.reshape(-1)).astype(np.int64)
                ck(lib.fg_set_node_subset(h, sid, P(ids), ids.size))
            tot, mx = ctypes.c_int64(), ctypes.c_int32()
            ck(lib.fg_neighbours(h, sid, ctypes.byref(tot), ctypes.byref(mx)))
            ck(lib.fg_shape_functions(h, sid, 0))
            off = np.empty(gs.shape[0] + 1, dtype=np.int64); dom = np.empty(tot.value, dtype=np.int64)
            phi = np.empty(tot.value); dphi = np.empty((tot.value, dim))
            ck(lib.fg_get_shapes(h, sid, P(off), P(dom), P(phi), P(dphi)))                   # the ragged containers of the FlexiSpace
            shapes[tag] = (off, dom, phi, dphi)
            sets[tag] = (sid, gs, tot.value)
            out.append(tag)
        return out

    def merged(tags):                                                                        # Flexi_Measure(dX, Space)
        if len(tags) == 1:
            return sets[tags[0]][0]
        key = "+".join(tags)
        if key not in sets:
            ids = (ctypes.c_int * len(tags))(*[sets[t][0] for t in tags])
            sid = len(sets)
            ck(lib.fg_concat_sets(h, ids, len(tags), sid))
            sets[key] = (sid, None, sum(sets[t][2] for t in tags))
        return sets[key][0]

    def bilinear(tags, kind, c0):                                                            # Bilinear_Assembler(f, Space)
        sid = merged(tags)
        nnz = ctypes.c_int64()
        ck(lib.fg_pattern(h, sid, ctypes.byref(nnz)))
        colptr = np.empty(nn + 1, dtype=np.int64); rowval = np.empty(nnz.value, dtype=np.int64); nz = np.empty(nnz.value)
        ck(lib.fg_get_pattern(h, sid, 1, P(colptr), P(rowval)))
        op = _lib.fg_operator(kind, 1, (ctypes.c_double * 4)(c0, 0, 0, 0), None, 0)
        ck(lib.fg_assemble_bilinear(h, sid, ctypes.byref(op), P(nz)))
        return oracle.csc_to_scipy(colptr, rowval, nz, nn), sid

    def linear(tags, c0):                                                                    # Linear_Assembler(f, Space)
        sid = merged(tags)
        F = np.empty(nn)
        op = _lib.fg_operator(_lib.LIN_SOURCE, 1, (ctypes.c_double * 4)(c0, 0, 0, 0), None, 0)
        ck(lib.fg_assemble_linear(h, sid, ctypes.byref(op), P(F)))
        return F

    dOmega = build_space(["Domain"], deg)
    A, sidA = bilinear(dOmega, _lib.OP_GRADGRAD, 1.0)                                        # A += Bilinear_Assembler(...)
    F = linear(dOmega, 5000.0)
    amax = ctypes.c_double()
    ck(lib.fg_values_max(h, sidA, ctypes.byref(amax)))                                       # alpha = maximum(A) * 100
    assert amax.value == max(A.max(), 0.0)
    alpha = amax.value * 100
    faces = ["Left", "Right", "Top", "Bottom", "Back", "Front"]
    Space_D = build_space(faces, deg)
    Fp = np.zeros(nn)
    for tag in faces:                                                                        # temp_space = one measure at a time
        Fp += linear([tag], alpha * 0.25)
    Ap, sidP = bilinear(Space_D, _lib.OP_MASS, alpha)
    K = (A + Ap).tocsc(); K.eliminate_zeros()
    b = F + Fp
    # the oracle's own CPU path of Linear_Problem on the same inputs
    omodel = oracle.create_model(model.domain, model.divisions)
    Ko, bo, info = oracle.linear_problem_scalar(omodel, dmax, deg, oracle.coef_gradgrad(3), oracle.coef_source(3, 5000.0), faces, [0.25] * 6)
    assert abs(alpha - info["alpha"]) <= 1e-10 * alpha
    assert np.array_equal(shapes["Domain"][0], info["off"]) and np.array_equal(shapes["Domain"][1], info["dom"])
    assert rel(shapes["Domain"][2], info["phi"]) <= 1e-9 and rel(shapes["Domain"][3], info["dphi"]) <= 1e-8   # FP64 restatement's own error ball
    assert abs(K - Ko).max() <= 1e-10 * abs(Ko).max() and rel(b, bo) <= 1e-10
    u, uo = fg.Solve(K, b), fg.Solve(Ko, bo)
    assert rel(u, uo) <= 1e-10
    # f3 with K resident: (A on the domain set) + (Ap on the merged boundary set), nothing uploaded but b
    ids = (ctypes.c_int * 2)(sidA, sidP)
    xs = np.empty(nn); it = ctypes.c_int(); res = ctypes.c_double()
    ck(lib.fg_solve_spd_sets(h, ids, None, 2, 1, P(b), P(xs), 1e-13, 20000, ctypes.byref(it), ctypes.byref(res)))
    assert res.value <= 1e-12 and rel(xs, u) <= 1e-9, (it.value, res.value)
    assert lib.fg_destroy(h) == 0


@pytest.mark.parametrize("n", [24, 48])
def test_resident_solve_iteration_counts(fg, n):
    """f3 at growing sizes (3D Poisson, penalty Dirichlet with alpha = 100 max(A)): Jacobi-PCG on the device-resident
    operators converges to 1e-10; iteration counts are printed (they go into DESIGN.md's table with the N = 100 figure from
    bench.py --config solve)."""
    model = fg.create_model((1.0, 1.0, 1.0), (n, n, n))
    tags = ["Left", "Right", "Top", "Bottom", "Back", "Front"]
    tri = fg.Triangulation(model, "Domain")
    recipe = fg.ApproxSpace(model, tri, D=1, dmax=1.35, Dirichlet_Boundaries=[fg.Triangulation(model, t) for t in tags],
                            Dirichlet_Values=[0.0] * 6)
    dO = fg.IntegrationSet(tri, 3)
    A, b, Spaces = fg.Linear_Problem([(fg.GradGrad(1.0), dO)], [(fg.Source(5000.0), dO)], recipe)
    u, it, res = fg.Solve_On_Device(Spaces, b, recipe, rtol=1e-10)
    print(f"resident PCG: N = {n}, {model.nnodes} unknowns, {it} iterations, residual {res:.2e}")
    assert res <= 1e-9 and it < 3000
    r = A @ u - b
    assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b)
    assert u.max() > 0
    recipe.cloud().close()


@pytest.mark.parametrize("dim", [2, 3])
def test_fem_branch_and_mixed_fem_efg_spaces(fg, oracle, dim):
    """SURVEY 8(f) f4, FEM branch of build_space (src/FlexiGal.jl:245-246, egauss_fem / egauss_bound_fem, src/Integration.jl:264-385):
    Gauss table, element shape functions (DOM = the cell's connectivity in vertex order) and sparsity (union of conn x conn) of a FEM
    set, the operators assembled on it, and a Cont_Cast2-shaped mixed problem: a FEM sub-domain and an EFG sub-domain whose operators
    add up in the numbering of the full cloud (examples/Cont_Cast2.jl:22-35), on a mapped (clustered) mesh."""
    if dim == 2:
        # examples/Cont_Cast2.jl:3-9: tanh clustering towards the right wall, supports scaled by the map's derivative
        model = fg.create_model((0.16, 1.0), (12, 12), map=lambda x: [0.16 * np.tanh(x[0] / 0.16 * 3.0) / np.tanh(3.0), x[1]])
        dm_map = lambda x0: [3.0 / (np.tanh(3.0) * np.cosh(x0[0] / 0.16 * 3.0) ** 2), 1.0]
        tags_b = ["Bottom", "Right"]
    else:
        model = fg.create_model((1.0, 1.0, 1.0), (5, 4, 4), map=lambda x: [x[0] ** 1.5, x[1], x[2] * (1 + 0.2 * x[2])])
        dm_map = None
        tags_b = ["Left", "Top"]
    x0max = model.x0[:, 0].max()
    assert fg.set_tag(model, lambda c: c[0] <= 0.45 * model.x[:, 0].max(), "Bulk") > 0 and fg.set_tag(model, lambda c: c[0] > 0.45 * model.x[:, 0].max(), "Shell") > 0
    tris = {t: fg.Triangulation(model, t) for t in ["Bulk", "Shell"] + tags_b}
    recipe = fg.ApproxSpace(model, list(tris.values()), D=1, dmax=1.75, dm_map=dm_map, methods={"Bulk": "FEM", tags_b[0]: "FEM"})
    dm = recipe.dm()
    # --- the FEM sets: cells and a boundary ---------------------------------------------------------------------------------
    for tag, deg in (("Bulk", 2), (tags_b[0], 3)):
        space = fg.build_space(recipe, [fg.IntegrationSet(tris[tag], deg)])
        S = space.merged()
        conn, _ = fg.geometry.get_entity_info(model, tag)
        gs, off, dom, phi, dphi = oracle.egauss_fem(model.x, conn, oracle.pgauss(deg))
        assert np.array_equal(S.offsets, off) and np.array_equal(S.DOM, dom)
        assert rel(S.gs, gs) <= 1e-14 and rel(S.PHI, phi) <= 1e-15 and rel(S.DPHI, dphi) <= 1e-13
        cp, rv = oracle.pattern(model.nnodes, off, dom)
        K = fg.Bilinear_Assembler(fg.GradGrad(30.0), space)
        assert np.array_equal(K.indptr + 1, cp) and np.array_equal(K.indices + 1, rv), "FEM sparsity differs"
        Kq = oracle.assemble_bilinear(gs, off, dom, phi, dphi, oracle.coef_gradgrad(dim, 30.0), 1, cp, rv, quad=True)
        assert rel(K.data, Kq) <= TOL
        v = np.arange(1.0, dim + 1)
        Ka = fg.Bilinear_Assembler(fg.Advection(v), space)
        Kaq = oracle.assemble_bilinear(gs, off, dom, phi, dphi, oracle.coef_advection(dim, v), 1, cp, rv, quad=True)
        assert rel(Ka.data, Kaq) <= TOL
        F = fg.Linear_Assembler(fg.Source(2.0), space)
        Fq = oracle.assemble_linear(gs, off, dom, phi, dphi, oracle.coef_source(dim, 2.0), 1, model.nnodes, quad=True)
        assert rel(F, Fq) <= TOL
        if tag == "Bulk":
            assert abs(K.sum()) <= 1e-9 * np.abs(K.data).max() * K.shape[0]            # element stiffness annihilates constants
            Kbulk, Kbulk_q = K, Kq
            bulk = (gs, off, dom, phi, dphi, cp, rv)
    # --- the EFG sub-domain next to it (Moving Kriging as in the example), same context, global numbering ------------------------
    recipe_mk = fg.ApproxSpace(model, list(tris.values()), D=1, dmax=1.75, dm_map=dm_map, technique="MK", methods={"Bulk": "FEM", tags_b[0]: "FEM"})
    sp_fem = fg.build_space(recipe_mk, [fg.IntegrationSet(tris["Bulk"], 2)])
    sp_efg = fg.build_space(recipe_mk, [fg.IntegrationSet(tris["Shell"], 3)])
    assert sp_fem.cloud is sp_efg.cloud
    A = fg.Bilinear_Assembler(fg.GradGrad(30.0), sp_fem) + fg.Bilinear_Assembler(fg.GradGrad(30.0), sp_efg)
    Se = sp_efg.merged()
    conn, _ = fg.geometry.get_entity_info(model, "Shell")
    ids = np.unique(conn.reshape(-1))
    gse = sp_efg.Measures[0].gs
    offe, dome, _, _, _, _ = _sub_reference(oracle, model, dm, gse, ids, technique="MK")
    assert np.array_equal(Se.offsets, offe) and np.array_equal(Se.DOM, dome) and Se.singular_count() == 0
    cpe, rve = oracle.pattern(model.nnodes, offe, dome)
    Keq = oracle.assemble_bilinear(gse, offe, dome, Se.PHI, Se.DPHI, oracle.coef_gradgrad(dim, 30.0), 1, cpe, rve, quad=True)
    Aq = oracle.csc_to_scipy(bulk[5], bulk[6], Kbulk_q, model.nnodes) + oracle.csc_to_scipy(cpe, rve, Keq, model.nnodes)
    assert abs(A - Aq).max() <= TOL * abs(Aq).max()
    assert np.abs(A @ np.ones(model.nnodes)).max() <= 1e-8 * abs(Aq).max()           # both discretisations reproduce constants
    # FEM and EFG measures cannot share one concatenation: the error says what to do instead
    with pytest.raises(fg.FlexiGalGPUError, match="FEM and EFG sets in one concatenation"):
        fg.build_space(recipe_mk, [fg.IntegrationSet(tris["Bulk"], 2), fg.IntegrationSet(tris["Shell"], 3)]).merged()


@pytest.mark.parametrize("case", ["lattice3d", "lattice2d", "jittered3d", "dm_map2d"])
def test_sparsity_from_the_cluster_index_equals_the_witness_search(fg, oracle, case):
    """fg_pattern after the set's own search derives the sparsity from the cluster index (point masks per local node: pair (a,b) of a
    cluster is present iff some point holds both) when every Gauss point sits in a 32-node cluster; otherwise, and with the option
    "pattern_from_clusters" 0, the witness search runs.  Both must give the oracle's union of DOM x DOM bit for bit."""
    if case == "lattice3d":
        model, dmax, deg = poisson3d(fg, 9); dm_map = None
    elif case == "lattice2d":
        model, dmax, deg = poisson2d(fg, 24, 1.75); dm_map = None
    elif case == "jittered3d":
        model, dmax, deg = poisson3d(fg, 8); jitter(model, amount=0.05); dm_map = None
    else:
        model, dmax, deg = poisson2d(fg, 20, 1.5); dm_map = lambda p: 1.0 + 0.3 * p[0]
    dm = fg.Influence_Domains(model, dmax, "rectangular", dm_map)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    got = {}
    for mode in (1, 0):
        cloud = fg.NodeCloud(model.x, dm)
        ctx, sid = cloud.ctx, cloud.ctx.new_set_id()
        ctx.set_option("pattern_from_clusters", mode)
        ctx.set_gauss(sid, gs)
        ctx.neighbours(sid)
        nnz = ctx.pattern(sid)
        colptr, rowval = ctx.get_pattern(sid, 1, nnz)
        assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv), (case, mode)
        # the assembly that follows uses the same index and pattern
        ctx.shape_functions(sid, "IMLS")
        K = np.empty(nnz); ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(model.dim, gs.shape[0]), K)
        got[mode] = K
        cloud.close()
    assert np.abs(got[1] - got[0]).max() <= 1e-13 * np.abs(got[0]).max()


@pytest.mark.parametrize("D", [1, 3])
def test_source_term_through_the_gather_map_equals_the_cluster_kernel(fg, oracle, D):
    """k_linear_pts (source terms on 32-node clusters through the inverse map of the point-batched assembly) against k_linear_cluster
    ("linear_pts" 0) and against the oracle's load vector, on a jittered cloud (clusters that overflow take the point-wise kernel)."""
    model, dmax, deg = poisson3d(fg, 9)
    jitter(model, amount=0.1)
    dm = fg.Influence_Domains(model, dmax)
    gs = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(model, "Domain"), deg)).gs
    ng = gs.shape[0]
    off, dom, _, _ = oracle.shape_fun(gs, model.x, dm, nthreads=8)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=8, quad=True)
    cloud = fg.NodeCloud(model.x, dm)
    ctx, sid = cloud.ctx, cloud.ctx.new_set_id()
    ctx.set_gauss(sid, gs)
    ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
    ctx.pattern(sid)
    ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), None)          # builds the cluster index
    src = (fg.Source(3.0) if D == 1 else fg.Source((1.0, -2.0, 0.5), D=3)).c_struct(3, ng)
    F1 = np.empty(model.nnodes * D); ctx.assemble_linear(sid, src, F1)
    ctx.set_option("linear_pts", 0)
    F0 = np.empty(model.nnodes * D); ctx.assemble_linear(sid, src, F0)
    assert np.abs(F1 - F0).max() <= 1e-13 * np.abs(F0).max()
    w = gs[:, 3] * gs[:, 4]
    f = np.zeros(model.nnodes)
    np.add.at(f, dom - 1, np.asarray(phiq, dtype=np.float64) * np.repeat(w, np.diff(off)))
    comp = [3.0] if D == 1 else [1.0, -2.0, 0.5]
    Fq = (f[:, None] * np.asarray(comp)[None, :]).reshape(-1)
    assert np.abs(F1 - Fq).max() <= TOL * np.abs(Fq).max()
    cloud.close()


@pytest.mark.parametrize("consumer", ["streamed_assembly", "getter", "linear_form", "plain_assembly"])
def test_deferred_shape_functions_give_the_same_numbers(fg, oracle, consumer):
    """Option "lazy_shapes": fg_shape_functions(IMLS) only sets the buffers up; the streamed assembly then computes the shape functions of
    the search clusters its next range needs between the assembly launches (shape functions, K and the download as one pipeline), and every
    other consumer materialises all of them on first use.  Same kernel on the same data: PHI / DPHI / DOM bit-identical to the eager
    call, K equal up to the order of the atomic sums, whichever consumer comes first."""
    import torch
    model = fg.create_model((1.0, 1.0, 1.0), (14, 14, 14), boundary=False)
    x = np.asfortranarray(model.x)
    dm = np.asfortranarray(np.full_like(x, 1.35 / 14))
    gs = fg.egauss(x, model.conn, fg.pgauss(3))
    ng = gs.shape[0]
    eager = fg.NodeCloud(x, dm, "rectangular", device=0)
    e, es = eager.ctx, eager.ctx.new_set_id()
    e.set_gauss(es, gs)
    tl, _ = e.neighbours(es); e.shape_functions(es, "IMLS")
    nnz = e.pattern(es)
    cop = fg.GradGrad(1.0).c_struct(3, ng)
    Kref = np.empty(nnz); e.assemble_bilinear(es, cop, Kref)
    src = fg.Source(3.0).c_struct(3, ng)
    Fref = np.empty(model.nnodes); e.assemble_linear(es, src, Fref)
    off_r, dom_r, phi_r, dphi_r = e.get_shapes(es, ng, tl)
    lazy = fg.NodeCloud(x, dm, "rectangular", device=0)
    c, sid = lazy.ctx, lazy.ctx.new_set_id()
    c.set_option("lazy_shapes", 1)
    c.set_gauss(sid, gs)
    c.neighbours(sid); c.shape_functions(sid, "IMLS")          # deferred
    assert c.pattern(sid) == nnz
    if consumer == "streamed_assembly":
        pinned = torch.empty(nnz, dtype=torch.float64).pin_memory(); pinned.fill_(float("nan"))
        K = pinned.numpy()
        c.assemble_bilinear_async(sid, cop, K); c.sync()
        assert c.assembly_stats(sid)["clusters"] >= 512
    elif consumer == "plain_assembly":
        K = np.empty(nnz); c.assemble_bilinear(sid, cop, K)
    elif consumer == "linear_form":
        F = np.empty(model.nnodes); c.assemble_linear(sid, src, F)
        assert np.abs(F - Fref).max() <= 1e-13 * np.abs(Fref).max()
        K = np.empty(nnz); c.assemble_bilinear(sid, cop, K)
    else:
        K = None
    off, dom, phi, dphi = c.get_shapes(sid, ng, tl)
    assert np.array_equal(off, off_r) and np.array_equal(dom, dom_r)
    assert np.array_equal(phi, phi_r) and np.array_equal(dphi, dphi_r)
    assert c.singular_count(sid) == 0
    if K is not None:
        assert np.isfinite(K).all() and np.abs(K - Kref).max() <= 1e-13 * np.abs(Kref).max()
    # a second pass through the same context (new search, deferred again) and the streamed pipeline after it
    c.neighbours(sid); c.shape_functions(sid, "IMLS")
    pinned = torch.empty(nnz, dtype=torch.float64).pin_memory(); pinned.fill_(float("nan"))
    K2 = pinned.numpy()
    c.assemble_bilinear_async(sid, cop, K2)
    F2 = np.empty(model.nnodes); c.assemble_linear(sid, src, F2)
    c.sync()
    assert np.abs(K2 - Kref).max() <= 1e-13 * np.abs(Kref).max() and np.abs(F2 - Fref).max() <= 1e-13 * np.abs(Fref).max()
    eager.close(); lazy.close()


@pytest.mark.parametrize("numbering", ["lattice", "shuffled"])
def test_streamed_download_equals_the_plain_download(fg, oracle, numbering):
    """fg_assemble_bilinear_async: the clusters are assembled in ranges and the columns no later cluster touches are mirrored and
    copied out early.  The values on the host equal those of fg_assemble_bilinear + fg_get_values (up to the order of the atomic
    sums), for the lattice numbering (most columns leave early) and for a shuffled node numbering (nothing is final until the
    last range: the bounds collapse and the download happens at the end)."""
    import torch
    model = fg.create_model((1.0, 1.0, 1.0), (14, 14, 14), boundary=False)
    x = np.asfortranarray(model.x)
    if numbering == "shuffled":
        perm = np.random.default_rng(5).permutation(model.nnodes)
        inv = np.empty_like(perm); inv[perm] = np.arange(model.nnodes)
        x = np.asfortranarray(x[perm])                          # new node k = old node perm[k]
        conn = inv[model.conn - 1] + 1
    else:
        conn = model.conn
    dm = np.asfortranarray(np.full_like(x, 1.35 / 14))
    gs = fg.egauss(x, conn, fg.pgauss(3))
    cloud = fg.NodeCloud(x, dm, "rectangular", device=0)
    ctx = cloud.ctx
    sid = ctx.new_set_id()
    ctx.set_gauss(sid, gs)
    ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
    nnz = ctx.pattern(sid)
    for op in (fg.GradGrad(1.0), fg.Mass(2.0)):
        cop = op.c_struct(3, gs.shape[0])
        ref = np.empty(nnz)
        ctx.assemble_bilinear(sid, cop, ref)
        assert ctx.assembly_stats(sid)["clusters"] >= 512
        pinned = torch.empty(nnz, dtype=torch.float64).pin_memory()
        pinned.fill_(float("nan"))
        out = pinned.numpy()
        ctx.assemble_bilinear_async(sid, cop, out)
        ctx.sync()
        assert np.isfinite(out).all()
        assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
        pageable = np.full(nnz, np.nan)
        ctx.assemble_bilinear_async(sid, cop, pageable)
        ctx.sync()
        assert np.abs(pageable - ref).max() <= 1e-13 * np.abs(ref).max()
    # symmetric: the early mirror of every column block used final upper entries
    colptr, rowval = ctx.get_pattern(sid, 1, nnz)
    import scipy.sparse as sp
    K = sp.csc_matrix((out, rowval - 1, colptr - 1), shape=(model.nnodes, model.nnodes))
    assert abs(K - K.T).max() == 0.0
    cloud.close()

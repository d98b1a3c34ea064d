"""CPU tests of the checker itself: the oracle against its committed golden vectors, against its own
__float128 build, against a NumPy/BLAS second opinion, and against the analytic MLS invariants
(the reference ships no fixtures -- SURVEY 8c -- so these are the pins that exist)."""
import hashlib
import os

import numpy as np
import pytest

from cases import rel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["poisson2d_source", "poisson3d_source"])
def test_oracle_reproduces_golden_vectors(oracle, name):
    G = np.load(os.path.join(GOLD, name + ".npz"))
    dim = len(G["domain"])
    model = oracle.create_model(tuple(G["domain"]), tuple(int(v) for v in G["divisions"]))
    dm = oracle.influence_domains(model, float(G["dmax"]))
    gs, _ = oracle.background_integration(model, "Domain", int(G["degree"]))
    assert sha(gs) == str(G["gs_sha"]), "Gauss table changed"
    off, dom, phi, dphi = oracle.shape_fun(gs, model.x, dm, nthreads=os.cpu_count() or 1)
    assert int(off[-1]) == int(G["total_L"]) and np.array_equal(np.diff(off), G["counts"])
    assert sha(dom) == str(G["dom_sha"]), "neighbour lists changed"
    ent = G["sample_entries"]
    assert np.array_equal(dom[ent], G["sample_dom"])
    assert np.array_equal(phi[ent], G["sample_phi64"]) and np.array_equal(dphi[ent], G["sample_dphi64"])   # same binary, same bits
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    assert rv.size == int(G["nnz"]) and sha(cp) == str(G["colptr_sha"]) and sha(rv) == str(G["rowval_sha"])
    K = oracle.assemble_bilinear(gs, off, dom, phi, dphi, oracle.coef_gradgrad(dim), 1, cp, rv)
    assert np.array_equal(K[G["k_sel"]], G["k64"])
    # the literal FP64 restatement sits inside its documented error ball around the __float128 truth
    assert rel(G["sample_phi64"], G["sample_phiq"]) < 2e-10 and rel(G["k64"], G["kq"]) < 2e-10


def test_survey_counts(oracle):
    """Sizes established by the survey: L histograms, sum L^2, nnz, max nnz per column."""
    G2, G3 = (np.load(os.path.join(GOLD, n + ".npz")) for n in ("poisson2d_source", "poisson3d_source"))
    c2 = np.bincount(G2["counts"]); c3 = np.bincount(G3["counts"])
    assert {k: int(c2[k]) for k in np.nonzero(c2)[0]} == {4: 220, 6: 3694, 8: 116, 9: 17601, 12: 854, 16: 15}
    assert {k: int(c3[k]) for k in np.nonzero(c3)[0]} == {8: 2744, 12: 12936, 18: 20328, 27: 10648}
    assert int((G2["counts"].astype(np.int64) ** 2).sum()) == 1696425 and int((G3["counts"].astype(np.int64) ** 2).sum()) == 16387064
    assert int(G2["nnz"]) == 66873 and int(G3["nnz"]) == 205379


@pytest.mark.parametrize("dim", [2, 3])
def test_c_loops_agree_with_numpy_blas_and_float128(oracle, dim):
    model = oracle.create_model((1.0,) * dim, (9,) * dim)
    dm = oracle.influence_domains(model, 1.6)
    gs, _ = oracle.background_integration(model, "Domain", 2)
    off, dom, phi, dphi = oracle.shape_fun(gs, model.x, dm)
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, quad=True)
    assert rel(phi, phiq) < 1e-11 and rel(dphi, dphiq) < 1e-10
    for g in (0, gs.shape[0] // 2, gs.shape[0] - 1):
        v = dom[off[g]:off[g + 1]]
        assert np.array_equal(v, oracle.domain(gs[g, :dim], model.x, dm))
        p, dp = oracle.imls_numpy(gs[g, :dim], model.x, v, dm)
        assert np.abs(p - phi[off[g]:off[g + 1]]).max() < 1e-11 and np.abs(dp - dphi[off[g]:off[g + 1]]).max() < 1e-9


def test_mls_invariants_and_operator_identities(oracle):
    model = oracle.create_model((1.0, 1.0, 1.0), (5, 5, 5))
    dm = oracle.influence_domains(model, 1.35)
    gs, _ = oracle.background_integration(model, "Domain", 3)
    off, dom, phi, dphi = oracle.shape_fun(gs, model.x, dm, quad=True)
    seg = off[:-1]
    X = model.x[dom - 1]
    assert np.abs(np.add.reduceat(phi, seg) - 1).max() < 1e-13
    assert np.abs(np.add.reduceat(dphi, seg, axis=0)).max() < 1e-11
    assert np.abs(np.add.reduceat(phi * X[:, 0] * X[:, 1] * X[:, 2], seg) - gs[:, 0] * gs[:, 1] * gs[:, 2]).max() < 1e-13
    cp, rv = oracle.pattern(model.nnodes, off, dom)
    K = oracle.csc_to_scipy(cp, rv, oracle.assemble_bilinear(gs, off, dom, phi, dphi, oracle.coef_gradgrad(3), 1, cp, rv, quad=True), model.nnodes)
    M = oracle.csc_to_scipy(cp, rv, oracle.assemble_bilinear(gs, off, dom, phi, dphi, oracle.coef_mass(3), 1, cp, rv, quad=True), model.nnodes)
    assert abs(K.sum()) < 1e-10 and abs(M.sum() - 1.0) < 1e-13 and abs(K - K.T).max() < 1e-13
    # the COO + sparse() restatement and the pattern-based assembler are the same operator
    cp2, rv2, nz2 = oracle.assemble_gradgrad_coo(gs, off, dom, dphi, model.nnodes)
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    assert np.abs(nz2 - K.data).max() < 1e-12 * np.abs(K.data).max()


def test_moving_kriging_restatement(oracle):
    """MK (src/Shape_Functions.jl:164-239) interpolates: phi_a(x_b) = delta_ab, and reproduces the basis."""
    model = oracle.create_model((1.0, 1.0), (8, 8))
    dm = oracle.influence_domains(model, 1.75)
    gs, _ = oracle.background_integration(model, "Domain", 2)
    off, dom, phi, dphi = oracle.shape_fun(gs, model.x, dm, technique="MK")
    phiq, dphiq = oracle.shape_fun_given_dom(gs, model.x, dm, off, dom, technique="MK", quad=True)
    assert rel(phi, phiq) < 1e-9 and rel(dphi, dphiq) < 1e-8
    seg = off[:-1]
    X = model.x[dom - 1]
    assert np.abs(np.add.reduceat(phiq, seg) - 1).max() < 1e-12
    assert np.abs(np.add.reduceat(phiq * X[:, 0] * X[:, 1], seg) - gs[:, 0] * gs[:, 1]).max() < 1e-12
    node = np.array([[model.x[10, 0], model.x[10, 1], 1.0, 1.0]])
    o2, d2, p2, _ = oracle.shape_fun(node, model.x, dm, technique="MK", quad=True)
    assert abs(p2[list(d2).index(11)] - 1.0) < 1e-10 and np.abs(np.delete(p2, list(d2).index(11))).max() < 1e-10

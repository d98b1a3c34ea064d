#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle (oracle/efg_oracle.{c,py}).

The reference ships no golden vectors and Julia is not installed, so these fixtures pin the ORACLE (the
literal restatement of the reference algorithm, and its __float128 evaluation) against accidental change;
they are what the GPU tests and the oracle tests compare with.  Run: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import efg_oracle as o  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(name, domain, divisions, dmax, degree, src, tags):
    dim = len(domain)
    model = o.create_model(domain, divisions)
    dm = o.influence_domains(model, dmax)
    gs, _ = o.background_integration(model, "Domain", degree)
    nt = os.cpu_count() or 1
    off, dom, phi, dphi = o.shape_fun(gs, model.x, dm, nthreads=nt)
    phiq, dphiq = o.shape_fun_given_dom(gs, model.x, dm, off, dom, nthreads=nt, quad=True)
    cp, rv = o.pattern(model.nnodes, off, dom)
    K = o.assemble_bilinear(gs, off, dom, phi, dphi, o.coef_gradgrad(dim), 1, cp, rv)
    Kq = o.assemble_bilinear(gs, off, dom, phiq, dphiq, o.coef_gradgrad(dim), 1, cp, rv, quad=True)
    Fq = o.assemble_linear(gs, off, dom, phiq, dphiq, o.coef_source(dim, src), 1, model.nnodes, quad=True)
    A, b, _ = o.linear_problem_scalar(model, dmax, degree, o.coef_gradgrad(dim), o.coef_source(dim, src), tags, [0.0] * len(tags))
    from scipy.sparse.linalg import spsolve
    u = spsolve(A, b)
    ng = gs.shape[0]
    pts = np.unique(np.concatenate([np.arange(0, min(ng, 60)), np.arange(ng // 2, ng // 2 + 60), np.arange(ng - 60, ng),
                                    np.random.default_rng(0).integers(0, ng, 120)]))
    ent = np.concatenate([np.arange(off[g], off[g + 1]) for g in pts])
    ksel = np.arange(0, K.size, 97)
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        domain=np.array(domain), divisions=np.array(divisions), dmax=dmax, degree=degree, src=src,
                        ngauss=ng, total_L=int(off[-1]), counts=np.diff(off).astype(np.int16), dom_sha=sha(dom), gs_sha=sha(gs),
                        sample_points=pts, sample_entries=ent, sample_dom=dom[ent], sample_phi64=phi[ent], sample_dphi64=dphi[ent],
                        sample_phiq=phiq[ent], sample_dphiq=dphiq[ent],
                        nnz=rv.size, colptr_sha=sha(cp), rowval_sha=sha(rv), k_sel=ksel, k64=K[ksel], kq=Kq[ksel], fq=Fq, u=u,
                        k_sum_abs=np.abs(Kq).sum())
    print(name, "ngauss", ng, "sum L", int(off[-1]), "nnz", rv.size)


if __name__ == "__main__":
    o.build()
    make("poisson2d_source", (1.0, 1.0), (50, 50), 1.5, 3, 50.0, ["Left", "Right", "Top", "Bottom"])
    make("poisson3d_source", (1.0, 1.0, 1.0), (12, 12, 12), 1.35, 3, 5000.0, ["Left", "Right", "Top", "Bottom", "Back", "Front"])

"""CPU oracle for the FlexiGal EFG hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``flexigal_b200/`` does.

PARITY UNPINNED: the reference is pure Julia, ships no golden vectors / tests, and Julia is
not installed in this image, so this restatement cannot be checked against outputs of the
reference itself.  Every function cites the reference file:line (under /root/reference) it
follows; the loops that matter live in ``efg_oracle.c`` (double build = literal restatement,
__float128 build = extended-precision "truth" of the very same formulas).

Conventions are Julia's: arrays are column-major (numpy ``order='F'``), ids are 1-based Int64.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
RECT, CIRC = 0, 1
IMLS, MK = 0, 1
_SHAPES = {"rectangular": RECT, "circular": CIRC, "spherical": CIRC}
_TECH = {"IMLS": IMLS, "MK": MK}


def build(force: bool = False) -> None:
    """Compile efg_oracle.c (double + __float128 builds) with the committed Makefile."""
    libs = [os.path.join(_BUILD, n) for n in ("libefg_oracle.so", "libefg_oracle_q.so")]
    src = os.path.join(_HERE, "efg_oracle.c")
    fresh = all(os.path.exists(p) and os.path.getmtime(p) >= os.path.getmtime(src) for p in libs)
    if force or not fresh:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)


_libs: dict = {}


def _lib(quad: bool = False):
    key = "q" if quad else "d"
    if key not in _libs:
        path = os.path.join(_BUILD, "libefg_oracle_q.so" if quad else "libefg_oracle.so")
        if not os.path.exists(path):
            build()
        _libs[key] = ctypes.CDLL(path)
    return _libs[key]


def _fn(name: str, quad: bool = False):
    return getattr(_lib(quad), ("fgoq_" if quad else "fgo_") + name)


def _p(a, ct=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(ct))


_i64 = ctypes.c_int64


# ------------------------------------------------------------------------------------------------
# L1 setup restated (inputs of the hot path): mesh, tags, support sizes, quadrature
# ------------------------------------------------------------------------------------------------
def generate_cartesian_mesh(domain, divisions):
    """src/Meshing.jl:1-60 -- node idx = i+1+(Nx+1)j(+(Nx+1)(Ny+1)k), x = (L/N)*i."""
    D = len(domain)
    assert D in (2, 3)
    if D == 2:
        Nx, Ny = divisions
        Lx, Ly = domain
        x = np.zeros(((Nx + 1) * (Ny + 1), 2), order="F")
        for j in range(Ny + 1):
            for i in range(Nx + 1):
                idx = i + (Nx + 1) * j
                x[idx, 0] = (Lx / Nx) * i
                x[idx, 1] = (Ly / Ny) * j
        conn = np.zeros((Nx * Ny, 4), dtype=np.int64, order="F")
        for j in range(Ny):
            for i in range(Nx):
                e = i + j * Nx
                n1 = i + 1 + j * (Nx + 1)
                n2 = n1 + 1
                n3 = n2 + Nx + 1
                n4 = n3 - 1
                conn[e] = (n1, n2, n3, n4)
    else:
        Nx, Ny, Nz = divisions
        Lx, Ly, Lz = domain
        x = np.zeros(((Nx + 1) * (Ny + 1) * (Nz + 1), 3), order="F")
        for k in range(Nz + 1):
            for j in range(Ny + 1):
                for i in range(Nx + 1):
                    idx = i + (Nx + 1) * j + (Nx + 1) * (Ny + 1) * k
                    x[idx, 0] = (Lx / Nx) * i
                    x[idx, 1] = (Ly / Ny) * j
                    x[idx, 2] = (Lz / Nz) * k
        conn = np.zeros((Nx * Ny * Nz, 8), dtype=np.int64, order="F")
        lay = (Nx + 1) * (Ny + 1)
        for k in range(Nz):
            for j in range(Ny):
                for i in range(Nx):
                    e = i + j * Nx + k * Nx * Ny
                    n0 = i + 1 + j * (Nx + 1) + k * lay
                    n1 = n0 + 1
                    n2 = n1 + Nx + 1
                    n3 = n0 + Nx + 1
                    conn[e] = (n0, n1, n2, n3, n0 + lay, n1 + lay, n2 + lay, n3 + lay)
    return x, conn


def _external_entities(conn, local):
    """FindEdgesQuad / FindFacesandEdgesHexa (src/Meshing.jl:120-212) reduced to what they return
    for the external set: sub-entities that occur exactly once, in order of appearance
    (element-major, local entity minor), node order as listed (unique keeps first occurrences,
    matching_cols_dict maps to the last one; for single occurrences they coincide)."""
    allr = np.concatenate([conn[:, loc][:, None, :] for loc in local], axis=1).reshape(-1, len(local[0]))
    keys = np.sort(allr, axis=1)
    _, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    return allr[cnt[inv.reshape(-1)] == 1]


def create_model(domain, divisions, map=None):
    """src/Geometry.jl:20-276: mesh + tagged boundary entities (tags are decided on the UNMAPPED
    coordinates, the optional coordinate map is applied afterwards, :259-266)."""
    D = len(domain)
    x, conn = generate_cartesian_mesh(domain, divisions)
    tol = 1e-6
    ent = {}
    if D == 2:
        Lx, Ly = domain
        ext = _external_entities(conn, [[0, 1], [1, 2], [2, 3], [3, 0]])
        tags = {"Left": [], "Right": [], "Bottom": [], "Top": []}
        for n1, n2 in ext:
            xm = (x[n1 - 1, 0] + x[n2 - 1, 0]) * 0.5
            ym = (x[n1 - 1, 1] + x[n2 - 1, 1]) * 0.5
            if xm < tol:
                tags["Left"].append((n1, n2))
            elif xm > Lx - tol:
                tags["Right"].append((n1, n2))
            elif ym < tol:
                tags["Bottom"].append((n1, n2))
            elif ym > Ly - tol:
                tags["Top"].append((n1, n2))
        ent["ext_edges"] = {k: np.array(v, dtype=np.int64).reshape(-1, 2) for k, v in tags.items()}
        ent["ext_edges"]["Boundary"] = ext
        ent["faces"] = {"Domain": conn}
    else:
        Lx, Ly, Lz = domain
        ext = _external_entities(conn, [[0, 3, 7, 4], [1, 2, 6, 5], [0, 1, 5, 4], [2, 6, 7, 3], [0, 1, 2, 3], [4, 5, 6, 7]])
        tags = {"Left": [], "Right": [], "Bottom": [], "Top": [], "Back": [], "Front": []}
        for f in ext:
            xm = (x[f[0] - 1, 0] + x[f[1] - 1, 0] + x[f[2] - 1, 0] + x[f[3] - 1, 0]) * 0.25
            ym = (x[f[0] - 1, 1] + x[f[1] - 1, 1] + x[f[2] - 1, 1] + x[f[3] - 1, 1]) * 0.25
            zm = (x[f[0] - 1, 2] + x[f[1] - 1, 2] + x[f[2] - 1, 2] + x[f[3] - 1, 2]) * 0.25
            if xm < tol:
                tags["Left"].append(f)
            elif xm > Lx - tol:
                tags["Right"].append(f)
            elif zm < tol:
                tags["Bottom"].append(f)
            elif zm > Lz - tol:
                tags["Top"].append(f)
            elif ym < tol:
                tags["Back"].append(f)
            elif ym > Ly - tol:
                tags["Front"].append(f)
        ent["ext_faces"] = {k: np.array(v, dtype=np.int64).reshape(-1, 4) for k, v in tags.items()}
        ent["ext_faces"]["Boundary"] = ext
        ent["volumes"] = {"Domain": conn}
    x0 = x.copy(order="F")
    if map is not None:
        for i in range(x.shape[0]):
            x[i, :] = map(x0[i, :])
    return SimpleNamespace(domain=tuple(domain), divisions=tuple(divisions), x=x, x0=x0, conn=conn,
                           entities=ent, dim=D, ncells=conn.shape[0], nnodes=x.shape[0])


def get_entity_info(model, tag):
    """src/Geometry.jl:288-301."""
    order = ["faces", "ext_edges"] if model.dim == 2 else ["volumes", "ext_faces"]
    for cat in order:
        if cat in model.entities and tag in model.entities[cat]:
            return model.entities[cat][tag], cat
    raise KeyError(tag)


def influence_domains(model, dmax, shape="rectangular", dm_map=None):
    """Influence_Domains, src/FlexiGal.jl:164-199."""
    dim = model.dim
    dvec = [dmax] * dim if np.isscalar(dmax) else list(dmax)
    if shape == "rectangular":
        dm = np.zeros((model.nnodes, dim), order="F")
        for d in range(dim):
            dm[:, d] = dvec[d] * model.domain[d] / model.divisions[d]
        if dm_map is not None:
            for i in range(model.nnodes):
                s = dm_map(model.x0[i, :])
                if np.isscalar(s):
                    dm[i, :] *= s
                else:
                    for d in range(dim):
                        dm[i, d] *= s[d]
        return dm
    if shape in ("circular", "spherical"):
        spacing = sum(model.domain[d] / model.divisions[d] for d in range(dim)) / dim
        return np.full(model.nnodes, dvec[0] * spacing)
    raise ValueError(f"Unknown shape of influence domain: {shape}")


def pgauss(k):
    """src/Integration.jl:1-23."""
    if k == 1:
        return np.array([[0.0], [2.0]])
    if k == 2:
        return np.array([[-1 / np.sqrt(3.0), 1 / np.sqrt(3.0)], [1.0, 1.0]])
    if k == 3:
        x1 = -np.sqrt(0.6)
        return np.array([[x1, 0.0, -x1], [5 / 9, 8 / 9, 5 / 9]])
    if k == 4:
        x1, x2 = -0.8611363115940526, -0.3399810435848563
        w1, w2 = 0.3478548451374539, 0.6521451548625461
        return np.array([[x1, x2, -x2, -x1], [w1, w2, w2, w1]])
    raise ValueError("Gauss integrations just implemented for ngpts = 1, 2, 3 o 4")


def _coords_by_ele(x, conn):
    return x[conn - 1, :]  # (nele, nvert, dim), src/Meshing.jl:62-73


def _measure2d(xi, eta, xe):
    """Measure2D_Cell, src/Integration.jl:25-74 (sequential accumulation over the 4 vertices)."""
    ncell, _, dim = xe.shape
    phi = [0.25 * (1 - xi) * (1 - eta), 0.25 * (1 + xi) * (1 - eta), 0.25 * (1 + xi) * (1 + eta), 0.25 * (1 - xi) * (1 + eta)]
    dxi = [-0.25 * (1 - eta), 0.25 * (1 - eta), 0.25 * (1 + eta), -0.25 * (1 + eta)]
    det = [-0.25 * (1 - xi), -0.25 * (1 + xi), 0.25 * (1 + xi), 0.25 * (1 - xi)]
    xg = np.zeros((ncell, dim)); a = np.zeros((ncell, dim)); b = np.zeros((ncell, dim))
    for i in range(dim):
        for k in range(4):
            xg[:, i] += phi[k] * xe[:, k, i]
            a[:, i] += dxi[k] * xe[:, k, i]
            b[:, i] += det[k] * xe[:, k, i]
    if dim == 2:
        jac = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    else:
        sx = a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2]
        sy = b[:, 0] * a[:, 2] - a[:, 0] * b[:, 2]
        sz = a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]
        jac = np.sqrt(sx ** 2 + sy ** 2 + sz ** 2)
    return xg, jac


def _measure3d(xi, eta, zeta, xe):
    """Measure3D_Cell, src/Integration.jl:76-146."""
    ncell = xe.shape[0]
    sg = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    phi = [(1 / 8) * (1 + s * xi) * (1 + t * eta) * (1 + u * zeta) for s, t, u in sg]
    dl = [[s * (1 / 8) * (1 + t * eta) * (1 + u * zeta) for s, t, u in sg],
          [t * (1 / 8) * (1 + s * xi) * (1 + u * zeta) for s, t, u in sg],
          [u * (1 / 8) * (1 + s * xi) * (1 + t * eta) for s, t, u in sg]]
    J = np.zeros((ncell, 3, 3)); xg = np.zeros((ncell, 3))
    for i in range(3):
        for j in range(3):
            for k in range(8):
                J[:, i, j] += dl[j][k] * xe[:, k, i]
                if j == 0:
                    xg[:, i] += phi[k] * xe[:, k, i]
    jac = (J[:, 0, 0] * (J[:, 1, 1] * J[:, 2, 2] - J[:, 2, 1] * J[:, 1, 2])
           - J[:, 0, 1] * (J[:, 1, 0] * J[:, 2, 2] - J[:, 2, 0] * J[:, 1, 2])
           + J[:, 0, 2] * (J[:, 1, 0] * J[:, 2, 1] - J[:, 2, 0] * J[:, 1, 1]))
    return xg, jac


def egauss(xc, conn, gauss):
    """src/Integration.jl:148-182: gs[(e-1)*l^dim + count] = [x_g..., w, detJ], i fastest."""
    dim = xc.shape[1]
    l = gauss.shape[1]
    npts = l ** dim
    gs = np.zeros((conn.shape[0] * npts, dim + 2), order="F")
    xe = _coords_by_ele(xc, conn)
    count = 0
    if dim == 2:
        for j in range(l):
            for i in range(l):
                xg, jac = _measure2d(gauss[0, i], gauss[0, j], xe)
                gs[count::npts, 0:2] = xg
                gs[count::npts, 2] = gauss[1, i] * gauss[1, j]
                gs[count::npts, 3] = jac
                count += 1
    else:
        for k in range(l):
            for j in range(l):
                for i in range(l):
                    xg, jac = _measure3d(gauss[0, i], gauss[0, j], gauss[0, k], xe)
                    gs[count::npts, 0:3] = xg
                    gs[count::npts, 3] = gauss[1, i] * gauss[1, j] * gauss[1, k]
                    gs[count::npts, 4] = jac
                    count += 1
    return gs


def egauss_bound(xc, conn, gauss):
    """src/Integration.jl:184-218."""
    dim = xc.shape[1]
    l = gauss.shape[1]
    npts = l ** (dim - 1)
    gs = np.zeros((conn.shape[0] * npts, dim + 2), order="F")
    xe = _coords_by_ele(xc, conn)
    count = 0
    if dim == 2:
        for i in range(l):
            xi = gauss[0, i]
            gs[count::l, 0] = (1 - xi) / 2 * xe[:, 0, 0] + (1 + xi) / 2 * xe[:, 1, 0]
            gs[count::l, 1] = (1 - xi) / 2 * xe[:, 0, 1] + (1 + xi) / 2 * xe[:, 1, 1]
            gs[count::l, 2] = gauss[1, i]
            gs[count::l, 3] = np.sqrt((xe[:, 0, 0] - xe[:, 1, 0]) ** 2 + (xe[:, 1, 1] - xe[:, 0, 1]) ** 2) / 2
            count += 1
    else:
        for j in range(l):
            for i in range(l):
                xg, jac = _measure2d(gauss[0, i], gauss[0, j], xe)
                gs[count::npts, 0:3] = xg
                gs[count::npts, 3] = gauss[1, i] * gauss[1, j]
                gs[count::npts, 4] = jac
                count += 1
    return gs


# FEM branch of build_space: BackgroundIntegration(iset, :FEM), src/Integration.jl:220-385 ------------------------------------------
def _fem_shape_line2(xi):
    """fem_shape_line2, src/Integration.jl:220-224."""
    return np.array([0.5 * (1 - xi), 0.5 * (1 + xi)]), np.array([[-0.5, 0.5]])


def _fem_shape_quad4(xi, eta):
    """fem_shape_quad4, src/Integration.jl:227-239."""
    phi = np.array([0.25 * (1 - xi) * (1 - eta), 0.25 * (1 + xi) * (1 - eta), 0.25 * (1 + xi) * (1 + eta), 0.25 * (1 - xi) * (1 + eta)])
    d = np.array([[-0.25 * (1 - eta), 0.25 * (1 - eta), 0.25 * (1 + eta), -0.25 * (1 + eta)],
                  [-0.25 * (1 - xi), -0.25 * (1 + xi), 0.25 * (1 + xi), 0.25 * (1 - xi)]])
    return phi, d


def _fem_shape_hex8(xi, eta, zeta):
    """fem_shape_hex8, src/Integration.jl:242-258."""
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]); sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0]); sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    fx, fy, fz = 1 + sx * xi, 1 + sy * eta, 1 + sz * zeta
    phi = (1 / 8) * fx * fy * fz
    d = np.stack([sx * (1 / 8) * fy * fz, sy * (1 / 8) * fx * fz, sz * (1 / 8) * fx * fy])
    return phi, d


def egauss_fem(xc, conn, gauss):
    """egauss_fem / egauss_bound_fem, src/Integration.jl:264-385 -> (gs, offsets 0-based, dom 1-based in vertex order, phi, dphi[sumL, dim]).
    The gradient is inv(J)' * local gradient with J[p, d] = d x_d / d xi_p, exactly as the reference writes it (:292-295, :314-317);
    boundary entities use the tangent / metric-tensor formulas of :341-376."""
    xc = np.asarray(xc, dtype=np.float64); conn = np.asarray(conn, dtype=np.int64)
    dim = xc.shape[1]; nv = conn.shape[1]; l = gauss.shape[1]
    pd = {(2, 4): 2, (3, 8): 3, (2, 2): 1, (3, 4): 2}[(dim, nv)]
    npts = l ** pd
    nc = conn.shape[0]; ng = nc * npts
    gs = np.zeros((ng, dim + 2), order="F")
    phi = np.zeros(ng * nv); dphi = np.zeros((ng * nv, dim)); dom = np.zeros(ng * nv, dtype=np.int64)
    idx = 0
    for e in range(nc):
        xe = xc[conn[e] - 1, :]
        for q in range(npts):
            i, j, k = q % l, (q // l) % l, q // (l * l)
            if pd == 1:
                N, dN = _fem_shape_line2(gauss[0, i]); w = gauss[1, i]
            elif pd == 2:
                N, dN = _fem_shape_quad4(gauss[0, i], gauss[0, j]); w = gauss[1, i] * gauss[1, j]
            else:
                N, dN = _fem_shape_hex8(gauss[0, i], gauss[0, j], gauss[0, k]); w = gauss[1, i] * gauss[1, j] * gauss[1, k]
            xg = np.array([sum(N[m] * xe[m, d] for m in range(nv)) for d in range(dim)])
            T = np.array([[sum(dN[p, m] * xe[m, d] for m in range(nv)) for d in range(dim)] for p in range(pd)])
            if pd == dim:
                meas = np.linalg.det(T) if dim == 3 else T[0, 0] * T[1, 1] - T[0, 1] * T[1, 0]
                invJT = np.linalg.inv(T).T
                G = np.array([invJT @ dN[:, m] for m in range(nv)])
            elif pd == 1:
                tx, ty = T[0, 0], T[0, 1]
                meas = np.sqrt(tx ** 2 + ty ** 2)
                G = np.array([[dN[0, m] * tx / meas ** 2, dN[0, m] * ty / meas ** 2] for m in range(nv)])
            else:
                t1, t2 = T[0], T[1]
                g11, g12, g22 = t1 @ t1, t1 @ t2, t2 @ t2
                detG = g11 * g22 - g12 ** 2
                meas = np.sqrt(detG)
                invG = np.array([[g22 / detG, -g12 / detG], [-g12 / detG, g11 / detG]])
                G = np.array([(invG @ dN[:, m])[0] * t1 + (invG @ dN[:, m])[1] * t2 for m in range(nv)])
            gs[idx, :dim] = xg; gs[idx, dim] = w; gs[idx, dim + 1] = meas
            sl = slice(idx * nv, (idx + 1) * nv)
            phi[sl] = N; dphi[sl, :] = G; dom[sl] = conn[e]
            idx += 1
    off = np.arange(ng + 1, dtype=np.int64) * nv
    return gs, off, dom, phi, dphi


def background_integration(model, tag, degree):
    """BackgroundIntegration(:EFG), src/FlexiGal.jl:27-52."""
    conn, cat = get_entity_info(model, tag)
    is_boundary = cat != ("faces" if model.dim == 2 else "volumes")
    g = pgauss(degree)
    return (egauss_bound if is_boundary else egauss)(model.x, conn, g), is_boundary


# ------------------------------------------------------------------------------------------------
# hot path restated (C loops)
# ------------------------------------------------------------------------------------------------
def _prep(x, dm, gs):
    x = np.asfortranarray(x, dtype=np.float64)
    dm = np.asfortranarray(dm, dtype=np.float64)
    gs = np.asfortranarray(gs, dtype=np.float64)
    return x, dm, gs


def domain(gpos, x, dm, shape="rectangular"):
    """domain(), src/Shape_Functions.jl:3-27 -> ascending 1-based ids."""
    x, dm, _ = _prep(x, dm, np.zeros((1, 1)))
    nn, dim = x.shape
    g = np.ascontiguousarray(gpos, dtype=np.float64)
    out = np.empty(nn, dtype=np.int64)
    f = _fn("domain"); f.restype = _i64
    L = f(dim, _i64(nn), _p(x), _p(dm), _SHAPES[shape], _p(g), _p(out, _i64))
    return out[:L].copy()


def shape_fun(gs, x, dm, shape="rectangular", technique="IMLS", nthreads=1, quad=False):
    """SHAPE_FUN(), src/Shape_Functions.jl:241-263 -> (offsets[ngauss+1] 0-based, dom 1-based,
    phi[sumL], dphi[sumL, dim]).  quad=True evaluates the same formulas in __float128 (the
    neighbour predicate stays FP64) and rounds the result to double."""
    x, dm, gs = _prep(x, dm, gs)
    nn, dim = x.shape
    ng = gs.shape[0]
    off = np.zeros(ng + 1, dtype=np.int64)
    _fn("shape_fun_count", quad)(dim, _i64(nn), _p(x), _p(dm), _SHAPES[shape], _i64(ng), _p(gs), _p(off, _i64), nthreads)
    tot = int(off[-1])
    dom = np.empty(max(tot, 1), dtype=np.int64)
    phi = np.empty(max(tot, 1)); dphi = np.empty((max(tot, 1), dim))
    rc = _fn("shape_fun_fill", quad)(dim, _i64(nn), _p(x), _p(dm), _SHAPES[shape], _TECH[technique], _i64(ng), _p(gs),
                                     _p(off, _i64), _p(dom, _i64), _p(phi), _p(dphi), nthreads)
    assert rc == 0
    return off, dom[:tot], phi[:tot], dphi[:tot]


def shape_fun_given_dom(gs, x, dm, off, dom, shape="rectangular", technique="IMLS", nthreads=1, quad=False):
    x, dm, gs = _prep(x, dm, gs)
    nn, dim = x.shape
    ng = gs.shape[0]
    off = np.ascontiguousarray(off, dtype=np.int64); dom = np.ascontiguousarray(dom, dtype=np.int64)
    tot = int(off[-1])
    phi = np.empty(max(tot, 1)); dphi = np.empty((max(tot, 1), dim))
    _fn("shape_fun_given_dom", quad)(dim, _i64(nn), _p(x), _p(dm), _SHAPES[shape], _TECH[technique], _i64(ng), _p(gs),
                                     _p(off, _i64), _p(dom, _i64), _p(phi), _p(dphi), nthreads)
    return phi[:tot], dphi[:tot]


def imls_numpy(gpos, x, v, dm):
    """Improved_Moving_Least_Squares (rectangular), src/Shape_Functions.jl:88-162, written with the
    same array expressions as the Julia source so that numpy's BLAS plays the role of OpenBLAS.
    A second opinion on the C loops; v is 1-based."""
    gpos = np.asarray(gpos, dtype=float)
    dim = gpos.size
    nv = x[v - 1, :]
    L = len(v)
    if dim == 2:
        q = np.column_stack([np.ones(L), nv[:, 0], nv[:, 1], nv[:, 0] * nv[:, 1]])
    else:
        X, Y, Z = nv[:, 0], nv[:, 1], nv[:, 2]
        q = np.column_stack([np.ones(L), X, Y, Z, X * Y, X * Z, Y * Z, X * Y * Z])
    dif = gpos[None, :] - nv
    dmv = dm[v - 1, :]
    dr = np.sign(dif) / dmv
    r = np.abs(dif) / dmv
    wt = np.where(r > 0.5, (4 / 3) - 4 * r + 4 * r ** 2 - (4 / 3) * r ** 3, (2 / 3) - 4 * r ** 2 + 4 * r ** 3)
    dwt = np.where(r > 0.5, (-4 + 8 * r - 4 * r ** 2) * dr, (-8 * r + 12 * r ** 2) * dr)
    w = np.prod(wt, axis=1)
    dw = np.empty((L, dim))
    for d in range(dim):
        t = wt.copy(); t[:, d] = dwt[:, d]
        dw[:, d] = np.prod(t, axis=1)
    g = gpos
    if dim == 2:
        QQg = np.array([[1, g[0], g[1], g[0] * g[1]], [0, 1, 0, g[1]], [0, 0, 1, g[0]]], dtype=float)
    else:
        QQg = np.array([[1, g[0], g[1], g[2], g[0] * g[1], g[0] * g[2], g[1] * g[2], g[0] * g[1] * g[2]],
                        [0, 1, 0, 0, g[1], g[2], 0, g[1] * g[2]],
                        [0, 0, 1, 0, g[0], 0, g[2], g[0] * g[2]],
                        [0, 0, 0, 1, 0, g[0], g[1], g[0] * g[1]]], dtype=float)
    m = q.shape[1]
    p = q.copy(); PPg = QQg.copy()
    for i in range(1, m):
        for k in range(i):
            c1 = (q[:, i] @ (w * p[:, k])) / (p[:, k] @ (w * p[:, k]))
            p[:, i] -= c1 * p[:, k]
            PPg[:, i] -= c1 * PPg[:, k]
    aa = p.T @ (w[:, None] * p)
    aia = 1.0 / np.diag(aa)
    gam = np.zeros((m, dim + 1))
    gam[:, 0] = aia * PPg[0, :]
    for d in range(dim):
        daa = p.T @ (dw[:, d][:, None] * p)
        gam[:, d + 1] = aia * (PPg[d + 1, :] - daa @ gam[:, 0])
    B = p * w[:, None]
    phi = B @ gam[:, 0]
    dphi = np.column_stack([B @ gam[:, d + 1] + (p * dw[:, d][:, None]) @ gam[:, 0] for d in range(dim)])
    return phi, dphi


def pattern(nnodes, off, dom):
    """Sparsity of sparse(row, col, val) for the COO of src/Assembling_Operators.jl:14-22:
    column j = sorted union of DOM_g over the Gauss points containing j.  1-based CSC."""
    off = np.ascontiguousarray(off, dtype=np.int64); dom = np.ascontiguousarray(dom, dtype=np.int64)
    ng = off.size - 1
    colptr = np.empty(nnodes + 1, dtype=np.int64)
    f = _fn("pattern"); f.restype = _i64
    nnz = f(_i64(nnodes), _i64(ng), _p(off, _i64), _p(dom, _i64), _p(colptr, _i64), None)
    rowval = np.empty(max(nnz, 1), dtype=np.int64)
    f(_i64(nnodes), _i64(ng), _p(off, _i64), _p(dom, _i64), _p(colptr, _i64), _p(rowval, _i64))
    return colptr, rowval[:nnz]


def _coef(coef, ng, n):
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    if coef.ndim == 1 or coef.shape[0] != ng or coef.size == n:
        assert coef.size == n, (coef.shape, n)
        return coef.reshape(-1), 0
    assert coef.size == ng * n
    return coef.reshape(-1), n


def assemble_bilinear(gs, off, dom, phi, dphi, coef, D, colptr, rowval, quad=False):
    """Bilinear_Assembler, src/Assembling_Operators.jl:81-104 (+ :2-26, :28-79) for the closure
    described by the coefficient tensor ``coef`` ((D*nb)^2 constant, or per Gauss point)."""
    gs = np.asfortranarray(gs, dtype=np.float64)
    ng, dim = gs.shape[0], gs.shape[1] - 2
    n = D * (1 + dim)
    c, stride = _coef(coef, ng, n * n)
    nnz = rowval.size
    out = np.zeros(nnz * D * D)
    phi = np.ascontiguousarray(phi); dphi = np.ascontiguousarray(dphi)
    off = np.ascontiguousarray(off, dtype=np.int64); dom = np.ascontiguousarray(dom, dtype=np.int64)
    _fn("assemble_bilinear", quad)(dim, D, _i64(ng), _p(gs), _p(off, _i64), _p(dom, _i64), _p(phi), _p(dphi), _p(c), _i64(stride),
                                   _p(colptr, _i64), _p(rowval, _i64), _p(out), _i64(nnz))
    return out


def assemble_linear(gs, off, dom, phi, dphi, coef, D, nnodes, quad=False):
    """Linear_Assembler, src/Assembling_Operators.jl:147-165 (+ :105-145)."""
    gs = np.asfortranarray(gs, dtype=np.float64)
    ng, dim = gs.shape[0], gs.shape[1] - 2
    n = D * (1 + dim)
    c, stride = _coef(coef, ng, n)
    out = np.zeros(nnodes * D)
    phi = np.ascontiguousarray(phi); dphi = np.ascontiguousarray(dphi)
    off = np.ascontiguousarray(off, dtype=np.int64); dom = np.ascontiguousarray(dom, dtype=np.int64)
    _fn("assemble_linear", quad)(dim, D, _i64(ng), _p(gs), _p(off, _i64), _p(dom, _i64), _p(phi), _p(dphi), _p(c), _i64(stride),
                                 _p(out), _i64(nnodes))
    return out


def assemble_gradgrad_coo(gs, off, dom, dphi, nnodes):
    """Literal COO push + sparse() for the Poisson stiffness (timed CPU-baseline leg)."""
    gs = np.asfortranarray(gs, dtype=np.float64)
    ng, dim = gs.shape[0], gs.shape[1] - 2
    off = np.ascontiguousarray(off, dtype=np.int64); dom = np.ascontiguousarray(dom, dtype=np.int64)
    L = np.diff(off)
    cap = int(min(int((L * L).sum()), nnodes * 4096)) + 1
    colptr = np.empty(nnodes + 1, dtype=np.int64); rowval = np.empty(cap, dtype=np.int64); nz = np.empty(cap)
    f = _fn("assemble_gradgrad_coo"); f.restype = _i64
    nnz = f(dim, _i64(nnodes), _i64(ng), _p(gs), _p(off, _i64), _p(dom, _i64), _p(np.ascontiguousarray(dphi)),
            _p(colptr, _i64), _p(rowval, _i64), _p(nz))
    return colptr, rowval[:nnz].copy(), nz[:nnz].copy()


def dof_csc(colptr, rowval, D):
    """Expand the node-level CSC pattern to the DOF-level one of _assemble_vector!
    (global dof = (node-1)*D + i, src/Assembling_Operators.jl:65-68); nzval layout matches
    efg_oracle.c:assemble_bilinear."""
    if D == 1:
        return colptr.copy(), rowval.copy()
    nn = colptr.size - 1
    cnt = np.diff(colptr)
    cp = np.empty(nn * D + 1, dtype=np.int64)
    cp[0] = 1
    cp[1:] = 1 + np.cumsum(np.repeat(cnt * D, D))
    rv = np.empty(rowval.size * D * D, dtype=np.int64)
    for j in range(nn):
        rows = rowval[colptr[j] - 1:colptr[j + 1] - 1]
        drows = ((rows - 1)[:, None] * D + np.arange(1, D + 1)[None, :]).reshape(-1)
        base = (colptr[j] - 1) * D * D
        for jc in range(D):
            rv[base + jc * drows.size: base + (jc + 1) * drows.size] = drows
    return cp, rv


# coefficient tensors of the canonical closures (arithmetic from src/Fields_Operations.jl) ----------
def coef_gradgrad(dim, k=1.0, D=1):
    """∇(δT)⋅∇(T) (Fields_Operations.jl:321-323, :1002) ; vector: ∇δu⊙∇u (:552-556)."""
    nb = 1 + dim; n = D * nb
    C = np.zeros((n, n))
    for i in range(D):
        for d in range(dim):
            C[i * nb + 1 + d, i * nb + 1 + d] = k
    return C


def coef_mass(dim, c=1.0, D=1):
    """δu*u (:1004) ; vector δu⋅u (:546-550)."""
    nb = 1 + dim; n = D * nb
    C = np.zeros((n, n))
    for i in range(D):
        C[i * nb, i * nb] = c
    return C


def coef_advection(dim, v):
    """δT*(v⋅∇T) (examples/Poisson2D_No_Source.jl:30); v: (dim,) or (ngauss, dim)."""
    v = np.asarray(v, dtype=float)
    nb = 1 + dim
    if v.ndim == 1:
        C = np.zeros((nb, nb)); C[0, 1:] = v
        return C
    C = np.zeros((v.shape[0], nb, nb)); C[:, 0, 1:] = v
    return C.reshape(v.shape[0], -1)


def coef_elasticity(dim, G, lam):
    """∇(δu)⊙σ(u), σ = G(∇u+∇uᵀ)+λ(∇⋅u)Id (examples/Elasticity_Test.jl:25-26):
    K_ab[i,j] = G δ_ij ∇φ_a·∇φ_b + G ∂_jφ_a ∂_iφ_b + λ ∂_iφ_a ∂_jφ_b."""
    D = dim; nb = 1 + dim; n = D * nb
    C = np.zeros((n, n))
    for i in range(D):
        for j in range(D):
            for al in range(dim):
                for be in range(dim):
                    c = 0.0
                    if i == j and al == be:
                        c += G
                    if al == j and be == i:
                        c += G
                    if al == i and be == j:
                        c += lam
                    C[i * nb + 1 + al, j * nb + 1 + be] = c
    return C


def coef_source(dim, c, D=1):
    """Linear form c*δT (only .phi of the returned measure is used, Assembling_Operators.jl:203-209)."""
    nb = 1 + dim
    c = np.asarray(c, dtype=float)
    if D == 1:
        if c.ndim == 0:
            out = np.zeros(nb); out[0] = c
            return out
        out = np.zeros((c.shape[0], nb)); out[:, 0] = c
        return out
    if c.ndim == 1:
        out = np.zeros(D * nb); out[0::nb] = c
        return out
    out = np.zeros((c.shape[0], D * nb)); out[:, 0::nb] = c
    return out


def point_values(u, off, dom, phi, D=1):
    """Get_Point_Values(::FlexiFunction), src/Fields_Operations.jl:40-64: u_h(g) = sum_a phi_a u[dom_a] (sequential)."""
    u = np.asarray(u, dtype=float).reshape(-1, D)
    out = np.zeros((off.size - 1, D))
    for g in range(off.size - 1):
        for k in range(off[g], off[g + 1]):
            out[g] += phi[k] * u[dom[k] - 1]
    return out


def point_gradients(u, off, dom, dphi, D=1):
    """Get_Point_Values(::VecFlexiFunction{:grad}), src/Fields_Operations.jl:84-112: sum_a u_a (x) dphi_a,
    returned as [g][d][i] (the memory order of SMatrix{D,dim})."""
    u = np.asarray(u, dtype=float).reshape(-1, D)
    dim = dphi.shape[1]
    out = np.zeros((off.size - 1, dim, D))
    for g in range(off.size - 1):
        for k in range(off[g], off[g + 1]):
            out[g] += np.outer(dphi[k], u[dom[k] - 1])
    return out


def csc_to_scipy(colptr, rowval, nzval, n):
    import scipy.sparse as sp
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(n, n))


def linear_problem_scalar(model, dmax, degree, bil_coef, lin_coef, dirichlet_tags, dirichlet_values=None,
                          shape="rectangular", technique="IMLS", dm_map=None):
    """Linear_Problem for a scalar field with penalty Dirichlet, src/Assembling_Operators.jl:167-267
    (α = maximum(A)*100, :238; boundary sets search ALL nodes, FlexiGal.jl:254)."""
    import scipy.sparse as sp
    dim = model.dim
    dm = influence_domains(model, dmax, shape, dm_map)
    gs, _ = background_integration(model, "Domain", degree)
    off, dom, phi, dphi = shape_fun(gs, model.x, dm, shape, technique, nthreads=os.cpu_count() or 1)
    cp, rv = pattern(model.nnodes, off, dom)
    A = csc_to_scipy(cp, rv, assemble_bilinear(gs, off, dom, phi, dphi, bil_coef, 1, cp, rv), model.nnodes)
    F = assemble_linear(gs, off, dom, phi, dphi, lin_coef, 1, model.nnodes) if lin_coef is not None else np.zeros(model.nnodes)
    alpha = max(A.max(), 0.0) * 100
    Fp = np.zeros(model.nnodes)
    parts = []
    vals = dirichlet_values if dirichlet_values else []
    for t, tag in enumerate(dirichlet_tags):
        gsb, _ = background_integration(model, tag, degree)
        ob, db, pb, dpb = shape_fun(gsb, model.x, dm, shape, technique, nthreads=os.cpu_count() or 1)
        parts.append((gsb, ob, db, pb, dpb))
        if t < len(vals):
            v = vals[t]
            cv = np.array([alpha * v(g[:dim]) for g in gsb]) if callable(v) else alpha * v
            Fp += assemble_linear(gsb, ob, db, pb, dpb, coef_source(dim, cv), 1, model.nnodes)
    if parts:
        gsb = np.asfortranarray(np.vstack([p[0] for p in parts]))
        ob = np.concatenate([[0]] + [p[1][1:] + sum(q[1][-1] for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.int64)
        db = np.concatenate([p[2] for p in parts]); pb = np.concatenate([p[3] for p in parts]); dpb = np.vstack([p[4] for p in parts])
        cpb, rvb = pattern(model.nnodes, ob, db)
        Ap = csc_to_scipy(cpb, rvb, assemble_bilinear(gsb, ob, db, pb, dpb, coef_mass(dim, alpha), 1, cpb, rvb), model.nnodes)
    else:
        Ap = sp.csc_matrix((model.nnodes, model.nnodes))
    K = (A + Ap).tocsc()
    K.eliminate_zeros()   # Julia's sparse `+` drops entries whose sum is exactly 0.0 (Assembling_Operators.jl:266)
    return K, F + Fp, dict(gs=gs, off=off, dom=dom, phi=phi, dphi=dphi, dm=dm, alpha=alpha)

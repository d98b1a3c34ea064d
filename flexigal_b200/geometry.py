"""Host-side setup that feeds the hot path: node cloud, tagged entities, support sizes, Gauss tables.

In a FlexiGal deployment these stay Julia (Meshing.jl / Geometry.jl / Integration.jl and
Influence_Domains run once and hand plain Float64 arrays across the C ABI).  This module is the
Python stand-in used by the tests and bench.py: vectorised NumPy, same numbering and the same
floating-point operation order as the reference so that the arrays crossing the ABI are the ones
Julia would pass (citations are file:line under /root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Model:
    """model{D}, src/Geometry.jl:5-18."""
    domain: tuple
    divisions: tuple
    x: np.ndarray        # nnodes x dim, column-major (possibly mapped)
    x0: np.ndarray       # unmapped coordinates
    conn: np.ndarray     # ncells x 2^dim, 1-based
    entities: dict = field(default_factory=dict)

    @property
    def dim(self):
        return self.x.shape[1]

    @property
    def nnodes(self):
        return self.x.shape[0]

    @property
    def ncells(self):
        return self.conn.shape[0]


def generate_cartesian_mesh(domain, divisions):
    """src/Meshing.jl:1-60: node idx = i + (Nx+1) j (+ (Nx+1)(Ny+1) k), coordinate (L/N)*i;
    cells x-fastest with the vertex order of :22-25 (quad) / :45-53 (hexa)."""
    dim = len(domain)
    if dim not in (2, 3):
        raise ValueError(f"Only 2D or 3D problems are supported. Detected dimension: {dim}.")
    n = [int(v) for v in divisions]
    step = [domain[d] / n[d] for d in range(dim)]
    axes = [step[d] * np.arange(n[d] + 1, dtype=np.float64) for d in range(dim)]
    if dim == 2:
        X = np.empty(((n[0] + 1) * (n[1] + 1), 2), order="F")
        X[:, 0] = np.tile(axes[0], n[1] + 1)
        X[:, 1] = np.repeat(axes[1], n[0] + 1)
        i, j = np.meshgrid(np.arange(n[0]), np.arange(n[1]), indexing="xy")
        n1 = (i + 1 + j * (n[0] + 1)).reshape(-1)
        conn = np.stack([n1, n1 + 1, n1 + 1 + n[0] + 1, n1 + n[0] + 1], axis=1).astype(np.int64)
    else:
        lay = (n[0] + 1) * (n[1] + 1)
        X = np.empty((lay * (n[2] + 1), 3), order="F")
        X[:, 0] = np.tile(axes[0], (n[1] + 1) * (n[2] + 1))
        X[:, 1] = np.tile(np.repeat(axes[1], n[0] + 1), n[2] + 1)
        X[:, 2] = np.repeat(axes[2], lay)
        k, j, i = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
        n0 = (i + 1 + j * (n[0] + 1) + k * lay).reshape(-1)
        n1, n3 = n0 + 1, n0 + n[0] + 1
        n2 = n1 + n[0] + 1
        conn = np.stack([n0, n1, n2, n3, n0 + lay, n1 + lay, n2 + lay, n3 + lay], axis=1).astype(np.int64)
    return X, np.asfortranarray(conn)


def _boundary_entities(conn, local_faces):
    """External sub-entities = those owned by exactly one cell (src/Meshing.jl:120-212), kept in
    order of appearance (cell-major, local face minor) with the node order of the owning cell."""
    faces = np.stack([conn[:, f] for f in local_faces], axis=1).reshape(-1, len(local_faces[0]))
    key = np.sort(faces, axis=1)
    # hash rows to one integer for a fast unique (node ids < 2^31, up to 4 per face -> use structured view)
    view = np.ascontiguousarray(key).view([("", key.dtype)] * key.shape[1]).reshape(-1)
    _, inv, cnt = np.unique(view, return_inverse=True, return_counts=True)
    return faces[cnt[inv] == 1]


def create_model(domain, divisions, map=None, boundary=True):
    """create_model, src/Geometry.jl:20-276: mesh, boundary tags decided on the UNMAPPED box
    (tolerance 1e-6), optional coordinate map applied afterwards.  ``boundary=False`` skips the
    boundary-entity extraction (only the "Domain" tag exists then)."""
    x, conn = generate_cartesian_mesh(domain, divisions)
    dim = len(domain)
    tol = 1e-6
    ent = {}
    if not boundary:
        ent["faces" if dim == 2 else "volumes"] = {"Domain": conn}
    elif dim == 2:
        ext = _boundary_entities(conn, [[0, 1], [1, 2], [2, 3], [3, 0]])
        mid = 0.5 * (x[ext[:, 0] - 1] + x[ext[:, 1] - 1])
        sel = np.full(len(ext), -1)
        for code, m in enumerate([mid[:, 0] < tol, mid[:, 0] > domain[0] - tol, mid[:, 1] < tol, mid[:, 1] > domain[1] - tol]):
            sel[(sel < 0) & m] = code
        names = ["Left", "Right", "Bottom", "Top"]
        ent["ext_edges"] = {nm: ext[sel == c] for c, nm in enumerate(names)}
        ent["ext_edges"]["Boundary"] = ext
        ent["faces"] = {"Domain": conn}
    else:
        ext = _boundary_entities(conn, [[0, 3, 7, 4], [1, 2, 6, 5], [0, 1, 5, 4], [2, 6, 7, 3], [0, 1, 2, 3], [4, 5, 6, 7]])
        mid = 0.25 * (x[ext[:, 0] - 1] + x[ext[:, 1] - 1] + x[ext[:, 2] - 1] + x[ext[:, 3] - 1])
        sel = np.full(len(ext), -1)
        tests = [mid[:, 0] < tol, mid[:, 0] > domain[0] - tol, mid[:, 2] < tol, mid[:, 2] > domain[2] - tol,
                 mid[:, 1] < tol, mid[:, 1] > domain[1] - tol]
        for code, m in enumerate(tests):
            sel[(sel < 0) & m] = code
        names = ["Left", "Right", "Bottom", "Top", "Back", "Front"]
        ent["ext_faces"] = {nm: ext[sel == c] for c, nm in enumerate(names)}
        ent["ext_faces"]["Boundary"] = ext
        ent["volumes"] = {"Domain": conn}
    x0 = x.copy(order="F")
    if map is not None:
        x = np.asfortranarray(np.stack([np.asarray(map(x0[i, :]), dtype=float) for i in range(x0.shape[0])]))
    return Model(tuple(domain), tuple(divisions), x, x0, conn, ent)


def set_tag(model: Model, region, name="NewTag"):
    """set_tag!, src/Geometry.jl:303-328: a sub-domain tag = the cells whose vertex centroid (sequential sum over the
    cell's vertices, then one division) satisfies ``region``; kept in cell order."""
    cat = "faces" if model.dim == 2 else "volumes"
    nv = model.conn.shape[1]
    cen = np.zeros((model.ncells, model.dim))
    for v in range(nv):
        cen += model.x[model.conn[:, v] - 1, :]
    cen /= nv
    sel = np.array([bool(region(cen[e, :])) for e in range(model.ncells)])
    if not sel.any():
        return None
    model.entities.setdefault(cat, {})[name] = model.conn[sel, :]
    return int(sel.sum())


def get_entity_info(model: Model, tag: str):
    """src/Geometry.jl:288-301 -> (connectivity, category)."""
    order = ["faces", "ext_edges"] if model.dim == 2 else ["volumes", "ext_faces"]
    for cat in order:
        if cat in model.entities and tag in model.entities[cat]:
            return model.entities[cat][tag], cat
    raise KeyError(f"tag '{tag}' not found")


def Influence_Domains(model: Model, dmax, shape="rectangular", dm_map=None):
    """Influence_Domains, src/FlexiGal.jl:164-199."""
    dim = model.dim
    dvec = [float(dmax)] * dim if np.isscalar(dmax) else [float(v) for v in dmax]
    if shape == "rectangular":
        dm = np.empty((model.nnodes, dim), order="F")
        for d in range(dim):
            dm[:, d] = dvec[d] * model.domain[d] / model.divisions[d]
        if dm_map is not None:
            for i in range(model.nnodes):
                s = dm_map(model.x0[i, :])
                dm[i, :] *= (s if np.isscalar(s) else np.asarray(s, dtype=float))
        return dm
    if shape in ("circular", "spherical"):
        spacing = sum(model.domain[d] / model.divisions[d] for d in range(dim)) / dim
        return np.full(model.nnodes, dvec[0] * spacing)
    raise ValueError(f"Unknown shape of influence domain: {shape}. Shapes supported are: rectangular, circular, spherical, cylindrical.")


def pgauss(k: int):
    """Gauss-Legendre points/weights, src/Integration.jl:1-23."""
    if k == 1:
        return np.array([[0.0], [2.0]])
    if k == 2:
        a = 1 / np.sqrt(3.0)
        return np.array([[-a, a], [1.0, 1.0]])
    if k == 3:
        a = np.sqrt(0.6)
        return np.array([[-a, 0.0, a], [5 / 9, 8 / 9, 5 / 9]])
    if k == 4:
        return np.array([[-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526],
                         [0.3478548451374539, 0.6521451548625461, 0.6521451548625461, 0.3478548451374539]])
    raise ValueError("Gauss integrations just implemented for ngpts = 1, 2, 3 o 4")


def _quad_measure(xi, eta, xe):
    """Bilinear map of a quad cell/face at (xi,eta): point and measure, src/Integration.jl:25-74
    (the four vertex terms are accumulated in vertex order, as the reference's `.+=` does)."""
    N = (0.25 * (1 - xi) * (1 - eta), 0.25 * (1 + xi) * (1 - eta), 0.25 * (1 + xi) * (1 + eta), 0.25 * (1 - xi) * (1 + eta))
    dxi = (-0.25 * (1 - eta), 0.25 * (1 - eta), 0.25 * (1 + eta), -0.25 * (1 + eta))
    det = (-0.25 * (1 - xi), -0.25 * (1 + xi), 0.25 * (1 + xi), 0.25 * (1 - xi))
    ncell, _, dim = xe.shape
    xg = np.zeros((ncell, dim)); a = np.zeros((ncell, dim)); b = np.zeros((ncell, dim))
    for v in range(4):
        xg += N[v] * xe[:, v, :]
        a += dxi[v] * xe[:, v, :]
        b += det[v] * xe[:, v, :]
    if dim == 2:
        jac = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    else:
        sx = a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2]
        sy = b[:, 0] * a[:, 2] - a[:, 0] * b[:, 2]
        sz = a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]
        jac = np.sqrt(sx ** 2 + sy ** 2 + sz ** 2)
    return xg, jac


_HEX_SIGNS = ((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1))


def _hex_measure(xi, eta, zeta, xe):
    """Trilinear map of a hexahedral cell, src/Integration.jl:76-146.  ``xe`` is (8, 3, ncell)
    (vertex, coordinate, cell) so that every accumulation below streams over contiguous memory;
    the eight vertex terms are added in vertex order, as the reference's `.+=` does."""
    N = [(1 / 8) * (1 + s * xi) * (1 + t * eta) * (1 + u * zeta) for s, t, u in _HEX_SIGNS]
    dN = ([s * (1 / 8) * (1 + t * eta) * (1 + u * zeta) for s, t, u in _HEX_SIGNS],
          [t * (1 / 8) * (1 + s * xi) * (1 + u * zeta) for s, t, u in _HEX_SIGNS],
          [u * (1 / 8) * (1 + s * xi) * (1 + t * eta) for s, t, u in _HEX_SIGNS])
    ncell = xe.shape[2]
    xg = np.zeros((3, ncell)); J = np.zeros((3, 3, ncell)); tmp = np.empty(ncell)
    for v in range(8):
        for i in range(3):
            np.multiply(xe[v, i], N[v], out=tmp); xg[i] += tmp
            for j in range(3):
                np.multiply(xe[v, i], dN[j][v], out=tmp); J[i, j] += tmp
    jac = (J[0, 0] * (J[1, 1] * J[2, 2] - J[2, 1] * J[1, 2]) - J[0, 1] * (J[1, 0] * J[2, 2] - J[2, 0] * J[1, 2])
           + J[0, 2] * (J[1, 0] * J[2, 1] - J[2, 0] * J[1, 1]))
    return xg, jac


def egauss(x, conn, gauss, out=None):
    """Gauss table of domain cells, src/Integration.jl:148-182: row (e-1)*l^dim + count holds
    [x_g..., w, detJ]; count runs i fastest, then j, then k."""
    dim, l = x.shape[1], gauss.shape[1]
    npts = l ** dim
    gs = np.empty((conn.shape[0] * npts, dim + 2), order="F") if out is None else out
    c = 0
    if dim == 2:
        xe = x[conn - 1, :]
        for j in range(l):
            for i in range(l):
                xg, jac = _quad_measure(gauss[0, i], gauss[0, j], xe)
                gs[c::npts, :2] = xg; gs[c::npts, 2] = gauss[1, i] * gauss[1, j]; gs[c::npts, 3] = jac
                c += 1
    else:
        xe = np.ascontiguousarray(np.stack([x[conn[:, v] - 1, :].T for v in range(8)]))   # (8, 3, ncell)
        for k in range(l):
            for j in range(l):
                for i in range(l):
                    xg, jac = _hex_measure(gauss[0, i], gauss[0, j], gauss[0, k], xe)
                    for d in range(3):
                        gs[c::npts, d] = xg[d]
                    gs[c::npts, 3] = gauss[1, i] * gauss[1, j] * gauss[1, k]; gs[c::npts, 4] = jac
                    c += 1
    return gs


def egauss_bound(x, conn, gauss):
    """Gauss table of boundary edges/faces, src/Integration.jl:184-218."""
    dim, l = x.shape[1], gauss.shape[1]
    npts = l ** (dim - 1)
    gs = np.empty((conn.shape[0] * npts, dim + 2), order="F")
    xe = x[conn - 1, :]
    c = 0
    if dim == 2:
        for i in range(l):
            xi = gauss[0, i]
            gs[c::l, 0] = (1 - xi) / 2 * xe[:, 0, 0] + (1 + xi) / 2 * xe[:, 1, 0]
            gs[c::l, 1] = (1 - xi) / 2 * xe[:, 0, 1] + (1 + xi) / 2 * xe[:, 1, 1]
            gs[c::l, 2] = gauss[1, i]
            gs[c::l, 3] = np.sqrt((xe[:, 0, 0] - xe[:, 1, 0]) ** 2 + (xe[:, 1, 1] - xe[:, 0, 1]) ** 2) / 2
            c += 1
    else:
        for j in range(l):
            for i in range(l):
                xg, jac = _quad_measure(gauss[0, i], gauss[0, j], xe)
                gs[c::npts, :3] = xg; gs[c::npts, 3] = gauss[1, i] * gauss[1, j]; gs[c::npts, 4] = jac
                c += 1
    return gs


@dataclass(frozen=True)
class Triangulation:
    """src/FlexiGal.jl:12-15."""
    model: Model = field(compare=False)
    tag: str = "Domain"

    def __hash__(self):
        return hash((id(self.model), self.tag))

    def __eq__(self, o):
        return isinstance(o, Triangulation) and o.model is self.model and o.tag == self.tag


@dataclass(frozen=True)
class IntegrationSet:
    """src/FlexiGal.jl:16-19."""
    tri: Triangulation
    degree: int


@dataclass
class DomainMeasure:
    """src/FlexiGal.jl:21-25."""
    tag: str
    gs: np.ndarray
    degree: int
    is_boundary: bool = False


def BackgroundIntegration(iset: IntegrationSet, method: str = "EFG") -> DomainMeasure:
    """BackgroundIntegration(:EFG), src/FlexiGal.jl:27-52."""
    if method != "EFG":
        raise ValueError(f"Método desconocido: {method} (only the EFG path is accelerated)")
    model = iset.tri.model
    conn, cat = get_entity_info(model, iset.tri.tag)
    boundary = cat != ("faces" if model.dim == 2 else "volumes")
    g = pgauss(iset.degree)
    gs = (egauss_bound if boundary else egauss)(model.x, conn, g)
    return DomainMeasure(iset.tri.tag, gs, iset.degree, boundary)

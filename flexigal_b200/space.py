"""Host-side mirror of the reference's hot-path interface, on top of the C ABI.

Same names and argument meaning as the Julia functions they stand in for (file:line under
/root/reference):

    SHAPE_FUN(gs, x, dm; shape, technique) -> (PHI, DPHI, DOM)     src/Shape_Functions.jl:241
    build_space(recipe, isets) -> FlexiSpace                       src/FlexiGal.jl:226
    Bilinear_Assembler(f, Space) -> SparseMatrixCSC                src/Assembling_Operators.jl:81
    Linear_Assembler(f, Space) -> Vector                           src/Assembling_Operators.jl:147
    Linear_Problem(a, b, recipe) -> (A, F, Spaces)                 src/Assembling_Operators.jl:167

Differences forced by the device boundary: user closures cannot run on the
GPU, so ``f`` is an operator descriptor (``operators.py``: closed forms, a coefficient tensor, or
a Python closure that is *probed* into a coefficient tensor); and the ragged
Vector{Vector{..}} containers are flat arrays + offsets that stay on the device until asked for.

No CPU fallback exists: every function here ends in a CUDA kernel launch through
``libflexigal_gpu.so`` and raises if the library or the GPU is missing.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .geometry import (BackgroundIntegration, DomainMeasure, Influence_Domains, IntegrationSet, Model, Triangulation,
                       get_entity_info)
from .operators import BilinearOp, LinearOp, Mass, Source, as_bilinear, as_linear


class NodeCloud:
    """(x, dm, shape) resident on one GPU, with its bin grid: what SHAPE_FUN receives."""

    def __init__(self, x, dm, shape="rectangular", device=0, stream=None):
        self.ctx = _lib.Context(device)
        if stream is not None:
            self.ctx.set_stream(stream)
        self.ctx.set_nodes(x, dm, shape)
        self.shape = shape
        self.nnodes, self.dim = self.ctx.nnodes, self.ctx.dim

    def close(self):
        self.ctx.close()


class ShapeSet:
    """(PHI, DPHI, DOM) of one Gauss table, device resident.  ``offsets`` (0-based, ngauss+1),
    ``DOM`` (1-based Int64, ascending per point), ``PHI``, ``DPHI`` (total_L x dim) are fetched
    lazily; ``PHI_list()`` etc. give the reference's ragged Vector{Vector} view."""

    def __init__(self, cloud: NodeCloud, set_id: int, ngauss: int, total_L: int, max_L: int, gs=None):
        self.cloud, self.set_id, self.ngauss, self.total_L, self.max_L = cloud, set_id, ngauss, total_L, max_L
        self.gs = gs
        self._host = None
        self._pattern = {}

    def _fetch(self):
        if self._host is None:
            self._host = self.cloud.ctx.get_shapes(self.set_id, self.ngauss, self.total_L)
        return self._host

    offsets = property(lambda s: s._fetch()[0])
    DOM = property(lambda s: s._fetch()[1])
    PHI = property(lambda s: s._fetch()[2])
    DPHI = property(lambda s: s._fetch()[3])

    def ragged(self):
        off, dom, phi, dphi = self._fetch()
        sl = [slice(off[g], off[g + 1]) for g in range(self.ngauss)]
        return [phi[s] for s in sl], [dphi[s] for s in sl], [dom[s] for s in sl]

    def singular_count(self):
        return self.cloud.ctx.singular_count(self.set_id)

    def pattern(self, D=1):
        """DOF-level CSC pattern (1-based colptr, rowval) of the operators assembled on this set."""
        if D not in self._pattern:
            ctx = self.cloud.ctx
            nnz = ctx.pattern(self.set_id)
            self._pattern[D] = ctx.get_pattern(self.set_id, D, nnz)
        return self._pattern[D]


def SHAPE_FUN(gs, x=None, dm=None, shape="rectangular", technique="IMLS", cloud: NodeCloud | None = None, device=0,
              node_ids=None) -> ShapeSet:
    """SHAPE_FUN, src/Shape_Functions.jl:241-263: neighbour search + shape functions for every row
    of ``gs``.  Pass ``cloud`` to reuse a node cloud already on the GPU.  ``node_ids`` (1-based) is build_space's
    ``SHAPE_FUN(gs, x[ids,:], dm[ids,:])`` + ``DOM = ids[DOM_local]`` (src/FlexiGal.jl:250-252) done on the device:
    only those nodes are candidates, DOM comes back in the numbering of the full cloud."""
    if cloud is None:
        cloud = NodeCloud(x, dm, shape, device)
    ctx = cloud.ctx
    if technique not in _lib.TECHNIQUES:
        ctx.shape_functions(-1, technique)   # raises the reference's message
    sid = ctx.new_set_id()
    ng = ctx.set_gauss(sid, gs)
    if node_ids is not None:
        ctx.set_node_subset(sid, node_ids)
    total_L, max_L = ctx.neighbours(sid)
    ctx.shape_functions(sid, technique)
    return ShapeSet(cloud, sid, ng, total_L, max_L, gs=np.asfortranarray(gs, dtype=np.float64))


@dataclass
class ApproxSpace:
    """ApproxSpace recipe, src/FlexiGal.jl:63-105 (Field_Type reduced to the field dimension D)."""
    model: Model
    triangulations: list
    D: int = 1
    dmax: object = 1.5
    shape: str = "rectangular"
    technique: str = "IMLS"
    dm_map: object = None
    Dirichlet_Boundaries: list = field(default_factory=list)
    Dirichlet_Values: list = field(default_factory=list)
    methods: dict = field(default_factory=dict)     # tag -> "FEM" | "EFG" (default EFG): ApproxSpace(...; method=[...]), src/FlexiGal.jl:63-105
    device: int = 0
    _cloud: object = None
    _dm: object = None

    def __post_init__(self):
        if isinstance(self.triangulations, Triangulation):
            self.triangulations = [self.triangulations]
        if isinstance(self.Dirichlet_Boundaries, Triangulation):
            self.Dirichlet_Boundaries = [self.Dirichlet_Boundaries]

    def dm(self):
        if self._dm is None:
            self._dm = Influence_Domains(self.model, self.dmax, self.shape, self.dm_map)
        return self._dm

    def cloud(self) -> NodeCloud:
        if self._cloud is None:
            self._cloud = NodeCloud(self.model.x, self.dm(), self.shape, self.device)
        return self._cloud


@dataclass
class FlexiSpace:
    """FlexiSpace, src/FlexiGal.jl:107-121: per-tag shape functions + the measures, in order."""
    sets: dict                   # tag -> ShapeSet
    D: int
    Measures: list               # DomainMeasure, in the order given
    nnodes: int
    cloud: NodeCloud
    _merged: object = None

    def merged(self) -> ShapeSet:
        """Flexi_Measure, src/FlexiGal.jl:207-224: the measures' shape functions concatenated."""
        if self._merged is None:
            parts = [self.sets[m.tag] for m in self.Measures]
            if len(parts) == 1:
                self._merged = parts[0]
            else:
                ctx = self.cloud.ctx
                dst = ctx.new_set_id()
                ctx.concat_sets([p.set_id for p in parts], dst)
                gs = np.asfortranarray(np.vstack([p.gs for p in parts]))
                self._merged = ShapeSet(self.cloud, dst, sum(p.ngauss for p in parts), sum(p.total_L for p in parts),
                                        max(p.max_L for p in parts), gs=gs)
        return self._merged


def build_space(recipe: ApproxSpace, isets) -> FlexiSpace:
    """build_space, src/FlexiGal.jl:226-262: ONE device context per recipe (``recipe.cloud()``), one set id per tag.
    Domain sets restrict the candidate nodes to those of the tagged cells (:250-252, ``fg_set_node_subset``: the search
    runs on the subset, DOM and the sparsity come back in global numbering); boundary sets search all nodes (:254)."""
    m = recipe.model
    cloud = recipe.cloud()
    sets, measures = {}, []
    for iset in isets:
        tag = iset.tri.tag
        if recipe.methods.get(tag, "EFG") == "FEM":
            # FEM branch (src/FlexiGal.jl:245-246): element shape functions at the cell's Gauss points, built on the device
            conn, _ = get_entity_info(m, tag)
            ctx = cloud.ctx
            sid = ctx.new_set_id()
            from .geometry import pgauss
            ng, total_L, gs = ctx.generate_fem(sid, conn, pgauss(iset.degree))
            sets[tag] = ShapeSet(cloud, sid, ng, total_L, conn.shape[1], gs=gs)
            measures.append(DomainMeasure(tag, gs, iset.degree, conn.shape[1] != 2 ** m.dim))
            continue
        meas = BackgroundIntegration(iset, "EFG")
        ids = None
        if not meas.is_boundary:
            conn, _ = get_entity_info(m, tag)
            ids = np.unique(conn.reshape(-1))            # sort!(unique(view(conn, :)))
            if ids.size == m.nnodes:
                ids = None                               # every node: same search as an unrestricted one
        sets[tag] = SHAPE_FUN(meas.gs, technique=recipe.technique, cloud=cloud, node_ids=ids)
        measures.append(meas)
    return FlexiSpace(sets, recipe.D, measures, m.nnodes, cloud)


def _to_csc(colptr, rowval, nzval, n):
    import scipy.sparse as sp
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(n, n))


def Bilinear_Assembler(f, Space: FlexiSpace, out=None):
    """Bilinear_Assembler, src/Assembling_Operators.jl:81-104 -> scipy CSC matrix (ndofs x ndofs),
    explicit zeros kept like sparse()."""
    return _assemble_bilinear_set(f, Space.merged(), Space.D)


def _assemble_bilinear_set(f, S: ShapeSet, D):
    ctx = S.cloud.ctx
    op: BilinearOp = as_bilinear(f, S.cloud.dim, D, S.gs)
    colptr, rowval = S.pattern(D)
    nz = np.empty(rowval.size)
    ctx.assemble_bilinear_async(S.set_id, op.c_struct(S.cloud.dim, S.ngauss), nz)     # download pipelined behind the assembly where the path allows
    ctx.sync()
    return _to_csc(colptr, rowval, nz, S.cloud.nnodes * D)


def Linear_Assembler(f, Space: FlexiSpace):
    """Linear_Assembler, src/Assembling_Operators.jl:147-165 -> ndofs vector."""
    return _assemble_linear_set(f, Space.merged(), Space.D)


def _assemble_linear_set(f, S: ShapeSet, D):
    ctx = S.cloud.ctx
    op: LinearOp = as_linear(f, S.cloud.dim, D, S.gs)
    F = np.empty(S.cloud.nnodes * D)
    ctx.assemble_linear(S.set_id, op.c_struct(S.cloud.dim, S.ngauss), F)
    return F


def Linear_Problem(bilinear_terms, linear_terms, recipe: ApproxSpace, alpha=None):
    """Linear_Problem, src/Assembling_Operators.jl:167-267.  ``*_terms`` are lists of
    (operator, IntegrationSet) -- the Integrated / MultiIntegrated terms of the weak form.
    Penalty Dirichlet: alpha = maximum(A)*100 (:238), boundary sets search all nodes, sparse `+`
    drops exact-zero sums (:266)."""
    import scipy.sparse as sp
    D, ndofs = recipe.D, recipe.D * recipe.model.nnodes
    Spaces: dict = {}
    max_deg = 1

    def space_of(iset):
        if iset not in Spaces:
            Spaces[iset] = build_space(recipe, [iset])
        return Spaces[iset]

    A = sp.csc_matrix((ndofs, ndofs))
    for f, iset in bilinear_terms:
        max_deg = max(max_deg, iset.degree)
        A = A + Bilinear_Assembler(f, space_of(iset))
    F = np.zeros(ndofs)
    for f, iset in (linear_terms or []):
        max_deg = max(max_deg, iset.degree)
        F = F + Linear_Assembler(f, space_of(iset))
    Ap = sp.csc_matrix((ndofs, ndofs))
    Fp = np.zeros(ndofs)
    if alpha is None:                       # Linear_Problem: maximum(A)*100 (:238); NL_Solver passes its fixed 1e6 (src/Solvers.jl:139)
        alpha = max(A.max(), 0.0) * 100 if A.nnz else 0.0
    if recipe.Dirichlet_Boundaries:
        dsets = [IntegrationSet(tri, max_deg) for tri in recipe.Dirichlet_Boundaries]
        if dsets[0] not in Spaces:
            Space_D = build_space(recipe, dsets)
            for s in dsets:
                Spaces[s] = Space_D
        else:
            Space_D = Spaces[dsets[0]]
        for diri, meas in zip(recipe.Dirichlet_Values, Space_D.Measures):
            tmp = FlexiSpace(Space_D.sets, D, [meas], Space_D.nnodes, Space_D.cloud)
            if callable(diri):
                vals = np.array([diri(g) for g in meas.gs[:, :recipe.model.dim]], dtype=float)
                src = Source(alpha * vals)
            else:
                src = Source(alpha * np.asarray(diri, dtype=float))
            Fp = Fp + Linear_Assembler(src, tmp)
        Ap = Bilinear_Assembler(Mass(alpha), Space_D)
    K = (A + Ap).tocsc()
    K.eliminate_zeros()
    return K, F + Fp, Spaces


def Solve(A, F, cloud: NodeCloud | None = None, rtol=1e-13, max_iter=50000):
    """Solve, src/Solvers.jl:24-31: u = A \\ F.  On the host (SciPy direct solve) like the reference, or -- SURVEY 8(f)
    row f3 -- with ``cloud`` given, Jacobi-preconditioned CG on that cloud's GPU (symmetric positive definite A)."""
    if cloud is not None:
        x, it, res = cloud.ctx.solve_spd(A, F, rtol, max_iter)
        if not res <= max(10 * rtol, 1e-9):
            raise RuntimeError(f"device CG stopped at relative residual {res:.3e} after {it} iterations")
        return x
    from scipy.sparse.linalg import spsolve
    return spsolve(A.tocsc(), F)


def Solve_On_Device(Spaces, F, recipe: ApproxSpace, rtol=1e-13, max_iter=50000):
    """SURVEY 8(f) f3 with K resident: u = (sum of the operators LAST assembled on each space of ``Spaces``) \\ F.
    After ``Linear_Problem`` with one bilinear term per integration set (the Poisson / elasticity examples) that sum is
    exactly ``A + Ap`` (src/Assembling_Operators.jl:266): the domain operator lives on the domain set, the penalty mass
    matrix on the merged Dirichlet set.  Patterns and values never leave the device; only F goes up and u comes down."""
    seen, sids = set(), []
    for sp_ in (Spaces.values() if isinstance(Spaces, dict) else Spaces):
        if id(sp_) in seen:
            continue
        seen.add(id(sp_))
        sids.append(sp_.merged().set_id)
    x, it, res = recipe.cloud().ctx.solve_spd_sets(sids, F, recipe.D, None, rtol, max_iter)
    if not res <= max(10 * rtol, 1e-9):
        raise RuntimeError(f"device CG stopped at relative residual {res:.3e} after {it} iterations")
    return x, it, res


def Get_Point_Values(u_nodal, Space: FlexiSpace, gradient=False):
    """Get_Point_Values of FlexiFunction(u, Space, Measure) / of its gradient (src/Fields_Operations.jl:40-112):
    the field (ngauss x D) or its gradient (ngauss x dim x D, SMatrix{D,dim} memory order) at the Gauss points."""
    S = Space.merged()
    val, grad = S.cloud.ctx.evaluate_field(S.set_id, S.ngauss, u_nodal, Space.D, values=not gradient, gradients=gradient)
    return grad if gradient else val

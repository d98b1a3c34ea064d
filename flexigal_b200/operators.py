"""Integrand descriptors: what crosses the C ABI in place of the reference's Julia closures.

The reference calls a user closure ``f_pure(sa, sb)`` for every neighbour pair of every Gauss
point (src/Assembling_Operators.jl:18, :62) and ``f_pure(sa)`` for linear forms (:118, :137).
Every closure the weak-form DSL can build is bilinear (linear) in the per-node operands
B = [phi, d1 phi, ..], so it is fully described by a coefficient tensor per Gauss point
(SURVEY 7.2 #3).  The closed forms below are the closures of the shipped examples
(arithmetic from src/Fields_Operations.jl); ``Closure``/``LinearClosure`` probe an arbitrary
Python closure with unit operands into a tensor, which is how the Julia shim handles a closure
it does not recognise.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib


class BilinearOp:
    kind = _lib.OP_GENERIC

    def __init__(self, D=1):
        self.D = D

    def constants(self, dim):
        return [0.0] * 4

    def coefficients(self, dim, ngauss):
        return None, 0

    def c_struct(self, dim, ngauss):
        op = _lib.fg_operator()
        op.kind, op.D = self.kind, self.D
        for k, v in enumerate(self.constants(dim)):
            op.c[k] = v
        coef, stride = self.coefficients(dim, ngauss)
        self._keep = coef
        op.coef = coef.ctypes.data if coef is not None else None
        op.coef_stride = stride
        return op


class GradGrad(BilinearOp):
    """k * grad(dT).grad(T)  (Fields_Operations.jl:321-323, :1002); vector fields: :552-556."""
    kind = _lib.OP_GRADGRAD

    def __init__(self, k=1.0, D=1):
        super().__init__(D)
        self.k = float(k)

    def constants(self, dim):
        return [self.k, 0.0, 0.0, 0.0]


class Mass(BilinearOp):
    """c * dT*T  /  c * (du . u)  (:1004, :546-550): the penalty term of Linear_Problem."""
    kind = _lib.OP_MASS

    def __init__(self, c=1.0, D=1):
        super().__init__(D)
        self.c = float(c)

    def constants(self, dim):
        return [self.c, 0.0, 0.0, 0.0]


class Advection(BilinearOp):
    """dT * (v . grad T) with a constant velocity (examples/Poisson2D_No_Source.jl:30)."""
    kind = _lib.OP_ADVECTION

    def __init__(self, v):
        super().__init__(1)
        self.v = [float(t) for t in v]

    def constants(self, dim):
        return (self.v + [0.0] * 4)[:4]


class Elasticity(BilinearOp):
    """grad(du) (.) sigma(u), sigma = G (grad u + grad u^T) + lambda (div u) Id
    (examples/Elasticity_Test.jl:25-26)."""
    kind = _lib.OP_ELASTICITY

    def __init__(self, G, lam, D):
        super().__init__(D)
        self.G, self.lam = float(G), float(lam)

    def constants(self, dim):
        return [self.G, self.lam, 0.0, 0.0]


class Generic(BilinearOp):
    """Coefficient tensor C[(i,al),(j,be)]: shape (n,n) constant or (ngauss,n,n), n = D*(1+dim)."""

    def __init__(self, C, D=1):
        super().__init__(D)
        self.C = np.ascontiguousarray(C, dtype=np.float64)

    def coefficients(self, dim, ngauss):
        n = self.D * (1 + dim)
        if self.C.size == n * n:
            return self.C.reshape(-1), 0
        if self.C.size != ngauss * n * n:
            raise ValueError(f"coefficient tensor has {self.C.size} entries, expected {n * n} or {ngauss}*{n * n}")
        return self.C.reshape(-1), n * n


@dataclass
class _Operand:
    """SingleFlexiMeasure (src/Fields_Operations.jl:304-309) as seen by a probed closure."""
    phi: float
    dphi: np.ndarray
    coord: np.ndarray
    ind: int


class Closure(BilinearOp):
    """A Python closure f(sa, sb) -> scalar (D=1) probed into a coefficient tensor with unit
    operands, once per Gauss point when ``per_point`` (the closure looks at .coord/.ind)."""

    def __init__(self, f, per_point=False):
        super().__init__(1)
        self.f, self.per_point = f, per_point
        self.gs = None

    def coefficients(self, dim, ngauss):
        nb = 1 + dim

        def unit(k, coord, ind):
            return _Operand(1.0 if k == 0 else 0.0, np.eye(dim)[k - 1] if k else np.zeros(dim), coord, ind)

        def tensor(coord, ind):
            return np.array([[self.f(unit(a, coord, ind), unit(b, coord, ind)) for b in range(nb)] for a in range(nb)], dtype=float)

        if not self.per_point:
            return np.ascontiguousarray(tensor(np.zeros(dim), 0)).reshape(-1), 0
        C = np.stack([tensor(self.gs[g, :dim], g) for g in range(ngauss)])
        return np.ascontiguousarray(C).reshape(-1), nb * nb


class LinearOp:
    kind = _lib.LIN_GENERIC

    def __init__(self, D=1):
        self.D = D

    def constants(self, dim):
        return [0.0] * 4

    def coefficients(self, dim, ngauss):
        return None, 0

    c_struct = BilinearOp.c_struct


class Source(LinearOp):
    """c * dT (examples/Poisson3D_Source.jl:13); c scalar / length-D constant, or per Gauss point
    (ngauss,) / (ngauss, D) -- e.g. the penalty load alpha*g(x) (Assembling_Operators.jl:250-259)."""

    def __init__(self, c, D=1):
        super().__init__(D)
        self.c = np.asarray(c, dtype=np.float64)

    def _constant(self):
        return self.c.ndim == 0 or (self.D > 1 and self.c.shape == (self.D,))

    @property
    def kind(self):
        return _lib.LIN_SOURCE if self._constant() else _lib.LIN_GENERIC

    def constants(self, dim):
        if not self._constant():
            return [0.0] * 4
        v = np.broadcast_to(self.c, (self.D,)).tolist()
        return (v + [0.0] * 4)[:4]

    def coefficients(self, dim, ngauss):
        if self._constant():
            return None, 0
        nb = 1 + dim
        c = self.c.reshape(ngauss, -1)
        if c.shape[1] not in (1, self.D):
            raise ValueError("per-point source must be (ngauss,) or (ngauss, D)")
        out = np.zeros((ngauss, self.D * nb))
        for i in range(self.D):
            out[:, i * nb] = c[:, i if c.shape[1] > 1 else 0]
        return np.ascontiguousarray(out).reshape(-1), self.D * nb


class GenericLinear(LinearOp):
    """Coefficient vector c[(i,al)]: (D*nb,) constant or (ngauss, D*nb)."""

    def __init__(self, c, D=1):
        super().__init__(D)
        self.c = np.ascontiguousarray(c, dtype=np.float64)

    def coefficients(self, dim, ngauss):
        n = self.D * (1 + dim)
        if self.c.size == n:
            return self.c.reshape(-1), 0
        if self.c.size != ngauss * n:
            raise ValueError(f"coefficient vector has {self.c.size} entries, expected {n} or {ngauss}*{n}")
        return self.c.reshape(-1), n


def as_bilinear(f, dim, D, gs) -> BilinearOp:
    if isinstance(f, BilinearOp):
        if isinstance(f, Closure):
            f.gs = gs
        if f.D != D and not isinstance(f, (Generic, Closure)):
            f.D = D
        return f
    if callable(f):
        c = Closure(f)
        c.gs = gs
        return c
    raise TypeError("bilinear integrand must be an operator descriptor or a closure f(sa, sb)")


def as_linear(f, dim, D, gs) -> LinearOp:
    if isinstance(f, LinearOp):
        if f.D != D:
            f.D = D
        return f
    if np.isscalar(f):
        return Source(f, D)
    raise TypeError("linear integrand must be an operator descriptor or a constant")

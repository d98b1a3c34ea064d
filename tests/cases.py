"""Shared problem definitions of the parity tests (BASELINE.json configs at oracle-friendly sizes)."""
import numpy as np


def poisson2d(fg, n=50, dmax=1.5):
    """C1: examples/Poisson2D_Source.jl:2-10 (Domain (1,1), Divisions (50,50), dmax 1.5, degree 3)."""
    model = fg.create_model((1.0, 1.0), (n, n))
    return model, dmax, 3


def poisson3d(fg, n=12, dmax=1.35):
    """C2: examples/Poisson3D_Source.jl:2-10 (Domain (1,1,1), Divisions (12,12,12), dmax 1.35, degree 3)."""
    model = fg.create_model((1.0, 1.0, 1.0), (n, n, n))
    return model, dmax, 3


def jitter(model, amount=0.2, seed=0):
    """Irregular variant: interior nodes moved by amount*h*U(-1,1)."""
    rng = np.random.default_rng(seed)
    x = model.x.copy(order="F")
    h = np.array([model.domain[d] / model.divisions[d] for d in range(model.dim)])
    interior = np.all((model.x0 > 1e-12) & (model.x0 < np.array(model.domain) - 1e-12), axis=1)
    x[interior] += amount * h * rng.uniform(-1, 1, size=(int(interior.sum()), model.dim))
    model.x = np.asfortranarray(x)
    return model


def rel(a, b):
    """max |a-b| relative to max |b| (the norm SURVEY 7.2 #1 reports)."""
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))

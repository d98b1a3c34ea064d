"""flexigal_b200 -- the Element-Free-Galerkin hot path of FlexiGal on NVIDIA B200 (sm_100a).

csrc/      hand-written CUDA kernels + the C ABI (include/flexigal_gpu.h) -> lib/libflexigal_gpu.so
_lib.py    ctypes binding of exactly those symbols (what the Julia shim `ccall`s)
space.py   host-side mirror of the reference interface: SHAPE_FUN, build_space,
           Bilinear_Assembler, Linear_Assembler, Linear_Problem
operators.py  integrand descriptors (the reference's closures as coefficient tensors)
geometry.py   setup that stays Julia in a FlexiGal deployment (mesh, dm, Gauss tables)

There is no CPU implementation in this package; importing it without the built library, or using
it without a B200-class GPU, raises.
"""
from . import geometry  # noqa: F401
from ._lib import Context, FlexiGalGPUError, load  # noqa: F401
from .geometry import (BackgroundIntegration, DomainMeasure, Influence_Domains, IntegrationSet, Model,  # noqa: F401
                       Triangulation, create_model, egauss, egauss_bound, pgauss, set_tag)
from .operators import (Advection, Closure, Elasticity, Generic, GenericLinear, GradGrad, Mass, Source)  # noqa: F401
from .space import (ApproxSpace, Bilinear_Assembler, FlexiSpace, Linear_Assembler, Linear_Problem, NodeCloud,  # noqa: F401
                    SHAPE_FUN, Get_Point_Values, ShapeSet, Solve, Solve_On_Device, build_space)

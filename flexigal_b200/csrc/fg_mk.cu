// fg_mk.cu -- Moving-Kriging shape functions (src/Shape_Functions.jl:164-239).  Placeholder until
// the kernel lands: fails loudly, never falls back.
#include "fg_internal.cuh"

int fg_impl_mk(fg_ctx *ctx, GaussSet &S) {
    (void)S;
    return fg_fail(ctx, FG_E_ARG, "technique MK is not implemented in this build");
}

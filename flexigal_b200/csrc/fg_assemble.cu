// fg_assemble.cu -- kernel 3: element contributions scattered into the prebuilt CSC pattern.
//
// Replaces Bilinear_Assembler / Linear_Assembler and their inner loops
// (src/Assembling_Operators.jl:2-26, :28-79, :81-104, :105-165): for every Gauss point and every
// ordered pair (a,b) of its neighbours, K[dom[a], dom[b]] += f(a,b) * (jac*w).  The user closure
// f is bilinear in B_a=[phi_a, grad phi_a] and B_b, so it arrives as a coefficient tensor
// (FG_OP_GENERIC) or as one of the closed forms the reference's examples use.  No COO triplets:
// values go straight into nzval of the pattern built by fg_pattern.cu (trial index b owns the
// CSC column, like sparse(row, col, ...)).
//
// v1 mapping: one warp per Gauss point, lanes over the trial index b, loop over the test index a.
// dom is ascending and so is each column's row list, so the position of row dom[a] in column
// dom[b] only moves forward: a bounded binary search from the previous hit.
#include "fg_internal.cuh"

struct AsmArgs {
    int64_t nnodes, ngauss;
    const double *gs;
    const int64_t *off;
    const int *dom;
    const double *phi, *dphi;
    const int64_t *colptr;
    const int *rowval;
    double *nzval;
    const double *coef;     // device
    int64_t coef_stride;
    double c[4];
};

template <int DIM, int D, int KIND>
__device__ __forceinline__ void pair_block(const AsmArgs &A, const double *__restrict__ C, const double (&Ba)[1 + DIM],
                                           const double (&Bb)[1 + DIM], double wj, double (&out)[D * D]) {
    // out[jc*D + ir] = K_ab[ir][jc]
    constexpr int nb = 1 + DIM;
    if (KIND == FG_OP_GRADGRAD || KIND == FG_OP_MASS) {
        double s;
        if (KIND == FG_OP_MASS) s = Ba[0] * Bb[0];
        else {
            s = Ba[1] * Bb[1];
#pragma unroll
            for (int d = 1; d < DIM; ++d) s += Ba[1 + d] * Bb[1 + d];
        }
        s *= A.c[0] * wj;
#pragma unroll
        for (int jc = 0; jc < D; ++jc)
#pragma unroll
            for (int ir = 0; ir < D; ++ir) out[jc * D + ir] = ir == jc ? s : 0.0;
    } else if (KIND == FG_OP_ADVECTION) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += A.c[d] * Bb[1 + d];
        out[0] = Ba[0] * s * wj;
    } else if (KIND == FG_OP_ELASTICITY) {
        // K_ab[i][j] = G d_ij grad_a.grad_b + G d_j phi_a d_i phi_b + lambda d_i phi_a d_j phi_b
        const double G = A.c[0], lam = A.c[1];
        double gg = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gg += Ba[1 + d] * Bb[1 + d];
#pragma unroll
        for (int jc = 0; jc < D; ++jc)
#pragma unroll
            for (int ir = 0; ir < D; ++ir) {
                double s = G * (Ba[1 + (jc < DIM ? jc : 0)] * Bb[1 + (ir < DIM ? ir : 0)]) + lam * (Ba[1 + (ir < DIM ? ir : 0)] * Bb[1 + (jc < DIM ? jc : 0)]);
                if (ir == jc) s += G * gg;
                out[jc * D + ir] = s * wj;
            }
    } else {
        constexpr int n = D * nb;
#pragma unroll
        for (int jc = 0; jc < D; ++jc)
#pragma unroll
            for (int ir = 0; ir < D; ++ir) {
                double s = 0.0;
#pragma unroll
                for (int al = 0; al < nb; ++al) {
                    double t = 0.0;
#pragma unroll
                    for (int be = 0; be < nb; ++be) t += C[(ir * nb + al) * n + (jc * nb + be)] * Bb[be];
                    s += Ba[al] * t;
                }
                out[jc * D + ir] = s * wj;
            }
    }
}

template <int DIM, int D, int KIND>
__global__ void __launch_bounds__(128) k_bilinear(AsmArgs A) {
    constexpr int nb = 1 + DIM, n = D * nb;
    __shared__ double sC[4][n * n];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (g >= A.ngauss) return;
    const double wj = A.gs[(DIM + 1) * A.ngauss + g] * A.gs[DIM * A.ngauss + g];   // gs[ind,end]*gs[ind,end-1]
    if (KIND == FG_OP_GENERIC) {
        const double *src = A.coef + g * A.coef_stride;
        for (int t = lane; t < n * n; t += 32) sC[wid][t] = src[t];
        __syncwarp();
    }
    const int64_t o0 = A.off[g];
    const int L = (int)(A.off[g + 1] - o0);
    for (int b = lane; b < L; b += 32) {
        double Bb[nb];
        Bb[0] = A.phi[o0 + b];
#pragma unroll
        for (int d = 0; d < DIM; ++d) Bb[1 + d] = A.dphi[(o0 + b) * DIM + d];
        const int j = A.dom[o0 + b];
        const int64_t c0 = A.colptr[j];
        const int nj = (int)(A.colptr[j + 1] - c0);
        const int *rows = A.rowval + c0;
        int pos = 0;
        for (int a = 0; a < L; ++a) {
            const int i = A.dom[o0 + a];
            int lo = pos, hi = nj - 1;       // rows[lo..hi] contains i
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows[mid] < i) lo = mid + 1; else hi = mid; }
            pos = lo;
            double Ba[nb];
            Ba[0] = A.phi[o0 + a];
#pragma unroll
            for (int d = 0; d < DIM; ++d) Ba[1 + d] = A.dphi[(o0 + a) * DIM + d];
            double out[D * D];
            pair_block<DIM, D, KIND>(A, sC[wid], Ba, Bb, wj, out);
            // DOF-level CSC position: (colptr[j])*D*D + jc*(nj*D) + k*D + ir
            double *dst = A.nzval + c0 * (D * D) + (int64_t)pos * D;
#pragma unroll
            for (int jc = 0; jc < D; ++jc)
#pragma unroll
                for (int ir = 0; ir < D; ++ir) {
                    if ((KIND == FG_OP_GRADGRAD || KIND == FG_OP_MASS) && ir != jc) continue;
                    atomicAdd(dst + (int64_t)jc * nj * D + ir, out[jc * D + ir]);
                }
        }
    }
}

template <int DIM, int D>
static int launch_kind(fg_ctx *ctx, const AsmArgs &A, int kind, int64_t ngauss) {
    const int bd = 128;
    const unsigned grid = (unsigned)((ngauss + 3) / 4);
    switch (kind) {
    case FG_OP_GENERIC: k_bilinear<DIM, D, FG_OP_GENERIC><<<grid, bd, 0, ctx->stream>>>(A); break;
    case FG_OP_GRADGRAD: k_bilinear<DIM, D, FG_OP_GRADGRAD><<<grid, bd, 0, ctx->stream>>>(A); break;
    case FG_OP_MASS: k_bilinear<DIM, D, FG_OP_MASS><<<grid, bd, 0, ctx->stream>>>(A); break;
    case FG_OP_ADVECTION: if (D == 1) k_bilinear<DIM, 1, FG_OP_ADVECTION><<<grid, bd, 0, ctx->stream>>>(A); break;
    case FG_OP_ELASTICITY: if (D == DIM) k_bilinear<DIM, DIM, FG_OP_ELASTICITY><<<grid, bd, 0, ctx->stream>>>(A); break;
    }
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

static int upload_coef(fg_ctx *ctx, const fg_operator *op, int64_t ngauss, int ncoef, const double **dev, int64_t *stride) {
    *dev = nullptr; *stride = 0;
    if (op->kind != 0) return FG_OK;
    const int64_t count = op->coef_stride == 0 ? ncoef : (ngauss - 1) * op->coef_stride + ncoef;
    if (op->coef_stride != 0 && op->coef_stride < ncoef) return fg_fail(ctx, FG_E_ARG, "coef_stride %lld smaller than the tensor (%d)", (long long)op->coef_stride, ncoef);
    FG_CUDA(ctx, ctx->coef.reserve(count));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->coef.p, op->coef, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    *dev = ctx->coef.p; *stride = op->coef_stride;
    return FG_OK;
}

int fg_impl_bilinear(fg_ctx *ctx, GaussSet &S, const fg_operator *op) {
    NodeSet &N = ctx->nodes;
    Pattern &P = S.pat;
    const int D = op->D, dim = N.dim, nb = 1 + dim;
    const int64_t total = P.nnz * D * D;
    FG_CUDA(ctx, P.nzval.reserve(total));
    FG_CUDA(ctx, cudaMemsetAsync(P.nzval.p, 0, sizeof(double) * total, ctx->stream));
    P.D = D;
    if (S.n == 0 || S.total_L == 0) return FG_OK;
    AsmArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p; A.phi = S.phi.p; A.dphi = S.dphi.p;
    A.colptr = P.colptr.p; A.rowval = P.rowval.p; A.nzval = P.nzval.p;
    for (int k = 0; k < 4; ++k) A.c[k] = op->c[k];
    int rc = upload_coef(ctx, op, S.n, (D * nb) * (D * nb), &A.coef, &A.coef_stride);
    if (rc) return rc;
    if (dim == 2) {
        if (D == 1) rc = launch_kind<2, 1>(ctx, A, op->kind, S.n);
        else if (D == 2) rc = launch_kind<2, 2>(ctx, A, op->kind, S.n);
        else rc = launch_kind<2, 3>(ctx, A, op->kind, S.n);
    } else {
        if (D == 1) rc = launch_kind<3, 1>(ctx, A, op->kind, S.n);
        else if (D == 2) rc = launch_kind<3, 2>(ctx, A, op->kind, S.n);
        else rc = launch_kind<3, 3>(ctx, A, op->kind, S.n);
    }
    return rc;
}

// --------------------------------------------------------------------------------------------
// linear forms: F[(node-1)*D + i] += (sum_al c[i*nb+al] B_a[al]) * (jac*w)
// --------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_linear(AsmArgs A, int D, int kind, double *__restrict__ F) {
    constexpr int nb = 1 + DIM;
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= A.ngauss) return;
    const double wj = A.gs[(DIM + 1) * A.ngauss + g] * A.gs[DIM * A.ngauss + g];
    const double *C = kind == FG_LIN_GENERIC ? A.coef + g * A.coef_stride : nullptr;
    const int64_t o0 = A.off[g], o1 = A.off[g + 1];
    for (int64_t k = o0 + lane; k < o1; k += 32) {
        double B[nb];
        B[0] = A.phi[k];
#pragma unroll
        for (int d = 0; d < DIM; ++d) B[1 + d] = A.dphi[k * DIM + d];
        const int64_t node = A.dom[k];
        for (int i = 0; i < D; ++i) {
            double s;
            if (kind == FG_LIN_SOURCE) s = A.c[i] * B[0];
            else {
                s = 0.0;
#pragma unroll
                for (int al = 0; al < nb; ++al) s += C[i * nb + al] * B[al];
            }
            atomicAdd(F + node * D + i, s * wj);
        }
    }
}

int fg_impl_linear(fg_ctx *ctx, GaussSet &S, const fg_operator *op) {
    NodeSet &N = ctx->nodes;
    const int D = op->D, dim = N.dim, nb = 1 + dim;
    FG_CUDA(ctx, S.fvec.reserve(N.n * 3));
    FG_CUDA(ctx, cudaMemsetAsync(S.fvec.p, 0, sizeof(double) * N.n * D, ctx->stream));
    S.fvec_D = D;
    if (S.n == 0 || S.total_L == 0) return FG_OK;
    AsmArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p; A.phi = S.phi.p; A.dphi = S.dphi.p;
    A.colptr = nullptr; A.rowval = nullptr; A.nzval = nullptr;
    for (int k = 0; k < 4; ++k) A.c[k] = op->c[k];
    int rc = upload_coef(ctx, op, S.n, D * nb, &A.coef, &A.coef_stride);
    if (rc) return rc;
    const unsigned grid = (unsigned)((S.n * 32 + 127) / 128);
    if (dim == 2) k_linear<2><<<grid, 128, 0, ctx->stream>>>(A, D, op->kind, S.fvec.p);
    else k_linear<3><<<grid, 128, 0, ctx->stream>>>(A, D, op->kind, S.fvec.p);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

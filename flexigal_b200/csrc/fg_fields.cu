// fg_fields.cu -- SURVEY 8(f) rows f2 and f3: what sits right after the assembly in a Newton / Picard loop.
//
// f2  Field evaluation at Gauss points: Get_Point_Values (src/Fields_Operations.jl:40-112) and the gathers of
//     evaluate_at_point (:340-407):   u_h(g) = sum_a phi_a u[dom_a],   grad u_h(g) = sum_a u[dom_a] (x) grad phi_a.
//     One warp per Gauss point, lanes over its neighbours, warp-shuffle reductions; PHI/DPHI/DOM are already on
//     the device, so a Newton step only uploads the nodal vector.
// f3  Solve A u = F on the device for the symmetric positive definite systems of the Poisson / elasticity
//     examples (Solve, src/Solvers.jl:24-31 does A \ F on the host): Jacobi-preconditioned conjugate gradients
//     on the assembled CSC arrays (symmetric => the CSC columns are the CSR rows).
#include "fg_internal.cuh"
#include <math.h>
#include <algorithm>
#include <string.h>
#include <vector>

template <int DIM>
__global__ void __launch_bounds__(128) k_eval_field(int64_t ngauss, int D, const int64_t *__restrict__ off, const int *__restrict__ dom,
                                                    const double *__restrict__ phi, const double *__restrict__ dphi, int64_t tl,
                                                    const double *__restrict__ u, double *__restrict__ val, double *__restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= ngauss) return;
    const int64_t o0 = off[g], o1 = off[g + 1];
    for (int i = 0; i < D; ++i) {
        double acc[1 + DIM];
#pragma unroll
        for (int k = 0; k <= DIM; ++k) acc[k] = 0.0;
        for (int64_t e = o0 + lane; e < o1; e += 32) {
            const double un = u[(int64_t)dom[e] * D + i];
            if (val) acc[0] += phi[e] * un;
            if (grad) {
#pragma unroll
                for (int d = 0; d < DIM; ++d) acc[1 + d] += dphi[d * tl + e] * un;
            }
        }
#pragma unroll
        for (int k = 0; k <= DIM; ++k)
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], sft);
        if (lane == 0) {
            if (val) val[g * D + i] = acc[0];
            if (grad) {
                // SMatrix{D,DIM} memory order (column-major): element (i, d) at i + D*d
#pragma unroll
                for (int d = 0; d < DIM; ++d) grad[g * D * DIM + d * D + i] = acc[1 + d];
            }
        }
    }
}

extern "C" int fg_evaluate_field(fg_ctx *ctx, int set_id, int D, const double *u_nodal, double *values, double *gradients) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_evaluate_field: shape functions of set %d not computed", set_id);
    if (D < 1 || D > 3 || !u_nodal) return fg_fail(ctx, FG_E_ARG, "fg_evaluate_field: bad D / nodal vector");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rcd = fg_ensure_shapes(ctx, *S); if (rcd) return rcd; }
    { int rcd = fg_ensure_dom(ctx, *S); if (rcd) return rcd; }
    const int dim = ctx->nodes.dim;
    const int64_t nn = ctx->nodes.n, ng = S->n;
    // staging lives in the context (grow-only): this is the per-Newton-step path, and a cudaMalloc/cudaFree pair per call
    // costs 10-400 ms and synchronises the device
    DevBuf<double> &du = ctx->field_u, &dv = ctx->field_val, &dg = ctx->field_grad;
    FG_CUDA(ctx, du.reserve(nn * D));
    if (values) FG_CUDA(ctx, dv.reserve(ng * D));
    if (gradients) FG_CUDA(ctx, dg.reserve(ng * D * dim));
    FG_CUDA(ctx, cudaMemcpyAsync(du.p, u_nodal, sizeof(double) * nn * D, cudaMemcpyHostToDevice, ctx->stream));
    if (ng > 0) {
        const unsigned grid = (unsigned)((ng * 32 + 127) / 128);
        if (dim == 2) k_eval_field<2><<<grid, 128, 0, ctx->stream>>>(ng, D, S->off.p, S->dom.p, S->phi.p, S->dphi.p, S->total_L, du.p, values ? dv.p : nullptr, gradients ? dg.p : nullptr);
        else k_eval_field<3><<<grid, 128, 0, ctx->stream>>>(ng, D, S->off.p, S->dom.p, S->phi.p, S->dphi.p, S->total_L, du.p, values ? dv.p : nullptr, gradients ? dg.p : nullptr);
        FG_CHECK_LAUNCH(ctx);
    }
    if (values && ng) FG_CUDA(ctx, cudaMemcpyAsync(values, dv.p, sizeof(double) * ng * D, cudaMemcpyDeviceToHost, ctx->stream));
    if (gradients && ng) FG_CUDA(ctx, cudaMemcpyAsync(gradients, dg.p, sizeof(double) * ng * D * dim, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// f3: preconditioned CG for symmetric positive definite systems.  Two sources of the matrix:
//   fg_solve_spd       Julia's SparseMatrixCSC arrays (1-based Int64, host) -- uploaded once per call into grow-only buffers
//   fg_solve_spd_sets  K = sum_k scale_k * (matrix last assembled on set k), values and pattern RESIDENT on the device
//                      (node-level CSC + DOF block layout of fg_get_pattern): nothing but the right-hand side is uploaded.
// Symmetric => (K x)_c = column c of K dotted with x: one warp per DOF column, coalesced reads of the column.
// --------------------------------------------------------------------------------------------
__global__ void k_spmv_sym(int64_t n, const int64_t *__restrict__ ptr, const int64_t *__restrict__ idx, const double *__restrict__ a,
                           const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    double s = 0.0;
    for (int64_t k = ptr[r] - 1 + lane; k < ptr[r + 1] - 1; k += 32) s += a[k] * x[idx[k] - 1];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if (lane == 0) y[r] = s;
}

__global__ void k_diag_inv(int64_t n, const int64_t *__restrict__ ptr, const int64_t *__restrict__ idx, const double *__restrict__ a, double *__restrict__ diag) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0.0;
    for (int64_t k = ptr[r] - 1; k < ptr[r + 1] - 1; ++k) if (idx[k] - 1 == r) d += a[k];
    diag[r] = d;
}

// resident format: y[c] (+)= scale * sum_t a[col block] x[row dof]; DOF column c = (j, jc), entries t = k*D + ir
__global__ void k_spmv_set(int64_t nnodes, int D, const int64_t *__restrict__ colptr, const int *__restrict__ rowval, const double *__restrict__ a,
                           double scale, const double *__restrict__ x, double *__restrict__ y, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= nnodes * D) return;
    const int64_t j = c / D; const int jc = (int)(c - j * D);
    const int64_t c0 = colptr[j]; const int nj = (int)(colptr[j + 1] - c0);
    const double *col = a + c0 * D * D + (int64_t)jc * nj * D;
    double s = 0.0;
    for (int t = lane; t < nj * D; t += 32) { const int k = t / D, ir = t - k * D; s += col[t] * x[(int64_t)rowval[c0 + k] * D + ir]; }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if (lane == 0) y[c] = accumulate ? y[c] + scale * s : scale * s;
}

__global__ void k_diag_set(int64_t nnodes, int D, const int64_t *__restrict__ colptr, const int *__restrict__ rowval, const double *__restrict__ a,
                           double scale, double *__restrict__ diag) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nnodes * D) return;
    const int64_t j = c / D; const int jc = (int)(c - j * D);
    const int64_t c0 = colptr[j]; const int nj = (int)(colptr[j + 1] - c0);
    int lo = 0, hi = nj - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rowval[c0 + mid] < j) lo = mid + 1; else hi = mid; }
    if (nj > 0 && rowval[c0 + lo] == j) diag[c] += scale * a[c0 * D * D + (int64_t)jc * nj * D + (int64_t)lo * D + jc];
}

__global__ void k_invert_diag(int64_t n, double *__restrict__ d) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) d[i] = d[i] != 0.0 ? 1.0 / d[i] : 1.0;
}

// out[slot] += sum_i a[i]*b[i]  (block partials + atomicAdd; enough accuracy for the scalar recurrences of CG)
__global__ void k_dot(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i] * b[i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w]; atomicAdd(out, t); }
}

// scal layout: [0] rz_old, [1] pAp, [2] rz_new, [3] rr
__global__ void k_cg_update_xr(int64_t n, const double *__restrict__ scal, const double *__restrict__ p, const double *__restrict__ Ap,
                               const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double alpha = scal[0] / scal[1];
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    z[i] = dinv[i] * ri;
}

__global__ void k_cg_update_p(int64_t n, const double *__restrict__ scal, const double *__restrict__ z, double *__restrict__ p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double beta = scal[2] / scal[0];
    p[i] = z[i] + beta * p[i];
}

__global__ void k_cg_shift(double *scal) { scal[0] = scal[2]; scal[1] = 0.0; scal[2] = 0.0; scal[3] = 0.0; }

__global__ void k_residual(int64_t n, const double *__restrict__ b, const double *__restrict__ Ax, const double *__restrict__ dinv,
                           double *__restrict__ r, double *__restrict__ z, double *__restrict__ p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ri = b[i] - Ax[i];
    r[i] = ri; z[i] = dinv[i] * ri; p[i] = z[i];
}

// The CG iteration itself; `matvec(x, y)` enqueues y = K x on the context's stream, dinv holds the inverted diagonal.
// Work vectors: ctx->solve.v[0..6] = b, x, r, z, p, Ap, dinv (+ 8 scalars at the end of v[6]).
template <typename MatVec>
static int pcg(fg_ctx *ctx, int64_t n, MatVec matvec, double *x_host, double rtol, int max_iter, int *iterations, double *rel_residual) {
    auto &W = ctx->solve;
    double *db = W.v[0].p, *dx = W.v[1].p, *dr = W.v[2].p, *dz = W.v[3].p, *dp = W.v[4].p, *dAp = W.v[5].p, *dinv = W.v[6].p, *dscal = W.v[6].p + n;
    cudaStream_t st = ctx->stream;
    const unsigned gv = (unsigned)((n + 255) / 256), gd = (unsigned)std::min<int64_t>(1024, (n + 255) / 256);
    FG_CUDA(ctx, cudaMemsetAsync(dx, 0, sizeof(double) * n, st));
    FG_CUDA(ctx, cudaMemsetAsync(dscal, 0, sizeof(double) * 8, st));
    FG_CUDA(ctx, cudaMemsetAsync(dAp, 0, sizeof(double) * n, st));
    k_residual<<<gv, 256, 0, st>>>(n, db, dAp, dinv, dr, dz, dp);          // x0 = 0: r = b
    k_dot<<<gd, 256, 0, st>>>(n, dr, dz, dscal + 0);
    k_dot<<<gd, 256, 0, st>>>(n, db, db, dscal + 4);
    ctx->launches += 3;
    double h[8];
    FG_CUDA(ctx, cudaMemcpyAsync(h, dscal, sizeof h, cudaMemcpyDeviceToHost, st));
    FG_CUDA(ctx, cudaStreamSynchronize(st));
    const double bb = h[4];
    int it = 0; double rr = bb;
    if (bb > 0.0) {
        for (it = 1; it <= max_iter; ++it) {
            int rc = matvec(dp, dAp); if (rc) return rc;
            k_dot<<<gd, 256, 0, st>>>(n, dp, dAp, dscal + 1);
            k_cg_update_xr<<<gv, 256, 0, st>>>(n, dscal, dp, dAp, dinv, dx, dr, dz);
            k_dot<<<gd, 256, 0, st>>>(n, dr, dz, dscal + 2);
            k_dot<<<gd, 256, 0, st>>>(n, dr, dr, dscal + 3);
            k_cg_update_p<<<gv, 256, 0, st>>>(n, dscal, dz, dp);
            ctx->launches += 5;
            if (it % 25 == 0 || it == max_iter) {          // look at the residual every 25 iterations only
                FG_CUDA(ctx, cudaMemcpyAsync(h, dscal, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
                FG_CUDA(ctx, cudaStreamSynchronize(st));
                rr = h[3];
                if (!(rr == rr)) break;
                if (sqrt(rr / bb) <= rtol) { k_cg_shift<<<1, 1, 0, st>>>(dscal); break; }
            }
            k_cg_shift<<<1, 1, 0, st>>>(dscal);
        }
    }
    // true residual of the returned iterate
    { int rc = matvec(dx, dAp); if (rc) return rc; }
    k_residual<<<gv, 256, 0, st>>>(n, db, dAp, dinv, dr, dz, dp);
    FG_CUDA(ctx, cudaMemsetAsync(dscal, 0, sizeof(double) * 8, st));
    k_dot<<<gd, 256, 0, st>>>(n, dr, dr, dscal + 3);
    FG_CHECK_LAUNCH(ctx);
    FG_CUDA(ctx, cudaMemcpyAsync(h, dscal, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    if (x_host) FG_CUDA(ctx, cudaMemcpyAsync(x_host, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    FG_CUDA(ctx, cudaStreamSynchronize(st));
    if (iterations) *iterations = it > max_iter ? max_iter : it;
    if (rel_residual) *rel_residual = bb > 0.0 ? sqrt(h[3] / bb) : 0.0;
    return FG_OK;
}

static int solve_reserve(fg_ctx *ctx, int64_t n) {
    for (int k = 0; k < 7; ++k) FG_CUDA(ctx, ctx->solve.v[k].reserve(n + (k == 6 ? 8 : 0)));
    return FG_OK;
}

extern "C" int fg_solve_spd(fg_ctx *ctx, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval, const double *rhs,
                            double *x, double rtol, int max_iter, int *iterations, double *rel_residual) {
    if (!ctx) return FG_E_ARG;
    if (n <= 0 || !colptr || !rowval || !nzval || !rhs || !x) return fg_fail(ctx, FG_E_ARG, "fg_solve_spd: bad arguments");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nnz = colptr[n] - 1;
    int rc = solve_reserve(ctx, n); if (rc) return rc;
    auto &W = ctx->solve;
    FG_CUDA(ctx, W.cp.reserve(n + 1 + nnz));          // colptr, then rowval (both Int64 1-based, as Julia holds them)
    FG_CUDA(ctx, ctx->coef.reserve(nnz));
    cudaStream_t st = ctx->stream;
    int64_t *dptr = W.cp.p, *didx = W.cp.p + n + 1; double *da = ctx->coef.p;
    FG_CUDA(ctx, cudaMemcpyAsync(dptr, colptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    FG_CUDA(ctx, cudaMemcpyAsync(didx, rowval, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, st));
    FG_CUDA(ctx, cudaMemcpyAsync(da, nzval, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
    FG_CUDA(ctx, cudaMemcpyAsync(W.v[0].p, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    const unsigned gv = (unsigned)((n + 255) / 256), gw = (unsigned)((n * 32 + 255) / 256);
    k_diag_inv<<<gv, 256, 0, st>>>(n, dptr, didx, da, W.v[6].p);
    k_invert_diag<<<gv, 256, 0, st>>>(n, W.v[6].p);
    FG_CHECK_LAUNCH(ctx);
    auto matvec = [&](const double *xin, double *yout) {
        k_spmv_sym<<<gw, 256, 0, st>>>(n, dptr, didx, da, xin, yout);
        FG_CHECK_LAUNCH(ctx);
        return FG_OK;
    };
    return pcg(ctx, n, matvec, x, rtol, max_iter, iterations, rel_residual);
}

extern "C" int fg_solve_spd_sets(fg_ctx *ctx, const int *set_ids, const double *scales, int nsets, int D, const double *rhs, double *x,
                                 double rtol, int max_iter, int *iterations, double *rel_residual) {
    if (!ctx) return FG_E_ARG;
    if (!set_ids || nsets <= 0 || D < 1 || D > 3 || !rhs || !x) return fg_fail(ctx, FG_E_ARG, "fg_solve_spd_sets: bad arguments");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<GaussSet *> sets;
    for (int k = 0; k < nsets; ++k) {
        GaussSet *S = fg_find_set(ctx, set_ids[k]);
        if (!S || !S->pat.ready || S->pat.D != D) return fg_fail(ctx, FG_E_STATE, "fg_solve_spd_sets: set %d holds no assembled matrix with D = %d", set_ids[k], D);
        sets.push_back(S);
    }
    const int64_t nn = ctx->nodes.n, n = nn * D;
    int rc = solve_reserve(ctx, n); if (rc) return rc;
    auto &W = ctx->solve;
    cudaStream_t st = ctx->stream;
    FG_CUDA(ctx, cudaMemcpyAsync(W.v[0].p, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    FG_CUDA(ctx, cudaMemsetAsync(W.v[6].p, 0, sizeof(double) * n, st));
    const unsigned gv = (unsigned)((n + 255) / 256), gw = (unsigned)((n * 32 + 255) / 256);
    for (int k = 0; k < nsets; ++k)
        k_diag_set<<<gv, 256, 0, st>>>(nn, D, sets[k]->pat.colptr.p, sets[k]->pat.rowval.p, sets[k]->pat.nzval.p, scales ? scales[k] : 1.0, W.v[6].p);
    k_invert_diag<<<gv, 256, 0, st>>>(n, W.v[6].p);
    FG_CHECK_LAUNCH(ctx);
    auto matvec = [&](const double *xin, double *yout) {
        for (int k = 0; k < nsets; ++k)
            k_spmv_set<<<gw, 256, 0, st>>>(nn, D, sets[k]->pat.colptr.p, sets[k]->pat.rowval.p, sets[k]->pat.nzval.p, scales ? scales[k] : 1.0, xin, yout, k > 0);
        ctx->launches += nsets;
        FG_CUDA(ctx, cudaGetLastError());
        return FG_OK;
    };
    return pcg(ctx, n, matvec, x, rtol, max_iter, iterations, rel_residual);
}

// maximum(A) of the matrix last assembled on a set, as Julia's maximum over a SparseMatrixCSC sees it: implicit zeros count,
// so the result is max(0, largest stored value) unless the matrix is completely full (Linear_Problem's alpha, Assembling_Operators.jl:238).
__global__ void k_max_value(int64_t n, const double *__restrict__ a, double *__restrict__ out) {
    double m = -1.0 / 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, a[i]);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, sft));
    if ((threadIdx.x & 31) == 0) {          // max over non-negative-or-any doubles through the ordered-int trick
        long long bits = __double_as_longlong(m);
        bits = bits >= 0 ? bits : bits ^ 0x7fffffffffffffffll;
        atomicMax((long long *)out, bits);
    }
}

extern "C" int fg_values_max(fg_ctx *ctx, int set_id, double *max_value) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready || S->pat.D == 0 || !max_value) return fg_fail(ctx, FG_E_STATE, "fg_values_max: nothing assembled for set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = S->pat.nnz * S->pat.D * S->pat.D, ndof = ctx->nodes.n * S->pat.D;
    FG_CUDA(ctx, ctx->flags.reserve(64));
    long long init = (long long)0x8000000000000000ull;      // below every encoded double
    long long *slot = (long long *)(ctx->flags.p + 20);
    FG_CUDA(ctx, cudaMemcpyAsync(slot, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    if (total) { k_max_value<<<(unsigned)std::min<int64_t>(2048, (total + 255) / 256), 256, 0, ctx->stream>>>(total, S->pat.nzval.p, (double *)slot); FG_CHECK_LAUNCH(ctx); }
    long long bits = 0;
    FG_CUDA(ctx, cudaMemcpyAsync(&bits, slot, sizeof bits, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bits = bits >= 0 ? bits : bits ^ 0x7fffffffffffffffll;
    double m; memcpy(&m, &bits, sizeof m);
    if (total == 0) m = 0.0;
    else if ((double)total < (double)ndof * (double)ndof) m = m > 0.0 ? m : 0.0;
    *max_value = m;
    return FG_OK;
}

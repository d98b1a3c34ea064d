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
    const int dim = ctx->nodes.dim;
    const int64_t nn = ctx->nodes.n, ng = S->n;
    DevBuf<double> du, dv, dg;
    int rc = FG_OK;
    do {
        if (du.reserve(nn * D) != cudaSuccess || (values && dv.reserve(ng * D) != cudaSuccess) || (gradients && dg.reserve(ng * D * dim) != cudaSuccess)) {
            rc = fg_fail(ctx, FG_E_OOM, "fg_evaluate_field: device allocation failed"); break;
        }
        cudaMemcpyAsync(du.p, u_nodal, sizeof(double) * nn * D, cudaMemcpyHostToDevice, ctx->stream);
        if (ng > 0) {
            const unsigned grid = (unsigned)((ng * 32 + 127) / 128);
            if (dim == 2) k_eval_field<2><<<grid, 128, 0, ctx->stream>>>(ng, D, S->off.p, S->dom.p, S->phi.p, S->dphi.p, S->total_L, du.p, values ? dv.p : nullptr, gradients ? dg.p : nullptr);
            else k_eval_field<3><<<grid, 128, 0, ctx->stream>>>(ng, D, S->off.p, S->dom.p, S->phi.p, S->dphi.p, S->total_L, du.p, values ? dv.p : nullptr, gradients ? dg.p : nullptr);
            ++ctx->launches;
        }
        if (values && ng) cudaMemcpyAsync(values, dv.p, sizeof(double) * ng * D, cudaMemcpyDeviceToHost, ctx->stream);
        if (gradients && ng) cudaMemcpyAsync(gradients, dg.p, sizeof(double) * ng * D * dim, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = fg_fail(ctx, FG_E_CUDA, "fg_evaluate_field: %s", cudaGetErrorString(e));
    } while (0);
    du.release(); dv.release(); dg.release();
    return rc;
}

// --------------------------------------------------------------------------------------------
// f3: Jacobi-preconditioned CG on a symmetric CSC/CSR matrix (1-based Int64 arrays, as Julia holds them)
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

__global__ void k_diag_inv(int64_t n, const int64_t *__restrict__ ptr, const int64_t *__restrict__ idx, const double *__restrict__ a, double *__restrict__ dinv) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0.0;
    for (int64_t k = ptr[r] - 1; k < ptr[r + 1] - 1; ++k) if (idx[k] - 1 == r) d += a[k];
    dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
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

extern "C" int fg_solve_spd(fg_ctx *ctx, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval, const double *rhs,
                            double *x, double rtol, int max_iter, int *iterations, double *rel_residual) {
    if (!ctx) return FG_E_ARG;
    if (n <= 0 || !colptr || !rowval || !nzval || !rhs || !x) return fg_fail(ctx, FG_E_ARG, "fg_solve_spd: bad arguments");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nnz = colptr[n] - 1;
    DevBuf<int64_t> dptr, didx;
    DevBuf<double> da, db, dx, dr, dz, dp, dAp, dinv, dscal;
    int rc = FG_OK;
    auto fail_alloc = [&]() { return dptr.reserve(n + 1) || didx.reserve(nnz) || da.reserve(nnz) || db.reserve(n) || dx.reserve(n) || dr.reserve(n) ||
                                     dz.reserve(n) || dp.reserve(n) || dAp.reserve(n) || dinv.reserve(n) || dscal.reserve(8); };
    do {
        if (fail_alloc()) { rc = fg_fail(ctx, FG_E_OOM, "fg_solve_spd: device allocation failed"); break; }
        cudaStream_t st = ctx->stream;
        cudaMemcpyAsync(dptr.p, colptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(didx.p, rowval, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(da.p, nzval, sizeof(double) * nnz, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(db.p, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(dx.p, 0, sizeof(double) * n, st);
        cudaMemsetAsync(dscal.p, 0, sizeof(double) * 8, st);
        const unsigned gv = (unsigned)((n + 255) / 256), gw = (unsigned)((n * 32 + 255) / 256), gd = (unsigned)std::min<int64_t>(1024, (n + 255) / 256);
        k_diag_inv<<<gv, 256, 0, st>>>(n, dptr.p, didx.p, da.p, dinv.p);
        cudaMemsetAsync(dAp.p, 0, sizeof(double) * n, st);
        k_residual<<<gv, 256, 0, st>>>(n, db.p, dAp.p, dinv.p, dr.p, dz.p, dp.p);          // x0 = 0: r = b
        k_dot<<<gd, 256, 0, st>>>(n, dr.p, dz.p, dscal.p + 0);
        k_dot<<<gd, 256, 0, st>>>(n, db.p, db.p, dscal.p + 4);
        double h[8];
        cudaMemcpyAsync(h, dscal.p, sizeof h, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "fg_solve_spd: setup failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        const double bb = h[4];
        int it = 0; double rr = bb;
        if (bb > 0.0) {
            for (it = 1; it <= max_iter; ++it) {
                k_spmv_sym<<<gw, 256, 0, st>>>(n, dptr.p, didx.p, da.p, dp.p, dAp.p);
                k_dot<<<gd, 256, 0, st>>>(n, dp.p, dAp.p, dscal.p + 1);
                k_cg_update_xr<<<gv, 256, 0, st>>>(n, dscal.p, dp.p, dAp.p, dinv.p, dx.p, dr.p, dz.p);
                k_dot<<<gd, 256, 0, st>>>(n, dr.p, dz.p, dscal.p + 2);
                k_dot<<<gd, 256, 0, st>>>(n, dr.p, dr.p, dscal.p + 3);
                k_cg_update_p<<<gv, 256, 0, st>>>(n, dscal.p, dz.p, dp.p);
                ctx->launches += 6;
                if (it % 25 == 0 || it == max_iter) {          // look at the residual every 25 iterations only
                    cudaMemcpyAsync(h, dscal.p, sizeof(double) * 4, cudaMemcpyDeviceToHost, st);
                    cudaStreamSynchronize(st);
                    rr = h[3];
                    if (!(rr == rr)) break;
                    if (sqrt(rr / bb) <= rtol) { k_cg_shift<<<1, 1, 0, st>>>(dscal.p); break; }
                }
                k_cg_shift<<<1, 1, 0, st>>>(dscal.p);
            }
        }
        // true residual of the returned iterate
        k_spmv_sym<<<gw, 256, 0, st>>>(n, dptr.p, didx.p, da.p, dx.p, dAp.p);
        k_residual<<<gv, 256, 0, st>>>(n, db.p, dAp.p, dinv.p, dr.p, dz.p, dp.p);
        cudaMemsetAsync(dscal.p, 0, sizeof(double) * 8, st);
        k_dot<<<gd, 256, 0, st>>>(n, dr.p, dr.p, dscal.p + 3);
        cudaMemcpyAsync(h, dscal.p, sizeof(double) * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(x, dx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "fg_solve_spd: %s", cudaGetErrorString(e)); break; }
        if (iterations) *iterations = it > max_iter ? max_iter : it;
        if (rel_residual) *rel_residual = bb > 0.0 ? sqrt(h[3] / bb) : 0.0;
    } while (0);
    dptr.release(); didx.release(); da.release(); db.release(); dx.release(); dr.release(); dz.release(); dp.release(); dAp.release(); dinv.release(); dscal.release();
    return rc;
}

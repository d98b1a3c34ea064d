// fg_mk.cu -- Moving-Kriging shape functions (technique = :MK), src/Shape_Functions.jl:164-239.
//
// Per Gauss point with neighbours v (L of them), Cq = 4.3, theta = 0.1 Cq^2:
//   centre xye = mean x[v], scale den_d = mean dm[v,d] * Cq                                   (:168-170)
//   P[i,:]  = bilinear/trilinear basis of n_i = (x_i - xye)/den, p = basis at (g - xye)/den    (:171-197)
//   R[i,j]  = exp(-theta |(x_i - x_j)/den|^2),  r_i = exp(-theta |(g - x_i)/den|^2), dr = d r / d g   (:199-225)
//   A = (P^T R^-1 P)^-1 P^T R^-1,  B = R^-1 (I - P A),  phi = p^T A + r^T B,  d_d phi = dp_d^T A + dr_d^T B   (:226-232)
// The reference factorises R with LU and forms A and B explicitly.  R is symmetric positive definite, so this
// kernel uses a Cholesky factorisation and never forms the L x L matrix B:
//   Z = R^-1 [P | r | dr_1..dr_dim],  M = P^T Z_P,  y_c = M^-1 (p_c - P^T s_c),  phi_c = s_c + Z_P y_c
// (c = 0: values, c = d: d-th derivative; s_c = Z[:, m+c]) -- the same numbers in exact arithmetic.
//
// One warp per Gauss point; R, P and the right-hand sides live in shared memory (sized by the longest
// neighbour list), lanes = rows for the factorisation and the final products, lanes = right-hand sides for
// the triangular solves.
#include "fg_internal.cuh"

struct MkArgs {
    int64_t nnodes, ngauss;
    const double *x, *dm, *gs;
    const int64_t *off;
    const int *dom;
    double *phi, *dphi;   // dphi[d * tl + e]
    int64_t tl;
    int *singular;
    int maxL;
};

template <int M>
__device__ __forceinline__ bool small_cholesky(double (&A)[M * (M + 1) / 2], double (&inv)[M]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double s = A[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < j) s -= A[j * (j + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
        if (!(s > 0.0)) { ok = false; s = 1.0; }
        const double is = rsqrt(s);
        inv[j] = is;
        A[j * (j + 1) / 2 + j] = s * is;
#pragma unroll
        for (int i = 0; i < M; ++i)
            if (i > j) {
                double t = A[i * (i + 1) / 2 + j];
#pragma unroll
                for (int k = 0; k < M; ++k)
                    if (k < j) t -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
                A[i * (i + 1) / 2 + j] = t * is;
            }
    }
    return ok;
}

template <int M>
__device__ __forceinline__ void small_solve(const double (&Lm)[M * (M + 1) / 2], const double (&inv)[M], double (&b)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double t = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < i) t -= Lm[i * (i + 1) / 2 + k] * b[k];
        b[i] = t * inv[i];
    }
#pragma unroll
    for (int ii = 0; ii < M; ++ii) {
        const int i = M - 1 - ii;
        double t = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k > i) t -= Lm[k * (k + 1) / 2 + i] * b[k];
        b[i] = t * inv[i];
    }
}

template <int DIM>
__device__ __forceinline__ void mk_basis(const double (&n)[DIM], double (&b)[DIM == 2 ? 4 : 8]) {
    b[0] = 1.0; b[1] = n[0]; b[2] = n[1];
    if (DIM == 2) b[3] = n[0] * n[1];
    else { b[3] = n[DIM - 1]; b[4] = n[0] * n[1]; b[5] = n[0] * n[DIM - 1]; b[6] = n[1] * n[DIM - 1]; b[7] = n[0] * n[1] * n[DIM - 1]; }
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(128) k_mk(MkArgs A) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NR = M + 1 + DIM;            // right-hand sides: P | r | dr_d
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (g >= A.ngauss) return;
    const int Lm = A.maxL, ldr = Lm | 1;
    double *R = sm + (size_t)wid * ((size_t)Lm * ldr + (size_t)Lm * (NR + M + DIM));   // Lm x ldr
    double *W = R + (size_t)Lm * ldr;                                                   // Lm x NR (right-hand sides -> Z)
    double *P = W + (size_t)Lm * NR;                                                    // Lm x M
    double *X = P + (size_t)Lm * M;                                                     // Lm x DIM node coordinates
    const int64_t o0 = A.off[g];
    const int L = (int)(A.off[g + 1] - o0);
    if (L == 0) return;
    const double Cq = 4.3, theta = 0.1 * (Cq * Cq);
    double gp[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gp[d] = A.gs[d * A.ngauss + g];
    // centre and scale (sums in neighbour order, like sum(x[v,d])/L)
    double xye[DIM], den[DIM];
    {
        double sx[DIM], sh[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { sx[d] = 0.0; sh[d] = 0.0; }
        for (int a = lane; a < L; a += 32) {
            const int id = A.dom[o0 + a];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double xa = A.x[d * A.nnodes + id];
                X[a * DIM + d] = xa;
                sx[d] += xa;
                sh[d] += SHAPE == FG_SHAPE_RECTANGULAR ? A.dm[d * A.nnodes + id] : A.dm[id];
            }
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) { sx[d] += __shfl_xor_sync(0xffffffffu, sx[d], sft); sh[d] += __shfl_xor_sync(0xffffffffu, sh[d], sft); }
            xye[d] = sx[d] / L;
            den[d] = (sh[d] / L) * Cq;
        }
    }
    __syncwarp();
    double gn[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gn[d] = (gp[d] - xye[d]) / den[d];
    // rows of P, R, r, dr
    for (int i = lane; i < L; i += 32) {
        double n[DIM], b[M];
#pragma unroll
        for (int d = 0; d < DIM; ++d) n[d] = (X[i * DIM + d] - xye[d]) / den[d];
        mk_basis<DIM>(n, b);
#pragma unroll
        for (int k = 0; k < M; ++k) { P[i * M + k] = b[k]; W[i * NR + k] = b[k]; }
        double arg = 0.0, ds[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { ds[d] = (gp[d] - X[i * DIM + d]) / den[d]; arg += ds[d] * ds[d]; }
        const double rv = exp(-theta * arg);
        W[i * NR + M] = rv;
#pragma unroll
        for (int d = 0; d < DIM; ++d) W[i * NR + M + 1 + d] = rv * (-2.0 * theta * ds[d] / den[d]);
        for (int j = 0; j <= i; ++j) {
            double a2 = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) { const double t = (X[i * DIM + d] - X[j * DIM + d]) / den[d]; a2 += t * t; }
            R[i * ldr + j] = exp(-theta * a2);
        }
    }
    __syncwarp();
    // Cholesky R = C C^T (lower, in place), right-looking
    bool ok = true;
    for (int j = 0; j < L; ++j) {
        const double djj = R[j * ldr + j];
        if (!(djj > 0.0)) { ok = false; break; }
        const double inv = 1.0 / sqrt(djj);
        __syncwarp();
        for (int i = j + lane; i < L; i += 32) R[i * ldr + j] *= inv;          // includes the diagonal: sqrt(djj)
        __syncwarp();
        for (int i = j + 1 + lane; i < L; i += 32) {
            const double cij = R[i * ldr + j];
            for (int k = j + 1; k <= i; ++k) R[i * ldr + k] -= cij * R[k * ldr + j];
        }
        __syncwarp();
    }
    ok = __all_sync(0xffffffffu, ok);
    // Z = R^-1 W : one right-hand side per lane
    if (ok && lane < NR) {
        for (int i = 0; i < L; ++i) {
            double t = W[i * NR + lane];
            for (int k = 0; k < i; ++k) t -= R[i * ldr + k] * W[k * NR + lane];
            W[i * NR + lane] = t / R[i * ldr + i];
        }
        for (int i = L - 1; i >= 0; --i) {
            double t = W[i * NR + lane];
            for (int k = i + 1; k < L; ++k) t -= R[k * ldr + i] * W[k * NR + lane];
            W[i * NR + lane] = t / R[i * ldr + i];
        }
    }
    __syncwarp();
    // lanes 0..DIM: y_c = M^-1 (p_c - P^T s_c), M = P^T Z_P
    double y[M];
#pragma unroll
    for (int k = 0; k < M; ++k) y[k] = 0.0;
    if (ok && lane <= DIM) {
        double Mm[M * (M + 1) / 2];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j)
                if (j <= i) {
                    double s = 0.0;
                    for (int a = 0; a < L; ++a) s += P[a * M + i] * W[a * NR + j];
                    Mm[i * (i + 1) / 2 + j] = s;
                }
        double pc[M], b0[M];
        mk_basis<DIM>(gn, b0);
#pragma unroll
        for (int k = 0; k < M; ++k) pc[k] = 0.0;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < M; ++k) pc[k] = b0[k];
        } else {                       // d p / d g_d, d = lane-1 (:182-197)
            const int d = lane - 1;
            double e[DIM];
#pragma unroll
            for (int t = 0; t < DIM; ++t) e[t] = t == d ? 1.0 / den[t] : 0.0;
            pc[1] = e[0]; pc[2] = e[1];
            if (DIM == 2) pc[3] = e[0] * gn[1] + gn[0] * e[1];
            else {
                pc[3] = e[DIM - 1];
                pc[4] = e[0] * gn[1] + gn[0] * e[1];
                pc[5] = e[0] * gn[DIM - 1] + gn[0] * e[DIM - 1];
                pc[6] = e[1] * gn[DIM - 1] + gn[1] * e[DIM - 1];
                pc[7] = e[0] * gn[1] * gn[DIM - 1] + gn[0] * e[1] * gn[DIM - 1] + gn[0] * gn[1] * e[DIM - 1];
            }
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
            double s = 0.0;
            for (int a = 0; a < L; ++a) s += P[a * M + k] * W[a * NR + M + lane];
            y[k] = pc[k] - s;
        }
        double inv[M];
        ok = small_cholesky<M>(Mm, inv);
        small_solve<M>(Mm, inv, y);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok && lane == 0) atomicAdd(A.singular, 1);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double *Y = R;                     // R is no longer needed: park the (1+DIM) x M solutions there
    if (lane <= DIM) {
#pragma unroll
        for (int k = 0; k < M; ++k) Y[lane * M + k] = y[k];
    }
    __syncwarp();
    // phi_c[a] = s_c[a] + Z_P[a,:] . y_c
    for (int a = lane; a < L; a += 32) {
        double out[1 + DIM];
#pragma unroll
        for (int c = 0; c <= DIM; ++c) out[c] = W[a * NR + M + c];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const double z = W[a * NR + k];
#pragma unroll
            for (int c = 0; c <= DIM; ++c) out[c] += z * Y[c * M + k];
        }
        A.phi[o0 + a] = ok ? out[0] : nan;
#pragma unroll
        for (int d = 0; d < DIM; ++d) A.dphi[d * A.tl + o0 + a] = ok ? out[1 + d] : nan;
    }
}

// --------------------------------------------------------------------------------------------
// Register-resident variant for lists of at most 32 neighbours (default).  k_mk above spends ~6 000 FP64 warp instructions per
// Gauss point at half-empty warps: every lane builds its row of R with a trip count of its own (the warp pays the longest), the
// triangular solves run with one lane per right-hand side (7 of 32 lanes, two dependent shared-memory loads per step), and every
// coordinate difference is a division.  Here
//   * SG = 16 or 32 lanes own one Gauss point (two points per warp when no list is longer than 16), lane = row;
//   * the strict lower triangle of R is built PAIR-parallel (lane = pair (i,j), coordinates fetched with indexed shuffles), staged
//     in shared memory and read back row-wise into registers;
//   * the Cholesky factorisation is right-looking on register rows (column j of the factor is broadcast lane by lane), and the
//     forward substitution of all right-hand sides [P | r | dr] rides in the same column loop: Y = C^-1 W;
//   * M = P^T R^-1 P = Y_P^T Y_P and P^T R^-1 r_c = Y_P^T y_c need only the forward-solved rows (sub-group reductions), and
//     phi_c = R^-1 (r_c + P y_c)-type terms collapse to ONE backward substitution per output component,
//     phi_c = C^-T (y_c' + Y_P y_c), done right-looking with lane = row on the transposed factor (through shared memory).
// Same formulas as the kernel above in exact arithmetic.
// --------------------------------------------------------------------------------------------
template <int SG>
__device__ __forceinline__ double sg_sum(double v) {
#pragma unroll
    for (int sft = SG / 2; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
    return v;
}

// One column step of the right-looking Cholesky on register rows (lane = row, SG lanes per point), with the forward substitution
// of the NR right-hand sides W riding along.  The row is kept ROTATED: r[0] is always the current column j, and the trailing
// update writes r[k-1] = r[k] - c_ij c_(j+k)j, so the column index of the loop may be a run-time value although r[] lives in
// registers (a fully unrolled factorisation is 40 KB of straight-line code per list-length class: ncu showed 32 % of the stall
// samples waiting for instructions).  WIDTH = compile-time bound on the number of columns right of j; finished columns go to the
// shared-memory copy of the factor (row-major: the backward substitution reads it transposed).
template <int SG, int NR, int WIDTH>
__device__ __forceinline__ void mk_col_step(double (&r)[SG], double (&W)[NR], int sl, int j, bool &ok, double &myinv, double *__restrict__ myrow) {
    const double piv = __shfl_sync(0xffffffffu, r[0], j, SG);
    if (!(piv > 0.0)) ok = false;
    const double inv = rsqrt(piv);
    const double cij = sl >= j ? r[0] * inv : 0.0;          // lane j: sqrt(piv); lanes above the diagonal: 0
    myrow[j] = cij;
    if (sl == j) myinv = inv;
#pragma unroll
    for (int c = 0; c < NR; ++c) {
        const double yj = __shfl_sync(0xffffffffu, W[c], j, SG) * inv;
        W[c] = sl == j ? yj : fma(-cij, yj, W[c]);
    }
#pragma unroll
    for (int k = 1; k <= WIDTH; ++k) {
        const double ckj = __shfl_sync(0xffffffffu, cij, j + k, SG);     // j + k >= SG: wraps onto a dead entry
        r[k - 1] = fma(-cij, ckj, r[k]);
    }
}

// Columns [J0, J0 + 4) with the trailing width that J0 allows, then the next group of four (rolled loops, decreasing static widths).
template <int SG, int NR, int J0>
__device__ __forceinline__ void mk_chol_fwd(double (&r)[SG], double (&W)[NR], int sl, int Lmax, bool &ok, double &myinv, double *__restrict__ myrow) {
    if constexpr (J0 < SG) {
        const int jend = min(J0 + 4, Lmax);
#pragma unroll 1
        for (int j = J0; j < jend; ++j) mk_col_step<SG, NR, SG - 1 - J0>(r, W, sl, j, ok, myinv, myrow);
        if (Lmax > J0 + 4) mk_chol_fwd<SG, NR, J0 + 4>(r, W, sl, Lmax, ok, myinv, myrow);
    }
}

template <int DIM, int SHAPE, int SG>
__global__ void __launch_bounds__(128) k_mk_rows(MkArgs A) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NR = M + 1 + DIM;            // right-hand sides: P | r | dr_d
    constexpr int PPW = 32 / SG;               // Gauss points per warp
    constexpr int LD = SG + 1;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sl = lane & (SG - 1), sub = lane / SG;
    const int64_t g = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wid) * PPW + sub;
    double *Rs = sm + ((size_t)wid * PPW + sub) * (SG * LD);
    int64_t o0 = 0; int L = 0;
    if (g < A.ngauss) { o0 = A.off[g]; L = (int)(A.off[g + 1] - o0); }
    const int Lmax = __reduce_max_sync(0xffffffffu, L);     // redux.sync: the compiler knows the result is warp-uniform (no WARPSYNC wrappers around the shuffles below)
    if (Lmax == 0) return;                     // warp-uniform: every shuffle below is executed by all 32 lanes
    const double Cq = 4.3, theta = 0.1 * (Cq * Cq);
    const bool row = sl < L;
    // ---- my neighbour: coordinates, support; centre and scale of the point (:168-170) ------------------------------------
    double x[DIM], xye[DIM], den[DIM], iden[DIM], gp[DIM];
    {
        const int id = row ? A.dom[o0 + sl] : 0;
        double sx[DIM], sh[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            x[d] = row ? A.x[d * A.nnodes + id] : 0.0;
            sx[d] = sg_sum<SG>(x[d]);
            sh[d] = sg_sum<SG>(row ? (SHAPE == FG_SHAPE_RECTANGULAR ? A.dm[d * A.nnodes + id] : A.dm[id]) : 0.0);
            const double Ld = (double)max(L, 1);
            xye[d] = sx[d] / Ld;
            den[d] = (sh[d] / Ld) * Cq;
            if (L == 0) den[d] = 1.0;
            iden[d] = 1.0 / den[d];
            gp[d] = g < A.ngauss ? A.gs[d * A.ngauss + g] : 0.0;
        }
    }
    // ---- my row of [P | r | dr] --------------------------------------------------------------------------------------------
    double gn[DIM], W[NR];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gn[d] = (gp[d] - xye[d]) / den[d];
    {
        double n[DIM], b[M];
#pragma unroll
        for (int d = 0; d < DIM; ++d) n[d] = (x[d] - xye[d]) / den[d];
        mk_basis<DIM>(n, b);
        double arg = 0.0, ds[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { ds[d] = (gp[d] - x[d]) / den[d]; arg += ds[d] * ds[d]; }
        const double rv = exp(-theta * arg);
#pragma unroll
        for (int k = 0; k < M; ++k) W[k] = row ? b[k] : 0.0;
        W[M] = row ? rv : 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) W[M + 1 + d] = row ? rv * (-2.0 * theta * ds[d] / den[d]) : 0.0;
    }
    // ---- strict lower triangle of R, one pair per lane and round ---------------------------------------------------------------
    {
        const int npairs = L * (L - 1) / 2, npairs_max = Lmax * (Lmax - 1) / 2;
        for (int t0 = 0; t0 < npairs_max; t0 += SG) {
            const int t = t0 + sl;
            const bool valid = t < npairs;
            int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
            if (i * (i - 1) / 2 > t) --i;
            if ((i + 1) * i / 2 <= t) ++i;
            int j = t - i * (i - 1) / 2;
            if (!valid) { i = 0; j = 0; }
            double a2 = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double xi = __shfl_sync(0xffffffffu, x[d], i, SG), xj = __shfl_sync(0xffffffffu, x[d], j, SG);
                const double tt = (xi - xj) * iden[d];
                a2 += tt * tt;
            }
            if (valid) Rs[i * LD + j] = exp(-theta * a2);
        }
    }
    __syncwarp();
    double r[SG];                                // my row of R, then of its Cholesky factor (entries k <= sl; the rest stays 0)
#pragma unroll
    for (int k = 0; k < SG; ++k) {             // unconditional loads + selects: a guarded loop would index r[] dynamically (local memory)
        const double t = Rs[sl * LD + k];      // k >= sl: never written, never used
        r[k] = (row && k < sl) ? t : (k == sl ? 1.0 : 0.0);                                            // rows >= L: identity
    }
    // ---- right-looking Cholesky on (rotating) register rows + forward substitution of the right-hand sides; the finished columns
    // land in Rs (the staging copy of R is dead once every lane holds its row) ----------------------------------------------------
    bool ok = true;
    double myinv = 1.0;
    __syncwarp();
    mk_chol_fwd<SG, NR, 0>(r, W, sl, Lmax, ok, myinv, Rs + sl * LD);
    // ---- M = Y_P^T Y_P, q_c = Y_P^T y_c; y_c = M^-1 (p_c - q_c) (every lane, redundantly) ----------------------------------------
    double Mm[M * (M + 1) / 2];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j <= i) Mm[i * (i + 1) / 2 + j] = sg_sum<SG>(W[i] * W[j]);
    double inv4[M];
    if (!small_cholesky<M>(Mm, inv4)) ok = false;
    double b0[M];
    mk_basis<DIM>(gn, b0);
    double v[1 + DIM];
#pragma unroll
    for (int c = 0; c <= DIM; ++c) {
        double pc[M];
#pragma unroll
        for (int k = 0; k < M; ++k) pc[k] = 0.0;
        if (c == 0) {
#pragma unroll
            for (int k = 0; k < M; ++k) pc[k] = b0[k];
        } else {                       // d p / d g_d, d = c-1 (:182-197)
            double e[DIM];
#pragma unroll
            for (int t = 0; t < DIM; ++t) e[t] = t == c - 1 ? 1.0 / den[t] : 0.0;
            pc[1] = e[0]; pc[2] = e[1];
            if (DIM == 2) pc[3] = e[0] * gn[1] + gn[0] * e[1];
            else {
                pc[3] = e[DIM - 1];
                pc[4] = e[0] * gn[1] + gn[0] * e[1];
                pc[5] = e[0] * gn[DIM - 1] + gn[0] * e[DIM - 1];
                pc[6] = e[1] * gn[DIM - 1] + gn[1] * e[DIM - 1];
                pc[7] = e[0] * gn[1] * gn[DIM - 1] + gn[0] * e[1] * gn[DIM - 1] + gn[0] * gn[1] * e[DIM - 1];
            }
        }
        double y[M];
#pragma unroll
        for (int k = 0; k < M; ++k) y[k] = pc[k] - sg_sum<SG>(W[k] * W[M + c]);
        small_solve<M>(Mm, inv4, y);
        double t = W[M + c];
#pragma unroll
        for (int k = 0; k < M; ++k) t = fma(W[k], y[k], t);
        v[c] = t;                                  // my row of y_c' + Y_P y_c
    }
    // ---- phi_c = C^-T v_c: right-looking backward substitution, lane = row, on the transposed factor --------------------------
    __syncwarp();                                  // the factor's columns were written by their row lanes during the factorisation
    for (int i = Lmax - 1; i >= 0; --i) {
        const double invi = __shfl_sync(0xffffffffu, myinv, i, SG);
        const double cik = (sl < i && i < L) ? Rs[i * LD + sl] : 0.0;      // C[i][sl]; rows >= L are identity rows
#pragma unroll
        for (int c = 0; c <= DIM; ++c) {
            const double zi = __shfl_sync(0xffffffffu, v[c], i, SG) * invi;
            v[c] = sl == i ? zi : fma(-cik, zi, v[c]);
        }
    }
    // sub-group-uniform verdict (every lane saw the same pivots), one count per point
    if (!ok && sl == 0 && L > 0) atomicAdd(A.singular, 1);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (row) {
        A.phi[o0 + sl] = ok ? v[0] : nan;
#pragma unroll
        for (int d = 0; d < DIM; ++d) A.dphi[d * A.tl + o0 + sl] = ok ? v[1 + d] : nan;
    }
}

template <int DIM, int SHAPE, int SG>
static int run_rows(fg_ctx *ctx, GaussSet &S, MkArgs &A) {
    constexpr int PPW = 32 / SG, WARPS = 4;
    const size_t smem = sizeof(double) * WARPS * PPW * SG * (SG + 1);
    const int64_t per_block = (int64_t)WARPS * PPW;
    FG_CUDA(ctx, cudaFuncSetAttribute(k_mk_rows<DIM, SHAPE, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mk_rows<DIM, SHAPE, SG><<<(unsigned)((S.n + per_block - 1) / per_block), WARPS * 32, smem, ctx->stream>>>(A);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

template <int DIM, int SHAPE>
static int run(fg_ctx *ctx, GaussSet &S, MkArgs &A) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NR = M + 1 + DIM;
    const int Lm = A.maxL;
    if (Lm <= 32 && fg_opt(ctx, "mk_rows", 1) != 0) return Lm <= 16 ? run_rows<DIM, SHAPE, 16>(ctx, S, A) : run_rows<DIM, SHAPE, 32>(ctx, S, A);
    const size_t per_warp = sizeof(double) * ((size_t)Lm * (Lm | 1) + (size_t)Lm * (NR + M + DIM));
    int warps = 4;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    if (per_warp * warps > 200 * 1024)
        return fg_fail(ctx, FG_E_CAPACITY, "fg_shape_functions(MK): a Gauss point has %d neighbours; the kernel holds at most ~150", Lm);
    FG_CUDA(ctx, cudaFuncSetAttribute(k_mk<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * warps)));
    k_mk<DIM, SHAPE><<<(unsigned)((S.n + warps - 1) / warps), warps * 32, per_warp * warps, ctx->stream>>>(A);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

int fg_impl_mk(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    S.singular = 0;
    if (S.n == 0 || S.total_L == 0) return FG_OK;
    FG_CUDA(ctx, S.singular_dev.reserve(1));
    FG_CUDA(ctx, cudaMemsetAsync(S.singular_dev.p, 0, sizeof(int), ctx->stream));
    MkArgs A;
    { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }
    A.nnodes = N.n; A.ngauss = S.n; A.x = N.x.p; A.dm = N.dm.p; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p;
    A.phi = S.phi.p; A.dphi = S.dphi.p; A.tl = S.total_L; A.singular = S.singular_dev.p; A.maxL = S.max_L;
    if (N.dim == 2) return N.shape == FG_SHAPE_RECTANGULAR ? run<2, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<2, FG_SHAPE_CIRCULAR>(ctx, S, A);
    return N.shape == FG_SHAPE_RECTANGULAR ? run<3, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<3, FG_SHAPE_CIRCULAR>(ctx, S, A);
}

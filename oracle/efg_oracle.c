/*
 * efg_oracle.c -- CPU restatement of FlexiGal's Element-Free-Galerkin hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  Nothing under flexigal_b200/ links or imports it.
 *
 * PARITY UNPINNED: the reference (pure Julia, /root/reference) ships no golden
 * vectors, no assertions and no test directory (only a CI job that runs four
 * examples), and Julia is not installed here, so this restatement cannot be
 * checked against reference outputs.  It follows the reference source line by
 * line (citations below, file:line under /root/reference) and is additionally
 * validated by (i) a second build of the very same code in __float128
 * ("truth", -DFGO_QUAD) and (ii) the analytic MLS invariants in tests/.
 *
 * Third-party arithmetic the reference leans on and that is NOT in the tree:
 * OpenBLAS 0.3.27 ddot/dgemm/dgemv (Manifest.toml:1218-1221) for the dots in
 * src/Shape_Functions.jl:138,143,148-158, LAPACK dgetrf/dgetrs for :226-232,
 * SparseArrays 1.11.0 sparse() (Manifest.toml:1716-1719).  Their summation
 * order is restated here as the plain textual left-to-right order.
 *
 * Data conventions are Julia's: column-major, x[nnodes][dim] stored as
 * x[d*nnodes + i], gs[ngauss][dim+2] stored as gs[c*ngauss + g]; node ids in
 * DOM are 1-based and ascending (findall).
 *
 * Build (see oracle/Makefile):
 *   gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC           -> libefg_oracle.so
 *   gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC -DFGO_QUAD -> libefg_oracle_q.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef FGO_QUAD
typedef __float128 real;
/* Newton square root in __float128 (avoids a libquadmath dependency). */
static real r_sqrt(real a) {
    if (a <= 0) return 0;
    real y = (real)sqrt((double)a);
    for (int it = 0; it < 4; ++it) y = (y + a / y) / 2;
    return y;
}
/* exp in __float128: argument reduction by 2^-k + Taylor series, squared back. */
static real r_exp(real a) {
    int neg = a < 0; if (neg) a = -a;
    int k = 0; while (a > (real)0.0625) { a /= 2; ++k; }
    real term = 1, sum = 1;
    for (int n = 1; n < 40; ++n) { term = term * a / n; sum += term; }
    for (int i = 0; i < k; ++i) sum = sum * sum;
    return neg ? 1 / sum : sum;
}
#define FGO_NAME(n) fgoq_##n
#else
typedef double real;
static real r_sqrt(real a) { return sqrt(a); }
static real r_exp(real a) { return exp(a); }
#define FGO_NAME(n) fgo_##n
#endif

#define FGO_RECT 0
#define FGO_CIRC 1
#define FGO_IMLS 0
#define FGO_MK 1
#define FGO_MAXPOLY 8

static real r_abs(real a) { return a < 0 ? -a : a; }
static real r_sign(real a) { return a > 0 ? (real)1 : (a < 0 ? (real)-1 : (real)0); } /* Julia sign(0)=0 */

/* ------------------------------------------------------------------------------------------
 * domain(): src/Shape_Functions.jl:3-27.  Brute-force scan over all nodes; ascending 1-based
 * ids (findall).  The predicate is evaluated in double in BOTH builds: the neighbour list is
 * an integer result defined by the reference's FP64 comparison, not by exact arithmetic.
 * Returns L; writes ids if out != NULL.
 * ---------------------------------------------------------------------------------------- */
int64_t FGO_NAME(domain)(int dim, int64_t nnodes, const double *x, const double *dm, int shape,
                         const double *gpos, int64_t *out)
{
    int64_t L = 0;
    for (int64_t i = 0; i < nnodes; ++i) {
        int inside;
        if (shape == FGO_RECT) {
            inside = 1;
            for (int j = 0; j < dim; ++j)                                   /* :7-14 */
                inside &= fabs(x[j * nnodes + i] - gpos[j]) <= dm[j * nnodes + i];
        } else {
            double dist_sq = 0.0;                                           /* :16-22 */
            for (int j = 0; j < dim; ++j) {
                double t = x[j * nnodes + i] - gpos[j];
                dist_sq += t * t;
            }
            inside = dist_sq <= dm[i] * dm[i];
        }
        if (inside) { if (out) out[L] = i + 1; ++L; }
    }
    return L;
}

/* ------------------------------------------------------------------------------------------
 * cubwgt(): src/Shape_Functions.jl:29-86.  dif is L x dim (column-major, ld = L).
 * ---------------------------------------------------------------------------------------- */
static void cubwgt(int dim, int L, const real *dif, const int64_t *v, int64_t nnodes, const double *dm,
                   int shape, real *w, real *dw)
{
    const real c43 = (real)4 / 3, c23 = (real)2 / 3;
    for (int i = 0; i < L; ++i) { w[i] = 0; for (int d = 0; d < dim; ++d) dw[d * L + i] = 0; }
    if (shape == FGO_RECT) {
        for (int i = 0; i < L; ++i) {
            real wt[3], dwt[3];
            for (int d = 0; d < dim; ++d) {
                real dmv = (real)dm[d * nnodes + (v[i] - 1)];
                real dr = r_sign(dif[d * L + i]) / dmv;                      /* :38 */
                real r = r_abs(dif[d * L + i]) / dmv;                        /* :39 */
                if (r > (real)0.5) {                                        /* :40-42 */
                    wt[d] = c43 - 4 * r + 4 * (r * r) - c43 * (r * r * r);
                    dwt[d] = (-4 + 8 * r - 4 * (r * r)) * dr;
                } else {                                                    /* :43-45 */
                    wt[d] = c23 - 4 * (r * r) + 4 * (r * r * r);
                    dwt[d] = (-8 * r + 12 * (r * r)) * dr;
                }
            }
            if (dim == 2) {                                                 /* :48-51 */
                w[i] = wt[0] * wt[1];
                dw[0 * L + i] = dwt[0] * wt[1];
                dw[1 * L + i] = wt[0] * dwt[1];
            } else {                                                        /* :52-57 */
                w[i] = wt[0] * wt[1] * wt[2];
                dw[0 * L + i] = dwt[0] * wt[1] * wt[2];
                dw[1 * L + i] = wt[0] * dwt[1] * wt[2];
                dw[2 * L + i] = wt[0] * wt[1] * dwt[2];
            }
        }
    } else {
        for (int i = 0; i < L; ++i) {                                       /* :60-83 */
            real dmi = (real)dm[v[i] - 1];
            real dist = 0;
            for (int d = 0; d < dim; ++d) dist += dif[d * L + i] * dif[d * L + i];
            dist = r_sqrt(dist);
            real r = dist / dmi;
            if (r <= 1) {
                real f, df;
                if (r <= (real)0.5) {
                    f = c23 - 4 * (r * r) + 4 * (r * r * r);
                    df = -8 * r + 12 * (r * r);
                } else {
                    f = c43 - 4 * r + 4 * (r * r) - c43 * (r * r * r);
                    df = -4 + 8 * r - 4 * (r * r);
                }
                w[i] = f;
                if (dist > (real)1e-12)
                    for (int d = 0; d < dim; ++d) dw[d * L + i] = df * (1 / dmi) * (dif[d * L + i] / dist);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Improved_Moving_Least_Squares(): src/Shape_Functions.jl:88-162.  Literal: unshifted
 * absolute-coordinate basis, weighted CLASSICAL Gram-Schmidt with q[:,i] in the numerator
 * (:138), only diag(aa) used (:144).  phi[L], dphi[L][dim] (row-major per neighbour, the
 * SVector layout of :156-159).
 * ---------------------------------------------------------------------------------------- */
static void imls(int dim, int L, const double *gpos_d, int64_t nnodes, const double *x, const int64_t *v,
                 const double *dm, int shape, double *phi_out, double *dphi_out, real *work)
{
    const int m = dim == 2 ? 4 : 8;
    real *q = work;              /* L x m column-major */
    real *p = q + (size_t)L * m;
    real *dif = p + (size_t)L * m;   /* L x dim */
    real *w = dif + (size_t)L * 3;
    real *dw = w + L;            /* L x dim */
    real gpos[3];
    for (int d = 0; d < dim; ++d) gpos[d] = (real)gpos_d[d];
    for (int i = 0; i < L; ++i) {                                           /* :92-112 */
        real x_ = (real)x[0 * nnodes + v[i] - 1], y_ = (real)x[1 * nnodes + v[i] - 1];
        if (dim == 2) {
            q[0 * L + i] = 1; q[1 * L + i] = x_; q[2 * L + i] = y_; q[3 * L + i] = x_ * y_;
        } else {
            real z_ = (real)x[2 * nnodes + v[i] - 1];
            q[0 * L + i] = 1; q[1 * L + i] = x_; q[2 * L + i] = y_; q[3 * L + i] = z_;
            q[4 * L + i] = x_ * y_; q[5 * L + i] = x_ * z_; q[6 * L + i] = y_ * z_; q[7 * L + i] = x_ * y_ * z_;
        }
        for (int d = 0; d < dim; ++d) dif[d * L + i] = gpos[d] - (real)x[d * nnodes + v[i] - 1];   /* :116-121 */
    }
    cubwgt(dim, L, dif, v, nnodes, dm, shape, w, dw);                       /* :122 */
    real PPg[4][FGO_MAXPOLY];                                               /* :123-132 */
    memset(PPg, 0, sizeof PPg);
    if (dim == 2) {
        PPg[0][0] = 1; PPg[0][1] = gpos[0]; PPg[0][2] = gpos[1]; PPg[0][3] = gpos[0] * gpos[1];
        PPg[1][1] = 1; PPg[1][3] = gpos[1];
        PPg[2][2] = 1; PPg[2][3] = gpos[0];
    } else {
        PPg[0][0] = 1; PPg[0][1] = gpos[0]; PPg[0][2] = gpos[1]; PPg[0][3] = gpos[2];
        PPg[0][4] = gpos[0] * gpos[1]; PPg[0][5] = gpos[0] * gpos[2]; PPg[0][6] = gpos[1] * gpos[2];
        PPg[0][7] = gpos[0] * gpos[1] * gpos[2];
        PPg[1][1] = 1; PPg[1][4] = gpos[1]; PPg[1][5] = gpos[2]; PPg[1][7] = gpos[1] * gpos[2];
        PPg[2][2] = 1; PPg[2][4] = gpos[0]; PPg[2][6] = gpos[2]; PPg[2][7] = gpos[0] * gpos[2];
        PPg[3][3] = 1; PPg[3][5] = gpos[0]; PPg[3][6] = gpos[1]; PPg[3][7] = gpos[0] * gpos[1];
    }
    memcpy(p, q, sizeof(real) * (size_t)L * m);                             /* :134 */
    for (int i = 1; i < m; ++i) {                                           /* :136-142 */
        for (int k = 0; k < i; ++k) {
            real num = 0, den = 0;
            for (int a = 0; a < L; ++a) {
                real wp = w[a] * p[k * L + a];
                num += q[i * L + a] * wp;
                den += p[k * L + a] * wp;
            }
            real c1 = num / den;
            for (int a = 0; a < L; ++a) p[i * L + a] -= c1 * p[k * L + a];
            for (int r = 0; r <= dim; ++r) PPg[r][i] -= c1 * PPg[r][k];
        }
    }
    real aia[FGO_MAXPOLY], gam[4][FGO_MAXPOLY];
    for (int k = 0; k < m; ++k) {                                           /* :143-146 */
        real s = 0;
        for (int a = 0; a < L; ++a) s += p[k * L + a] * (w[a] * p[k * L + a]);
        aia[k] = 1 / s;
        gam[0][k] = aia[k] * PPg[0][k];
    }
    for (int d = 0; d < dim; ++d) {                                         /* :147-150 */
        for (int r = 0; r < m; ++r) {
            real t = 0;
            for (int c = 0; c < m; ++c) {
                real daa = 0;
                for (int a = 0; a < L; ++a) daa += p[r * L + a] * (dw[d * L + a] * p[c * L + a]);
                t += daa * gam[0][c];
            }
            gam[d + 1][r] = aia[r] * (PPg[d + 1][r] - t);
        }
    }
    for (int a = 0; a < L; ++a) {                                           /* :151-158 */
        real ph = 0;
        for (int k = 0; k < m; ++k) ph += (p[k * L + a] * w[a]) * gam[0][k];
        phi_out[a] = (double)ph;
        for (int d = 0; d < dim; ++d) {
            real t1 = 0, t2 = 0;
            for (int k = 0; k < m; ++k) t1 += (p[k * L + a] * w[a]) * gam[d + 1][k];
            for (int k = 0; k < m; ++k) t2 += (p[k * L + a] * dw[d * L + a]) * gam[0][k];
            dphi_out[a * dim + d] = (double)(t1 + t2);
        }
    }
}

/* Dense LU with partial pivoting (LAPACK dgetrf semantics, unblocked) and solves; n x n
 * column-major, in place.  Used by Moving_Kriging (src/Shape_Functions.jl:226-230). */
static void lu_factor(int n, real *A, int *piv)
{
    for (int k = 0; k < n; ++k) {
        int pv = k; real mx = r_abs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (r_abs(A[k * n + i]) > mx) { mx = r_abs(A[k * n + i]); pv = i; }
        piv[k] = pv;
        if (pv != k) for (int j = 0; j < n; ++j) { real t = A[j * n + k]; A[j * n + k] = A[j * n + pv]; A[j * n + pv] = t; }
        real inv = 1 / A[k * n + k];
        for (int i = k + 1; i < n; ++i) A[k * n + i] *= inv;
        for (int j = k + 1; j < n; ++j) {
            real akj = A[j * n + k];
            for (int i = k + 1; i < n; ++i) A[j * n + i] -= A[k * n + i] * akj;
        }
    }
}
static void lu_solve(int n, const real *A, const int *piv, real *b, int nrhs)
{
    for (int c = 0; c < nrhs; ++c) {
        real *y = b + (size_t)c * n;
        for (int k = 0; k < n; ++k) if (piv[k] != k) { real t = y[k]; y[k] = y[piv[k]]; y[piv[k]] = t; }
        for (int k = 0; k < n; ++k) for (int i = k + 1; i < n; ++i) y[i] -= A[k * n + i] * y[k];
        for (int k = n - 1; k >= 0; --k) { y[k] /= A[k * n + k]; for (int i = 0; i < k; ++i) y[i] -= A[k * n + i] * y[k]; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Moving_Kriging(): src/Shape_Functions.jl:164-239 (Cq = 4.3, theta = 0.1 Cq^2).
 * ---------------------------------------------------------------------------------------- */
static void mk(int dim, int L, const double *gpos_d, int64_t nnodes, const double *x, const int64_t *v,
               const double *dm, int shape, double *phi_out, double *dphi_out)
{
    const int m = dim == 2 ? 4 : 8;
    const real Cq = (real)4.3;                 /* the reference's Float64 literal 4.3 */
    const real theta = (real)0.1 * (Cq * Cq);  /* :167 */
    real xye[3], den[3], gn[3];
    for (int d = 0; d < dim; ++d) {                                          /* :168-170 */
        real s = 0, h = 0;
        for (int i = 0; i < L; ++i) s += (real)x[d * nnodes + v[i] - 1];
        xye[d] = s / L;
        if (shape == FGO_RECT) { for (int i = 0; i < L; ++i) h += (real)dm[d * nnodes + v[i] - 1]; }
        else { for (int i = 0; i < L; ++i) h += (real)dm[v[i] - 1]; }
        den[d] = (h / L) * Cq;
        gn[d] = ((real)gpos_d[d] - xye[d]) / den[d];                        /* :175 */
    }
    real *P = malloc(sizeof(real) * ((size_t)L * m * 3 + (size_t)L * L * 2 + (size_t)L * 4 + 64));
    real *R1P = P + (size_t)L * m;           /* L x m */
    real *A = R1P + (size_t)L * m;           /* m x L */
    real *R = A + (size_t)L * m;             /* L x L */
    real *Bm = R + (size_t)L * L;            /* L x L */
    real *r = Bm + (size_t)L * L, *dr = r + L;                               /* dr: L x dim */
    real pv[FGO_MAXPOLY], dp[3][FGO_MAXPOLY];
    memset(dp, 0, sizeof dp);
    for (int i = 0; i < L; ++i) {                                            /* :176-191 */
        real n[3];
        for (int d = 0; d < dim; ++d) n[d] = ((real)x[d * nnodes + v[i] - 1] - xye[d]) / den[d];
        P[0 * L + i] = 1;
        if (dim == 2) { P[1 * L + i] = n[0]; P[2 * L + i] = n[1]; P[3 * L + i] = n[0] * n[1]; }
        else {
            P[1 * L + i] = n[0]; P[2 * L + i] = n[1]; P[3 * L + i] = n[2];
            P[4 * L + i] = n[0] * n[1]; P[5 * L + i] = n[0] * n[2]; P[6 * L + i] = n[1] * n[2];
            P[7 * L + i] = n[0] * n[1] * n[2];
        }
    }
    pv[0] = 1;
    if (dim == 2) {                                                          /* :182-184 */
        pv[1] = gn[0]; pv[2] = gn[1]; pv[3] = gn[0] * gn[1];
        dp[0][1] = 1 / den[0]; dp[0][3] = gn[1] / den[0];
        dp[1][2] = 1 / den[1]; dp[1][3] = gn[0] / den[1];
    } else {                                                                 /* :192-197 */
        pv[1] = gn[0]; pv[2] = gn[1]; pv[3] = gn[2];
        pv[4] = gn[0] * gn[1]; pv[5] = gn[0] * gn[2]; pv[6] = gn[1] * gn[2]; pv[7] = gn[0] * gn[1] * gn[2];
        dp[0][1] = 1 / den[0]; dp[0][4] = gn[1] / den[0]; dp[0][5] = gn[2] / den[0]; dp[0][7] = gn[1] * gn[2] / den[0];
        dp[1][2] = 1 / den[1]; dp[1][4] = gn[0] / den[1]; dp[1][6] = gn[2] / den[1]; dp[1][7] = gn[0] * gn[2] / den[1];
        dp[2][3] = 1 / den[2]; dp[2][5] = gn[0] / den[2]; dp[2][6] = gn[1] / den[2]; dp[2][7] = gn[0] * gn[1] / den[2];
    }
    for (int i = 0; i < L; ++i)                                              /* :199-208 */
        for (int j = 0; j < L; ++j) {
            real arg = 0;
            for (int d = 0; d < dim; ++d) {
                real t = ((real)x[d * nnodes + v[i] - 1] - (real)x[d * nnodes + v[j] - 1]) / den[d];
                arg += t * t;
            }
            R[j * L + i] = r_exp(-theta * arg);
        }
    for (int i = 0; i < L; ++i) {                                            /* :209-225 */
        real arg = 0, ds[3];
        for (int d = 0; d < dim; ++d) {
            ds[d] = ((real)gpos_d[d] - (real)x[d * nnodes + v[i] - 1]) / den[d];
            arg += ds[d] * ds[d];
        }
        real val = r_exp(-theta * arg);
        r[i] = val;
        for (int d = 0; d < dim; ++d) dr[d * L + i] = val * (-2 * theta * ds[d] / den[d]);
    }
    int *piv = malloc(sizeof(int) * (L + 16)), pivM[FGO_MAXPOLY];
    lu_factor(L, R, piv);                                                    /* :226 */
    memcpy(R1P, P, sizeof(real) * (size_t)L * m);
    lu_solve(L, R, piv, R1P, m);                                             /* :227  R1P = R \ P */
    real M[FGO_MAXPOLY * FGO_MAXPOLY];
    for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) {                /* :228  M = P' R1P */
        real s = 0; for (int a = 0; a < L; ++a) s += P[i * L + a] * R1P[j * L + a];
        M[j * m + i] = s;
    }
    /* A = M \ R1P'  (m x L, column-major ld m), :229 */
    for (int a = 0; a < L; ++a) for (int i = 0; i < m; ++i) A[a * m + i] = R1P[i * L + a];
    lu_factor(m, M, pivM);
    lu_solve(m, M, pivM, A, L);
    for (int j = 0; j < L; ++j) for (int i = 0; i < L; ++i) {                /* :230  B = R \ (I - P A) */
        real s = 0; for (int k = 0; k < m; ++k) s += P[k * L + i] * A[j * m + k];
        Bm[j * L + i] = (i == j ? (real)1 : (real)0) - s;
    }
    lu_solve(L, R, piv, Bm, L);
    for (int a = 0; a < L; ++a) {                                            /* :231-232 */
        real s = 0, s2 = 0;
        for (int k = 0; k < m; ++k) s += pv[k] * A[a * m + k];
        for (int i = 0; i < L; ++i) s2 += r[i] * Bm[a * L + i];
        phi_out[a] = (double)(s + s2);
        for (int d = 0; d < dim; ++d) {
            real t = 0, t2 = 0;
            for (int k = 0; k < m; ++k) t += dp[d][k] * A[a * m + k];
            for (int i = 0; i < L; ++i) t2 += dr[d * L + i] * Bm[a * L + i];
            dphi_out[a * dim + d] = (double)(t + t2);
        }
    }
    free(piv); free(P);
}

/* ------------------------------------------------------------------------------------------
 * SHAPE_FUN(): src/Shape_Functions.jl:241-263, two-phase for a C caller.
 *   count phase  : offsets[g+1]-offsets[g] = L_g (offsets 0-based, length ngauss+1)
 *   fill phase   : dom (1-based ascending), phi, dphi[sumL][dim]
 * The loop over Gauss points (:248) carries no state, so it is OpenMP-parallel here; with
 * nthreads = 1 it is the reference's serial loop.
 * ---------------------------------------------------------------------------------------- */
int FGO_NAME(shape_fun_count)(int dim, int64_t nnodes, const double *x, const double *dm, int shape,
                              int64_t ngauss, const double *gs, int64_t *offsets, int nthreads)
{
    offsets[0] = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t g = 0; g < ngauss; ++g) {
        double gpos[3];
        for (int d = 0; d < dim; ++d) gpos[d] = gs[d * ngauss + g];
        offsets[g + 1] = FGO_NAME(domain)(dim, nnodes, x, dm, shape, gpos, NULL);
    }
    for (int64_t g = 0; g < ngauss; ++g) offsets[g + 1] += offsets[g];
    return 0;
}

int FGO_NAME(shape_fun_fill)(int dim, int64_t nnodes, const double *x, const double *dm, int shape, int technique,
                             int64_t ngauss, const double *gs, const int64_t *offsets,
                             int64_t *dom, double *phi, double *dphi, int nthreads)
{
    int bad = 0;
#pragma omp parallel num_threads(nthreads)
    {
        size_t cap = 0; real *work = NULL;
#pragma omp for schedule(dynamic, 16)
        for (int64_t g = 0; g < ngauss; ++g) {
            double gpos[3];
            for (int d = 0; d < dim; ++d) gpos[d] = gs[d * ngauss + g];
            int64_t off = offsets[g];
            int L = (int)FGO_NAME(domain)(dim, nnodes, x, dm, shape, gpos, dom + off);   /* :250 */
            if (L != offsets[g + 1] - off) { bad = 1; continue; }
            if ((size_t)L > cap) { cap = (size_t)L + 64; free(work); work = malloc(sizeof(real) * cap * 32); }
            if (technique == FGO_IMLS)                                                   /* :251-252 */
                imls(dim, L, gpos, nnodes, x, dom + off, dm, shape, phi + off, dphi + off * dim, work);
            else                                                                         /* :253-254 */
                mk(dim, L, gpos, nnodes, x, dom + off, dm, shape, phi + off, dphi + off * dim);
        }
        free(work);
    }
    return bad;
}

/* Shape functions for given neighbour lists only (no search): lets the tests evaluate the
 * __float128 truth on exactly the DOM the FP64 predicate selected. */
int FGO_NAME(shape_fun_given_dom)(int dim, int64_t nnodes, const double *x, const double *dm, int shape, int technique,
                                  int64_t ngauss, const double *gs, const int64_t *offsets, const int64_t *dom,
                                  double *phi, double *dphi, int nthreads)
{
#pragma omp parallel num_threads(nthreads)
    {
        size_t cap = 0; real *work = NULL;
#pragma omp for schedule(dynamic, 16)
        for (int64_t g = 0; g < ngauss; ++g) {
            double gpos[3];
            for (int d = 0; d < dim; ++d) gpos[d] = gs[d * ngauss + g];
            int64_t off = offsets[g];
            int L = (int)(offsets[g + 1] - off);
            if ((size_t)L > cap) { cap = (size_t)L + 64; free(work); work = malloc(sizeof(real) * cap * 32); }
            if (technique == FGO_IMLS) imls(dim, L, gpos, nnodes, x, dom + off, dm, shape, phi + off, dphi + off * dim, work);
            else mk(dim, L, gpos, nnodes, x, dom + off, dm, shape, phi + off, dphi + off * dim);
        }
        free(work);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Bilinear assembly: src/Assembling_Operators.jl:2-26 (scalar), :28-79 (vector), :81-104.
 *
 * The user's closure f_pure(sa, sb) is bilinear in B_a = [phi_a, d1 phi_a, ..] and B_b, so it is
 * restated as a coefficient tensor C[(i,al),(j,be)]:
 *     K_ab[i][j] = ( sum_{al,be} B_a[al] * C[(i*nb+al)*(D*nb) + (j*nb+be)] * B_b[be] ) * weightjac
 * with nb = 1+dim; a = test (row), b = trial (column).  coef_stride = 0: one constant tensor;
 * otherwise coef + g*coef_stride is the tensor of Gauss point g.
 *
 * COO triplets are generated in the reference's order (Gauss point, a, b; vector case
 * column-of-block outer, row inner, :65-75) and reduced like SparseArrays.sparse(): duplicates
 * summed in input order, structural zeros kept.  Two-phase: pattern (colptr/rowval, 1-based,
 * Int64), then values.
 * ---------------------------------------------------------------------------------------- */
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }

/* Node-level pattern: column j holds the sorted union of DOM_g over all g containing j. */
int64_t FGO_NAME(pattern)(int64_t nnodes, int64_t ngauss, const int64_t *offsets, const int64_t *dom,
                          int64_t *colptr, int64_t *rowval)
{
    int64_t *cnt = calloc((size_t)nnodes + 1, sizeof(int64_t));
    for (int64_t g = 0; g < ngauss; ++g) {
        int64_t L = offsets[g + 1] - offsets[g];
        for (int64_t k = offsets[g]; k < offsets[g + 1]; ++k) cnt[dom[k]] += L;       /* upper bound */
    }
    int64_t *start = malloc(sizeof(int64_t) * ((size_t)nnodes + 2));
    start[0] = 0;
    for (int64_t j = 1; j <= nnodes; ++j) start[j] = start[j - 1] + cnt[j];
    int64_t *buf = malloc(sizeof(int64_t) * (size_t)(start[nnodes] + 1));
    int64_t *fill = calloc((size_t)nnodes + 1, sizeof(int64_t));
    for (int64_t g = 0; g < ngauss; ++g)
        for (int64_t kb = offsets[g]; kb < offsets[g + 1]; ++kb) {
            int64_t j = dom[kb];
            for (int64_t ka = offsets[g]; ka < offsets[g + 1]; ++ka) buf[start[j - 1] + fill[j]++] = dom[ka];
        }
    int64_t nnz = 0;
    colptr[0] = 1;
    for (int64_t j = 1; j <= nnodes; ++j) {
        int64_t *b = buf + start[j - 1], n = fill[j];
        qsort(b, (size_t)n, sizeof(int64_t), cmp_i64);
        int64_t last = -1;
        for (int64_t k = 0; k < n; ++k) if (b[k] != last) { last = b[k]; if (rowval) rowval[nnz] = last; ++nnz; }
        colptr[j] = nnz + 1;
    }
    free(cnt); free(start); free(buf); free(fill);
    return nnz;
}

static int64_t find_row(const int64_t *rowval, int64_t lo, int64_t hi, int64_t row)
{
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (rowval[mid] < row) lo = mid + 1; else hi = mid; }
    return lo;
}

/* nzval has D*D*nnz_node entries laid out as the DOF-level CSC of the reference
 * (global_col = (node_j-1)*D + jc, global_row = (node_i-1)*D + ir, :65-68):
 *   pos(j, jc, k, ir) = (colptr[j]-1)*D*D + jc*(n_j*D) + k*D + ir,  n_j = colptr[j+1]-colptr[j],
 * k = index of node_i within column j. */
int FGO_NAME(assemble_bilinear)(int dim, int D, int64_t ngauss, const double *gs, const int64_t *offsets,
                                const int64_t *dom, const double *phi, const double *dphi,
                                const double *coef, int64_t coef_stride,
                                const int64_t *colptr, const int64_t *rowval, double *nzval_out, int64_t nnz_node)
{
    const int nb = 1 + dim, n = D * nb;
    real *nz = calloc((size_t)nnz_node * D * D, sizeof(real));
    for (int64_t g = 0; g < ngauss; ++g) {
        const double *C = coef + g * coef_stride;
        real weightjac = (real)gs[(dim + 1) * ngauss + g] * (real)gs[dim * ngauss + g];     /* :9 */
        int64_t off = offsets[g], L = offsets[g + 1] - off;
        for (int64_t a = 0; a < L; ++a) {
            real Ba[4]; Ba[0] = (real)phi[off + a];
            for (int d = 0; d < dim; ++d) Ba[1 + d] = (real)dphi[(off + a) * dim + d];
            for (int64_t b = 0; b < L; ++b) {
                real Bb[4]; Bb[0] = (real)phi[off + b];
                for (int d = 0; d < dim; ++d) Bb[1 + d] = (real)dphi[(off + b) * dim + d];
                int64_t j = dom[off + b], i = dom[off + a];
                int64_t c0 = colptr[j - 1] - 1, nj = colptr[j] - colptr[j - 1];
                int64_t k = find_row(rowval, c0, c0 + nj, i) - c0;
                for (int jc = 0; jc < D; ++jc)
                    for (int ir = 0; ir < D; ++ir) {
                        real s = 0;
                        for (int al = 0; al < nb; ++al)
                            for (int be = 0; be < nb; ++be) {
                                real c = (real)C[(ir * nb + al) * n + (jc * nb + be)];
                                if (c != 0) s += Ba[al] * c * Bb[be];
                            }
                        nz[c0 * D * D + jc * (nj * D) + k * D + ir] += s * weightjac;            /* :20, :62 */
                    }
            }
        }
    }
    for (int64_t t = 0; t < nnz_node * D * D; ++t) nzval_out[t] = (double)nz[t];
    free(nz);
    return 0;
}

/* Linear form: src/Assembling_Operators.jl:105-122 (scalar), :124-145 (vector), :147-165.
 *   F[(node-1)*D + i] += ( sum_al c[i*nb+al] * B_a[al] ) * weightjac, sequential += in Gauss order. */
int FGO_NAME(assemble_linear)(int dim, int D, int64_t ngauss, const double *gs, const int64_t *offsets,
                              const int64_t *dom, const double *phi, const double *dphi,
                              const double *coef, int64_t coef_stride, double *F_out, int64_t nnodes)
{
    const int nb = 1 + dim;
    real *F = calloc((size_t)nnodes * D, sizeof(real));
    for (int64_t g = 0; g < ngauss; ++g) {
        const double *c = coef + g * coef_stride;
        real weightjac = (real)gs[(dim + 1) * ngauss + g] * (real)gs[dim * ngauss + g];
        for (int64_t k = offsets[g]; k < offsets[g + 1]; ++k) {
            for (int i = 0; i < D; ++i) {
                real s = 0;
                if (c[i * nb] != 0) s += (real)c[i * nb] * (real)phi[k];
                for (int d = 0; d < dim; ++d) if (c[i * nb + 1 + d] != 0) s += (real)c[i * nb + 1 + d] * (real)dphi[k * dim + d];
                F[(dom[k] - 1) * D + i] += s * weightjac;                                       /* :119, :141 */
            }
        }
    }
    for (int64_t t = 0; t < nnodes * D; ++t) F_out[t] = (double)F[t];
    free(F);
    return 0;
}

/* Literal COO push + counting-sort reduction, exactly as _assemble_scalar! + sparse() do for
 * the Poisson stiffness closure (dphi_a . dphi_b) -- used only as the timed CPU baseline leg
 * (it materialises sum L^2 triplets like the reference) and to cross-check the pattern path. */
int64_t FGO_NAME(assemble_gradgrad_coo)(int dim, int64_t nnodes, int64_t ngauss, const double *gs,
                                        const int64_t *offsets, const int64_t *dom, const double *dphi,
                                        int64_t *colptr, int64_t *rowval, double *nzval)
{
    int64_t total = 0;
    for (int64_t g = 0; g < ngauss; ++g) { int64_t L = offsets[g + 1] - offsets[g]; total += L * L; }   /* :92 */
    int64_t *row = malloc(sizeof(int64_t) * (size_t)(total + 1)), *col = malloc(sizeof(int64_t) * (size_t)(total + 1));
    double *val = malloc(sizeof(double) * (size_t)(total + 1));
    int64_t pos = 0;
    for (int64_t g = 0; g < ngauss; ++g) {                                                          /* :6-25 */
        double weightjac = gs[(dim + 1) * ngauss + g] * gs[dim * ngauss + g];
        int64_t off = offsets[g], L = offsets[g + 1] - off;
        for (int64_t a = 0; a < L; ++a)
            for (int64_t b = 0; b < L; ++b) {
                double s = dphi[(off + a) * dim] * dphi[(off + b) * dim];
                for (int d = 1; d < dim; ++d) s += dphi[(off + a) * dim + d] * dphi[(off + b) * dim + d];
                row[pos] = dom[off + a]; col[pos] = dom[off + b]; val[pos] = s * weightjac; ++pos;
            }
    }
    /* sparse(): stable counting sort by column, then by row inside each column, duplicates
     * summed in input order, explicit zeros kept. */
    int64_t *cstart = calloc((size_t)nnodes + 2, sizeof(int64_t));
    for (int64_t t = 0; t < total; ++t) cstart[col[t] + 1]++;
    cstart[0] = 0; cstart[1] = 0;
    for (int64_t j = 1; j <= nnodes; ++j) cstart[j + 1] += cstart[j];
    int64_t *perm = malloc(sizeof(int64_t) * (size_t)(total + 1));
    int64_t *cfill = calloc((size_t)nnodes + 2, sizeof(int64_t));
    for (int64_t t = 0; t < total; ++t) { int64_t j = col[t]; perm[cstart[j] + cfill[j]++] = t; }
    int64_t nnz = 0;
    int64_t *slot = malloc(sizeof(int64_t) * ((size_t)nnodes + 2));
    for (int64_t i = 0; i <= nnodes; ++i) slot[i] = -1;
    colptr[0] = 1;
    for (int64_t j = 1; j <= nnodes; ++j) {
        int64_t s = cstart[j], e = cstart[j] + cfill[j], first = nnz;
        for (int64_t t = s; t < e; ++t) {               /* distinct rows of this column */
            int64_t r = row[perm[t]];
            if (slot[r] < 0) { slot[r] = 1; rowval[nnz++] = r; }
        }
        qsort(rowval + first, (size_t)(nnz - first), sizeof(int64_t), cmp_i64);
        for (int64_t k = first; k < nnz; ++k) { slot[rowval[k]] = k; nzval[k] = 0.0; }
        for (int64_t t = s; t < e; ++t) nzval[slot[row[perm[t]]]] += val[perm[t]];   /* input order */
        for (int64_t k = first; k < nnz; ++k) slot[rowval[k]] = -1;
        colptr[j] = nnz + 1;
    }
    free(row); free(col); free(val); free(cstart); free(perm); free(cfill); free(slot);
    return nnz;
}

// fg_mls.cu -- kernel 2: moving-least-squares shape functions and gradients (FP64).
//
// Replaces Improved_Moving_Least_Squares (src/Shape_Functions.jl:88-162).  The reference
// orthogonalises the bilinear/trilinear basis [1,x,y,(z),xy,(xz,yz,xyz)] (:92-112) in absolute
// coordinates with weighted classical Gram-Schmidt (:133-142) so that the moment matrix is
// diagonal.  MLS only depends on the SPAN of the basis, so this kernel evaluates the same
// shape functions in the closed form
//     A = sum_a W_a q_a q_a^T,   A g0 = b(g),   A g_d = d_d b(g) - (sum_a d_dW_a q_a q_a^T) g0,
//     phi_a = W_a q_a.g0,        d_d phi_a = W_a q_a.g_d + d_dW_a q_a.g0
// with the basis written in the local coordinate u = (y - g) * (1/dm): b(g) = e_0 and
// d_d b(g) = e_{1+d}/dm_d, the moment matrix is O(1)-conditioned and a Cholesky factorisation in
// registers replaces Gram-Schmidt.  In exact arithmetic the result equals the reference's; in
// FP64 it sits ~1e-15 from the extended-precision truth where the reference's own unshifted
// Gram-Schmidt is 1e-12 .. 1e-6 away ("parity").
//
// Phase A, one thread per Gauss point: two sweeps over the neighbours (moments; A_,d g0) and
// 1+dim triangular solves, everything in registers.  Phase B, one thread per (point,neighbour)
// ENTRY of the block's contiguous output slice: recompute the weight/basis of that entry, dot
// with the point's g0/g_d (parked in shared memory) and store phi/dphi fully coalesced.
#include "fg_internal.cuh"
#include <algorithm>

struct MlsArgs {
    int64_t nnodes, ngauss;
    const double *nrec, *gs;   // nrec: per node [x_0..x_{dim-1}, 1/dm_0..1/dm_{dim-1}] (AoS: one record = 2*dim doubles)
    const int64_t *off;
    const int *dom;
    double *phi, *dphi;   // dphi component-major: dphi[d * tl + e] (coalesced stores here, coalesced loads in the assembly)
    int64_t tl;           // total_L
    int *singular;   // device counter
};

// cubic-spline weight of cubwgt (src/Shape_Functions.jl:29-86); dif = g - x_a.
// 4/3-4r+4r^2-(4/3)r^3 = (4/3)(1-r)^3 and -4+8r-4r^2 = -4(1-r)^2 (the outer branch, r > 1/2).
__device__ __forceinline__ void spline1d(double dif, double idm, double &w, double &dw) {
    const double r = fabs(dif) * idm;
    const double sg = copysign(idm, dif);       // sign(dif)/dm; at dif = 0 the inner branch gives dw = 0 * (...) = 0 anyway
#if !defined(FG_SPLINE_FACTORED) && !defined(FG_SPLINE_HORNER)
    // Both branches share one shape:  w = c0 + c1 a^2 (1 - r),  dw/dr = e2 a^2 + e1 r  with a = r, (c0, c1, e2, e1) = (2/3, -4, 12, -8) inside
    // (2/3 - 4r^2 + 4r^3 = 2/3 - 4 r^2 (1 - r)) and a = 1 - r, (0, 4/3, -4, 0) outside: 8 FP64 operations instead of 13, five selects.
    // Measured on k_imls_p at the full size: 12.60 -> 12.40 ms, errors against the __float128 truth unchanged (phi 9.5e-16, grad 9.2e-16).
    const double om = 1.0 - r;
    const bool outer = r > 0.5;
    const double a = outer ? om : r;
    const double a2 = a * a;
    w = fma(outer ? 4.0 / 3.0 : -4.0, a2 * om, outer ? 0.0 : 2.0 / 3.0);
    dw = fma(outer ? -4.0 : 12.0, a2, (outer ? 0.0 : -8.0) * r) * sg;
#elif defined(FG_SPLINE_FACTORED)
    // both branches in the reference's factored form, then a select (13 FP64 operations): the round's form until the third session
    const double om = 1.0 - r, om2 = om * om;
    const double w_out = ((4.0 / 3.0) * om) * om2, d_out = -4.0 * om2;
    const double w_in = fma(r * r, fma(4.0, r, -4.0), 2.0 / 3.0), d_in = r * fma(12.0, r, -8.0);
    const bool outer = r > 0.5;
    w = outer ? w_out : w_in;
    dw = (outer ? d_out : d_in) * sg;
#else
    // ONE cubic in r with the branch's coefficients selected: 7 FP64 operations instead of 13, but 12 more selects and a serial
    // Horner chain.  Measured on k_imls_p at the full size: 13.48 ms against 13.38 for the factored form (round 3) -- kept for reference.
    const bool outer = r > 0.5;
    const double c3 = outer ? -4.0 / 3.0 : 4.0, c2 = outer ? 4.0 : -4.0, c1 = outer ? -4.0 : 0.0, c0 = outer ? 4.0 / 3.0 : 2.0 / 3.0;
    const double e2 = outer ? -4.0 : 12.0, e1 = outer ? 8.0 : -8.0;
    w = fma(fma(fma(c3, r, c2), r, c1), r, c0);
    dw = fma(fma(e2, r, e1), r, c1) * sg;
#endif
}

template <int DIM>
__device__ __forceinline__ void load_rec(const MlsArgs &A, int id, double (&rec)[2 * DIM]) {
    const double2 *p = reinterpret_cast<const double2 *>(A.nrec + (size_t)id * (2 * DIM));   // 16-byte aligned records
#pragma unroll
    for (int k = 0; k < DIM; ++k) { const double2 t = __ldg(p + k); rec[2 * k] = t.x; rec[2 * k + 1] = t.y; }
}

template <int DIM, int SHAPE>
__device__ __forceinline__ void eval_rec(const double (&rec)[2 * DIM], const double (&g)[DIM], const double (&nsc)[DIM],
                                         double &w, double (&dw)[DIM], double (&u)[DIM]) {
    double dif[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        dif[d] = g[d] - rec[d];
        u[d] = dif[d] * nsc[d];            // local coordinate (x_a - g) / dm: nsc = -1/dm
    }
    if (SHAPE == FG_SHAPE_RECTANGULAR) {
        double wt[DIM], dwt[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) spline1d(dif[d], rec[DIM + d], wt[d], dwt[d]);
        if (DIM == 2) {
            w = wt[0] * wt[1];
            dw[0] = dwt[0] * wt[1];
            dw[1] = wt[0] * dwt[1];
        } else {
            const double w01 = wt[0] * wt[1], w2 = wt[DIM - 1];
            w = w01 * w2;
            dw[0] = dwt[0] * wt[1] * w2;
            dw[1] = wt[0] * dwt[1] * w2;
            dw[DIM - 1] = w01 * dwt[DIM - 1];
        }
    } else {
        const double idmi = rec[DIM];
        double dist = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) dist += dif[d] * dif[d];
        dist = sqrt(dist);
        const double r = dist * idmi;
        w = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) dw[d] = 0.0;
        if (r <= 1.0) {
            double df;
            if (r <= 0.5) { w = 2.0 / 3.0 + r * r * (4.0 * r - 4.0); df = r * (12.0 * r - 8.0); }
            else { const double om = 1.0 - r; w = (4.0 / 3.0) * om * om * om; df = -4.0 * om * om; }
            if (dist > 1e-12) {
                const double f = df * idmi / dist;
#pragma unroll
                for (int d = 0; d < DIM; ++d) dw[d] = f * dif[d];
            }
        }
    }
}

// The bilinear / trilinear basis q = [1, ux, uy, (uz), ux uy, (ux uz, uy uz, ux uy uz)]: basis function i is the
// monomial with exponent bits basis_bits(i) (bit 0 = ux, bit 1 = uy, bit 2 = uz).
template <int DIM>
__host__ __device__ constexpr int basis_bits(int i) { return DIM == 2 ? i : (i == 3 ? 4 : (i == 4 ? 3 : i)); }

template <int DIM>
__device__ __forceinline__ void basis(const double (&u)[DIM], double (&q)[DIM == 2 ? 4 : 8]) {
    q[0] = 1.0; q[1] = u[0]; q[2] = u[1];
    if (DIM == 2) q[3] = u[0] * u[1];
    else {
        q[3] = u[DIM - 1];
        q[4] = u[0] * u[1]; q[5] = u[0] * u[DIM - 1]; q[6] = u[1] * u[DIM - 1]; q[7] = q[4] * u[DIM - 1];
    }
}

// q . c for a coefficient vector c (stride `st`): the basis is multilinear, so the contraction nests -- 3 FMAs in
// 2D, 7 in 3D, and the products ux uy, ... are never formed.
template <int DIM>
__device__ __forceinline__ double basis_dot(const double (&u)[DIM], const double *c, int st) {
    if (DIM == 2) {
        const double a = fma(c[st], u[0], c[0]), b = fma(c[3 * st], u[0], c[2 * st]);
        return fma(b, u[1], a);
    } else {
        const double a0 = fma(c[st], u[0], c[0]), a1 = fma(c[4 * st], u[0], c[2 * st]);          // 1, ux | uy, ux uy
        const double b0 = fma(c[5 * st], u[0], c[3 * st]), b1 = fma(c[7 * st], u[0], c[6 * st]);  // uz, ux uz | uy uz, ux uy uz
        const double lo = fma(a1, u[1], a0), hi = fma(b1, u[1], b0);
        return fma(hi, u[DIM - 1], lo);
    }
}

// Moments of the basis: q_i q_j = ux^a uy^b uz^c with a, b, c in {0,1,2} -- 9 (2D) / 27 (3D) distinct sums
// instead of the 10 / 36 entries of the symmetric matrix.  mom[(a*3 + b)*3 + c] (2D: c = 0).
template <int DIM>
__device__ __forceinline__ void add_moments(double w, const double (&u)[DIM], double (&mom)[DIM == 2 ? 9 : 27]) {
    const double ta[3] = {w, w * u[0], (w * u[0]) * u[0]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double t1 = ta[a] * u[1];
        const double tb[3] = {ta[a], t1, t1 * u[1]};
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (DIM == 2) mom[a * 3 + b] += tb[b];
            else {
                const double uz = u[DIM - 1];
                mom[(a * 3 + b) * 3] += tb[b];
                mom[(a * 3 + b) * 3 + 1] = fma(tb[b], uz, mom[(a * 3 + b) * 3 + 1]);
                mom[(a * 3 + b) * 3 + 2] = fma(tb[b] * uz, uz, mom[(a * 3 + b) * 3 + 2]);
            }
        }
    }
}

template <int DIM>
__host__ __device__ constexpr int moment_index(int i, int j) {
    const int bi = basis_bits<DIM>(i), bj = basis_bits<DIM>(j);
    const int a = (bi & 1) + (bj & 1), b = ((bi >> 1) & 1) + ((bj >> 1) & 1), c = ((bi >> 2) & 1) + ((bj >> 2) & 1);
    return DIM == 2 ? a * 3 + b : (a * 3 + b) * 3 + c;
}

// packed lower triangle
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// All loops below have compile-time bounds with `if (k < j)`-style guards so that full unrolling
// turns every index into a constant and the matrices live in registers.
template <int M>
__device__ __forceinline__ bool cholesky(double (&A)[M * (M + 1) / 2], double (&inv)[M]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double s = A[TRI(j, j)];
        const double orig = s;
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < j) s -= A[TRI(j, k)] * A[TRI(j, k)];
        if (!(s > 1e-13 * orig)) { ok = false; s = 1.0; }
        const double is = rsqrt(s);
        inv[j] = is;
        A[TRI(j, j)] = s * is;
#pragma unroll
        for (int i = 0; i < M; ++i)
            if (i > j) {
                double t = A[TRI(i, j)];
#pragma unroll
                for (int k = 0; k < M; ++k)
                    if (k < j) t -= A[TRI(i, k)] * A[TRI(j, k)];
                A[TRI(i, j)] = t * is;
            }
    }
    return ok;
}

template <int M>
__device__ __forceinline__ void chol_solve(const double (&Lm)[M * (M + 1) / 2], const double (&inv)[M], double (&b)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double t = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < i) t -= Lm[TRI(i, k)] * b[k];
        b[i] = t * inv[i];
    }
#pragma unroll
    for (int ii = 0; ii < M; ++ii) {
        const int i = M - 1 - ii;
        double t = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k > i) t -= Lm[TRI(k, i)] * b[k];
        b[i] = t * inv[i];
    }
}

#define IMLS_BLOCK 128
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(IMLS_BLOCK, 4) k_imls(MlsArgs A, int maxL) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NG = (1 + DIM) * M;          // g0 and g_d
    constexpr int NP = NG + DIM;               // + scale (-1/dm)
    constexpr int NT = M * (M + 1) / 2;
    constexpr int NROW = NP;                   // rows of the table
    constexpr int NPARK = NT < NP ? NT : NP;   // entries of the Cholesky factor that phase A parks in the table during sweep 2
    extern __shared__ double sm[];
    constexpr int bd = IMLS_BLOCK;              // compile-time: the shared-memory parameter table is addressed with immediates
    const int tid = threadIdx.x;
    double *spar = sm;                         // NP x bd, parameter-major, indexed by the point's slot in the block
    int *soff = (int *)(sm + (size_t)NROW * bd); // bd + 1 entry offsets relative to the block's first entry
    unsigned char *sowner = (unsigned char *)(soff + bd + 4);   // maxL * bd: owner slot of every entry of the block
    // scratch of the counting sort: lives in the (still unused) table, so that four blocks stay within the 164 KB
    // shared-memory configuration and the SM keeps 64 KB of L1 for the node records
    int *sperm = (int *)sm;                    // bd: points of the block ordered by decreasing neighbour count
    int *shist = sperm + bd;                   // 257 buckets
    const int64_t g0i = blockIdx.x * (int64_t)bd;
    const int npts = (int)min((int64_t)bd, A.ngauss - g0i);
    const int64_t base = A.off[g0i];

    // ---- order the block's points by neighbour count, so that the lanes of a warp loop equally long ----------
    for (int t = tid; t < 257; t += bd) shist[t] = 0;
    __syncthreads();
    int myL = -1, mybucket = 0, mypos = 0;
    if (tid < npts) {
        const int64_t o0 = A.off[g0i + tid];
        myL = (int)(A.off[g0i + tid + 1] - o0);
        soff[tid] = (int)(o0 - base);
        mybucket = 255 - min(myL, 255);        // descending
        mypos = atomicAdd(&shist[mybucket], 1);
        for (int a = 0; a < myL; ++a) sowner[(soff[tid]) + a] = (unsigned char)tid;
    }
    if (tid == 0) soff[npts] = (int)(A.off[g0i + npts] - base);
    __syncthreads();
    if (tid < 32) {                            // exclusive scan of 256 buckets by one warp
        int v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = shist[tid * 8 + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (tid >= sft) incl += t; }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { shist[tid * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    if (tid < npts) sperm[shist[mybucket] + mypos] = tid;
    __syncthreads();
    const int my_slot = tid < npts ? sperm[tid] : 0;
    __syncthreads();                           // the sort scratch is dead from here on: the table may be written

    // ---- Phase A: one thread per Gauss point ----------------------------------------------------------------------
    if (tid < npts) {
        const int slot = my_slot;
        const int64_t gi = g0i + slot;
        const int64_t o0 = base + soff[slot];
        const int L = soff[slot + 1] - soff[slot];
        double g[DIM], sc[DIM], nsc[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) g[d] = A.gs[d * A.ngauss + gi];
        double gam[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) gam[k] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { sc[d] = 1.0; nsc[d] = -1.0; }
        if (L > 0) {
            double rec[2 * DIM], nrec[2 * DIM];
            load_rec<DIM>(A, A.dom[o0], rec);
#pragma unroll
            for (int d = 0; d < DIM; ++d) { sc[d] = rec[DIM + (SHAPE == FG_SHAPE_RECTANGULAR ? d : 0)]; nsc[d] = -sc[d]; }
            // sweep 1: moments (the next neighbour's record is in flight while this one is used)
            double mom[DIM == 2 ? 9 : 27];
#pragma unroll
            for (int k = 0; k < (DIM == 2 ? 9 : 27); ++k) mom[k] = 0.0;
            int idn = A.dom[o0 + min(1, L - 1)];
            for (int a = 0; a < L; ++a) {
                load_rec<DIM>(A, idn, nrec);
                idn = A.dom[o0 + min(a + 2, L - 1)];
                double w, dw[DIM], u[DIM];
                eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
                add_moments<DIM>(w, u, mom);
#pragma unroll
                for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
            }
            double Am[M * (M + 1) / 2];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < M; ++j)
                    if (j <= i) Am[TRI(i, j)] = mom[moment_index<DIM>(i, j)];
            double inv[M];
            const bool ok = cholesky<M>(Am, inv);
            double g0v[M];
#pragma unroll
            for (int k = 0; k < M; ++k) g0v[k] = 0.0;
            g0v[0] = 1.0;                                   // b(g) = e_0
            chol_solve<M>(Am, inv, g0v);
            // the factor is not needed during sweep 2: it waits in this point's column of the shared table (which is
            // only filled at the end of phase A) instead of being spilled to local memory -- 8 GB of DRAM writes at 2.7e7 points
#pragma unroll
            for (int k = 0; k < NPARK; ++k) spar[k * bd + slot] = Am[k];
            // sweep 2: v_d = (sum_a d_dW_a q_a q_a^T) g0
            double v[DIM][M];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
#pragma unroll
                for (int k = 0; k < M; ++k) v[d][k] = 0.0;
            load_rec<DIM>(A, A.dom[o0], rec);
            idn = A.dom[o0 + min(1, L - 1)];
            for (int a = 0; a < L; ++a) {
                load_rec<DIM>(A, idn, nrec);
                idn = A.dom[o0 + min(a + 2, L - 1)];
                double w, dw[DIM], u[DIM], q[M];
                eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
                basis<DIM>(u, q);
                const double s = basis_dot<DIM>(u, g0v, 1);
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    const double c = dw[d] * s;
#pragma unroll
                    for (int k = 0; k < M; ++k) v[d][k] += c * q[k];
                }
#pragma unroll
                for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
            }
#pragma unroll
            for (int k = 0; k < M; ++k) gam[k] = g0v[k];
#pragma unroll
            for (int k = 0; k < NPARK; ++k) Am[k] = spar[k * bd + slot];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double rhs[M];
#pragma unroll
                for (int k = 0; k < M; ++k) rhs[k] = -v[d][k];
                rhs[1 + d] += sc[d];                        // d_d b(g) = e_{1+d} * du_d/dy_d
                chol_solve<M>(Am, inv, rhs);
#pragma unroll
                for (int k = 0; k < M; ++k) gam[(1 + d) * M + k] = rhs[k];
            }
            if (!ok) {
                atomicAdd(A.singular, 1);
                const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
                for (int k = 0; k < NG; ++k) gam[k] = nan;
            }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) spar[k * bd + slot] = gam[k];
#pragma unroll
        for (int d = 0; d < DIM; ++d) spar[(NG + d) * bd + slot] = nsc[d];
    }
    __syncthreads();

    // ---- Phase B: one thread per entry of the block's output slice ---------------------------------------------------
    const int total = soff[npts];
    // the record of the NEXT entry of this thread is in flight while the current one is evaluated (id two ahead)
    double rec[2 * DIM], nrec[2 * DIM];
    int idn = 0;
    if (tid < total) load_rec<DIM>(A, A.dom[base + tid], rec);
    if (tid + bd < total) idn = A.dom[base + tid + bd];
#pragma unroll 2
    for (int e = tid; e < total; e += bd) {
        if (e + bd < total) load_rec<DIM>(A, idn, nrec);
        if (e + 2 * bd < total) idn = A.dom[base + e + 2 * bd];
        const int p = sowner[e];
        double g[DIM], sc[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { g[d] = A.gs[d * A.ngauss + g0i + p]; sc[d] = spar[(NG + d) * bd + p]; }   // g: L1-resident, shared by the point's entries
        double w, dw[DIM], u[DIM];
        eval_rec<DIM, SHAPE>(rec, g, sc, w, dw, u);       // sc holds -1/dm here (parked negated by phase A)
        const double s = basis_dot<DIM>(u, spar + p, bd);
        A.phi[base + e] = w * s;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double t = basis_dot<DIM>(u, spar + (size_t)(1 + d) * M * bd + p, bd);
            A.dphi[d * A.tl + base + e] = w * t + dw[d] * s;
        }
#pragma unroll
        for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
    }
}

// --------------------------------------------------------------------------------------------
// Warp-autonomous variant (default).  k_imls_order sorts the Gauss points of every tile of 1024 by neighbour count, so that
// 32 CONSECUTIVE entries of the permutation have (nearly) equal list lengths; k_imls_w gives every warp 32 such points and
// runs both phases on them without ever meeting another warp: no counting sort inside the kernel, no block barrier between
// the phases (11 % + 9 % of the stall samples of k_imls), a 9.7 KB table per warp.  Phase B walks the entries of the warp's
// 32 points: each point's list is a contiguous run of the output arrays, so the stores still leave as full sectors.
// --------------------------------------------------------------------------------------------
#define IMLS_TILE 1024
__global__ void __launch_bounds__(IMLS_TILE) k_imls_order(int64_t ngauss, const int64_t *__restrict__ off, int *__restrict__ perm) {
    __shared__ int hist[264];
    const int tid = threadIdx.x;
    const int64_t base = blockIdx.x * (int64_t)IMLS_TILE, g = base + tid;
    for (int t = tid; t < 264; t += IMLS_TILE) hist[t] = 0;
    __syncthreads();
    int bucket = 0, pos = 0;
    const bool live = g < ngauss;
    if (live) {
        const int L = (int)(off[g + 1] - off[g]);
        bucket = 255 - min(L, 255);                 // longest lists first
        pos = atomicAdd(&hist[bucket], 1);
    }
    __syncthreads();
    if (tid < 32) {                                  // exclusive scan of 256 buckets by one warp
        int v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = hist[tid * 8 + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (tid >= sft) incl += t; }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hist[tid * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    if (live) perm[base + hist[bucket] + pos] = (int)g;
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(128, 4) k_imls_w(MlsArgs A, const int *__restrict__ perm) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NG = (1 + DIM) * M;          // g0 and g_d
    constexpr int NT = M * (M + 1) / 2;
    constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;      // gamma rows, -1/dm, the point itself; also parks the Cholesky factor
    constexpr int WBYTES = ROWS * 32 * 8 + 32 * 8 + 36 * 4;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *mine = smraw + (size_t)wid * WBYTES;
    double *spar = reinterpret_cast<double *>(mine);                      // ROWS x 32, parameter-major: spar[k * 32 + slot]
    long long *sgo = reinterpret_cast<long long *>(mine + ROWS * 32 * 8); // first entry of every slot's list in the output arrays
    int *soff = reinterpret_cast<int *>(mine + ROWS * 32 * 8 + 32 * 8);   // 33 prefix sums of the list lengths
    const int64_t w0 = (blockIdx.x * (int64_t)(blockDim.x >> 5) + wid) * 32;
    if (w0 >= A.ngauss) return;
    const int npts = (int)min((int64_t)32, A.ngauss - w0);

    // ---- Phase A: lane = Gauss point ---------------------------------------------------------------------------------
    int L = 0; int64_t o0 = 0; int64_t gi = 0;
    if (lane < npts) {
        gi = perm[w0 + lane];
        o0 = A.off[gi];
        L = (int)(A.off[gi + 1] - o0);
    }
    {
        int incl = L;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
        soff[lane + 1] = incl;
        if (lane == 0) soff[0] = 0;
        sgo[lane] = o0;
    }
    double g[DIM], sc[DIM], nsc[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { g[d] = lane < npts ? A.gs[d * A.ngauss + gi] : 0.0; sc[d] = 1.0; nsc[d] = -1.0; }
    double gam[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) gam[k] = 0.0;
    if (L > 0) {
        double rec[2 * DIM], nrec[2 * DIM];
        load_rec<DIM>(A, A.dom[o0], rec);
#pragma unroll
        for (int d = 0; d < DIM; ++d) { sc[d] = rec[DIM + (SHAPE == FG_SHAPE_RECTANGULAR ? d : 0)]; nsc[d] = -sc[d]; }
        // sweep 1: moments (the next neighbour's record is in flight while this one is used)
        double mom[DIM == 2 ? 9 : 27];
#pragma unroll
        for (int k = 0; k < (DIM == 2 ? 9 : 27); ++k) mom[k] = 0.0;
        int idn = A.dom[o0 + min(1, L - 1)];
        for (int a = 0; a < L; ++a) {
            load_rec<DIM>(A, idn, nrec);
            idn = A.dom[o0 + min(a + 2, L - 1)];
            double w, dw[DIM], u[DIM];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            add_moments<DIM>(w, u, mom);
#pragma unroll
            for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
        }
        double Am[NT];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j)
                if (j <= i) Am[TRI(i, j)] = mom[moment_index<DIM>(i, j)];
        double inv[M];
        const bool ok = cholesky<M>(Am, inv);
        double g0v[M];
#pragma unroll
        for (int k = 0; k < M; ++k) g0v[k] = 0.0;
        g0v[0] = 1.0;                                   // b(g) = e_0
        chol_solve<M>(Am, inv, g0v);
        // the factor waits in this point's column of the warp's table during sweep 2 (registers are the scarce resource)
#pragma unroll
        for (int k = 0; k < NT; ++k) spar[k * 32 + lane] = Am[k];
        // sweep 2: v_d = (sum_a d_dW_a q_a q_a^T) g0
        double v[DIM][M];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < M; ++k) v[d][k] = 0.0;
        load_rec<DIM>(A, A.dom[o0], rec);
        idn = A.dom[o0 + min(1, L - 1)];
        for (int a = 0; a < L; ++a) {
            load_rec<DIM>(A, idn, nrec);
            idn = A.dom[o0 + min(a + 2, L - 1)];
            double w, dw[DIM], u[DIM], q[M];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            basis<DIM>(u, q);
            const double s = basis_dot<DIM>(u, g0v, 1);
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double c = dw[d] * s;
#pragma unroll
                for (int k = 0; k < M; ++k) v[d][k] += c * q[k];
            }
#pragma unroll
            for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
        }
#pragma unroll
        for (int k = 0; k < M; ++k) gam[k] = g0v[k];
#pragma unroll
        for (int k = 0; k < NT; ++k) Am[k] = spar[k * 32 + lane];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double rhs[M];
#pragma unroll
            for (int k = 0; k < M; ++k) rhs[k] = -v[d][k];
            rhs[1 + d] += sc[d];                        // d_d b(g) = e_{1+d} * du_d/dy_d
            chol_solve<M>(Am, inv, rhs);
#pragma unroll
            for (int k = 0; k < M; ++k) gam[(1 + d) * M + k] = rhs[k];
        }
        if (!ok) {
            atomicAdd(A.singular, 1);
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
            for (int k = 0; k < NG; ++k) gam[k] = nan;
        }
    }
#pragma unroll
    for (int k = 0; k < NG; ++k) spar[k * 32 + lane] = gam[k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { spar[(NG + d) * 32 + lane] = nsc[d]; spar[(NG + DIM + d) * 32 + lane] = g[d]; }
    __syncwarp();

    // ---- Phase B: lane = entry of the warp's 32 lists ------------------------------------------------------------------
    const int total = soff[npts];
    int p = 0;                                             // slot of my current entry (only moves forward)
    double rec[2 * DIM], nrec[2 * DIM];
    int pn = 0; int64_t gidx = 0, gidx_n = 0;
    auto locate = [&](int e, int &slot, int64_t &idx) {    // entry e of the warp -> (slot, index in the output arrays)
        while (e >= soff[slot + 1]) ++slot;
        idx = sgo[slot] + (e - soff[slot]);
    };
    if (lane < total) { locate(lane, p, gidx); load_rec<DIM>(A, __ldcs(A.dom + gidx), rec); }
    pn = p;
    for (int e = lane; e < total; e += 32) {
        if (e + 32 < total) { locate(e + 32, pn, gidx_n); load_rec<DIM>(A, __ldcs(A.dom + gidx_n), nrec); }
        double gp[DIM], scp[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { scp[d] = spar[(NG + d) * 32 + p]; gp[d] = spar[(NG + DIM + d) * 32 + p]; }
        double w, dw[DIM], u[DIM];
        eval_rec<DIM, SHAPE>(rec, gp, scp, w, dw, u);       // scp holds -1/dm (parked negated by phase A)
        const double s = basis_dot<DIM>(u, spar + p, 32);
        __stcs(A.phi + gidx, w * s);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double t = basis_dot<DIM>(u, spar + (size_t)(1 + d) * M * 32 + p, 32);
            __stcs(A.dphi + d * A.tl + gidx, w * t + dw[d] * s);
        }
#pragma unroll
        for (int k = 0; k < 2 * DIM; ++k) rec[k] = nrec[k];
        p = pn; gidx = gidx_n;
    }
}

// --------------------------------------------------------------------------------------------
// Cluster-centric variant (default after a cluster search).  ncu on k_imls / k_imls_w: the LSU data pipe is 91 % busy -- every
// sweep iteration and every output entry gathers a 48-byte node record through a DOM id, 16 different cache lines per warp
// instruction -- and the FP64 pipe idles at 50 %.  The neighbour search already knows better: all Gauss points of one search
// cluster draw their neighbours from ONE id-sorted candidate list (<= 128 nodes) and each point keeps a hit mask over it.
// One block per search cluster: the candidates' records are parked in shared memory once (3-6 KB), a point walks the set bits
// of its mask instead of reading DOM, and every record read is a shared-memory read.  No global gathers are left.
//   block = 256 threads; the cluster's points are counting-sorted by neighbour count, each warp then owns 32 of them and runs
//   phase A (lane = point) and phase B (lane = entry of those 32 lists, candidate index from a byte list written by sweep 2)
//   on its own table without meeting the other warps.
// --------------------------------------------------------------------------------------------
struct ClusterView {
    const int *cstart, *csorted;     // search cluster -> Gauss points
    const int *cperm;                // the same points, every cluster's slice ordered by decreasing neighbour count
    const int *cand;                 // per cluster: [0] n, [1..capc] id-sorted candidate nodes
    const unsigned *masks;           // per Gauss point: capc/32 words
    int capc;
};

// Every search cluster's points ordered by neighbour count (longest lists first), and the largest cluster / candidate list.
__global__ void __launch_bounds__(128) k_cluster_order(int64_t c_first, const int *__restrict__ cstart, const int *__restrict__ csorted,
                                                       const int64_t *__restrict__ off, int *__restrict__ cperm) {
    __shared__ int hist[264];
    const int64_t c = c_first + blockIdx.x;
    const int p0 = cstart[c], p1 = cstart[c + 1], tid = threadIdx.x;
    if (p0 == p1) return;
    for (int t = tid; t < 264; t += 128) hist[t] = 0;
    __syncthreads();
    for (int p = p0 + tid; p < p1; p += 128) {
        const int g = csorted[p];
        atomicAdd(&hist[255 - min((int)(off[g + 1] - off[g]), 255)], 1);
    }
    __syncthreads();
    if (tid < 32) {
        int v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = hist[tid * 8 + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (tid >= sft) incl += t; }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hist[tid * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    for (int p = p0 + tid; p < p1; p += 128) {
        const int g = csorted[p];
        cperm[p0 + atomicAdd(&hist[255 - min((int)(off[g + 1] - off[g]), 255)], 1)] = g;
    }
}

// One WARP per block = 32 points of one cluster with (nearly) equal list lengths.  Single-warp blocks keep every warp slot of the
// SM busy: inside a multi-warp block the warps with short lists would idle until the 27-neighbour warp is done (ncu on the first,
// 256-thread version: 9.9 of 16 resident warps active).  The block parks its cluster's candidate records (3 KB for 64 nodes) itself.
#ifndef IMLS_C_BLOCKS
#define IMLS_C_BLOCKS 15
#endif
#ifndef IMLS_C_UNROLL
#define IMLS_C_UNROLL 1
#endif
constexpr int kImlsUnroll = IMLS_C_UNROLL;
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(32, IMLS_C_BLOCKS) k_imls_c(MlsArgs A, ClusterView V, int lstride, int G, int nrec_cap, int *__restrict__ dom_out) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NG = (1 + DIM) * M;
    constexpr int NT = M * (M + 1) / 2;
    constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x;
    const int capc = V.capc;
    double *srec = reinterpret_cast<double *>(smraw);                    // nrec_cap records of 2*DIM doubles
    unsigned char *mine = smraw + (size_t)nrec_cap * 2 * DIM * 8;
    double *spar = reinterpret_cast<double *>(mine);                     // ROWS x 32, parameter-major
    long long *sgo = reinterpret_cast<long long *>(mine + ROWS * 32 * 8);
    int *soff = reinterpret_cast<int *>(mine + ROWS * 32 * 8 + 32 * 8);
    unsigned char *slist = mine + ROWS * 32 * 8 + 32 * 8 + 36 * 4;       // 32 lists of candidate indices, lstride bytes each
    int *sid = reinterpret_cast<int *>(slist + 32 * lstride);            // candidate node ids (DOM is written here when the search deferred it)

    const int64_t c = blockIdx.x / G;
    const int j = (int)(blockIdx.x - c * G);
    const int p0 = V.cstart[c], p1 = V.cstart[c + 1];
    const int first = p0 + 32 * j;
    if (first >= p1) return;
    const int wn = min(32, p1 - first);
    const int *cand = V.cand + c * (int64_t)(1 + 2 * capc);
    // ---- first-level loads together (candidate count and ids, my point), then everything that depends on them --------------
    const int nc_raw = cand[0];
    int myid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) myid[k] = lane + 32 * k < nrec_cap ? cand[1 + lane + 32 * k] : 0;      // slots beyond the count hold stale ids: masked below
    int L = 0; int64_t o0 = 0; int64_t gi = 0;
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (lane < wn) gi = V.cperm[first + lane];
    const int nc = min(nc_raw, nrec_cap);
    if (lane < wn) {
        o0 = A.off[gi];
        L = (int)(A.off[gi + 1] - o0);
        const uint4 mk = __ldg(reinterpret_cast<const uint4 *>(V.masks + gi * 4));
        m0 = mk.x; m1 = mk.y; m2 = mk.z; m3 = mk.w;
    }
    // the cluster's candidate records -> shared memory
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = lane + 32 * k;
        if (e < nc) {
            sid[e] = myid[k];
            const double2 *src = reinterpret_cast<const double2 *>(A.nrec + (size_t)myid[k] * (2 * DIM));
            double2 *dst = reinterpret_cast<double2 *>(srec + (size_t)e * (2 * DIM));
#pragma unroll
            for (int kk = 0; kk < DIM; ++kk) dst[kk] = __ldg(src + kk);
        }
    }
    {
        int incl = L;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
        soff[lane + 1] = incl;
        if (lane == 0) soff[0] = 0;
        sgo[lane] = o0;
    }
    __syncwarp();
    // walk the set bits of the mask, lowest first (= ascending node id = the order of DOM)
    struct Bits {
        unsigned w[4]; unsigned cur; int k;
        __device__ __forceinline__ void reset(unsigned a, unsigned b, unsigned c2, unsigned d) { w[0] = a; w[1] = b; w[2] = c2; w[3] = d; cur = a; k = 0; }
        __device__ __forceinline__ int next() {
            while (cur == 0u && k < 3) { ++k; cur = k == 1 ? w[1] : (k == 2 ? w[2] : w[3]); }
            const int b = __ffs(cur) - 1;
            cur &= cur - 1;
            return 32 * k + b;
        }
    } bits;
    auto smem_rec = [&](int b, double (&rec)[2 * DIM]) {
        const double2 *p2 = reinterpret_cast<const double2 *>(srec + (size_t)b * (2 * DIM));
#pragma unroll
        for (int k = 0; k < DIM; ++k) { const double2 t = p2[k]; rec[2 * k] = t.x; rec[2 * k + 1] = t.y; }
    };
    double g[DIM], sc[DIM], nsc[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { g[d] = lane < wn ? A.gs[d * A.ngauss + gi] : 0.0; sc[d] = 1.0; nsc[d] = -1.0; }
    double gam[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) gam[k] = 0.0;
    if (L > 0) {
        double rec[2 * DIM];
        bits.reset(m0, m1, m2, m3);
        {
            Bits first_bit = bits;
            smem_rec(first_bit.next(), rec);
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) { sc[d] = rec[DIM + (SHAPE == FG_SHAPE_RECTANGULAR ? d : 0)]; nsc[d] = -sc[d]; }
        double mom[DIM == 2 ? 9 : 27];
#pragma unroll
        for (int k = 0; k < (DIM == 2 ? 9 : 27); ++k) mom[k] = 0.0;
#pragma unroll (kImlsUnroll)
        for (int a = 0; a < L; ++a) {                     // sweep 1: moments
            smem_rec(bits.next(), rec);
            double w, dw[DIM], u[DIM];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            add_moments<DIM>(w, u, mom);
        }
        double Am[NT];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int jj = 0; jj < M; ++jj)
                if (jj <= i) Am[TRI(i, jj)] = mom[moment_index<DIM>(i, jj)];
        double inv[M];
        const bool ok = cholesky<M>(Am, inv);
        double g0v[M];
#pragma unroll
        for (int k = 0; k < M; ++k) g0v[k] = 0.0;
        g0v[0] = 1.0;
        chol_solve<M>(Am, inv, g0v);
#pragma unroll
        for (int k = 0; k < NT; ++k) spar[k * 32 + lane] = Am[k];       // the factor waits in the table during sweep 2
        double v[DIM][M];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < M; ++k) v[d][k] = 0.0;
        bits.reset(m0, m1, m2, m3);
        unsigned char *mylist = slist + lane * lstride;
#pragma unroll (kImlsUnroll)
        for (int a = 0; a < L; ++a) {                     // sweep 2: v_d = (sum_a d_dW_a q_a q_a^T) g0; candidate indices for phase B
            const int b = bits.next();
            mylist[a] = (unsigned char)b;
            smem_rec(b, rec);
            double w, dw[DIM], u[DIM], q[M];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            basis<DIM>(u, q);
            const double s = basis_dot<DIM>(u, g0v, 1);
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double cc = dw[d] * s;
#pragma unroll
                for (int k = 0; k < M; ++k) v[d][k] += cc * q[k];
            }
        }
#pragma unroll
        for (int k = 0; k < M; ++k) gam[k] = g0v[k];
#pragma unroll
        for (int k = 0; k < NT; ++k) Am[k] = spar[k * 32 + lane];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double rhs[M];
#pragma unroll
            for (int k = 0; k < M; ++k) rhs[k] = -v[d][k];
            rhs[1 + d] += sc[d];
            chol_solve<M>(Am, inv, rhs);
#pragma unroll
            for (int k = 0; k < M; ++k) gam[(1 + d) * M + k] = rhs[k];
        }
        if (!ok) {
            atomicAdd(A.singular, 1);
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
            for (int k = 0; k < NG; ++k) gam[k] = nan;
        }
    }
#pragma unroll
    for (int k = 0; k < NG; ++k) spar[k * 32 + lane] = gam[k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { spar[(NG + d) * 32 + lane] = nsc[d]; spar[(NG + DIM + d) * 32 + lane] = g[d]; }
    __syncwarp();

    // ---- Phase B: lane = entry of the warp's 32 lists --------------------------------------------------------------
    const int total = soff[wn];
    int p = 0;
    for (int e = lane; e < total; e += 32) {
        while (e >= soff[p + 1]) ++p;
        const int a = e - soff[p];
        const int64_t gidx = sgo[p] + a;
        double rec[2 * DIM];
        const int li = slist[p * lstride + a];
        if (dom_out) __stcs(dom_out + gidx, sid[li]);          // the deferred fill pass of the search: same bits, same order
        smem_rec(li, rec);
        double gp[DIM], scp[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { scp[d] = spar[(NG + d) * 32 + p]; gp[d] = spar[(NG + DIM + d) * 32 + p]; }
        double w, dw[DIM], u[DIM];
        eval_rec<DIM, SHAPE>(rec, gp, scp, w, dw, u);
        const double s = basis_dot<DIM>(u, spar + p, 32);
        __stcs(A.phi + gidx, w * s);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double t = basis_dot<DIM>(u, spar + (size_t)(1 + d) * M * 32 + p, 32);
            __stcs(A.dphi + d * A.tl + gidx, w * t + dw[d] * s);
        }
    }
}

// --------------------------------------------------------------------------------------------
// Software-pipelined cluster kernel (default).  The source-level samples of k_imls_c (profiles/r02_ncu_kernels.md, capture d) put
// 31 % of the warp time at the HEAD of the three loops: walking the hit mask (a divergent loop: FLO, the word switch), turning the bit
// into a shared-memory address and waiting for the record (short scoreboard), with nothing else of the same warp in flight -- and
// phase B chased four dependent shared-memory loads (slot search in the prefix sums, list position, candidate index, record) per entry.
// Here the masks are walked ONCE, into an entry-indexed table of 16-bit descriptors (owner slot << 8 | candidate index), and every
// loop fetches the record of the NEXT neighbour / entry before it does the arithmetic of the current one (two record buffers,
// loops unrolled by two so that the buffers alternate without copies).  The inverse pivots ride in the diagonal slots of the parked
// Cholesky factor (the solves never read the diagonal), which frees 16 registers during sweep 2.
// --------------------------------------------------------------------------------------------
#ifndef IMLS_P_BLOCKS
#define IMLS_P_BLOCKS 12
#endif
template <int DIM, int SHAPE>
#ifdef IMLS_P_MAXREG
__global__ void __maxnreg__(IMLS_P_MAXREG) k_imls_p(
#else
__global__ void __launch_bounds__(32, IMLS_P_BLOCKS) k_imls_p(
#endif
    MlsArgs A, ClusterView V, int ecap, int G, int nrec_cap, int *__restrict__ dom_out, int64_t c_first, int nwork, int *__restrict__ wq) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NG = (1 + DIM) * M;
    constexpr int NT = M * (M + 1) / 2;
    constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x;
    const int capc = V.capc;
    double *srec = reinterpret_cast<double *>(smraw);                    // nrec_cap records of 2*DIM doubles
    unsigned char *mine = smraw + (size_t)nrec_cap * 2 * DIM * 8;
    double *spar = reinterpret_cast<double *>(mine);                     // ROWS x 32, parameter-major
    long long *sgo = reinterpret_cast<long long *>(mine + ROWS * 32 * 8); // per slot: first entry in the output arrays minus first entry in the warp
    int *sid = reinterpret_cast<int *>(mine + ROWS * 32 * 8 + 32 * 8);   // candidate node ids (DOM is written from here when the search deferred it)
    unsigned short *sent = reinterpret_cast<unsigned short *>(mine + ROWS * 32 * 8 + 32 * 8 + sizeof(int) * nrec_cap);   // ecap entry descriptors

    // Work item = (search cluster, group of 32 points).  The grid may be smaller than the number of items (persistent blocks, option
    // "imls_grid"): a block then strides over the items, and the FIRST level of every item's dependent load chain (cluster extent ->
    // candidate ids / point id) is requested one item ahead, during the arithmetic of the current one.
    struct Item { int p0, p1, nc_raw, gi; int id[4]; };
    auto item_l1 = [&](int w, Item &I) {
        I.p0 = I.p1 = 0; I.nc_raw = 0; I.gi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) I.id[k] = 0;
        if (w >= nwork) return;
        const int64_t cl = w / G;
        const int j = (int)(w - cl * G);
        const int64_t c = c_first + cl;                       // search clusters [c_first, c_first + nwork / G) (ranged launches: fg_impl_imls_range)
        const int q0 = V.cstart[c], q1 = V.cstart[c + 1];
        I.p0 = q0 + 32 * j; I.p1 = q1;
        if (I.p0 >= q1) return;
        const int *cand = V.cand + c * (int64_t)(1 + 2 * capc);
        I.nc_raw = cand[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) I.id[k] = lane + 32 * k < nrec_cap ? cand[1 + lane + 32 * k] : 0;
        if (I.p0 + lane < q1) I.gi = V.cperm[I.p0 + lane];
    };
    // items beyond the first come from a device counter when the grid is persistent (wq != nullptr: dynamic balance; the fetch for the
    // item after next is issued at the top of the body and consumed an iteration later), else from the block stride
    auto next_work = [&](int after) -> int {
        if (wq == nullptr) return after + (int)gridDim.x;
        int v = 0;
        if (lane == 0) v = atomicAdd(wq, 1);
        return (int)gridDim.x + __shfl_sync(0xffffffffu, v, 0);
    };
    Item cur, nxt;
    int w_cur = blockIdx.x, w_nxt = next_work(w_cur), w_nn;
    item_l1(w_cur, cur);
    for (; w_cur < nwork; cur = nxt, w_cur = w_nxt, w_nxt = w_nn) {
    item_l1(w_nxt, nxt);
    w_nn = next_work(w_nxt);
    const int first = cur.p0, p1 = cur.p1;
    if (first >= p1) continue;
    __syncwarp();                                             // the previous item's phase B has read the tables this one overwrites
    const int wn = min(32, p1 - first);
    const int nc_raw = cur.nc_raw;
    int myid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) myid[k] = cur.id[k];
    int L = 0; int64_t o0 = 0; int64_t gi = 0;
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (lane < wn) gi = cur.gi;
    const int nc = min(nc_raw, nrec_cap);
    if (lane < wn) {
        o0 = A.off[gi];
        L = (int)(A.off[gi + 1] - o0);
        const uint4 mk = __ldg(reinterpret_cast<const uint4 *>(V.masks + gi * 4));
        m0 = mk.x; m1 = mk.y; m2 = mk.z; m3 = mk.w;
    }
    double g[DIM], sc[DIM], nsc[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { g[d] = lane < wn ? A.gs[d * A.ngauss + gi] : 0.0; sc[d] = 1.0; nsc[d] = -1.0; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = lane + 32 * k;
        if (e < nc) {
            sid[e] = myid[k];
            const double2 *src = reinterpret_cast<const double2 *>(A.nrec + (size_t)myid[k] * (2 * DIM));
            double2 *dst = reinterpret_cast<double2 *>(srec + (size_t)e * (2 * DIM));
#pragma unroll
            for (int kk = 0; kk < DIM; ++kk) dst[kk] = __ldg(src + kk);
        }
    }
    int incl = L;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
    const int e0 = incl - L;                                  // my first entry among the warp's entries
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    sgo[lane] = o0 - e0;
    // The hit mask is walked inside sweep 1, one neighbour AHEAD of the arithmetic (lowest bit first = ascending node id = the order
    // of DOM): the bit scan, the descriptor store and the record fetch of neighbour a+1 hide under the moments of neighbour a.  Two
    // 64-bit words, so the walk needs a single predicated word switch and no loop.
    unsigned long long cur = ((unsigned long long)m1 << 32) | m0;
    const unsigned long long mhi = ((unsigned long long)m3 << 32) | m2;
    int bbase = 0;
    auto next_idx = [&]() -> int {
        if (cur == 0ull) { cur = mhi; bbase = 64; }
        const int b = bbase + __ffsll((long long)cur) - 1;
        cur &= cur - 1ull;
        return b;
    };
    __syncwarp();
    auto smem_rec = [&](int b, double (&rec)[2 * DIM]) {
        const double2 *p2 = reinterpret_cast<const double2 *>(srec + (size_t)b * (2 * DIM));
#pragma unroll
        for (int k = 0; k < DIM; ++k) { const double2 t = p2[k]; rec[2 * k] = t.x; rec[2 * k + 1] = t.y; }
    };
    auto fetch = [&](int e, double (&rec)[2 * DIM]) { smem_rec(sent[e] & 0xff, rec); };
    double gam[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) gam[k] = 0.0;
    if (L > 0) {
        double ra[2 * DIM], rb[2 * DIM];
        const unsigned short tag = (unsigned short)(lane << 8);
        int ia = next_idx();
        sent[e0] = tag | (unsigned short)ia;
        smem_rec(ia, ra);
#pragma unroll
        for (int d = 0; d < DIM; ++d) { sc[d] = ra[DIM + (SHAPE == FG_SHAPE_RECTANGULAR ? d : 0)]; nsc[d] = -sc[d]; }
        double mom[DIM == 2 ? 9 : 27];
#pragma unroll
        for (int k = 0; k < (DIM == 2 ? 9 : 27); ++k) mom[k] = 0.0;
        auto proc1 = [&](const double (&rec)[2 * DIM]) {
            double w, dw[DIM], u[DIM];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            add_moments<DIM>(w, u, mom);
        };
        const int elast = e0 + L - 1;
        {   // sweep 1: moments
            int a = 0;
#pragma unroll 1
            for (; a + 2 <= L; a += 2) {
                const int ib = next_idx();
                sent[e0 + a + 1] = tag | (unsigned short)ib;
                smem_rec(ib, rb);
                proc1(ra);
                if (a + 2 < L) { ia = next_idx(); sent[e0 + a + 2] = tag | (unsigned short)ia; }   // else: the last record again, discarded
                smem_rec(ia, ra);
                proc1(rb);
            }
            if (a < L) proc1(ra);
        }
        double Am[NT];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int jj = 0; jj < M; ++jj)
                if (jj <= i) Am[TRI(i, jj)] = mom[moment_index<DIM>(i, jj)];
        double inv[M];
        const bool ok = cholesky<M>(Am, inv);
        double g0v[M];
#pragma unroll
        for (int k = 0; k < M; ++k) g0v[k] = 0.0;
        g0v[0] = 1.0;
        chol_solve<M>(Am, inv, g0v);
        fetch(e0, ra);
        // the factor waits in the table during sweep 2; its diagonal slots carry the inverse pivots (the solves read nothing else of the diagonal)
#pragma unroll
        for (int i = 0; i < M; ++i) Am[TRI(i, i)] = inv[i];
#pragma unroll
        for (int k = 0; k < NT; ++k) spar[k * 32 + lane] = Am[k];
        double v[DIM][M];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < M; ++k) v[d][k] = 0.0;
        auto proc2 = [&](const double (&rec)[2 * DIM]) {
            double w, dw[DIM], u[DIM], q[M];
            eval_rec<DIM, SHAPE>(rec, g, nsc, w, dw, u);
            basis<DIM>(u, q);
            const double s = basis_dot<DIM>(u, g0v, 1);
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double cc = dw[d] * s;
#pragma unroll
                for (int k = 0; k < M; ++k) v[d][k] += cc * q[k];
            }
        };
        {   // sweep 2: v_d = (sum_a d_dW_a q_a q_a^T) g0
            int a = 0;
#pragma unroll 1
            for (; a + 2 <= L; a += 2) {
                fetch(e0 + a + 1, rb);
                proc2(ra);
                fetch(min(e0 + a + 2, elast), ra);
                proc2(rb);
            }
            if (a < L) proc2(ra);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) gam[k] = g0v[k];
#pragma unroll
        for (int k = 0; k < NT; ++k) Am[k] = spar[k * 32 + lane];
#pragma unroll
        for (int i = 0; i < M; ++i) inv[i] = Am[TRI(i, i)];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double rhs[M];
#pragma unroll
            for (int k = 0; k < M; ++k) rhs[k] = -v[d][k];
            rhs[1 + d] += sc[d];
            chol_solve<M>(Am, inv, rhs);
#pragma unroll
            for (int k = 0; k < M; ++k) gam[(1 + d) * M + k] = rhs[k];
        }
        if (!ok) {
            atomicAdd(A.singular, 1);
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
            for (int k = 0; k < NG; ++k) gam[k] = nan;
        }
    }
#pragma unroll
    for (int k = 0; k < NG; ++k) spar[k * 32 + lane] = gam[k];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { spar[(NG + d) * 32 + lane] = nsc[d]; spar[(NG + DIM + d) * 32 + lane] = g[d]; }
    __syncwarp();

    // ---- Phase B: lane = entry of the warp's 32 lists; the next entry's descriptor and record are in flight during the arithmetic ----
    auto procB = [&](int e, unsigned ent, const double (&rec)[2 * DIM]) {
        const int p = (int)(ent >> 8), li = (int)(ent & 0xff);
        const int64_t gidx = sgo[p] + e;
        if (dom_out) __stcs(dom_out + gidx, sid[li]);          // the deferred fill pass of the search: same bits, same order
        double gp[DIM], scp[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { scp[d] = spar[(NG + d) * 32 + p]; gp[d] = spar[(NG + DIM + d) * 32 + p]; }
        double w, dw[DIM], u[DIM];
        eval_rec<DIM, SHAPE>(rec, gp, scp, w, dw, u);
        const double s = basis_dot<DIM>(u, spar + p, 32);
        __stcs(A.phi + gidx, w * s);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const double t = basis_dot<DIM>(u, spar + (size_t)(1 + d) * M * 32 + p, 32);
            __stcs(A.dphi + d * A.tl + gidx, w * t + dw[d] * s);
        }
    };
    if (lane < total) {
        const int last = total - 1;                                   // clamped prefetches re-read a valid entry and are discarded
        double ra[2 * DIM], rb[2 * DIM];
        unsigned ea = sent[lane], eb;
        smem_rec(ea & 0xff, ra);
        int e = lane;
#pragma unroll 1
        for (; e + 32 < total; e += 64) {
            eb = sent[e + 32];
            smem_rec(eb & 0xff, rb);
            procB(e, ea, ra);
            ea = sent[min(e + 64, last)];
            smem_rec(ea & 0xff, ra);
            procB(e + 32, eb, rb);
        }
        if (e < total) procB(e, ea, ra);
    }
    }   // work items
}

// The default cluster path (k_imls_p) works on any contiguous range of SEARCH clusters: the whole set in one launch (fg_shape_functions),
// or range by range when the shape functions were deferred (option "lazy_shapes") and the streamed assembly asks for the clusters its
// next range needs (fg_assemble_bilinear_async: shape functions, assembly and download of finished columns as one pipeline).
bool fg_imls_rangeable(const fg_ctx *ctx, const GaussSet &S) {
    if (!(S.nb_clustered && S.nb_capc == 128 && S.max_L <= 96 && S.max_L > 0 && fg_opt(ctx, "imls_cluster", 1) != 0 && fg_opt(ctx, "imls_pipe", 1) != 0)) return false;
    const int G = (S.nb_max_points + 31) / 32;
    return G > 0 && (int64_t)S.nbcl.n * G < ((int64_t)1 << 31);
}

template <int DIM, int SHAPE>
static int run_range(fg_ctx *ctx, GaussSet &S, const MlsArgs &A, int64_t c_lo, int64_t c_hi) {
    constexpr int M = DIM == 2 ? 4 : 8;
    constexpr int NG = (1 + DIM) * M, NT = M * (M + 1) / 2;
    constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;
    if (c_hi <= c_lo) return FG_OK;
    // every cluster's points ordered by list length (the largest cluster and candidate list, known from the search's count
    // pass, size the grid and the record buffer: no host synchronisation in this call)
    FG_CUDA(ctx, S.imls_perm.reserve(S.n));
    k_cluster_order<<<(unsigned)(c_hi - c_lo), 128, 0, ctx->stream>>>(c_lo, S.nbcl.grid.start.p, S.nbcl.sorted.p, S.off.p, S.imls_perm.p);
    FG_CHECK_LAUNCH(ctx);
    const int G = (S.nb_max_points + 31) / 32, nrec_cap = std::max(8, (S.nb_max_cand + 7) & ~7);
    const int ecap = (32 * S.max_L + 7) & ~7;
    const size_t smem = (size_t)nrec_cap * 2 * DIM * 8 + (size_t)ROWS * 32 * 8 + 32 * 8 + sizeof(int) * nrec_cap + sizeof(unsigned short) * ecap;
    ClusterView V;
    V.cstart = S.nbcl.grid.start.p; V.csorted = S.nbcl.sorted.p; V.cperm = S.imls_perm.p; V.cand = S.cand.p; V.masks = S.masks.p; V.capc = S.nb_capc;
    FG_CUDA(ctx, cudaFuncSetAttribute(k_imls_p<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nwork = (c_hi - c_lo) * G;
    int64_t grid = nwork;
    // "imls_grid" = w > 0: w blocks per resident slot, block-strided items; -1 (default): one block per slot, items from a device counter
    const int waves = fg_opt(ctx, "imls_grid", -1);
    int *wq = nullptr;
    if (waves != 0) {
        int per_sm = 0;
        FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_imls_p<DIM, SHAPE>, 32, smem));
        grid = std::min<int64_t>(nwork, (int64_t)ctx->sm_count * std::max(per_sm, 1) * std::max(waves, 1));
        if (waves < 0) {
            FG_CUDA(ctx, ctx->flags.reserve(64));
            wq = ctx->flags.p + 60;
            FG_CUDA(ctx, cudaMemsetAsync(wq, 0, sizeof(int), ctx->stream));
        }
    }
    k_imls_p<DIM, SHAPE><<<(unsigned)grid, 32, smem, ctx->stream>>>(A, V, ecap, G, nrec_cap, S.dom_pending ? S.dom.p : nullptr, c_lo, (int)nwork, wq);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

static MlsArgs mls_args(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    MlsArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.nrec = N.nrec.p; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p;
    A.phi = S.phi.p; A.dphi = S.dphi.p; A.tl = S.total_L; A.singular = S.singular_dev.p;
    return A;
}

// IMLS of the search clusters [c_lo, c_hi) (requires fg_imls_rangeable; the singular counter accumulates over the ranges)
int fg_impl_imls_range(fg_ctx *ctx, GaussSet &S, int64_t c_lo, int64_t c_hi) {
    NodeSet &N = ctx->nodes;
    const MlsArgs A = mls_args(ctx, S);
    if (N.dim == 2) return N.shape == FG_SHAPE_RECTANGULAR ? run_range<2, FG_SHAPE_RECTANGULAR>(ctx, S, A, c_lo, c_hi) : run_range<2, FG_SHAPE_CIRCULAR>(ctx, S, A, c_lo, c_hi);
    return N.shape == FG_SHAPE_RECTANGULAR ? run_range<3, FG_SHAPE_RECTANGULAR>(ctx, S, A, c_lo, c_hi) : run_range<3, FG_SHAPE_CIRCULAR>(ctx, S, A, c_lo, c_hi);
}

// Deferred shape functions: everything that is still missing, now.
int fg_ensure_shapes(fg_ctx *ctx, GaussSet &S) {
    if (!S.sf_lazy || !S.have_sf) return FG_OK;
    const int rc = fg_impl_imls_range(ctx, S, S.sf_done, S.nbcl.n);
    if (rc) return rc;
    S.sf_done = S.nbcl.n; S.sf_lazy = false; S.dom_pending = false;
    return FG_OK;
}

// fg_shape_functions with the option "lazy_shapes": buffers and counter are set up, the kernels run when a consumer needs them
int fg_defer_imls(fg_ctx *ctx, GaussSet &S) {
    S.singular = 0;
    FG_CUDA(ctx, S.singular_dev.reserve(1));
    FG_CUDA(ctx, cudaMemsetAsync(S.singular_dev.p, 0, sizeof(int), ctx->stream));
    S.sf_lazy = true; S.sf_done = 0;
    return FG_OK;
}

template <int DIM, int SHAPE>
static int run(fg_ctx *ctx, GaussSet &S, const MlsArgs &A) {
    constexpr int M = DIM == 2 ? 4 : 8;
    if (fg_imls_rangeable(ctx, S)) {
        const int rc = run_range<DIM, SHAPE>(ctx, S, A, 0, S.nbcl.n);
        if (rc) return rc;
        S.dom_pending = false;
        return FG_OK;
    }
    if (S.nb_clustered && S.nb_capc == 128 && S.max_L <= 96 && fg_opt(ctx, "imls_cluster", 1) != 0) {
        constexpr int NG = (1 + DIM) * M, NT = M * (M + 1) / 2;
        constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;
        const int lstride = (S.max_L + 3) & ~3;
        // every cluster's points ordered by list length (the largest cluster and candidate list, known from the search's count
        // pass, size the grid and the record buffer: no host synchronisation in this call)
        FG_CUDA(ctx, S.imls_perm.reserve(S.n));
        k_cluster_order<<<(unsigned)S.nbcl.n, 128, 0, ctx->stream>>>(0, S.nbcl.grid.start.p, S.nbcl.sorted.p, S.off.p, S.imls_perm.p);
        FG_CHECK_LAUNCH(ctx);
        const int G = (S.nb_max_points + 31) / 32, nrec_cap = std::max(8, (S.nb_max_cand + 7) & ~7);
        const size_t smem = (size_t)nrec_cap * 2 * DIM * 8 + (size_t)ROWS * 32 * 8 + 32 * 8 + 36 * 4 + (size_t)32 * lstride + sizeof(int) * nrec_cap;
        if (G > 0 && (int64_t)S.nbcl.n * G < ((int64_t)1 << 31)) {
            ClusterView V;
            V.cstart = S.nbcl.grid.start.p; V.csorted = S.nbcl.sorted.p; V.cperm = S.imls_perm.p; V.cand = S.cand.p; V.masks = S.masks.p; V.capc = S.nb_capc;
            FG_CUDA(ctx, cudaFuncSetAttribute(k_imls_c<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_imls_c<DIM, SHAPE><<<(unsigned)(S.nbcl.n * G), 32, smem, ctx->stream>>>(A, V, lstride, G, nrec_cap, S.dom_pending ? S.dom.p : nullptr);
            FG_CHECK_LAUNCH(ctx);
            S.dom_pending = false;
            return FG_OK;
        }
    }
    { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }          // the kernels below read DOM
    if (fg_opt(ctx, "imls_warp", 0) != 0) {
        constexpr int NG = (1 + DIM) * M, NT = M * (M + 1) / 2;
        constexpr int ROWS = (NG + 2 * DIM) > NT ? (NG + 2 * DIM) : NT;
        const size_t smem = 4 * (size_t)(ROWS * 32 * 8 + 32 * 8 + 36 * 4);
        FG_CUDA(ctx, S.imls_perm.reserve(((S.n + IMLS_TILE - 1) / IMLS_TILE) * IMLS_TILE));
        k_imls_order<<<(unsigned)((S.n + IMLS_TILE - 1) / IMLS_TILE), IMLS_TILE, 0, ctx->stream>>>(S.n, S.off.p, S.imls_perm.p);
        FG_CHECK_LAUNCH(ctx);
        FG_CUDA(ctx, cudaFuncSetAttribute(k_imls_w<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_imls_w<DIM, SHAPE><<<(unsigned)((S.n + 127) / 128), 128, smem, ctx->stream>>>(A, S.imls_perm.p);
        FG_CHECK_LAUNCH(ctx);
        return FG_OK;
    }
    const int bd = IMLS_BLOCK;
    const size_t rows = (1 + DIM) * M + DIM;                                          // >= 3 KB: also holds the sort scratch
    const size_t smem = sizeof(double) * rows * bd + sizeof(int) * (bd + 4) + (size_t)S.max_L * bd + 16;
    if (smem > 200 * 1024) return fg_fail(ctx, FG_E_CAPACITY, "fg_shape_functions: a Gauss point has %d neighbours; the IMLS kernel holds at most %d", S.max_L, 1400);
    FG_CUDA(ctx, cudaFuncSetAttribute(k_imls<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_imls<DIM, SHAPE><<<(unsigned)((S.n + bd - 1) / bd), bd, smem, ctx->stream>>>(A, S.max_L);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

int fg_impl_imls(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    S.singular = 0; S.sf_lazy = false;
    if (S.n == 0 || S.total_L == 0) return FG_OK;
    FG_CUDA(ctx, S.singular_dev.reserve(1));
    FG_CUDA(ctx, cudaMemsetAsync(S.singular_dev.p, 0, sizeof(int), ctx->stream));
    const MlsArgs A = mls_args(ctx, S);
    int rc;
    if (N.dim == 2) rc = N.shape == FG_SHAPE_RECTANGULAR ? run<2, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<2, FG_SHAPE_CIRCULAR>(ctx, S, A);
    else rc = N.shape == FG_SHAPE_RECTANGULAR ? run<3, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<3, FG_SHAPE_CIRCULAR>(ctx, S, A);
    return rc;   // the singular counter stays on the device until fg_singular_count asks for it
}

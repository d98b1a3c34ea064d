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
#include <stdlib.h>
#include <string.h>
#include <algorithm>

struct AsmArgs {
    int64_t nnodes, ngauss;
    const double *gs;
    const int64_t *off;
    const int *dom;
    const double *phi, *dphi;   // dphi is component-major: dphi[d * tl + e]
    int64_t tl;                 // total_L (stride between the components of dphi)
    const int64_t *colptr;
    const int *rowval;
    double *nzval;
    const double *coef;     // device
    int64_t coef_stride;
    double c[4];
    int dom_sorted;         // neighbour lists ascending (EFG search) or in element vertex order (FEM sets)
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

// Direct path: one warp per Gauss point, global FP64 atomics.  With `plist` it walks the listed
// points only (plist[0] = count): the points of clusters too large for the shared-memory path.
template <int DIM, int D, int KIND>
__global__ void __launch_bounds__(128) k_bilinear(AsmArgs A, const int *__restrict__ plist) {
    constexpr int nb = 1 + DIM, n = D * nb;
    __shared__ double sC[4][n * n];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t wglobal = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t count = plist ? plist[0] : A.ngauss;
  for (int64_t it = wglobal; it < count; it += wstride) {
    const int64_t g = plist ? plist[1 + it] : it;
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
        for (int d = 0; d < DIM; ++d) Bb[1 + d] = A.dphi[d * A.tl + o0 + b];
        const int j = A.dom[o0 + b];
        const int64_t c0 = A.colptr[j];
        const int nj = (int)(A.colptr[j + 1] - c0);
        const int *rows = A.rowval + c0;
        int pos = 0;
        for (int a = 0; a < L; ++a) {
            const int i = A.dom[o0 + a];
            int lo = A.dom_sorted ? pos : 0, hi = nj - 1;       // rows[lo..hi] contains i
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows[mid] < i) lo = mid + 1; else hi = mid; }
            pos = lo;
            double Ba[nb];
            Ba[0] = A.phi[o0 + a];
#pragma unroll
            for (int d = 0; d < DIM; ++d) Ba[1 + d] = A.dphi[d * A.tl + o0 + a];
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
    __syncwarp();
  }
}

template <int DIM, int D>
static int launch_kind(fg_ctx *ctx, const AsmArgs &A, int kind, int64_t ngauss, const int *plist) {
    const int bd = 128;
    const unsigned grid = plist ? (unsigned)(ctx->sm_count * 8) : (unsigned)((ngauss + 3) / 4);
    switch (kind) {
    case FG_OP_GENERIC: k_bilinear<DIM, D, FG_OP_GENERIC><<<grid, bd, 0, ctx->stream>>>(A, plist); break;
    case FG_OP_GRADGRAD: k_bilinear<DIM, D, FG_OP_GRADGRAD><<<grid, bd, 0, ctx->stream>>>(A, plist); break;
    case FG_OP_MASS: k_bilinear<DIM, D, FG_OP_MASS><<<grid, bd, 0, ctx->stream>>>(A, plist); break;
    case FG_OP_ADVECTION: if (D == 1) k_bilinear<DIM, 1, FG_OP_ADVECTION><<<grid, bd, 0, ctx->stream>>>(A, plist); break;
    case FG_OP_ELASTICITY: if (D == DIM) k_bilinear<DIM, DIM, FG_OP_ELASTICITY><<<grid, bd, 0, ctx->stream>>>(A, plist); break;
    }
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// Cluster path: one CTA per cluster of Gauss points (fg_cluster.cu).  The cluster's |U| x |U|
// block of K is accumulated in shared memory with plain read-modify-writes: warp w owns the rows
// with local index == w (mod 8), so no two warps ever touch the same entry and the 8 warps stream
// through the cluster's points independently (lanes = trial index b, loop over the owned test
// rows a).  One flush per cluster: every warp takes columns of the block, reads the column's row
// list from the CSC pattern (coalesced), finds the local row by binary search in U and issues
// one RED per non-zero entry -- adjacent lanes hit adjacent addresses.
// --------------------------------------------------------------------------------------------
struct ClArgs {
    const int *cstart, *csorted, *nU, *nodes;
    const unsigned char *lidx;
    int debug_skip;          // profiling aid: 1 = skip the flush, 2 = skip the accumulation (results are wrong)
    const int64_t *col0;
    const int *coln;
    const unsigned char *tab;
    int *ovf_extra;      // = ovf_points: [0] count, then ids (appended to by the assembly kernel)
    int ucap;
    const unsigned *inv; // 32-node blocks: gather map of k_bilinear_pts (8 words per Gauss point)
};

#define CL_WARPS 8

// per-column operand T_b such that K_ab[ir][jc] = sum_al B_a[al] * T_b[al]
template <int DIM, int KIND>
__device__ __forceinline__ void column_operand(const AsmArgs &A, const double *__restrict__ C, int n, int ir, int jc,
                                               const double (&Bb)[1 + DIM], double wj, double (&T)[1 + DIM]) {
    constexpr int nb = 1 + DIM;
#pragma unroll
    for (int al = 0; al < nb; ++al) T[al] = 0.0;
    if (KIND == FG_OP_GRADGRAD) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) T[1 + d] = A.c[0] * wj * Bb[1 + d];
    } else if (KIND == FG_OP_MASS) {
        T[0] = A.c[0] * wj * Bb[0];
    } else if (KIND == FG_OP_ADVECTION) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += A.c[d] * Bb[1 + d];
        T[0] = wj * s;
    } else if (KIND == FG_OP_ELASTICITY) {
        const double G = A.c[0], lam = A.c[1];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double t = 0.0;
            if (d == jc) t += G * Bb[1 + ir];
            if (d == ir) t += lam * Bb[1 + jc];
            if (ir == jc) t += G * Bb[1 + d];
            T[1 + d] = wj * t;
        }
    } else {
#pragma unroll
        for (int al = 0; al < nb; ++al) {
            double t = 0.0;
#pragma unroll
            for (int be = 0; be < nb; ++be) t += C[(ir * nb + al) * n + (jc * nb + be)] * Bb[be];
            T[al] = wj * t;
        }
    }
}

// Warp-per-cluster kernel.  Each warp owns a private |U| x |U| block in shared memory (upper triangle
// only when the integrand is symmetric) and walks the cluster's Gauss points: lane a holds the test-side
// operand B_a of neighbour a and the address of its row; the trial-side operands T_b and the column
// offsets are parked in a 32-record buffer and broadcast; four columns are in flight per trip (their
// read-modify-writes hit distinct entries).  Point headers are fetched 32 at a time and the operands of
// the next two points are in flight while the current one is accumulated.  Flush: the precomputed table
// (fg_cluster.cu) says where block entry (i,j) lives in the CSC column of node U[j]; non-zero entries
// leave as REDs, adjacent lanes to adjacent rows.
struct PointOperands { int la; double B[4]; };

// FP64 m8n8k4 MMA: D(8x8) = A(8x4) * B(4x8) + C.  Lane l holds A[l>>2][l&3], B[l&3][l>>2], C/D[l>>2][2*(l&3)+{0,1}].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

template <int DIM, int KIND, bool SYM, bool MMA>
__global__ void __launch_bounds__(32) k_bilinear_warp(AsmArgs A, ClArgs Cl, int D, int64_t nclusters) {
    constexpr int nb = 1 + DIM;
    constexpr bool USE_PHI = KIND == FG_OP_MASS || KIND == FG_OP_ADVECTION || KIND == FG_OP_GENERIC;
    constexpr bool USE_GRAD = KIND != FG_OP_MASS && KIND != FG_OP_ADVECTION;
    extern __shared__ double smem_d[];
    const int lane = threadIdx.x;
    const int ucap = Cl.ucap;
    const int kcap = SYM ? ucap * (ucap + 1) / 2 : ucap * (ucap | 1);
    double *Kloc = smem_d;
    double *sT = Kloc + ((kcap + 1) & ~1);                 // 32 records of 4 doubles: trial operands T_b
    double *sBr = sT + 32 * 4;                             // 32 records of 4 doubles: test operands B_a (MMA path)
    int *sOfs = (int *)(sBr + 32 * 4);                     // 32 column offsets (in doubles)
    int *sRow = sOfs + 32;                                 // 32 row offsets (MMA path)
    const int n = D * nb;

    for (int64_t c = blockIdx.x; c < nclusters; c += gridDim.x) {
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int ld = nu | 1;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        // flush table of my two block columns (j = lane, lane + 32): 64 bytes each, loaded now, used after the
        // accumulation, so that their latency hides behind it
        const bool fast_tab = ucap == 64;
        uint4 tabreg[2][4];
        int64_t fc0[2] = {0, 0}; int fnj[2] = {0, 0};
        if (fast_tab) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                if (j < nu) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(Cl.tab + c * (int64_t)(64 * 64) + j * 64);
#pragma unroll
                    for (int w = 0; w < 4; ++w) tabreg[h][w] = __ldg(src + w);
                    fc0[h] = Cl.col0[c * 64 + j];
                    fnj[h] = Cl.coln[c * 64 + j];
                }
            }
        }
        for (int comp = 0; comp < D * D; ++comp) {
            const int jc = comp / D, ir = comp % D;
            if ((KIND == FG_OP_GRADGRAD || KIND == FG_OP_MASS) && ir != jc) continue;
            const int kn = SYM ? nu * (nu + 1) / 2 : nu * ld;
            __syncwarp();
            for (int t = lane; t < kn; t += 32) Kloc[t] = 0.0;
            for (int pb = p0; pb < p1 && FG_DBG(Cl.debug_skip) != 2; pb += 32) {
                // headers of up to 32 points, one per lane
                const int npt = min(32, p1 - pb);
                int hg = 0, hL = 0; int64_t ho = 0; double hw = 0.0;
                if (lane < npt) {
                    hg = Cl.csorted[pb + lane];
                    ho = A.off[hg];
                    hL = (int)(A.off[hg + 1] - ho);
                    hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
                }
                auto fetch = [&](int q, PointOperands &P) {
                    if (q >= npt) return;
                    const int L = __shfl_sync(0xffffffffu, hL, q);
                    const int64_t o0 = __shfl_sync(0xffffffffu, ho, q);
                    if (lane < L && L <= 32) {
                        P.la = Cl.lidx[o0 + lane];
                        P.B[0] = KIND == FG_OP_GRADGRAD || KIND == FG_OP_ELASTICITY ? 0.0 : A.phi[o0 + lane];
#pragma unroll
                        for (int d = 0; d < DIM; ++d) P.B[1 + d] = KIND == FG_OP_MASS ? 0.0 : A.dphi[d * A.tl + o0 + lane];
                    }
                };
                auto process = [&](int q, const PointOperands &P) {
                    const int L = __shfl_sync(0xffffffffu, hL, q);
                    const int g = __shfl_sync(0xffffffffu, hg, q);
                    const double wj = __shfl_sync(0xffffffffu, hw, q);
                    if (L > 32) {          // longer than a warp: hand the point to the direct path (once)
                        if (lane == 0 && comp == 0) { const int slot = atomicAdd(Cl.ovf_extra, 1); Cl.ovf_extra[1 + slot] = g; }
                        return;
                    }
                    const bool row_ok = lane < L;
                    const int la = P.la;
                    double Ba[nb], T[nb];
#pragma unroll
                    for (int al = 0; al < nb; ++al) Ba[al] = P.B[al];
                    if (row_ok) column_operand<DIM, KIND>(A, KIND == FG_OP_GENERIC ? A.coef + g * A.coef_stride : nullptr, n, ir, jc, Ba, wj, T);
                    __syncwarp();
                    if (MMA) {
                        // operand records of all 32 slots (zero beyond L), row / column offsets into the block
#pragma unroll
                        for (int al = 0; al < 4; ++al) {
                            sT[lane * 4 + al] = (row_ok && al < nb) ? T[al] : 0.0;
                            sBr[lane * 4 + al] = (row_ok && al < nb) ? Ba[al] : 0.0;
                        }
                        sOfs[lane] = SYM ? la * (la + 1) / 2 : la;
                        sRow[lane] = SYM ? la : la * ld;
                        __syncwarp();
                        const int gid = lane >> 2, tig = lane & 3;
                        const int nblk = (L + 7) >> 3;
                        double af[4], bf[4]; int ro[4]; int2 co[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < nblk) {
                                af[r] = sBr[(8 * r + gid) * 4 + tig];
                                bf[r] = sT[(8 * r + gid) * 4 + tig];
                                ro[r] = sRow[8 * r + gid];
                                co[r] = *reinterpret_cast<const int2 *>(sOfs + 8 * r + 2 * tig);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r >= nblk) break;
                            const int arow = 8 * r + gid;
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2) {
                                if (s2 >= nblk || (SYM && s2 < r)) continue;
                                double d0, d1;
                                dmma884(d0, d1, af[r], bf[s2]);
                                const int bcol = 8 * s2 + 2 * tig;
                                const bool ok0 = arow < L && bcol < L && (!SYM || arow <= bcol);
                                const bool ok1 = arow < L && bcol + 1 < L && (!SYM || arow <= bcol + 1);
                                double *k0 = Kloc + ro[r] + co[s2].x, *k1 = Kloc + ro[r] + co[s2].y;
                                double o0 = 0.0, o1 = 0.0;
                                if (ok0) o0 = *k0;
                                if (ok1) o1 = *k1;
                                if (ok0) *k0 = o0 + d0;
                                if (ok1) *k1 = o1 + d1;
                            }
                        }
                        return;
                    }
                    if (row_ok) {
#pragma unroll
                        for (int al = 0; al < nb; ++al) sT[lane * 4 + al] = T[al];
                        sOfs[lane] = SYM ? la * (la + 1) / 2 : la;          // column offset: tri(lb) or lb
                    }
                    __syncwarp();
                    // my row: full storage Kloc[la*ld + lb]; symmetric storage Kloc[tri(lb) + la] for la <= lb
                    double *row = Kloc + (SYM ? la : la * ld);
                    int b = 0;
                    for (; b + 4 <= L; b += 4) {
                        double v[4]; int o[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double *Tb = sT + (b + u) * 4;
                            o[u] = sOfs[b + u];
                            double t = 0.0;
                            if (USE_PHI) t = Ba[0] * Tb[0];
                            if (USE_GRAD) {
#pragma unroll
                                for (int d = 0; d < DIM; ++d) t += Ba[1 + d] * Tb[1 + d];
                            }
                            v[u] = t;
                        }
                        if (row_ok) {
                            double k[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) k[u] = (!SYM || lane <= b + u) ? row[o[u]] : 0.0;
#pragma unroll
                            for (int u = 0; u < 4; ++u) if (!SYM || lane <= b + u) row[o[u]] = k[u] + v[u];
                        }
                    }
                    for (; b < L; ++b) {
                        const double *Tb = sT + b * 4;
                        const int o = sOfs[b];
                        double t = 0.0;
                        if (USE_PHI) t = Ba[0] * Tb[0];
                        if (USE_GRAD) {
#pragma unroll
                            for (int d = 0; d < DIM; ++d) t += Ba[1 + d] * Tb[1 + d];
                        }
                        if (row_ok && (!SYM || lane <= b)) row[o] += t;
                    }
                };
                // ---- direct-fragment path (gradgrad / mass, MMA): T_b = scale * B_b, so both MMA operands of a point are
                //      the same numbers; every lane fetches ITS fragment elements and block offsets straight from global
                //      memory (two points ahead) and nothing is staged through shared memory -------------------------------
                constexpr bool DIRECT = MMA && (KIND == FG_OP_GRADGRAD || KIND == FG_OP_MASS);
                if (DIRECT) {
                    const int gid = lane >> 2, tig = lane & 3;
                    const unsigned kbase = (unsigned)__cvta_generic_to_shared(Kloc);
                    const bool dg0 = gid <= 2 * tig, dg1 = gid <= 2 * tig + 1;     // diagonal tiles: row <= column
                    struct Frag { double a[4]; int la; };
                    auto dfetch = [&](int q, Frag &F) {
                        if (q >= npt) return;
                        const int L = __shfl_sync(0xffffffffu, hL, q);
                        const int64_t o0 = __shfl_sync(0xffffffffu, ho, q);
                        if (L > 32) return;
                        F.la = lane < L ? (int)Cl.lidx[o0 + lane] : 0;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int a = 8 * r + gid;
                            F.a[r] = 0.0;
                            if (a < L) {
                                if (KIND == FG_OP_MASS) { if (tig == 0) F.a[r] = A.phi[o0 + a]; }
                                else if (tig >= 1 && tig <= DIM) F.a[r] = A.dphi[(tig - 1) * A.tl + o0 + a];
                            }
                        }
                    };
                    auto dprocess = [&](int q, const Frag &F) {
                        const int L = __shfl_sync(0xffffffffu, hL, q);
                        const int g = __shfl_sync(0xffffffffu, hg, q);
                        const double scale = A.c[0] * __shfl_sync(0xffffffffu, hw, q);
                        if (L > 32) {
                            if (lane == 0 && comp == 0) { const int slot = atomicAdd(Cl.ovf_extra, 1); Cl.ovf_extra[1 + slot] = g; }
                            return;
                        }
                        const int nblk = (L + 7) >> 3;
                        __syncwarp();          // the previous point's stores to the block are visible to every lane before it is read again
                        // byte offsets into the block: row part and column part, handed around with shuffles
                        const int rowpart = (SYM ? F.la : F.la * ld) * 8;
                        const int colpart = (SYM ? F.la * (F.la + 1) / 2 : F.la) * 8;
                        unsigned ro[4], c0[4], c1[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            ro[r] = kbase + __shfl_sync(0xffffffffu, rowpart, 8 * r + gid);
                            c0[r] = __shfl_sync(0xffffffffu, colpart, (8 * r + 2 * tig) & 31);
                            c1[r] = __shfl_sync(0xffffffffu, colpart, (8 * r + 2 * tig + 1) & 31);
                        }
                        // phase 1: current block entries of all my tile elements (0 where masked)
                        double k0[4][4], k1[4][4];
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2) {
                                k0[r][s2] = 0.0; k1[r][s2] = 0.0;
                                if (r < nblk && s2 < nblk && (!SYM || s2 >= r)) {
                                    const int arow = 8 * r + gid, bcol = 8 * s2 + 2 * tig;
                                    const bool ok0 = arow < L && bcol < L && (!SYM || s2 > r || dg0);
                                    const bool ok1 = arow < L && bcol + 1 < L && (!SYM || s2 > r || dg1);
                                    if (ok0) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(k0[r][s2]) : "r"(ro[r] + c0[s2]));
                                    if (ok1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(k1[r][s2]) : "r"(ro[r] + c1[s2]));
                                }
                            }
                        // phase 2: D = A * (scale B)^T + K   (the add rides in the MMA)
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2)
                                if (r < nblk && s2 < nblk && (!SYM || s2 >= r))
                                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                                 : "+d"(k0[r][s2]), "+d"(k1[r][s2]) : "d"(F.a[r]), "d"(scale * F.a[s2]));
                        // phase 3: store back
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2)
                                if (r < nblk && s2 < nblk && (!SYM || s2 >= r)) {
                                    const int arow = 8 * r + gid, bcol = 8 * s2 + 2 * tig;
                                    const bool ok0 = arow < L && bcol < L && (!SYM || s2 > r || dg0);
                                    const bool ok1 = arow < L && bcol + 1 < L && (!SYM || s2 > r || dg1);
                                    if (ok0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ro[r] + c0[s2]), "d"(k0[r][s2]) : "memory");
                                    if (ok1) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ro[r] + c1[s2]), "d"(k1[r][s2]) : "memory");
                                }
                    };
                    Frag F0, F1;
                    dfetch(0, F0);
                    dfetch(1, F1);
                    for (int q = 0; q < npt; q += 2) {
                        Frag Fc = F0;
                        dfetch(q + 2, F0);
                        dprocess(q, Fc);
                        if (q + 1 < npt) {
                            Fc = F1;
                            dfetch(q + 3, F1);
                            dprocess(q + 1, Fc);
                        }
                    }
                } else {
                PointOperands P0, P1;
                P0.la = P1.la = 0;
#pragma unroll
                for (int al = 0; al < 4; ++al) P0.B[al] = P1.B[al] = 0.0;
                fetch(0, P0);
                fetch(1, P1);
                for (int q = 0; q < npt; q += 2) {
                    PointOperands Pc = P0;
                    fetch(q + 2, P0);
                    process(q, Pc);
                    if (q + 1 < npt) {
                        Pc = P1;
                        fetch(q + 3, P1);
                        process(q + 1, Pc);
                    }
                }
                }
            }
            __syncwarp();
            // ---- flush: lane = block column, walk down its rows ---------------------------------------------------
            if (fast_tab && FG_DBG(Cl.debug_skip) != 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = lane + 32 * h;
                    const bool col_ok = j < nu;
                    double *dst = A.nzval + fc0[h] * (D * D) + (int64_t)jc * fnj[h] * D + ir;
                    const double *col = Kloc + (SYM ? j * (j + 1) / 2 : j);
                    const int iend = col_ok ? (SYM ? j + 1 : nu) : 0;
                    const int imax = SYM ? min(nu, 32 * h + 32) : nu;            // warp-uniform bound
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        if (16 * w >= imax) break;
                        const unsigned words[4] = {tabreg[h][w].x, tabreg[h][w].y, tabreg[h][w].z, tabreg[h][w].w};
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            const int i = 16 * w + t;
                            const int k = (words[t >> 2] >> (8 * (t & 3))) & 255;
                            if (i < iend && k != 255) {
                                const double v = SYM ? col[i] : col[i * ld];
                                if (v != 0.0) atomicAdd(dst + (int64_t)k * D, v);
                            }
                        }
                    }
                }
            } else {
                const unsigned char *tab = Cl.tab + c * (int64_t)ucap * ucap;
                for (int j0 = 0; j0 < nu && FG_DBG(Cl.debug_skip) != 1; j0 += 32) {
                    const int j = j0 + lane;
                    const bool col_ok = j < nu;
                    double *dst = A.nzval;
                    if (col_ok) {
                        const int64_t c0 = Cl.col0[c * ucap + j];
                        const int nj = Cl.coln[c * ucap + j];
                        dst += c0 * (D * D) + (int64_t)jc * nj * D + ir;
                    }
                    const double *col = Kloc + (SYM ? j * (j + 1) / 2 : j);
                    const int iend = col_ok ? (SYM ? j + 1 : nu) : 0;
                    const int imax = SYM ? min(nu, j0 + 32) : nu;
                    for (int i = 0; i < imax; ++i) {
                        if (i < iend) {
                            const int k = tab[j * ucap + i];
                            const double v = SYM ? col[i] : col[i * ld];
                            if (k != 255 && v != 0.0) atomicAdd(dst + (int64_t)k * D, v);
                        }
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// Register-block kernel (symmetric scalar forms, clusters of <= 64 nodes): the cluster's 64 x 64 block never
// touches shared memory.  One warp per cluster holds the 36 upper 8x8 tiles of the block as FP64 MMA accumulators
// (72 doubles per lane).  For every Gauss point the test operand is scattered into LOCAL node order -- lane
// (gid,tig) of band r holds component tig of the neighbour whose local index is 8r+gid, or 0 -- through a 64-byte
// inverse map in shared memory; then tile (r,s) += A_r (8x4) . (scale A_s)^T (4x8) for the bands that have members
// (warp-uniform mask).  No read-modify-write, no address arithmetic, no per-point offsets; the MMAs of a point are
// independent.  The flush reads the accumulators and issues REDs through the same position table.
// --------------------------------------------------------------------------------------------
// All MMAs of one Gauss point for a COMPILE-TIME band mask: no branches between the MMAs, exactly the tiles whose
// row and column bands have members.
template <int NB, unsigned BM>
__device__ __forceinline__ void mma_tiles(double (&acc)[NB * (NB + 1) / 2][2], const double (&af)[NB], const double (&bf)[NB]) {
    int t = 0;
#pragma unroll
    for (int r = 0; r < NB; ++r)
#pragma unroll
        for (int s2 = r; s2 < NB; ++s2, ++t)
            if (((BM >> r) & 1u) && ((BM >> s2) & 1u))
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(af[r]), "d"(bf[s2]));
}

// NB = bands of 8 local indices: 8 (clusters of <= 64 nodes, 36 tiles, 8 warps/SM) or 4 (<= 32 nodes, 10 tiles,
// 16+ warps/SM -- node-centred single-spacing cells of a structured cloud have 27-node unions).
template <int DIM, int KIND, int NB>
__global__ void __launch_bounds__(32, NB == 4 ? 20 : 1) k_bilinear_reg(AsmArgs A, ClArgs Cl, int64_t nclusters) {
    constexpr int U = 8 * NB, NT = NB * (NB + 1) / 2;
    __shared__ __align__(16) double sloc[2][U][4];           // operands in local node order, double buffered
    __shared__ __align__(16) unsigned char stab[U * U];      // flush table of the current cluster
    __shared__ long long scol0[U];
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    for (int t = lane; t < 2 * U * 4; t += 32) (&sloc[0][0][0])[t] = 0.0;
    int prev_la[2] = {-1, -1};
    __syncwarp();
    for (int64_t c = blockIdx.x; c < nclusters; c += gridDim.x) {
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        __syncwarp();
        {   // flush table + column starts stream into shared memory while the points are accumulated
            const unsigned char *src = Cl.tab + c * (int64_t)(U * U);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab);
            for (int t = lane * 16; t < nu * U; t += 32 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + t), "l"(src + t) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            for (int j = lane; j < nu; j += 32) scol0[j] = Cl.col0[c * U + j];
        }
        double acc[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        for (int pb = p0; pb < p1 && FG_DBG(Cl.debug_skip) != 2; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0, hL = 0; int64_t ho = 0; double hw = 0.0;
            if (lane < npt) {
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hL = (int)(A.off[hg + 1] - ho);
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
                if (hL > 32) { const int slot = atomicAdd(Cl.ovf_extra, 1); Cl.ovf_extra[1 + slot] = hg; }   // direct path
            }
            // operands of a point in ENTRY order (coalesced, independent of everything: fetched two points ahead)
            auto fetch = [&](int q, PointOperands &P) {
                if (q >= npt) return;
                const int L = __shfl_sync(0xffffffffu, hL, q);
                const int64_t o0 = __shfl_sync(0xffffffffu, ho, q);
                P.la = -1;
                if (FG_DBG(Cl.debug_skip) == 3) { if (lane < L) { P.la = lane; P.B[0] = 0.0; P.B[1] = 1.0; P.B[2] = 2.0; P.B[3] = 3.0; } return; }
                if (lane < L && L <= 32) {
                    P.la = Cl.lidx[o0 + lane];
                    P.B[0] = KIND == FG_OP_MASS ? A.phi[o0 + lane] : 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) P.B[1 + d] = KIND == FG_OP_MASS ? 0.0 : A.dphi[d * A.tl + o0 + lane];
                    if (DIM == 2) P.B[3] = 0.0;
                }
            };
            auto process = [&](int q, const PointOperands &P) {
                const int L = __shfl_sync(0xffffffffu, hL, q);
                if (L > 32) return;
                const double scale = A.c[0] * __shfl_sync(0xffffffffu, hw, q);
                const int buf = q & 1;
                // scatter into local node order: clear the slot I used last time in this buffer, then write mine
                if (prev_la[buf] >= 0) { double2 *z = reinterpret_cast<double2 *>(sloc[buf][prev_la[buf]]); z[0] = make_double2(0.0, 0.0); z[1] = make_double2(0.0, 0.0); }
                __syncwarp();
                if (P.la >= 0) {
                    double2 *w = reinterpret_cast<double2 *>(sloc[buf][P.la]);
                    w[0] = make_double2(P.B[0], P.B[1]); w[1] = make_double2(P.B[2], P.B[3]);
                }
                prev_la[buf] = P.la;
                // 4 bands: every tile is multiplied anyway and unused slots hold zeros, so no band mask is needed
                const unsigned bm = NB == 4 ? 0xFu : __reduce_or_sync(0xffffffffu, P.la >= 0 ? (1u << (P.la >> 3)) : 0u);
                __syncwarp();
                double af[NB], bf[NB];
#pragma unroll
                for (int r = 0; r < NB; ++r) { af[r] = ((bm >> r) & 1u) ? sloc[buf][8 * r + gid][tig] : 0.0; bf[r] = scale * af[r]; }
                // the band masks that occur in the interior of a structured cloud get branch-free MMA lists
                if (FG_DBG(Cl.debug_skip) == 4) { acc[0][0] += af[0] + af[NB - 1] + bf[3]; return; }
                if (NB == 4) {          // 10 tiles: all of them, absent bands hold zeros
                    mma_tiles<NB, 0xFu>(acc, af, bf);
                } else {
                    switch (bm) {
                    case 0x3Fu: mma_tiles<NB, 0x3Fu>(acc, af, bf); break;
                    case 0x3Cu: mma_tiles<NB, 0x3Cu>(acc, af, bf); break;
                    case 0xFCu: mma_tiles<NB, 0xFCu>(acc, af, bf); break;
                    case 0x0Fu: mma_tiles<NB, 0x0Fu>(acc, af, bf); break;
                    case 0xFFu: mma_tiles<NB, 0xFFu>(acc, af, bf); break;
                    default: {
                        int t = 0;
#pragma unroll
                        for (int r = 0; r < NB; ++r)
#pragma unroll
                            for (int s2 = r; s2 < NB; ++s2, ++t)
                                if (((bm >> r) & 1u) && ((bm >> s2) & 1u))
                                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                                 : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(af[r]), "d"(bf[s2]));
                    }
                    }
                }
            };
            PointOperands P0, P1;
            P0.la = P1.la = -1;
#pragma unroll
            for (int al = 0; al < 4; ++al) P0.B[al] = P1.B[al] = 0.0;
            fetch(0, P0);
            fetch(1, P1);
            for (int q = 0; q < npt; q += 2) {
                PointOperands Pc = P0;
                fetch(q + 2, P0);
                process(q, Pc);
                if (q + 1 < npt) {
                    Pc = P1;
                    fetch(q + 3, P1);
                    process(q + 1, Pc);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (FG_DBG(Cl.debug_skip) == 1) continue;
        // ---- flush: tile (r,s) element (8r+gid, 8s+2tig+e), upper triangle, through the position table ----------
        int t = 0;
#pragma unroll
        for (int r = 0; r < NB; ++r)
#pragma unroll
            for (int s2 = r; s2 < NB; ++s2, ++t) {
                const int i = 8 * r + gid;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * s2 + 2 * tig + e;
                    const double v = acc[t][e];
                    if (i <= j && j < nu && v != 0.0) {
                        const int k = stab[j * U + i];
                        if (k != 255) atomicAdd(A.nzval + scol0[j] + k, v);
                    }
                }
            }
    }
}

// --------------------------------------------------------------------------------------------
// Point-batched register-block kernel (32-node clusters): the k dimension of the FP64 MMA runs over FOUR GAUSS POINTS.
// For one operand component (d_i phi, or phi for the mass form) the cluster's contribution is  S = sum_g w_g v_g v_g^T  with
// v_g the component's values in local node order (zeros for non-neighbours): an (U x G)(G x U) product.  Lane (gid, tig) of
// band r holds v of local node 8r+gid at point tig of the current group of four -- fetched STRAIGHT from the entry-ordered
// arrays through the per-point inverse map (fg_cluster.cu: byte r of word gid = position of local node 8r+gid in the point's
// list, 255 = absent): no shared-memory scatter, no warp synchronisation, no per-point shuffles beyond the three header
// broadcasts per group.  tile(r,s) += A_r (8 nodes x 4 points) . (w A_s)^T: 10 MMAs per component per four points (7.5 per
// point for grad.grad, 2.5 for the mass form; the node-order kernel above issues 10 per point).
//   KIND = GRADGRAD / MASS, any D: the D x D block of a pair is s * I, so the scalar block is accumulated once and the flush
//   writes it to the D diagonal DOF positions.  The flush covers row node <= column node; k_mirror_dof fills the rest.
// --------------------------------------------------------------------------------------------
// Bulk asynchronous copy (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) of a contiguous global range into shared memory, completion
// signalled on an mbarrier: one elected lane issues the copy of a whole flush table instead of 32 lanes x cp.async x 16 bytes.
__device__ __forceinline__ void fg_mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fg_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// bounded wait (a mismatch of the expected byte count must not hang the device: the caller's result would be wrong and the parity tests fail)
__device__ __forceinline__ bool fg_mbar_wait(unsigned mbar, unsigned parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

template <int DIM, int KIND>
__global__ void __launch_bounds__(32, 20) k_bilinear_pts(AsmArgs A, ClArgs Cl, int D, int64_t nclusters) {
    constexpr int NB = 4, U = 32, NT = NB * (NB + 1) / 2;
    constexpr int NC = KIND == FG_OP_MASS ? 1 : DIM;
    __shared__ __align__(16) unsigned char stab[U * U];
    __shared__ long long scol0[U];
    __shared__ int scoln[U];
    __shared__ __align__(8) unsigned long long mbar;
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&mbar);
    unsigned mbar_phase = 0;
    if (lane == 0) fg_mbar_init(mbar_a, 1);
    __syncwarp();
    const double *__restrict__ vals = KIND == FG_OP_MASS ? A.phi : A.dphi;
    // Headers (point id, first entry, weight) of the first 32 points of a cluster, one point per lane.  The NEXT cluster's headers
    // are requested while the current cluster is accumulated and flushed: the csorted -> off / gs chain (three dependent global
    // loads, 30 % of the stall samples when it sat at the top of every cluster) is off the critical path.
    struct Head { int nu, p0, p1, hg; long long ho; double hw; };
    auto head_l1 = [&](int64_t c, Head &H) {                 // level 1: cluster extent
        H.nu = 0; H.p0 = H.p1 = 0; H.hg = 0; H.ho = 0; H.hw = 0.0;
        if (c < nclusters) { H.nu = Cl.nU[c]; H.p0 = Cl.cstart[c]; H.p1 = Cl.cstart[c + 1]; }
    };
    auto head_l2 = [&](Head &H) {                            // level 2: the lane's point
        if (H.nu > 0 && H.p0 + lane < H.p1) H.hg = Cl.csorted[H.p0 + lane];
    };
    auto head_l3 = [&](Head &H) {                            // level 3: its list start and weight
        if (H.nu > 0 && H.p0 + lane < H.p1) {
            H.ho = A.off[H.hg];
            H.hw = A.c[0] * (A.gs[(DIM + 1) * A.ngauss + H.hg] * A.gs[DIM * A.ngauss + H.hg]);
        }
    };
    Head cur_h, nxt_h;
    head_l1(blockIdx.x, cur_h); head_l2(cur_h); head_l3(cur_h);
    for (int64_t c = blockIdx.x; c < nclusters; c += gridDim.x) {
        head_l1(c + gridDim.x, nxt_h);
        const int nu = cur_h.nu;
        if (nu <= 0) { head_l2(nxt_h); head_l3(nxt_h); cur_h = nxt_h; continue; }
        const int p0 = cur_h.p0, p1 = cur_h.p1;
        // the previous cluster's flush has read the table (generic proxy): order those reads before the async-proxy write below
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        {   // flush table (nu x 32 bytes, contiguous): ONE bulk asynchronous copy issued by lane 0, landing while the points are accumulated
            if (lane == 0) fg_bulk_g2s((unsigned)__cvta_generic_to_shared(stab), Cl.tab + c * (int64_t)(U * U), (unsigned)(nu * U), mbar_a);
            if (lane < nu) { scol0[lane] = Cl.col0[c * U + lane]; scoln[lane] = Cl.coln[c * U + lane]; }
        }
        double acc[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = cur_h.hg; long long ho = cur_h.ho; double hw = cur_h.hw;
            if (pb > p0) {             // clusters of more than 32 points: later batches fetch their headers here
                hg = 0; ho = 0; hw = 0.0;
                if (lane < npt) {
                    hg = Cl.csorted[pb + lane];
                    ho = A.off[hg];
                    hw = A.c[0] * (A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg]);
                }
            }
            // Software pipeline at COMPONENT granularity, without extra registers: right after the ten MMAs of (group q, component c)
            // have issued, the same four registers are reloaded with component c of group q+1; those loads are in flight during the
            // MMAs of the other components (WAR on the operand registers is resolved at issue).  The inverse-map word of group q+2 is
            // requested one group earlier still.  (ncu history: all loads up front -> 40 % long-scoreboard stalls before the MMAs;
            // two full operand buffers -> 128 registers, 16 warps/SM, no faster.)
            auto map_word = [&](int q) -> unsigned {           // q: first point of a group (warp-uniform)
                const int src = q + tig;
                const int g = __shfl_sync(0xffffffffu, hg, src & 31);
                return (q < npt && src < npt) ? __ldg(Cl.inv + (int64_t)g * 8 + gid) : 0xFFFFFFFFu;
            };
            double buf[NC][NB];
            unsigned word = map_word(0), word_n = map_word(4);
            long long o0 = __shfl_sync(0xffffffffu, ho, tig);
            double w = __shfl_sync(0xffffffffu, hw, tig);
#pragma unroll
            for (int comp = 0; comp < NC; ++comp)
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    const unsigned pos = (word >> (8 * r)) & 255u;
                    buf[comp][r] = pos != 255u ? __ldg(vals + o0 + (int64_t)comp * A.tl + pos) : 0.0;
                }
#pragma unroll 1
            for (int q = 0; q < npt; q += 4) {
                const bool more = q + 4 < npt;
                const int srcn = (q + 4 + tig) & 31;
                const long long o0n = __shfl_sync(0xffffffffu, ho, srcn);
                const double wn = __shfl_sync(0xffffffffu, hw, srcn);
                const unsigned word_nn = map_word(q + 8);
#pragma unroll
                for (int comp = 0; comp < NC; ++comp) {
                    double bf[NB];
#pragma unroll
                    for (int r = 0; r < NB; ++r) bf[r] = w * buf[comp][r];
                    mma_tiles<NB, 0xFu>(acc, buf[comp], bf);
                    if (more) {
#pragma unroll
                        for (int r = 0; r < NB; ++r) {
                            const unsigned pos = (word_n >> (8 * r)) & 255u;
                            buf[comp][r] = pos != 255u ? __ldg(vals + o0n + (int64_t)comp * A.tl + pos) : 0.0;
                        }
                    }
                }
                w = wn; word_n = word_nn;
            }
            if (pb == p0) head_l2(nxt_h);      // the next cluster's point ids: requested before the flush ...
        }
        head_l3(nxt_h);                        // ... their list starts and weights during it
        fg_mbar_wait(mbar_a, mbar_phase);          // the table has landed
        mbar_phase ^= 1u;
        __syncwarp();
        // ---- flush: tile (r,s) element (8r+gid, 8s+2tig+e), row node <= column node, through the position table ----------
        int t = 0;
#pragma unroll
        for (int r = 0; r < NB; ++r)
#pragma unroll
            for (int s2 = r; s2 < NB; ++s2, ++t) {
                const int i = 8 * r + gid;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * s2 + 2 * tig + e;
                    const double v = acc[t][e];
                    if (i <= j && j < nu && v != 0.0) {
                        const int k = stab[j * U + i];
                        if (k != 255) {
                            if (D == 1) atomicAdd(A.nzval + scol0[j] + k, v);
                            else {
                                const int nj = scoln[j];
                                double *dst = A.nzval + scol0[j] * (D * D) + (int64_t)k * D;
                                for (int d = 0; d < D; ++d) atomicAdd(dst + (int64_t)d * nj * D + d, v);
                            }
                        }
                    }
                }
            }
        cur_h = nxt_h;
    }
}

// --------------------------------------------------------------------------------------------
// Linear elasticity, D = DIM (grad(du) (.) sigma(u), Fields_Operations.jl:552-567,834-838), with the same point-batched MMAs:
//     K_ab[p][q] = G d_pq sum_d S_dd^{ab} + G S_qp^{ab} + lambda S_pq^{ab},     S_pq^{ab} = sum_g w_g d_p phi_a d_q phi_b
// so the cluster only needs the DIM(DIM+1)/2 moment blocks S_pq, p <= q (S_qp^{ab} = S_pq^{ba}):
//   k_elast_diag: one warp per cluster holds S_dd (upper tiles, DIM x 10 accumulators) and flushes the DIM diagonal components;
//   k_elast_off:  one warp per (cluster, pair p < q) holds the full 32 x 32 block S_pq (16 tiles), transposes it through shared
//                 memory at the flush and writes K[p][q] and K[q][p] for row node < column node (and [p][q] on the node diagonal).
// 78 MMAs per four Gauss points in 3D (19.5 per point) instead of 9 passes of 16; k_mirror_dof fills the lower triangle.
// --------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32, DIM == 2 ? 16 : 12) k_elast_diag(AsmArgs A, ClArgs Cl, int64_t nclusters) {
    constexpr int NB = 4, U = 32, NT = NB * (NB + 1) / 2, D = DIM;
    __shared__ __align__(16) unsigned char stab[U * U];
    __shared__ long long scol0[U];
    __shared__ int scoln[U];
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    const double G = A.c[0], lam = A.c[1];
    for (int64_t c = blockIdx.x; c < nclusters; c += gridDim.x) {
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        __syncwarp();
        {
            const unsigned char *src = Cl.tab + c * (int64_t)(U * U);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab);
            for (int t = lane * 16; t < nu * U; t += 32 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + t), "l"(src + t) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (lane < nu) { scol0[lane] = Cl.col0[c * U + lane]; scoln[lane] = Cl.coln[c * U + lane]; }
        }
        double acc[DIM][NT][2];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int t = 0; t < NT; ++t) { acc[d][t][0] = 0.0; acc[d][t][1] = 0.0; }
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0; long long ho = 0; double hw = 0.0;
            if (lane < npt) {
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
            }
            // component-granular software pipeline, as in k_bilinear_pts: component d of the next group is requested right after the
            // MMAs of component d of this group have issued
            auto map_word = [&](int q) -> unsigned {
                const int src = q + tig;
                const int g = __shfl_sync(0xffffffffu, hg, src & 31);
                return (q < npt && src < npt) ? __ldg(Cl.inv + (int64_t)g * 8 + gid) : 0xFFFFFFFFu;
            };
            double buf[DIM][NB];
            unsigned word = map_word(0), word_n = map_word(4);
            long long o0 = __shfl_sync(0xffffffffu, ho, tig);
            double w = __shfl_sync(0xffffffffu, hw, tig);
#pragma unroll
            for (int d = 0; d < DIM; ++d)
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    const unsigned pos = (word >> (8 * r)) & 255u;
                    buf[d][r] = pos != 255u ? __ldg(A.dphi + o0 + (int64_t)d * A.tl + pos) : 0.0;
                }
#pragma unroll 1
            for (int q = 0; q < npt; q += 4) {
                const bool more = q + 4 < npt;
                const int srcn = (q + 4 + tig) & 31;
                const long long o0n = __shfl_sync(0xffffffffu, ho, srcn);
                const double wn = __shfl_sync(0xffffffffu, hw, srcn);
                const unsigned word_nn = map_word(q + 8);
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double bf[NB];
#pragma unroll
                    for (int r = 0; r < NB; ++r) bf[r] = w * buf[d][r];
                    mma_tiles<NB, 0xFu>(acc[d], buf[d], bf);
                    if (more) {
#pragma unroll
                        for (int r = 0; r < NB; ++r) {
                            const unsigned pos = (word_n >> (8 * r)) & 255u;
                            buf[d][r] = pos != 255u ? __ldg(A.dphi + o0n + (int64_t)d * A.tl + pos) : 0.0;
                        }
                    }
                }
                w = wn; word_n = word_nn;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        int t = 0;
#pragma unroll
        for (int r = 0; r < NB; ++r)
#pragma unroll
            for (int s2 = r; s2 < NB; ++s2, ++t) {
                const int i = 8 * r + gid;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * s2 + 2 * tig + e;
                    if (i <= j && j < nu) {
                        const int k = stab[j * U + i];
                        if (k != 255) {
                            double tr = 0.0;
#pragma unroll
                            for (int d = 0; d < DIM; ++d) tr += acc[d][t][e];
                            const int nj = scoln[j];
                            double *dst = A.nzval + scol0[j] * (D * D) + (int64_t)k * D;
#pragma unroll
                            for (int d = 0; d < DIM; ++d) {
                                const double v = G * tr + (G + lam) * acc[d][t][e];
                                if (v != 0.0) atomicAdd(dst + (int64_t)d * nj * D + d, v);
                            }
                        }
                    }
                }
            }
    }
}

template <int DIM>
__global__ void __launch_bounds__(32, 16) k_elast_off(AsmArgs A, ClArgs Cl, int64_t nclusters) {
    constexpr int NB = 4, U = 32, D = DIM, NPAIR = DIM * (DIM - 1) / 2;
    __shared__ __align__(16) unsigned char stab[U * U];
    __shared__ long long scol0[U];
    __shared__ int scoln[U];
    __shared__ double sS[U][U + 1];           // the block, for the transposed reads of the flush
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    const double G = A.c[0], lam = A.c[1];
    for (int64_t item = blockIdx.x; item < nclusters * NPAIR; item += gridDim.x) {
        const int64_t c = item / NPAIR;
        const int pr = (int)(item - c * NPAIR);
        // pairs p < q in the order (0,1), (0,2), (1,2)
        const int pi = pr == 2 ? 1 : 0, pj = pr == 0 ? 1 : 2;
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        __syncwarp();
        {
            const unsigned char *src = Cl.tab + c * (int64_t)(U * U);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab);
            for (int t = lane * 16; t < nu * U; t += 32 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + t), "l"(src + t) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (lane < nu) { scol0[lane] = Cl.col0[c * U + lane]; scoln[lane] = Cl.coln[c * U + lane]; }
        }
        double acc[NB * NB][2];
#pragma unroll
        for (int t = 0; t < NB * NB; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0; long long ho = 0; double hw = 0.0;
            if (lane < npt) {
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
            }
            // two operand buffers, loop unrolled by two groups (no register copies): the operands of the next group are in flight while
            // the sixteen MMAs of this one issue; the inverse-map words run one group further ahead
            auto map_word = [&](int q) -> unsigned {
                const int src = q + tig;
                const int g = __shfl_sync(0xffffffffu, hg, src & 31);
                return (q < npt && src < npt) ? __ldg(Cl.inv + (int64_t)g * 8 + gid) : 0xFFFFFFFFu;
            };
            auto operands = [&](int q, unsigned word, double (&a)[NB], double (&b)[NB]) {
                const int src = (q + tig) & 31;
                const long long o0 = __shfl_sync(0xffffffffu, ho, src);
                const double w = __shfl_sync(0xffffffffu, hw, src);
                const double *__restrict__ bi = A.dphi + (int64_t)pi * A.tl + o0, *__restrict__ bj = A.dphi + (int64_t)pj * A.tl + o0;
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    const unsigned pos = (word >> (8 * r)) & 255u;
                    a[r] = pos != 255u ? __ldg(bi + pos) : 0.0;
                    b[r] = pos != 255u ? w * __ldg(bj + pos) : 0.0;
                }
            };
            auto accumulate = [&](const double (&a)[NB], const double (&b)[NB]) {
#pragma unroll
                for (int r = 0; r < NB; ++r)
#pragma unroll
                    for (int s2 = 0; s2 < NB; ++s2)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(acc[r * NB + s2][0]), "+d"(acc[r * NB + s2][1]) : "d"(a[r]), "d"(b[s2]));
            };
            double aA[NB], bA[NB], aB[NB], bB[NB];
            unsigned wodd = map_word(4), weven = map_word(8);
            operands(0, map_word(0), aA, bA);
#pragma unroll 1
            for (int q = 0; q < npt; q += 8) {
                if (q + 4 < npt) operands(q + 4, wodd, aB, bB);
                wodd = map_word(q + 12);
                accumulate(aA, bA);
                if (q + 4 < npt) {
                    if (q + 8 < npt) operands(q + 8, weven, aA, bA);
                    weven = map_word(q + 16);
                    accumulate(aB, bB);
                }
            }
        }
        // park the block in shared memory: element (row node i, column node j) = S_pq^{ij}
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NB; ++r)
#pragma unroll
            for (int s2 = 0; s2 < NB; ++s2) {
                sS[8 * r + gid][8 * s2 + 2 * tig] = acc[r * NB + s2][0];
                sS[8 * r + gid][8 * s2 + 2 * tig + 1] = acc[r * NB + s2][1];
            }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        // flush: lane = column node j, rows i <= j.  K_ij[p][q] = lam S^{ij} + G S^{ji};  K_ij[q][p] = G S^{ij} + lam S^{ji}
        const int j = lane;
        if (j < nu) {
            const int nj = scoln[j];
            double *col = A.nzval + scol0[j] * (D * D);
            for (int i = 0; i <= j; ++i) {
                const int k = stab[j * U + i];
                if (k == 255) continue;
                const double sij = sS[i][j], sji = sS[j][i];
                const double vpq = lam * sij + G * sji;
                if (vpq != 0.0) atomicAdd(col + (int64_t)pj * nj * D + (int64_t)k * D + pi, vpq);          // row comp p, column comp q
                if (i < j) {
                    const double vqp = G * sij + lam * sji;
                    if (vqp != 0.0) atomicAdd(col + (int64_t)pi * nj * D + (int64_t)k * D + pj, vqp);      // row comp q, column comp p
                }
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------------
// Generic coefficient tensors (constant or per Gauss point: the Newton / Picard re-assembly path of the SVK and Navier-Stokes
// examples, SaintVenant_Kirchoff_Test.jl:28-34, Lid_Cavity.jl:35) on 32-node clusters, with the gather of k_bilinear_pts:
//     K_ab[i][j] = w sum_{al,be} B_a[al] C[(i,al),(j,be)] B_b[be]
// One warp per (cluster, i, j) holds the full 32 x 32 block of that component pair (16 accumulator tiles).  Here the k dimension of
// the MMA is the operand component (nb = 1+dim <= 4): per Gauss point, lane (gid, tig) of band r fetches all nb components of local
// node 8r+gid through the inverse map (the four lanes of a quad share the addresses), keeps B[tig] as its A fragment and forms
// T[tig] = w sum_be C[(i,tig),(j,be)] B[be] as its B fragment; tile(r,s) += A_r T_s^T, 16 MMAs per point.  One pass over the
// operands per component pair instead of D^2 re-streamed passes through a shared-memory block (k_bilinear_warp).
// --------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32, 16) k_generic_pts(AsmArgs A, ClArgs Cl, int D, int64_t nclusters) {
    constexpr int NB = 4, U = 32, nb = 1 + DIM;
    __shared__ __align__(16) unsigned char stab[U * U];
    __shared__ long long scol0[U];
    __shared__ int scoln[U];
    const int lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    const int n = D * nb, DD = D * D;
    for (int64_t item = blockIdx.x; item < nclusters * DD; item += gridDim.x) {
        const int64_t c = item / DD;
        const int comp = (int)(item - c * DD);
        const int jc = comp / D, ir = comp - jc * D;
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        __syncwarp();
        {
            const unsigned char *src = Cl.tab + c * (int64_t)(U * U);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab);
            for (int t = lane * 16; t < nu * U; t += 32 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + t), "l"(src + t) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (lane < nu) { scol0[lane] = Cl.col0[c * U + lane]; scoln[lane] = Cl.coln[c * U + lane]; }
        }
        double acc[NB * NB][2];
#pragma unroll
        for (int t = 0; t < NB * NB; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        const int crow = (ir * nb + (tig < nb ? tig : 0)) * n + jc * nb;      // my row of the tensor: (i, al = tig), columns (j, be)
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0; long long ho = 0; double hw = 0.0;
            if (lane < npt) {
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
            }
            unsigned word_n = __ldg(Cl.inv + (int64_t)__shfl_sync(0xffffffffu, hg, 0) * 8 + gid);
#pragma unroll 1
            for (int q = 0; q < npt; ++q) {
                const int g = __shfl_sync(0xffffffffu, hg, q);
                const long long o0 = __shfl_sync(0xffffffffu, ho, q);
                const double w = __shfl_sync(0xffffffffu, hw, q);
                const unsigned word = word_n;
                {
                    const int gn = __shfl_sync(0xffffffffu, hg, (q + 1) & 31);
                    if (q + 1 < npt) word_n = __ldg(Cl.inv + (int64_t)gn * 8 + gid);
                }
                double cf[nb];
                const double *__restrict__ Cg = A.coef + (int64_t)g * A.coef_stride + crow;
#pragma unroll
                for (int be = 0; be < nb; ++be) cf[be] = tig < nb ? w * __ldg(Cg + be) : 0.0;
                double af[NB], tf[NB];
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    const unsigned pos = (word >> (8 * r)) & 255u;
                    double B[nb];
                    if (pos != 255u) {
                        B[0] = __ldg(A.phi + o0 + pos);
#pragma unroll
                        for (int d = 0; d < DIM; ++d) B[1 + d] = __ldg(A.dphi + (int64_t)d * A.tl + o0 + pos);
                    } else {
#pragma unroll
                        for (int al = 0; al < nb; ++al) B[al] = 0.0;
                    }
                    double t = 0.0, a = 0.0;
#pragma unroll
                    for (int be = 0; be < nb; ++be) { t = fma(cf[be], B[be], t); if (be == tig) a = B[be]; }
                    af[r] = a; tf[r] = t;
                }
#pragma unroll
                for (int r = 0; r < NB; ++r)
#pragma unroll
                    for (int s2 = 0; s2 < NB; ++s2)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(acc[r * NB + s2][0]), "+d"(acc[r * NB + s2][1]) : "d"(af[r]), "d"(tf[s2]));
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        // ---- flush: tile (r,s) element (row node 8r+gid, column node 8s+2tig+e) -> K[(i-node, ir), (j-node, jc)] ----------------
#pragma unroll
        for (int r = 0; r < NB; ++r)
#pragma unroll
            for (int s2 = 0; s2 < NB; ++s2) {
                const int i = 8 * r + gid;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * s2 + 2 * tig + e;
                    const double v = acc[r * NB + s2][e];
                    if (i < nu && j < nu && v != 0.0) {
                        const int k = stab[j * U + i];
                        if (k != 255) atomicAdd(A.nzval + scol0[j] * DD + (int64_t)jc * scoln[j] * D + (int64_t)k * D + ir, v);
                    }
                }
            }
    }
}

// Symmetric integrands are accumulated on the upper triangle only (row <= column, node-level, D = 1);
// this fills the strict lower triangle: K[I,J] = K[J,I].
// The position of every transposed entry is found once per pattern (binary search in the partner column) ...
__global__ void k_mirror_map(int64_t nnodes, const int64_t *__restrict__ colptr, const int *__restrict__ rowval, int *__restrict__ tmap) {
    const int64_t J = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (J >= nnodes) return;
    const int64_t c0 = colptr[J], c1 = colptr[J + 1];
    for (int64_t k = c0 + lane; k < c1; k += 32) {
        const int I = rowval[k];
        int t = -1;
        if (I > J) {
            int64_t lo = colptr[I], hi = colptr[I + 1] - 1;          // find row J in column I
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < J) lo = mid + 1; else hi = mid; }
            t = (int)lo;
        }
        tmap[k] = t;
    }
}

// ... and every assembly mirrors with one gather: K[I,J] = K[J,I] for I > J.
__global__ void k_mirror_upper(int64_t e0, int64_t nnz, const int *__restrict__ tmap, double *__restrict__ nz) {
    const int64_t k = e0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // entries [e0, nnz)
    if (k >= nnz) return;
    const int t = tmap[k];
    if (t >= 0) nz[k] = nz[t];
}

// Streamed download: smallest node id held by the clusters of each of the FG_STREAM_RANGES equal cluster ranges (the node list of a
// cluster is ascending, so that is its first entry).  Columns below the suffix minimum over the ranges still to come receive nothing
// more and can leave the device.
__global__ void k_cluster_range_min(int64_t nclusters, int ucap, const int *__restrict__ nodes, const int *__restrict__ nU, int *__restrict__ rmin) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nclusters || nU[c] <= 0) return;
    int i = (int)(c * FG_STREAM_RANGES / nclusters);                     // the host launches range i over [n i / R, n (i+1) / R) (floors)
    if (nclusters * (i + 1) / FG_STREAM_RANGES <= c) ++i;
    const int m = nodes[c * (int64_t)ucap];
    const unsigned peers = __match_any_sync(__activemask(), i);          // one atomic per range present in the warp
    int best = m;
    for (unsigned rest = peers; rest; rest &= rest - 1) best = min(best, __shfl_sync(peers, m, __ffs(rest) - 1));
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(rmin + i, best);
}

// DOF-level mirror (D > 1): the accumulating kernels fill the entries with (row node < column node), any components, and, inside a
// node's own diagonal block, row component <= column component; every other entry is the transpose of one of those.
// One warp per node column; tmap is the node-level transposition map of k_mirror_map.
__global__ void k_mirror_dof(int64_t nnodes, int D, const int64_t *__restrict__ colptr, const int *__restrict__ rowval, const int *__restrict__ tmap,
                             double *__restrict__ nz) {
    const int64_t J = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (J >= nnodes) return;
    const int64_t c0 = colptr[J];
    const int nj = (int)(colptr[J + 1] - c0);
    double *colJ = nz + c0 * D * D;
    for (int k = lane; k < nj; k += 32) {
        const int I = rowval[c0 + k];
        if (I > J) {
            const int64_t t = tmap[c0 + k];                 // node-level index of (row J, column I)
            const int64_t ci = colptr[I];
            const int ni = (int)(colptr[I + 1] - ci);
            const double *colI = nz + ci * D * D;
            const int kk = (int)(t - ci);
            for (int jc = 0; jc < D; ++jc)
                for (int ir = 0; ir < D; ++ir)
                    colJ[(int64_t)jc * nj * D + (int64_t)k * D + ir] = colI[(int64_t)ir * ni * D + (int64_t)kk * D + jc];
        } else if (I == J) {
            for (int jc = 0; jc < D; ++jc)
                for (int ir = jc + 1; ir < D; ++ir)
                    colJ[(int64_t)jc * nj * D + (int64_t)k * D + ir] = colJ[(int64_t)ir * nj * D + (int64_t)k * D + jc];
        }
    }
}

static size_t warp_smem_bytes(int ucap, bool sym) {
    const size_t kcap = sym ? (size_t)ucap * (ucap + 1) / 2 : (size_t)ucap * (ucap | 1);
    return sizeof(double) * (((kcap + 1) & ~(size_t)1) + 2 * 32 * 4) + sizeof(int) * 64 + 32;
}

template <int DIM>
static int launch_cluster(fg_ctx *ctx, const AsmArgs &A, const ClArgs &Cl, int kind, int D, int64_t nclusters, bool sym) {
    const size_t smem = warp_smem_bytes(Cl.ucap, sym);
    const int per_sm = (int)std::min<size_t>(32, (227 * 1024) / (smem + 1024));
    const unsigned grid = (unsigned)std::min<int64_t>(nclusters, (int64_t)ctx->sm_count * per_sm);
    const bool mma = fg_opt(ctx, "assembly_mma", 1) != 0;
#define FG_LAUNCH_CW(K, S)                                                                                          \
    do {                                                                                                            \
        if (mma) {                                                                                                  \
            FG_CUDA(ctx, cudaFuncSetAttribute(k_bilinear_warp<DIM, K, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_bilinear_warp<DIM, K, S, true><<<grid, 32, smem, ctx->stream>>>(A, Cl, D, nclusters);                 \
        } else {                                                                                                    \
            FG_CUDA(ctx, cudaFuncSetAttribute(k_bilinear_warp<DIM, K, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_bilinear_warp<DIM, K, S, false><<<grid, 32, smem, ctx->stream>>>(A, Cl, D, nclusters);                \
        }                                                                                                           \
    } while (0)
    switch (kind) {
    case FG_OP_GENERIC: FG_LAUNCH_CW(FG_OP_GENERIC, false); break;
    case FG_OP_GRADGRAD: if (sym) FG_LAUNCH_CW(FG_OP_GRADGRAD, true); else FG_LAUNCH_CW(FG_OP_GRADGRAD, false); break;
    case FG_OP_MASS: if (sym) FG_LAUNCH_CW(FG_OP_MASS, true); else FG_LAUNCH_CW(FG_OP_MASS, false); break;
    case FG_OP_ADVECTION: FG_LAUNCH_CW(FG_OP_ADVECTION, false); break;
    case FG_OP_ELASTICITY: FG_LAUNCH_CW(FG_OP_ELASTICITY, false); break;
    }
#undef FG_LAUNCH_CW
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

static int upload_coef(fg_ctx *ctx, const fg_operator *op, int64_t ngauss, int ncoef, const double **dev, int64_t *stride) {
    *dev = nullptr; *stride = 0;
    if (op->kind != 0) return FG_OK;
    const int64_t count = op->coef_stride == 0 ? ncoef : (ngauss - 1) * op->coef_stride + ncoef;
    if (op->coef_stride != 0 && op->coef_stride < ncoef) return fg_fail(ctx, FG_E_ARG, "coef_stride %lld smaller than the tensor (%d)", (long long)op->coef_stride, ncoef);
    if (fg_opt(ctx, "coef_on_device", 0) != 0) {          // the caller computed the tensors on this GPU (a Newton loop that keeps u_h on the device)
        *dev = op->coef; *stride = op->coef_stride;
        return FG_OK;
    }
    FG_CUDA(ctx, ctx->coef.reserve(count));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->coef.p, op->coef, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    // host pointers belong to the caller for the duration of the call only: with pinned memory the copy above is truly
    // asynchronous, so it has to be complete before the entry point returns (the assembly itself stays asynchronous)
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *dev = ctx->coef.p; *stride = op->coef_stride;
    return FG_OK;
}

// The cluster index of a set, built on first use (by the assembly, or by the sparsity build that derives the pattern from it):
// 32-node blocks when single-spacing cells fit them, with the fall-backs for irregular clouds.
int fg_ensure_clusters(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    int rc;
    const int ucap = fg_default_ucap(ctx, N);                   // block size vs warps per SM vs REDs per point (kernel 3)
    if (!S.cl.ready || S.cl.want != ucap) {
        rc = fg_impl_clusters(ctx, S, ucap); if (rc) return rc;
        S.cl.want = ucap;
        int novf = 0;
        if (ucap == 32 && fg_opt(ctx, "cluster_ucap", 0) == 0) {
            // irregular clouds: when the 32-node blocks leave more than 5 % of the points to the direct kernel
            // (global atomics), 64-node blocks are the better index
            int many = 0;
            FG_CUDA(ctx, cudaMemcpyAsync(&novf, S.cl.ovf_points.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaMemcpyAsync(&many, ctx->flags.p + 10, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            S.cl.pmask_many = many;
            // irregular clouds: the cells sized for 32 lattice nodes hold more than 32 real ones.  Keep those cells and give the
            // blocks 64 slots (cells sized for 64 lattice nodes would overflow 64 just the same: measured 94 % direct points)
            if ((int64_t)novf * 20 > S.n) {
                rc = fg_impl_clusters(ctx, S, 64, 32); if (rc) return rc;
                FG_CUDA(ctx, cudaMemcpyAsync(&novf, S.cl.ovf_points.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if ((int64_t)novf * 20 > S.n) { rc = fg_impl_clusters(ctx, S, 128, 32); if (rc) return rc; novf = -1; }     // shared-memory block kernels
            }
            S.cl.stat_overflow_points = novf;                   // points outside the index (known without another round trip; -1 = not read)
        }
    }
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
    // Deferred shape functions (option "lazy_shapes"): only the streamed point-batched path below computes them range by range between
    // its assembly launches (it reads neither DOM nor shape functions of later ranges); every other path needs all of them, and DOM, now.
    const bool lazy_candidate = S.sf_lazy && ctx->stream_to != nullptr && op->D == 1 && (op->kind == FG_OP_GRADGRAD || op->kind == FG_OP_MASS) &&
                                fg_opt(ctx, "assembly_direct", 0) == 0 && fg_opt(ctx, "assembly_pts", 1) != 0 && fg_opt(ctx, "assembly_sym", 1) != 0;
    if (!lazy_candidate) {
        { int rcd = fg_ensure_shapes(ctx, S); if (rcd) return rcd; }
        { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }
    }
    AsmArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p; A.phi = S.phi.p; A.dphi = S.dphi.p; A.tl = S.total_L;
    A.colptr = P.colptr.p; A.rowval = P.rowval.p; A.nzval = P.nzval.p; A.dom_sorted = S.fem.empty() ? 1 : 0;
    for (int k = 0; k < 4; ++k) A.c[k] = op->c[k];
    int rc = upload_coef(ctx, op, S.n, (D * nb) * (D * nb), &A.coef, &A.coef_stride);
    if (rc) return rc;
    const bool direct = fg_opt(ctx, "assembly_direct", 0) != 0;          // forces the global-atomics kernel (A/B forensics)
    // symmetric scalar integrands: upper triangle only, mirrored afterwards
    const bool scalar_kind = op->kind == FG_OP_GRADGRAD || op->kind == FG_OP_MASS;
    // point-batched MMA kernel (32-node clusters): D x D blocks of these forms are s * I, accumulated once, upper triangle + mirror
    const bool elast = op->kind == FG_OP_ELASTICITY;       // D == dim (checked by the caller); symmetric at DOF level
    const bool want_pts = !direct && (scalar_kind || elast) && fg_opt(ctx, "assembly_pts", 1) != 0 && fg_opt(ctx, "assembly_sym", 1) != 0;
    bool sym = !direct && (scalar_kind || elast) && fg_opt(ctx, "assembly_sym", 1) != 0 && ((D == 1 && scalar_kind) || want_pts);
    const int *plist = nullptr;
    if (!direct) {
        rc = fg_ensure_clusters(ctx, S); if (rc) return rc;
        const bool use_pts = want_pts && S.cl.ucap == 32;
        if ((D > 1 || elast) && !use_pts) sym = false;            // the 64-node / shared-memory kernels only know the symmetric path for D = 1
        if (!S.cl.tab_ready || (!sym && S.cl.tab_upper)) { rc = fg_impl_cluster_tables(ctx, S, sym); if (rc) return rc; }
        ClArgs Cl;
        Cl.col0 = S.cl.col0.p; Cl.coln = S.cl.coln.p; Cl.tab = S.cl.tab.p;
        Cl.cstart = S.cl.grid.start.p; Cl.csorted = S.cl.sorted.p; Cl.nU = S.cl.nU.p; Cl.nodes = S.cl.nodes.p; Cl.lidx = S.cl.lidx.p; Cl.ucap = S.cl.ucap;
        Cl.ovf_extra = S.cl.ovf_points.p; Cl.inv = nullptr;
        if ((rc = fg_overflow_list_reset(ctx, S))) return rc;      // long-list points of an earlier call are not kept
        Cl.debug_skip = fg_opt(ctx, "debug_skip", 0);
        const bool generic_pts = op->kind == FG_OP_GENERIC && S.cl.ucap == 32 && fg_opt(ctx, "assembly_pts", 1) != 0 && !direct;
        if (S.sf_lazy && !(use_pts && !elast && !generic_pts)) {       // (a candidate whose cluster index did not come out as 32-node blocks)
            { int rcd = fg_ensure_shapes(ctx, S); if (rcd) return rcd; }
            { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }
        }
        if (generic_pts) {
            Cl.inv = S.cl.inv.p;
            int per_sm = 0;
            if (dim == 2) {
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_generic_pts<2>, 32, 0));
                k_generic_pts<2><<<(unsigned)std::min<int64_t>(S.cl.n * D * D, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, D, S.cl.n);
            } else {
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_generic_pts<3>, 32, 0));
                k_generic_pts<3><<<(unsigned)std::min<int64_t>(S.cl.n * D * D, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, D, S.cl.n);
            }
            FG_CHECK_LAUNCH(ctx);
        } else if (use_pts && elast) {
            Cl.inv = S.cl.inv.p;
            int per_sm = 0;
            if (dim == 2) {
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_elast_diag<2>, 32, 0));
                k_elast_diag<2><<<(unsigned)std::min<int64_t>(S.cl.n, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, S.cl.n);
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_elast_off<2>, 32, 0));
                k_elast_off<2><<<(unsigned)std::min<int64_t>(S.cl.n, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, S.cl.n);
            } else {
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_elast_diag<3>, 32, 0));
                k_elast_diag<3><<<(unsigned)std::min<int64_t>(S.cl.n, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, S.cl.n);
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_elast_off<3>, 32, 0));
                k_elast_off<3><<<(unsigned)std::min<int64_t>(S.cl.n * 3, (int64_t)ctx->sm_count * std::max(per_sm, 1)), 32, 0, ctx->stream>>>(A, Cl, S.cl.n);
            }
            ++ctx->launches;
            FG_CHECK_LAUNCH(ctx);
        } else if (use_pts) {
            Cl.inv = S.cl.inv.p;
            // Streamed download (fg_assemble_bilinear_async, D = 1): the clusters go out in FG_STREAM_RANGES launches; after each one the
            // columns no later cluster touches are mirrored and copied to the host on the copy stream, under the next launch.
            bool stream = ctx->stream_to != nullptr && D == 1 && sym && !elast && S.cl.n >= 64 * FG_STREAM_RANGES && P.nnz < ((int64_t)1 << 31);
            if (stream && (!S.cl.stream_ready || S.cl.stream_pattern != P.version)) {
                int h[FG_STREAM_RANGES + 1];
                int *rmin = ctx->flags.p + 40;
                FG_CUDA(ctx, cudaMemsetAsync(rmin, 0x7f, sizeof(int) * FG_STREAM_RANGES, ctx->stream));
                k_cluster_range_min<<<(unsigned)((S.cl.n + 255) / 256), 256, 0, ctx->stream>>>(S.cl.n, S.cl.ucap, S.cl.nodes.p, S.cl.nU.p, rmin);
                FG_CHECK_LAUNCH(ctx);
                FG_CUDA(ctx, cudaMemcpyAsync(h, rmin, sizeof(int) * FG_STREAM_RANGES, cudaMemcpyDeviceToHost, ctx->stream));
                FG_CUDA(ctx, cudaMemcpyAsync(h + FG_STREAM_RANGES, S.cl.ovf_points.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                S.cl.stream_novf = h[FG_STREAM_RANGES];
                S.cl.stream_cols.assign(FG_STREAM_RANGES + 1, 0);
                S.cl.stream_cols[FG_STREAM_RANGES] = N.n;
                int64_t suf = N.n;                                       // suffix minimum over the ranges still to run
                for (int i = FG_STREAM_RANGES - 1; i >= 1; --i) { suf = std::min<int64_t>(suf, h[i]); S.cl.stream_cols[i] = suf; }
                for (int i = 1; i <= FG_STREAM_RANGES; ++i) S.cl.stream_cols[i] = std::max(S.cl.stream_cols[i], S.cl.stream_cols[i - 1]);
                S.cl.stream_ent.assign(FG_STREAM_RANGES + 1, 0);
                rc = fg_colptr_at(ctx, P, S.cl.stream_cols.data(), FG_STREAM_RANGES + 1, S.cl.stream_ent.data()); if (rc) return rc;
                S.cl.stream_ready = true; S.cl.stream_pattern = P.version;
            }
            if (stream && S.cl.stream_novf != 0) stream = false;       // points on the direct kernel scatter anywhere: no column is final early
            if (S.sf_lazy && !stream) {                                // deferred shape functions without the pipeline: all of them now
                { int rcd = fg_ensure_shapes(ctx, S); if (rcd) return rcd; }
                { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }
            }
            if (stream && !P.tmap_ready) {
                FG_CUDA(ctx, P.tmap.reserve(P.nnz));
                k_mirror_map<<<(unsigned)((N.n * 32 + 255) / 256), 256, 0, ctx->stream>>>(N.n, P.colptr.p, P.rowval.p, P.tmap.p);
                FG_CHECK_LAUNCH(ctx);
                P.tmap_ready = true;
            }
#define FG_LAUNCH_PTS(DIM_, KIND_)                                                                                   \
            do {                                                                                                     \
                int per_sm = 0;                                                                                      \
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bilinear_pts<DIM_, KIND_>, 32, 0)); \
                const unsigned grid = (unsigned)std::min<int64_t>(c1 - c0, (int64_t)ctx->sm_count * std::max(per_sm, 1)); \
                k_bilinear_pts<DIM_, KIND_><<<grid, 32, 0, ctx->stream>>>(A, R, D, c1 - c0);                         \
            } while (0)
            const int nranges = stream ? FG_STREAM_RANGES : 1;
            for (int i = 0; i < nranges; ++i) {
                const int64_t c0 = S.cl.n * i / nranges, c1 = S.cl.n * (i + 1) / nranges;
                if (S.sf_lazy) {                                         // shape functions of the search clusters this range reads, not computed yet
                    const int64_t need = fg_search_clusters_below(ctx, S, c1);
                    if (need > S.sf_done) { rc = fg_impl_imls_range(ctx, S, S.sf_done, need); if (rc) return rc; S.sf_done = need; }
                    if (S.sf_done >= S.nbcl.n) { S.sf_lazy = false; S.dom_pending = false; }
                }
                ClArgs R = Cl;                                           // the kernel indexes clusters from 0: shift the per-cluster arrays
                R.nU += c0; R.cstart += c0; R.col0 += c0 * 32; R.coln += c0 * 32; R.tab += c0 * (32 * 32);
                if (dim == 2) {
                    if (op->kind == FG_OP_GRADGRAD) FG_LAUNCH_PTS(2, FG_OP_GRADGRAD); else FG_LAUNCH_PTS(2, FG_OP_MASS);
                } else {
                    if (op->kind == FG_OP_GRADGRAD) FG_LAUNCH_PTS(3, FG_OP_GRADGRAD); else FG_LAUNCH_PTS(3, FG_OP_MASS);
                }
                FG_CHECK_LAUNCH(ctx);
                if (!stream) break;
                const int64_t e0 = S.cl.stream_ent[i], e1 = S.cl.stream_ent[i + 1];
                if (e1 > e0) {
                    k_mirror_upper<<<(unsigned)((e1 - e0 + 255) / 256), 256, 0, ctx->stream>>>(e0, e1, P.tmap.p, P.nzval.p);
                    FG_CHECK_LAUNCH(ctx);
                    FG_CUDA(ctx, cudaEventRecord(ctx->copy_event, ctx->stream));
                    FG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
                    FG_CUDA(ctx, cudaMemcpyAsync(ctx->stream_to + e0, P.nzval.p + e0, sizeof(double) * (e1 - e0), cudaMemcpyDeviceToHost, ctx->copy_stream));
                }
            }
#undef FG_LAUNCH_PTS
            if (stream) {
                { int rcd = fg_ensure_shapes(ctx, S); if (rcd) return rcd; }      // (nothing left unless the last range ended short of the last search cluster)
                ctx->streamed = true; ctx->copy_pending = true; return FG_OK;
            }
        } else if (sym && (Cl.ucap == 64 || Cl.ucap == 32) && fg_opt(ctx, "assembly_reg", 1) != 0) {
            if ((rc = fg_ensure_lidx(ctx, S))) return rc;
            // register-block kernel: 36 MMA accumulator tiles per warp, no shared-memory block
            // persistent grid = exactly the number of resident warps (a larger grid would leave a tail wave)
#define FG_LAUNCH_REG_NB(DIM_, KIND_, NB_)                                                                           \
            do {                                                                                                     \
                int per_sm = 0;                                                                                      \
                FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bilinear_reg<DIM_, KIND_, NB_>, 32, 0)); \
                const unsigned grid = (unsigned)std::min<int64_t>(S.cl.n, (int64_t)ctx->sm_count * std::max(per_sm, 1)); \
                k_bilinear_reg<DIM_, KIND_, NB_><<<grid, 32, 0, ctx->stream>>>(A, Cl, S.cl.n);                       \
            } while (0)
#define FG_LAUNCH_REG(DIM_, KIND_) do { if (Cl.ucap == 32) FG_LAUNCH_REG_NB(DIM_, KIND_, 4); else FG_LAUNCH_REG_NB(DIM_, KIND_, 8); } while (0)
            if (dim == 2) {
                if (op->kind == FG_OP_GRADGRAD) FG_LAUNCH_REG(2, FG_OP_GRADGRAD); else FG_LAUNCH_REG(2, FG_OP_MASS);
            } else {
                if (op->kind == FG_OP_GRADGRAD) FG_LAUNCH_REG(3, FG_OP_GRADGRAD); else FG_LAUNCH_REG(3, FG_OP_MASS);
            }
#undef FG_LAUNCH_REG_NB
#undef FG_LAUNCH_REG
            FG_CHECK_LAUNCH(ctx);
        } else {
            if ((rc = fg_ensure_lidx(ctx, S))) return rc;
            rc = dim == 2 ? launch_cluster<2>(ctx, A, Cl, op->kind, D, S.cl.n, sym) : launch_cluster<3>(ctx, A, Cl, op->kind, D, S.cl.n, sym);
            if (rc) return rc;
        }
        plist = S.cl.ovf_points.p;                     // clusters / points that did not fit: direct path
    }
    if (dim == 2) {
        if (D == 1) rc = launch_kind<2, 1>(ctx, A, op->kind, S.n, plist);
        else if (D == 2) rc = launch_kind<2, 2>(ctx, A, op->kind, S.n, plist);
        else rc = launch_kind<2, 3>(ctx, A, op->kind, S.n, plist);
    } else {
        if (D == 1) rc = launch_kind<3, 1>(ctx, A, op->kind, S.n, plist);
        else if (D == 2) rc = launch_kind<3, 2>(ctx, A, op->kind, S.n, plist);
        else rc = launch_kind<3, 3>(ctx, A, op->kind, S.n, plist);
    }
    if (rc) return rc;
    if (sym) {
        if (!P.tmap_ready) {
            if (P.nnz >= ((int64_t)1 << 31)) return fg_fail(ctx, FG_E_CAPACITY, "fg_assemble_bilinear: more than 2^31 node-level non-zeros");
            FG_CUDA(ctx, P.tmap.reserve(P.nnz));
            k_mirror_map<<<(unsigned)((N.n * 32 + 255) / 256), 256, 0, ctx->stream>>>(N.n, P.colptr.p, P.rowval.p, P.tmap.p);
            FG_CHECK_LAUNCH(ctx);
            P.tmap_ready = true;
        }
        if (D == 1) k_mirror_upper<<<(unsigned)((P.nnz + 255) / 256), 256, 0, ctx->stream>>>(0, P.nnz, P.tmap.p, P.nzval.p);
        else k_mirror_dof<<<(unsigned)((N.n * 32 + 255) / 256), 256, 0, ctx->stream>>>(N.n, D, P.colptr.p, P.rowval.p, P.tmap.p, P.nzval.p);
        FG_CHECK_LAUNCH(ctx);
    }
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// linear forms: F[(node-1)*D + i] += (sum_al c[i*nb+al] B_a[al]) * (jac*w)
// --------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_linear(AsmArgs A, int D, int kind, double *__restrict__ F, const int *__restrict__ plist) {
    constexpr int nb = 1 + DIM;
    const int lane = threadIdx.x & 31;
    const int64_t count = plist ? (int64_t)plist[0] : A.ngauss;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; it < count; it += stride) {
        const int64_t g = plist ? plist[1 + it] : it;
        const double wj = A.gs[(DIM + 1) * A.ngauss + g] * A.gs[DIM * A.ngauss + g];
        const double *C = kind == FG_LIN_GENERIC ? A.coef + g * A.coef_stride : nullptr;
        const int64_t o0 = A.off[g], o1 = A.off[g + 1];
        for (int64_t k = o0 + lane; k < o1; k += 32) {
            double B[nb];
            B[0] = A.phi[k];
#pragma unroll
            for (int d = 0; d < DIM; ++d) B[1 + d] = A.dphi[d * A.tl + k];
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
}

// Cluster path of the linear forms (used when the set's cluster index exists): one warp per cluster keeps the
// cluster's |U| x D slice of F in shared memory; the lanes of a point own distinct local indices, so the updates
// are plain read-modify-writes, and each cluster issues |U| x D REDs instead of one per neighbour-list entry.
template <int DIM>
__global__ void __launch_bounds__(128) k_linear_cluster(AsmArgs A, ClArgs Cl, int D, int kind, double *__restrict__ F, int64_t nclusters) {
    constexpr int nb = 1 + DIM;
    extern __shared__ double s_f[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double *sF = s_f + (size_t)wid * Cl.ucap * D;
    for (int64_t c = blockIdx.x * (int64_t)nw + wid; c < nclusters; c += (int64_t)gridDim.x * nw) {
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        for (int t = lane; t < nu * D; t += 32) sF[t] = 0.0;
        __syncwarp();
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0, hL = 0; int64_t ho = 0; double hw = 0.0;
            if (lane < npt) {       // headers of 32 points at a time
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hL = (int)(A.off[hg + 1] - ho);
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
            }
            for (int q = 0; q < npt; ++q) {
                const int g = __shfl_sync(0xffffffffu, hg, q), L = __shfl_sync(0xffffffffu, hL, q);
                const int64_t o0 = __shfl_sync(0xffffffffu, ho, q);
                const double wj = __shfl_sync(0xffffffffu, hw, q);
                const double *C = kind == FG_LIN_GENERIC ? A.coef + g * A.coef_stride : nullptr;
                for (int k0 = 0; k0 < L; k0 += 32) {
                    const int k = k0 + lane;
                    if (k < L) {
                        const int la = Cl.lidx[o0 + k];
                        double B[nb];
                        B[0] = A.phi[o0 + k];
                        if (kind == FG_LIN_GENERIC) {
#pragma unroll
                            for (int d = 0; d < DIM; ++d) B[1 + d] = A.dphi[d * A.tl + o0 + k];
                        }
                        for (int i = 0; i < D; ++i) {
                            double s;
                            if (kind == FG_LIN_SOURCE) s = A.c[i] * B[0];
                            else {
                                s = 0.0;
#pragma unroll
                                for (int al = 0; al < nb; ++al) s += C[i * nb + al] * B[al];
                            }
                            sF[la * D + i] += s * wj;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        for (int t = lane; t < nu * D; t += 32) {
            const double v = sF[t];
            if (v != 0.0) atomicAdd(F + (int64_t)Cl.nodes[c * Cl.ucap + t / D] * D + t % D, v);
        }
        __syncwarp();
    }
}

// Source terms on 32-node clusters through the gather map of k_bilinear_pts: f_U = sum_g w_g phi(g) in local node order.  Lane
// (gid, tig) holds local node 8r+gid (r = 0..3) at point tig of the current group of four -- one map word and four gathers per
// group instead of one warp pass per Gauss point with shared-memory read-modify-writes (k_linear_cluster: 4.2 ms at 2.7e7 points,
// on the critical path of the end-to-end step once the download of K hides behind the assembly).  The four point lanes are summed
// with two shuffles at the end; lane (gid, tig) flushes node 8 tig + gid: 32 REDs per cluster and component.
template <int DIM>
__global__ void __launch_bounds__(128) k_linear_pts(AsmArgs A, ClArgs Cl, int D, double *__restrict__ F, int64_t nclusters) {
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); c < nclusters; c += nw) {
        const int nu = Cl.nU[c];
        if (nu <= 0) continue;
        const int p0 = Cl.cstart[c], p1 = Cl.cstart[c + 1];
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int pb = p0; pb < p1; pb += 32) {
            const int npt = min(32, p1 - pb);
            int hg = 0; long long ho = 0; double hw = 0.0;
            if (lane < npt) {
                hg = Cl.csorted[pb + lane];
                ho = A.off[hg];
                hw = A.gs[(DIM + 1) * A.ngauss + hg] * A.gs[DIM * A.ngauss + hg];
            }
            // all eight map words of the batch first, then the gathers: the kernel waits on memory latency (ncu: long scoreboard on 32 of
            // 33 stalled warps with two groups in flight), so every independent load is issued before the first one is used
            unsigned words[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int src = 4 * k + tig;
                const int g = __shfl_sync(0xffffffffu, hg, src);
                words[k] = src < npt ? __ldg(Cl.inv + (int64_t)g * 8 + gid) : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int src = 4 * k + tig;
                const long long o0 = __shfl_sync(0xffffffffu, ho, src);
                const double w = __shfl_sync(0xffffffffu, hw, src);
                double v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned pos = (words[k] >> (8 * r)) & 255u;
                    v[r] = pos != 255u ? __ldg(A.phi + o0 + pos) : 0.0;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = fma(w, v[r], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 2);
        }
        const double v = tig == 0 ? acc[0] : (tig == 1 ? acc[1] : (tig == 2 ? acc[2] : acc[3]));
        const int ln = 8 * tig + gid;
        if (ln < nu && v != 0.0) {
            const int64_t node = Cl.nodes[c * 32 + ln];
            for (int i = 0; i < D; ++i) atomicAdd(F + node * D + i, A.c[i] * v);
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
    { int rcd = fg_ensure_dom(ctx, S); if (rcd) return rcd; }
    AsmArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.gs = S.gs.p; A.off = S.off.p; A.dom = S.dom.p; A.phi = S.phi.p; A.dphi = S.dphi.p; A.tl = S.total_L;
    A.colptr = nullptr; A.rowval = nullptr; A.nzval = nullptr; A.dom_sorted = S.fem.empty() ? 1 : 0;
    for (int k = 0; k < 4; ++k) A.c[k] = op->c[k];
    int rc = upload_coef(ctx, op, S.n, D * nb, &A.coef, &A.coef_stride);
    if (rc) return rc;
    const int *plist = nullptr;
    int64_t npoints = S.n;
    if (S.cl.ready && fg_opt(ctx, "assembly_direct", 0) == 0) {
        // the cluster index of the bilinear forms exists: per-cluster accumulation, the index build's overflow points directly
        ClArgs Cl;
        Cl.col0 = nullptr; Cl.coln = nullptr; Cl.tab = nullptr; Cl.ovf_extra = nullptr; Cl.debug_skip = 0; Cl.inv = nullptr;
        Cl.cstart = S.cl.grid.start.p; Cl.csorted = S.cl.sorted.p; Cl.nU = S.cl.nU.p; Cl.nodes = S.cl.nodes.p; Cl.lidx = S.cl.lidx.p; Cl.ucap = S.cl.ucap;
        if ((rc = fg_overflow_list_reset(ctx, S))) return rc;
        const size_t smem = sizeof(double) * 4 * (size_t)Cl.ucap * D;
        const unsigned cgrid = (unsigned)std::min<int64_t>((S.cl.n + 3) / 4, (int64_t)ctx->sm_count * 16);
        if (op->kind == FG_LIN_SOURCE && Cl.ucap == 32 && S.cl.inv.p && fg_opt(ctx, "linear_pts", 1) != 0) {
            Cl.inv = S.cl.inv.p;
            // persistent grid = exactly the resident blocks (16 per SM left a second, quarter-filled wave)
            int per_sm = 0;
            if (dim == 2) FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linear_pts<2>, 128, 0));
            else FG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linear_pts<3>, 128, 0));
            const unsigned pgrid = (unsigned)std::min<int64_t>((S.cl.n + 3) / 4, (int64_t)ctx->sm_count * std::max(per_sm, 1));
            if (dim == 2) k_linear_pts<2><<<pgrid, 128, 0, ctx->stream>>>(A, Cl, D, S.fvec.p, S.cl.n);
            else k_linear_pts<3><<<pgrid, 128, 0, ctx->stream>>>(A, Cl, D, S.fvec.p, S.cl.n);
        } else {
            if ((rc = fg_ensure_lidx(ctx, S))) return rc;
            if (dim == 2) k_linear_cluster<2><<<cgrid, 128, smem, ctx->stream>>>(A, Cl, D, op->kind, S.fvec.p, S.cl.n);
            else k_linear_cluster<3><<<cgrid, 128, smem, ctx->stream>>>(A, Cl, D, op->kind, S.fvec.p, S.cl.n);
        }
        FG_CHECK_LAUNCH(ctx);
        plist = S.cl.ovf_points.p;
    }
    // with a point list the count lives on the device: a fixed grid strides over it
    const unsigned grid = plist ? (unsigned)(ctx->sm_count * 8) : (unsigned)((npoints * 32 + 127) / 128);
    if (dim == 2) k_linear<2><<<grid, 128, 0, ctx->stream>>>(A, D, op->kind, S.fvec.p, plist);
    else k_linear<3><<<grid, 128, 0, ctx->stream>>>(A, D, op->kind, S.fvec.p, plist);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

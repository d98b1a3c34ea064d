// fg_pattern.cu -- sparsity pattern of the assembled operators, built without COO triplets.
//
// The reference pushes sum L^2 (row,col,val) triplets and lets sparse() sort them
// (src/Assembling_Operators.jl:14-22, :92-101).  Entry (i,j) is structurally present iff some
// Gauss point g has both i and j in its neighbour list, i.e. iff g lies in supp(i) and supp(j).
// One warp per column j: candidate rows i are the nodes whose support box overlaps supp(j)
// (node bins); each lane then looks for a witness Gauss point in the bins covering
// supp(i) /\ supp(j), testing the reference's own FP64 inclusion predicate on both nodes, so
// the pattern is bit-identical to the union of DOM_g x DOM_g.  Rows are rank-sorted in shared
// memory (ascending, as sparse() stores them); count pass -> scan -> fill pass.
#include "fg_internal.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <math.h>

struct PatArgs {
    int64_t nnodes, ngauss;
    const double *x, *dm;          // original order
    const double *xs, *dms;        // bin order
    const int *sorted_id, *nbin_start;
    const double *gxs;             // Gauss-point coordinates in Gauss-bin order: gxs[d*ngauss + k]
    const int *gbin_start;
    BinGeom geom;                  // node bins
    BinGeom ggeom;                 // Gauss-point bins (finer: the witness search visits few points per bin)
    double reach[FG_MAXDIM];
    // node subsets of the witness set (fg_set_node_subset): a Gauss point only witnesses pairs of nodes of ITS subset
    const unsigned *masks;         // nsub masks of mask_words words, or nullptr
    const unsigned char *gsub;     // subset index of every Gauss point in Gauss-bin order, or nullptr (all 0)
    int64_t mask_words;
};

__device__ __forceinline__ bool pat_mask_bit(const PatArgs &A, int sub, int id) {
    return (__ldg(A.masks + sub * A.mask_words + (id >> 5)) >> (id & 31)) & 1u;
}

#define PAT_WARPS 4
#define PAT_CAP_MAX 6144 // candidate rows per column that fit in shared memory (2 lists x 4 warps)

template <int DIM, int SHAPE>
__device__ __forceinline__ bool inside_node(const PatArgs &A, const double (&xn)[DIM], const double (&dn)[DIM], const double (&g)[DIM]) {
    if (SHAPE == FG_SHAPE_RECTANGULAR) {
        bool in = true;
#pragma unroll
        for (int d = 0; d < DIM; ++d) in = in & fg_inside_1d(xn[d], g[d], dn[d]);
        return in;
    } else {
        double dist_sq = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { const double t = __dsub_rn(xn[d], g[d]); dist_sq = __dadd_rn(dist_sq, __dmul_rn(t, t)); }
        return dist_sq <= __dmul_rn(dn[0], dn[0]);
    }
}

// Is there a Gauss point inside supp(i) /\ supp(j)?
template <int DIM, int SHAPE>
__device__ bool witness(const PatArgs &A, const double (&xi)[DIM], const double (&di)[DIM], const double (&xj)[DIM], const double (&dj)[DIM],
                        int idi, int idj) {
    int lo[FG_MAXDIM] = {0, 0, 0}, hi[FG_MAXDIM] = {0, 0, 0}, mid[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double ri = SHAPE == FG_SHAPE_RECTANGULAR ? di[d] : di[0], rj = SHAPE == FG_SHAPE_RECTANGULAR ? dj[d] : dj[0];
        const double slack = 1e-9 * A.reach[d];
        const double l = fmax(xi[d] - ri, xj[d] - rj) - slack, h = fmin(xi[d] + ri, xj[d] + rj) + slack;
        if (l > h) return false;
        lo[d] = fg_bin_coord(l, A.ggeom.origin[d], A.ggeom.inv_size[d], A.ggeom.nb[d]);
        hi[d] = fg_bin_coord(h, A.ggeom.origin[d], A.ggeom.inv_size[d], A.ggeom.nb[d]);
        mid[d] = fg_bin_coord(0.5 * (l + h), A.ggeom.origin[d], A.ggeom.inv_size[d], A.ggeom.nb[d]);
    }
    auto scan = [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k) {
            double g[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) g[d] = A.gxs[d * A.ngauss + k];
            if (inside_node<DIM, SHAPE>(A, xi, di, g) && inside_node<DIM, SHAPE>(A, xj, dj, g)) {
                if (!A.masks) return true;
                const int sub = A.gsub ? (int)A.gsub[k] : 0;
                if (pat_mask_bit(A, sub, idi) && pat_mask_bit(A, sub, idj)) return true;
            }
        }
        return false;
    };
    {   // the bin under the centre of the box first: it is the one most likely to lie inside both supports
        const int row = (mid[2] * A.ggeom.nb[1] + mid[1]) * A.ggeom.nb[0] + mid[0];
        if (scan(A.gbin_start[row], A.gbin_start[row + 1])) return true;
    }
    for (int bz = lo[2]; bz <= hi[2]; ++bz)
        for (int by = lo[1]; by <= hi[1]; ++by) {
            const int row = (bz * A.ggeom.nb[1] + by) * A.ggeom.nb[0];
            if (scan(A.gbin_start[row + lo[0]], A.gbin_start[row + hi[0] + 1])) return true;
        }
    return false;
}

template <int DIM, int SHAPE, bool FILL>
__global__ void __launch_bounds__(PAT_WARPS * 32) k_pattern(PatArgs A, int *__restrict__ cnt, const int64_t *__restrict__ colptr,
                                                            int *__restrict__ rowval, int *__restrict__ overflow, int PAT_CAP,
                                                            int *__restrict__ tmp, int tcap) {
    extern __shared__ int s_lists[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t j = blockIdx.x * (int64_t)PAT_WARPS + wid;
    if (j >= A.nnodes) return;
    int *cand = s_lists + (size_t)(2 * wid) * PAT_CAP, *rows = cand + PAT_CAP;
    constexpr int NDM = SHAPE == FG_SHAPE_RECTANGULAR ? DIM : 1;
    double xj[DIM], dj[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { xj[d] = A.x[d * A.nnodes + j]; dj[d] = A.dm[(d < NDM ? d : 0) * A.nnodes + j]; }
    // phase 1: candidate rows = nodes whose support can overlap supp(j)
    int lo[FG_MAXDIM] = {0, 0, 0}, hi[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double r = A.reach[d] + dj[d] * (1.0 + 1e-9);
        lo[d] = fg_bin_coord(xj[d] - r, A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
        hi[d] = fg_bin_coord(xj[d] + r, A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
    }
    int ncand = 0;
    // the (y,z) bin rows of the box: their node ranges are fetched 32 rows at a time (one per lane), then the
    // concatenated ranges are walked 32 nodes at a time -- a few dependent global-load rounds per column
    const int nry = hi[1] - lo[1] + 1, nrows = nry * (hi[2] - lo[2] + 1);
    for (int r0 = 0; r0 < nrows; r0 += 32) {
        const int r = r0 + lane;
        int k0 = 0, len = 0;
        if (r < nrows) {
            const int by = lo[1] + r % nry, bz = lo[2] + r / nry;
            const int row = (bz * A.geom.nb[1] + by) * A.geom.nb[0];
            k0 = A.nbin_start[row + lo[0]];
            len = A.nbin_start[row + hi[0] + 1] - k0;
        }
        int incl = len;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - len;
        for (int i0 = 0; i0 < tot; i0 += 32) {
            const int i = i0 + lane;
            int rr = 0;                                   // row of flat index i: the last lane whose exclusive prefix is <= i
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int cand_r = rr + step;
                const int ex = __shfl_sync(0xffffffffu, excl, cand_r < 32 ? cand_r : 31);
                if (cand_r < 32 && ex <= i) rr = cand_r;
            }
            const int rk0 = __shfl_sync(0xffffffffu, k0, rr), rex = __shfl_sync(0xffffffffu, excl, rr);
            const int k = rk0 + (i - rex);
            bool keep = false;
            if (i < tot) {
                keep = true;
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    const double di = A.dms[(d < NDM ? d : 0) * A.nnodes + k];
                    keep = keep & (fabs(A.xs[d * A.nnodes + k] - xj[d]) <= (di + dj[d]) * (1.0 + 1e-9));
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) { const int pos = ncand + __popc(m & ((1u << lane) - 1)); if (pos < PAT_CAP) cand[pos] = k; }
            ncand += __popc(m);
        }
    }
    if (ncand > PAT_CAP) { if (lane == 0) atomicMax(overflow, ncand); return; }
    __syncwarp();
    // phase 2: witness search, one candidate per lane
    int nrow = 0;
    for (int cb = 0; cb < ncand; cb += 32) {
        const int c = cb + lane;
        bool found = false; int id = -1;
        if (c < ncand) {
            const int k = cand[c];
            double xi[DIM], di[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) { xi[d] = A.xs[d * A.nnodes + k]; di[d] = A.dms[(d < NDM ? d : 0) * A.nnodes + k]; }
            id = A.sorted_id[k];
            found = witness<DIM, SHAPE>(A, xi, di, xj, dj, id, (int)j);
        }
        const unsigned m = __ballot_sync(0xffffffffu, found);
        if (found) rows[nrow + __popc(m & ((1u << lane) - 1))] = id;
        nrow += __popc(m);
    }
    __syncwarp();
    if (!FILL) {
        if (lane == 0) cnt[j] = nrow;
        // the sorted rows are parked in a fixed-stride scratch so that the fill pass is a plain compaction
        // (no second witness search); a column longer than the stride raises overflow[2] and the caller falls back
        if (tmp == nullptr) return;
        if (nrow > tcap) { if (lane == 0) atomicMax(overflow + 2, nrow); return; }
        rowval = tmp;
    }
    // phase 3: rank sort (ids are distinct) and store ascending
    const int64_t c0 = FILL ? colptr[j] : j * (int64_t)tcap;
    for (int e = lane; e < nrow; e += 32) {
        const int v = rows[e];
        int rank = 0;
        for (int t = 0; t < nrow; ++t) rank += rows[t] < v;
        rowval[c0 + rank] = v;
    }
}

__global__ void k_pattern_compact(int64_t nnodes, const int64_t *__restrict__ colptr, const int *__restrict__ tmp, int tcap, int *__restrict__ rowval) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (j >= nnodes) return;
    const int64_t c0 = colptr[j];
    const int n = (int)(colptr[j + 1] - c0);
    for (int k = lane; k < n; k += 32) rowval[c0 + k] = tmp[j * (int64_t)tcap + k];
}

// Gauss-point coordinates in Gauss-bin order
__global__ void k_gather_gauss(int dim, int64_t n, const int *__restrict__ sorted, const double *__restrict__ gs, double *__restrict__ dst,
                               const unsigned char *__restrict__ psub, unsigned char *__restrict__ gsub) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = sorted[k];
    for (int d = 0; d < dim; ++d) dst[d * n + k] = gs[d * n + i];
    if (psub) gsub[k] = psub[i];
}

template <int DIM, int SHAPE>
static int run(fg_ctx *ctx, GaussSet &S, const PatArgs &A) {
    Pattern &P = S.pat;
    const int64_t nn = A.nnodes;
    DevBuf<int> &cnt = ctx->pat_cnt;
    FG_CUDA(ctx, cnt.reserve(nn));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p, 0, 4 * sizeof(int), ctx->stream));
    const unsigned grid = (unsigned)((nn + PAT_WARPS - 1) / PAT_WARPS);
    int rc = FG_OK;
    int cap = 256;          // candidate rows per column held in shared memory; a longer list retries with the size needed
    const int tcap = 192;
    DevBuf<int> &tmp = ctx->pat_rows;
    const bool one_pass = tmp.reserve((size_t)nn * tcap) == cudaSuccess;     // scratch for the single-pass build (optional)
    if (!one_pass) cudaGetLastError();
    size_t smem = sizeof(int) * 2 * PAT_WARPS * cap;
    do {
        k_pattern<DIM, SHAPE, false><<<grid, PAT_WARPS * 32, smem, ctx->stream>>>(A, cnt.p, nullptr, nullptr, ctx->flags.p, cap, one_pass ? tmp.p : nullptr, tcap);
        ++ctx->launches;
        if (cudaGetLastError() != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "pattern count launch failed"); break; }
        {   // candidate lists longer than the shared-memory lists: retry once with the size that was needed
            int need = 0;
            cudaMemcpyAsync(&need, ctx->flags.p, sizeof need, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "pattern count failed"); break; }
            if (need > cap) {
                if (need > PAT_CAP_MAX) { rc = fg_fail(ctx, FG_E_CAPACITY, "fg_pattern: a column has %d candidate rows; the limit is %d", need, PAT_CAP_MAX); break; }
                cap = need; smem = sizeof(int) * 2 * PAT_WARPS * cap;
                cudaFuncSetAttribute(k_pattern<DIM, SHAPE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(k_pattern<DIM, SHAPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaMemsetAsync(ctx->flags.p, 0, 4 * sizeof(int), ctx->stream);
                k_pattern<DIM, SHAPE, false><<<grid, PAT_WARPS * 32, smem, ctx->stream>>>(A, cnt.p, nullptr, nullptr, ctx->flags.p, cap, one_pass ? tmp.p : nullptr, tcap);
                if (cudaGetLastError() != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "pattern count launch failed"); break; }
            }
        }
        if ((rc = fg_exclusive_scan_i32_to_i64(ctx, cnt.p, P.colptr.p, nn))) break;
        size_t bytes = 0;
        cub::DeviceReduce::Max(nullptr, bytes, cnt.p, ctx->flags.p + 1, (int)nn, ctx->stream);
        if (ctx->temp.reserve(bytes) != cudaSuccess) { rc = fg_fail(ctx, FG_E_OOM, "pattern: temp"); break; }
        cub::DeviceReduce::Max(ctx->temp.p, bytes, cnt.p, ctx->flags.p + 1, (int)nn, ctx->stream);
        int h[3] = {0, 0, 0}; int64_t nnz = 0;
        cudaMemcpyAsync(h, ctx->flags.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(&nnz, P.colptr.p + nn, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "pattern count: %s", cudaGetErrorString(e)); break; }
        if (h[0] > cap) { rc = fg_fail(ctx, FG_E_CAPACITY, "fg_pattern: a column has %d candidate rows; the limit is %d", h[0], cap); break; }
        P.nnz = nnz; P.max_col = h[1];
        if (P.rowval.reserve(nnz) != cudaSuccess) { rc = fg_fail(ctx, FG_E_OOM, "pattern: rowval (%lld entries)", (long long)nnz); break; }
        if (one_pass && h[2] == 0)
            k_pattern_compact<<<(unsigned)((nn * 32 + 255) / 256), 256, 0, ctx->stream>>>(nn, P.colptr.p, tmp.p, tcap, P.rowval.p);
        else
            k_pattern<DIM, SHAPE, true><<<grid, PAT_WARPS * 32, smem, ctx->stream>>>(A, nullptr, P.colptr.p, P.rowval.p, ctx->flags.p, cap, nullptr, 0);
        ++ctx->launches;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "pattern fill: %s", cudaGetErrorString(e)); break; }
    } while (0);
    return rc;
}

// --------------------------------------------------------------------------------------------
// Sparsity from the assembly's cluster index (default when the set's own search ran before fg_pattern and every Gauss point sits in
// a 32-node cluster).  Entry (i,j) is present iff some Gauss point holds both nodes in its list; the points of one cluster draw their
// lists from the cluster's union U (<= 32 sorted ids), and the gather map of the point-batched assembly says which local nodes each
// point holds.  k_cluster_pmask turns that into one 32-bit POINT mask per local node (bit q = the cluster's q-th point holds the node);
// pair (a,b) of a cluster is present iff P_a & P_b != 0.  k_pattern_clusters, one warp per column j: the clusters whose cells meet
// supp(j) (3^dim or 4^dim cells), lane = local node: find j in U by ballot, broadcast P_j, and insert U[b] with P_b & P_j != 0 into a
// small hash set; rank-sort, park in the fixed-stride scratch of the witness kernel (same scan / compaction afterwards).  No geometry
// is evaluated: the result is the union of DOM x DOM by construction, and the kernel reads ~6 KB of L2-resident index per column where
// the witness search walks Gauss points lane by lane (5.4 ms + 1.7 ms of Gauss-point binning at 10^6 nodes).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_cluster_pmask(int64_t nclusters, const int *__restrict__ cstart, const int *__restrict__ csorted,
                                                       const int *__restrict__ nU, const unsigned *__restrict__ inv, unsigned *__restrict__ pm,
                                                       int *__restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t c = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nclusters) return;
    const int nu = nU[c];
    unsigned P = 0u;
    if (nu > 0) {
        const int p0 = cstart[c], p1 = cstart[c + 1];
        if (p1 - p0 > 32) { if (lane == 0) atomicExch(flag, 1); }            // more points than mask bits: the caller falls back
        const int npt = min(32, p1 - p0);
        const int g_mine = lane < npt ? csorted[p0 + lane] : 0;
        for (int q = 0; q < npt; ++q) {
            const int g = __shfl_sync(0xffffffffu, g_mine, q);
            const unsigned word = __ldg(inv + (int64_t)g * 8 + (lane & 7));        // byte r of word gid = position of local node 8r+gid, 255 = absent
            if (((word >> (8 * (lane >> 3))) & 255u) != 255u) P |= 1u << q;
        }
    } else if (nu < 0) { if (lane == 0) atomicExch(flag, 1); }               // a cluster outside the index
    pm[c * 32 + lane] = P;
}

#define PATC_HASH 512
template <int DIM>
__global__ void __launch_bounds__(PAT_WARPS * 32) k_pattern_clusters(int64_t nnodes, const double *__restrict__ x, const double *__restrict__ dm, int ndm,
                                                                     BinGeom cg, const int *__restrict__ nU, const int *__restrict__ nodes,
                                                                     const unsigned *__restrict__ pm, int *__restrict__ cnt, int *__restrict__ overflow,
                                                                     int *__restrict__ tmp, int tcap, unsigned char *__restrict__ tab) {
    __shared__ int s_hash[PAT_WARPS][PATC_HASH];
    __shared__ __align__(16) int s_rows[PAT_WARPS][PATC_HASH / 2];
    __shared__ unsigned short s_slot[PAT_WARPS][PATC_HASH / 2];     // hash slot of rows[k]
    __shared__ unsigned char s_hrank[PAT_WARPS][PATC_HASH];         // rank (= position in the column) of the id in hash slot h
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t j = blockIdx.x * (int64_t)PAT_WARPS + wid;
    if (j >= nnodes) return;
    int *hs = s_hash[wid], *rows = s_rows[wid];
    unsigned short *slot = s_slot[wid];
    unsigned char *hrank = s_hrank[wid];
    for (int t = lane; t < PATC_HASH; t += 32) hs[t] = -1;
    int lo[FG_MAXDIM] = {0, 0, 0}, hi[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        const double xj = x[d * nnodes + j], r = dm[(d < ndm ? d : 0) * nnodes + j] * (1.0 + 1e-9);
        lo[d] = fg_bin_coord(xj - r, cg.origin[d], cg.inv_size[d], cg.nb[d]);
        hi[d] = fg_bin_coord(xj + r, cg.origin[d], cg.inv_size[d], cg.nb[d]);
    }
    __syncwarp();
    int nrow = 0;                                   // distinct rows inserted (warp-uniform)
    // The cells of the box are visited in chunks of 32: lane = cell reads the cell's union size in ONE parallel load, then the unions and
    // point masks of four non-empty cells are requested together before any of them is used (the plain triple loop chained a dependent
    // L2 round trip per cell: 64 of them in a row per column -- the kernel ran at the latency of that chain).
    const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, ncell = nx * ny * (hi[2] - lo[2] + 1);
    auto for_cells = [&](bool want_pm, auto &&visit) {
        for (int base = 0; base < ncell; base += 32) {
            const int n = base + lane;
            long long c_l = 0; int nu_l = 0;
            if (n < ncell) {
                const int bx = n % nx, t = n / nx, by = t % ny, bz = t / ny;
                c_l = ((long long)(lo[2] + bz) * cg.nb[1] + (lo[1] + by)) * cg.nb[0] + (lo[0] + bx);
                nu_l = nU[c_l];
            }
            unsigned live = __ballot_sync(0xffffffffu, nu_l > 0);       // cells of this chunk inside the index
            while (live) {
                long long c4[4]; int nu4[4], id4[4]; unsigned P4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    nu4[q] = 0; id4[q] = -1; P4[q] = 0u; c4[q] = 0;
                    if (live) {
                        const int src = __ffs(live) - 1;
                        live &= live - 1u;
                        c4[q] = __shfl_sync(0xffffffffu, c_l, src);
                        nu4[q] = __shfl_sync(0xffffffffu, nu_l, src);
                        id4[q] = nodes[c4[q] * 32 + lane];
                        if (want_pm) P4[q] = pm[c4[q] * 32 + lane];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (nu4[q] > 0) { if (!visit(c4[q], lane < nu4[q] ? id4[q] : -1, P4[q])) return; }
            }
        }
    };
    for_cells(true, [&](long long, int id, unsigned P) -> bool {
        const unsigned mj = __ballot_sync(0xffffffffu, id == (int)j);
        if (mj == 0u) return true;                  // this cluster does not hold node j
        const unsigned Pj = __shfl_sync(0xffffffffu, P, __ffs(mj) - 1);
        bool fresh = false;
        if (id >= 0 && (P & Pj) != 0u) {            // some point of the cluster holds both nodes
            unsigned h = ((unsigned)id * 2654435761u) >> 23;          // 9 bits
            while (true) {
                const int old = atomicCAS(&hs[h], -1, id);
                if (old == -1) { fresh = true; break; }
                if (old == id) break;
                h = (h + 1) & (PATC_HASH - 1);
            }
        }
        nrow += __popc(__ballot_sync(0xffffffffu, fresh));
        return nrow <= PATC_HASH / 2;               // (load factor 1/2; reported below)
    });
    __syncwarp();
    if (nrow > PATC_HASH / 2 || nrow > tcap) { if (lane == 0) { atomicMax(overflow + 2, nrow); cnt[j] = 0; } return; }
    // compact the set, rank-sort (ids are distinct), park ascending
    int n = 0;
    for (int t0 = 0; t0 < PATC_HASH; t0 += 32) {
        const int v = hs[t0 + lane];
        const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
        if (v >= 0) { const int k = n + __popc(m & ((1u << lane) - 1)); rows[k] = v; slot[k] = (unsigned short)(t0 + lane); }
        n += __popc(m);
    }
    __syncwarp();
    if (lane == 0) cnt[j] = n;
    if (n <= 128) {
        // rank sort with the lane's (up to four) values in registers and the list read four entries per broadcast load
        if (lane < 4 && n + lane < PATC_HASH / 2) rows[n + lane] = 0x7fffffff;        // padding of the last vector load
        __syncwarp();
        int v[4], rk[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { const int e = lane + 32 * r; v[r] = e < n ? rows[e] : 0x7fffffff; rk[r] = 0; }
        for (int t = 0; t < n; t += 4) {
            const int4 x = *reinterpret_cast<const int4 *>(rows + t);
#pragma unroll
            for (int r = 0; r < 4; ++r) rk[r] += (int)(x.x < v[r]) + (int)(x.y < v[r]) + (int)(x.z < v[r]) + (int)(x.w < v[r]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = lane + 32 * r;
            if (e < n) { tmp[j * (int64_t)tcap + rk[r]] = v[r]; hrank[slot[e]] = (unsigned char)rk[r]; }
        }
    } else {
        for (int e = lane; e < n; e += 32) {
            const int v = rows[e];
            int rank = 0;
            for (int t = 0; t < n; ++t) rank += rows[t] < v;
            tmp[j * (int64_t)tcap + rank] = v;
            hrank[slot[e]] = (unsigned char)min(rank, 255);
        }
    }
    if (tab == nullptr) return;
    // ---- flush tables of the assembly (fg_cluster.cu: k_cluster_tables), written here while the column is at hand: for every cluster
    // that holds node j as local node jl, row jl of its table = position of each union node U[i] <= j inside column j (255 = absent).
    // The set above knows every row of the column and, since the rank-sort, its position; one 32-byte store per (column, cluster).
    if (n > 254) { if (lane == 0) atomicExch(overflow + 4, 1); return; }      // positions are bytes: the caller runs k_cluster_tables instead
    __syncwarp();
    for_cells(false, [&](long long c, int id, unsigned) -> bool {
        const unsigned mj = __ballot_sync(0xffffffffu, id == (int)j);
        if (mj == 0u) return true;
        unsigned char pos = 255;
        if (id >= 0 && id <= (int)j) {
            unsigned h = ((unsigned)id * 2654435761u) >> 23;
            while (true) {
                const int key = hs[h];
                if (key == id) { pos = hrank[h]; break; }
                if (key == -1) break;
                h = (h + 1) & (PATC_HASH - 1);
            }
        }
        tab[c * 1024 + (__ffs(mj) - 1) * 32 + lane] = pos;
        return true;
    });
}

// CSC start and length of every block column of the cluster index (the other half of the flush tables)
__global__ void k_cluster_cols(int64_t nclusters, const int *__restrict__ nU, const int *__restrict__ nodes, const int64_t *__restrict__ colptr,
                               int64_t *__restrict__ col0, int *__restrict__ coln) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t c = t >> 5;
    if (c >= nclusters) return;
    if ((int)(t & 31) >= nU[c]) return;
    const int J = nodes[t];
    const int64_t c0 = colptr[J];
    col0[t] = c0; coln[t] = (int)(colptr[J + 1] - c0);
}

// Returns FG_OK with *done = true when the pattern was built from the cluster index; *done = false: not applicable, the caller
// runs the witness search.
static int pattern_from_clusters(fg_ctx *ctx, GaussSet &S, bool *done, bool *tables) {
    *done = false; *tables = false;
    NodeSet &N = ctx->nodes;
    Pattern &P = S.pat;
    if (!(S.have_nb && S.nb_clustered && S.nsub <= 1 && S.fem.empty() && S.n > 0 && S.total_L > 0 && fg_opt(ctx, "pattern_from_clusters", 1) != 0)) return FG_OK;
    if (fg_opt(ctx, "assembly_direct", 0) != 0) return FG_OK;
    int rc = fg_ensure_clusters(ctx, S);
    if (rc) return rc;
    Clusters &C = S.cl;
    if (!(C.ready && C.ucap == 32 && C.inv.p && C.stat_overflow_points == 0 && C.n > 0)) return FG_OK;
    const int64_t nn = N.n;
    const int tcap = 192;
    DevBuf<int> &cnt = ctx->pat_cnt;
    DevBuf<int> &tmp = ctx->pat_rows;
    FG_CUDA(ctx, cnt.reserve(nn));
    if (tmp.reserve((size_t)nn * tcap) != cudaSuccess) { cudaGetLastError(); return FG_OK; }
    FG_CUDA(ctx, ctx->pat_pmask.reserve((size_t)C.n * 32));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p, 0, 5 * sizeof(int), ctx->stream));
    // the flush tables of the symmetric assembly kernels come out of the same pass (rows <= column; a later unsymmetric form rebuilds them)
    const bool with_tab = fg_opt(ctx, "pattern_tables", 1) != 0;
    if (with_tab) {
        FG_CUDA(ctx, C.col0.reserve(C.n * (int64_t)32));
        FG_CUDA(ctx, C.coln.reserve(C.n * (int64_t)32));
        FG_CUDA(ctx, C.tab.reserve(C.n * (int64_t)32 * 32));
    }
    unsigned char *tabp = with_tab ? C.tab.p : nullptr;
    if (ctx->pmask_last == &S && ctx->pmask_ptr == ctx->pat_pmask.p && C.pmask_many == 0) {
        // the index build wrote the point masks (k_cluster_build_masks) and nothing has touched the buffer since
    } else {
        k_cluster_pmask<<<(unsigned)((C.n + 3) / 4), 128, 0, ctx->stream>>>(C.n, C.grid.start.p, C.sorted.p, C.nU.p, C.inv.p, ctx->pat_pmask.p, ctx->flags.p + 3);
        FG_CHECK_LAUNCH(ctx);
        ctx->pmask_last = nullptr;
    }
    const unsigned grid = (unsigned)((nn + PAT_WARPS - 1) / PAT_WARPS);
    const int ndm = N.shape == FG_SHAPE_RECTANGULAR ? N.dim : 1;
    if (N.dim == 2) k_pattern_clusters<2><<<grid, PAT_WARPS * 32, 0, ctx->stream>>>(nn, N.x.p, N.dm.p, ndm, fg_geom(C.grid), C.nU.p, C.nodes.p, ctx->pat_pmask.p, cnt.p, ctx->flags.p, tmp.p, tcap, tabp);
    else k_pattern_clusters<3><<<grid, PAT_WARPS * 32, 0, ctx->stream>>>(nn, N.x.p, N.dm.p, ndm, fg_geom(C.grid), C.nU.p, C.nodes.p, ctx->pat_pmask.p, cnt.p, ctx->flags.p, tmp.p, tcap, tabp);
    FG_CHECK_LAUNCH(ctx);
    if ((rc = fg_exclusive_scan_i32_to_i64(ctx, cnt.p, P.colptr.p, nn))) return rc;
    size_t bytes = 0;
    cub::DeviceReduce::Max(nullptr, bytes, cnt.p, ctx->flags.p + 1, (int)nn, ctx->stream);
    FG_CUDA(ctx, ctx->temp.reserve(bytes));
    cub::DeviceReduce::Max(ctx->temp.p, bytes, cnt.p, ctx->flags.p + 1, (int)nn, ctx->stream);
    int h[5] = {0, 0, 0, 0, 0}; int64_t nnz = 0;
    FG_CUDA(ctx, cudaMemcpyAsync(h, ctx->flags.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(&nnz, P.colptr.p + nn, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[2] != 0 || h[3] != 0) return FG_OK;          // a column longer than the scratch stride / a cluster with more than 32 points: witness search
    P.nnz = nnz; P.max_col = h[1];
    if (P.rowval.reserve(nnz) != cudaSuccess) return fg_fail(ctx, FG_E_OOM, "pattern: rowval (%lld entries)", (long long)nnz);
    k_pattern_compact<<<(unsigned)((nn * 32 + 255) / 256), 256, 0, ctx->stream>>>(nn, P.colptr.p, tmp.p, tcap, P.rowval.p);
    FG_CHECK_LAUNCH(ctx);
    if (with_tab && h[4] == 0) {                        // every column fits byte positions: the tables are complete
        k_cluster_cols<<<(unsigned)((C.n * 32 + 255) / 256), 256, 0, ctx->stream>>>(C.n, C.nU.p, C.nodes.p, P.colptr.p, C.col0.p, C.coln.p);
        FG_CHECK_LAUNCH(ctx);
        *tables = true;
    }
    *done = true;
    return FG_OK;
}

// Pattern of set S built from the Gauss positions of set W (W = S normally; a halo-extended
// superset on a multi-GPU slab so that shared columns have the same layout on both ranks).
int fg_impl_pattern(fg_ctx *ctx, GaussSet &S, GaussSet &W) {
    NodeSet &N = ctx->nodes;
    Pattern &P = S.pat;
    P.ready = false; P.D = 0; P.colptr_host.clear(); ++P.version;
    S.cl.tab_ready = false;          // flush tables index this pattern
    FG_CUDA(ctx, P.colptr.reserve(N.n + 1));
    if (&S == &W) {                  // the set's own Gauss points: the cluster index knows every (point, node) incidence already
        bool done = false, tables = false;
        const int rcc = pattern_from_clusters(ctx, S, &done, &tables);
        if (rcc) return rcc;
        if (done) {
            P.ready = true; P.tmap_ready = false;
            if (tables) { S.cl.tab_ready = true; S.cl.tab_upper = true; }
            return FG_OK;
        }
    }
    if (!W.gbins_ready) {
        // Gauss bins for the witness search: same origin as the node grid, cells three times finer per dimension
        // (a failing pair then tests a handful of points instead of whole support-sized bins)
        double inv[FG_MAXDIM] = {1, 1, 1}; int nbg[FG_MAXDIM] = {1, 1, 1};
        double total = 1.0;
        for (int d = 0; d < N.dim; ++d) {
            inv[d] = N.bins.inv_size[d] * 3.0;
            nbg[d] = (int)std::min(4096.0, floor(N.ext[d] * inv[d]) + 1.0);
            total *= nbg[d];
        }
        if (total > 2.0e8) { for (int d = 0; d < N.dim; ++d) { inv[d] = N.bins.inv_size[d]; nbg[d] = N.bins.nb[d]; } }
        int rc = fg_build_bins(ctx, N.dim, W.n, W.gs.p, N.bins.origin, inv, nbg, W.gbins, W.gsorted);
        if (rc) return rc;
        W.gbins_ready = true;
    }
    FG_CUDA(ctx, ctx->pat_gxs.reserve((size_t)N.dim * W.n));
    const bool multi_sub = W.nsub > 1;
    if (multi_sub) FG_CUDA(ctx, ctx->pat_gsub.reserve(W.n));
    if (W.n > 0) {
        k_gather_gauss<<<(unsigned)((W.n + 255) / 256), 256, 0, ctx->stream>>>(N.dim, W.n, W.gsorted.p, W.gs.p, ctx->pat_gxs.p,
                                                                              multi_sub ? W.psub.p : nullptr, ctx->pat_gsub.p);
        FG_CHECK_LAUNCH(ctx);
    }
    PatArgs A;
    A.nnodes = N.n; A.ngauss = W.n; A.x = N.x.p; A.dm = N.dm.p; A.xs = N.xs.p; A.dms = N.dms.p;
    A.sorted_id = N.sorted_id.p; A.nbin_start = N.bins.start.p; A.gxs = ctx->pat_gxs.p; A.gbin_start = W.gbins.start.p;
    A.geom = fg_geom(N.bins);
    A.ggeom = fg_geom(W.gbins);
    A.masks = W.nsub > 0 ? W.nmask.p : nullptr; A.gsub = multi_sub ? ctx->pat_gsub.p : nullptr; A.mask_words = W.mask_words;
    for (int d = 0; d < FG_MAXDIM; ++d) A.reach[d] = N.reach[d];
    int rc;
    if (N.dim == 2) rc = N.shape == FG_SHAPE_RECTANGULAR ? run<2, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<2, FG_SHAPE_CIRCULAR>(ctx, S, A);
    else rc = N.shape == FG_SHAPE_RECTANGULAR ? run<3, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<3, FG_SHAPE_CIRCULAR>(ctx, S, A);
    if (rc) return rc;
    P.ready = true;
    P.tmap_ready = false;
    return FG_OK;
}

// fg_cluster.cu -- spatial clusters of Gauss points for the shared-memory assembly (kernel 3).
//
// All Gauss points that fall into one cell of a coarse "cluster grid" share most of their
// neighbours: on the Poisson3D grid a 2x2x2-cell cluster has 216 points whose 85k (a,b)
// contributions land on only ~6.9k distinct matrix entries.  For every cluster this file
// computes the sorted union U of its neighbour lists and, for every DOM entry, its index in U,
// so that the assembly kernel can accumulate a dense |U| x |U| block in shared memory and
// flush it once.  Works for arbitrary clouds; a cluster whose union exceeds `ucap` is marked
// and its points take the direct path.
#include "fg_internal.cuh"
#include <math.h>
#include <algorithm>

#define CL_HASH 2048
static inline const NodeSet &N_(const fg_ctx *ctx) { return ctx->nodes; }

// hbits: log2 of the hash-set size actually used (>= 4 x ucap slots, at most CL_HASH): small sets are cleared and
// compacted in a few instructions per thread.
__global__ void __launch_bounds__(128) k_cluster_build(int64_t nclusters, const int *__restrict__ cstart, const int *__restrict__ csorted,
                                                       const int64_t *__restrict__ off, const int *__restrict__ dom, int ucap, int hbits,
                                                       int *__restrict__ nodes, int *__restrict__ nU, unsigned char *__restrict__ lidx,
                                                       int *__restrict__ ovf_points, int *__restrict__ max_u, unsigned *__restrict__ inv) {
    __shared__ int ht[CL_HASH];
    __shared__ unsigned char hrank[CL_HASH];     // rank of the id in slot h within the sorted union
    __shared__ int list[256];
    __shared__ int lslot[256];
    __shared__ int cnt, overflow;
    __shared__ __align__(4) unsigned char sinv[4][32];      // per warp: the inverse map of the point being indexed
    const int c = blockIdx.x;
    const int p0 = cstart[c], p1 = cstart[c + 1];
    if (p0 == p1) { if (threadIdx.x == 0) nU[c] = 0; return; }
    const int hsize = 1 << hbits, hshift = 32 - hbits;
    for (int t = threadIdx.x; t < hsize; t += blockDim.x) ht[t] = -1;
    if (threadIdx.x == 0) { cnt = 0; overflow = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // union of the neighbour lists (open-addressing hash set)
    for (int p = p0 + wid; p < p1; p += nw) {
        const int g = csorted[p];
        const int64_t o0 = off[g], o1 = off[g + 1];
        for (int64_t e = o0 + lane; e < o1; e += 32) {
            if (overflow) break;
            const int id = dom[e];
            unsigned h = ((unsigned)id * 2654435761u) >> hshift;
            while (true) {
                const int old = atomicCAS(&ht[h], -1, id);
                if (old == id) break;
                if (old == -1) { if (atomicAdd(&cnt, 1) >= ucap) overflow = 1; break; }
                h = (h + 1) & (hsize - 1);
                if (overflow) break;
            }
        }
    }
    __syncthreads();
    if (overflow) {     // too many distinct nodes: these points take the direct path
        if (threadIdx.x == 0) nU[c] = -1;
        for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) { const int slot = atomicAdd(ovf_points, 1); ovf_points[1 + slot] = csorted[p]; }
        return;
    }
    const int n = cnt;
    __syncthreads();
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < hsize; t += blockDim.x) { const int id = ht[t]; if (id >= 0) { const int k = atomicAdd(&cnt, 1); list[k] = id; lslot[k] = t; } }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {     // rank sort (ids distinct)
        const int v = list[t];
        int rank = 0;
        for (int k = 0; k < n; ++k) rank += list[k] < v;
        nodes[(int64_t)c * ucap + rank] = v;
        hrank[lslot[t]] = (unsigned char)rank;
    }
    if (threadIdx.x == 0) { nU[c] = n; atomicMax(max_u, n); }
    __syncthreads();
    // local index of every DOM entry: its id's slot in the hash set knows the rank
    for (int p = p0 + wid; p < p1; p += nw) {
        const int g = csorted[p];
        const int64_t o0 = off[g], o1 = off[g + 1];
        if (inv) { sinv[wid][lane] = 255; __syncwarp(); }
        for (int64_t e = o0 + lane; e < o1; e += 32) {
            const int id = dom[e];
            unsigned h = ((unsigned)id * 2654435761u) >> hshift;
            while (ht[h] != id) h = (h + 1) & (hsize - 1);
            const int rank = hrank[h];
            lidx[e] = (unsigned char)rank;
            // gather map of the point-batched MMA kernel: byte (rank >> 3) of word (rank & 7) = position in the list
            if (inv && rank < 32 && e - o0 < 255) sinv[wid][(rank & 7) * 4 + (rank >> 3)] = (unsigned char)(e - o0);
        }
        if (inv) {
            __syncwarp();
            if (lane < 8) inv[(int64_t)g * 8 + lane] = reinterpret_cast<const unsigned *>(sinv[wid])[lane];
            __syncwarp();
        }
    }
}

// Mean node spacing per dimension: h_d = s * reach_d with prod(ext_d/h_d + 1) = n  (bisection on s).
static void mean_spacing(const NodeSet &N, double *h) {
    double lo = 1e-9, hi = 1e9;
    for (int it = 0; it < 200; ++it) {
        const double s = sqrt(lo * hi);
        double cnt = 1.0;
        for (int d = 0; d < N.dim; ++d) cnt *= N.ext[d] / (s * N.reach[d]) + 1.0;
        if (cnt > (double)N.n) lo = s; else hi = s;
    }
    const double s = sqrt(lo * hi);
    for (int d = 0; d < N.dim; ++d) h[d] = s * N.reach[d];
}

// Nodes of a lattice with unit spacing that can lie in the support of a point of the cell [a, a+k], support
// half-width r:  a = 0 (cell between nodes) or a = -1/2 (cell centred on a node).  For r = 1.35 and k = 1 the
// centred cell sees 3 nodes per dimension, the aligned one 4: 27 instead of 64 nodes per cluster in 3D.
static double lattice_count(int k, double r, bool centred) {
    const double e = 1e-9;
    return centred ? floor(k - 0.5 + r + e) + floor(0.5 + r + e) + 1.0 : floor(k + r + e) + floor(r + e) + 1.0;
}

// Cluster-grid cell sizes and alignment: the largest whole number of mean node spacings per dimension such that a
// cell plus one support half-width on either side still holds at most `ucap` nodes, each dimension aligned
// (shift 0) or centred on the node lattice (shift -h/2), whichever sees fewer nodes.  Returns the node estimate.
static double choose_cluster_size(const NodeSet &N, int ucap, double *size, double *shift) {
    const int dim = N.dim;
    double h[FG_MAXDIM] = {1, 1, 1}, r[FG_MAXDIM] = {0, 0, 0};
    mean_spacing(N, h);
    for (int d = 0; d < dim; ++d) r[d] = N.reach[d] / h[d];
    int k[FG_MAXDIM] = {1, 1, 1};
    auto count_d = [&](int d, int kk) { return std::min(lattice_count(kk, r[d], false), lattice_count(kk, r[d], true)); };
    auto nodes_in = [&](const int *kk) {
        double u = 1.0;
        for (int d = 0; d < dim; ++d) u *= count_d(d, kk[d]);
        return u;
    };
    // greedy per-dimension growth: keep adding one spacing to the next dimension while the worst-case
    // node count of (cell + support margins) stays within ucap
    for (int round = 0; round < 64; ++round) {
        bool grew = false;
        for (int d = 0; d < dim; ++d) {
            int t[FG_MAXDIM] = {k[0], k[1], k[2]};
            t[d] += 1;
            if (nodes_in(t) <= ucap) { k[d] = t[d]; grew = true; }
        }
        if (!grew) break;
    }
    for (int d = 0; d < dim; ++d) {
        size[d] = k[d] * h[d] * (1.0 - 1e-7);   // a hair under k spacings: grid-aligned clouds stay inside
        shift[d] = lattice_count(k[d], r[d], true) < lattice_count(k[d], r[d], false) ? -0.5 * h[d] : 0.0;
    }
    return nodes_in(k);
}

// Block size of the cluster index: 32 nodes when single-spacing cells fit (10 accumulator tiles per warp in the
// register-block kernel, twice the warps per SM), else 64.  A function of the node set only, so that the neighbour
// search and the assembly agree on the cluster grid.
int fg_default_ucap(const fg_ctx *ctx, const NodeSet &N) {
    const int forced = fg_opt(ctx, "cluster_ucap", 0);
    if (forced > 0) return forced;
    double size[FG_MAXDIM], shift[FG_MAXDIM];
    return choose_cluster_size(N, 32, size, shift) <= 32.0 ? 32 : 64;
}

// Flush tables (tab[j*ucap + i]): for every cluster and every block column j, where the rows of U sit inside the CSC
// column of node U[j].  One warp per (cluster, column): stream the column's row list, look each row up in
// U through a small hash map (id -> local index), record its position (< 255) in a shared-memory copy of the table
// that is written out coalesced.  A column longer than 254 rows disables the cluster.
#define CT_HASH 512
__global__ void __launch_bounds__(128) k_cluster_tables(int64_t nclusters, int ucap, const int *__restrict__ nodes, int *__restrict__ nU,
                                                        const int64_t *__restrict__ colptr, const int *__restrict__ rowval,
                                                        int64_t *__restrict__ col0, int *__restrict__ coln, unsigned char *__restrict__ tab,
                                                        const int *__restrict__ cstart, const int *__restrict__ csorted, int *__restrict__ ovf_points,
                                                        int upper_only) {
    extern __shared__ __align__(16) unsigned char sT[];      // the cluster's table, written out coalesced at the end
    __shared__ int hkey[CT_HASH];                            // node id -> local index (open addressing, load <= 1/2)
    __shared__ unsigned char hval[CT_HASH];
    __shared__ int sU[256];
    __shared__ int bad;
    const int64_t c = blockIdx.x;
    const int nu = nU[c];
    if (nu <= 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = threadIdx.x; t < CT_HASH; t += blockDim.x) hkey[t] = -1;
    for (int t = threadIdx.x; t < nu * ucap / 4; t += blockDim.x) reinterpret_cast<unsigned *>(sT)[t] = 0xFFFFFFFFu;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < nu; t += blockDim.x) {
        const int id = nodes[c * ucap + t];
        sU[t] = id;
        unsigned h = ((unsigned)id * 2654435761u) >> 23;     // 9 bits
        while (atomicCAS(&hkey[h], -1, id) != -1) h = (h + 1) & (CT_HASH - 1);
        hval[h] = (unsigned char)t;
    }
    __syncthreads();
    for (int j = wid; j < nu; j += nw) {
        const int J = sU[j];
        const int64_t c0 = colptr[J];
        const int nj = (int)(colptr[J + 1] - c0);
        if (lane == 0) { col0[c * ucap + j] = c0; coln[c * ucap + j] = nj; }
        if (nj > 254) { bad = 1; continue; }
        for (int k = lane; k < nj; k += 32) {
            const int r = rowval[c0 + k];
            if (upper_only && r > J) break;                  // rows ascend: the symmetric kernels never look below the diagonal
            unsigned h = ((unsigned)r * 2654435761u) >> 23;
            while (true) {
                const int key = hkey[h];
                if (key == r) { sT[j * ucap + hval[h]] = (unsigned char)k; break; }   // column-major: column j owns 'ucap' contiguous bytes
                if (key == -1) break;
                h = (h + 1) & (CT_HASH - 1);
            }
        }
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) nU[c] = -1;
        for (int p = cstart[c] + threadIdx.x; p < cstart[c + 1]; p += blockDim.x) { const int slot = atomicAdd(ovf_points, 1); ovf_points[1 + slot] = csorted[p]; }
        return;
    }
    unsigned *T = reinterpret_cast<unsigned *>(tab + c * (int64_t)ucap * ucap);
    for (int t = threadIdx.x; t < nu * ucap / 4; t += blockDim.x) T[t] = reinterpret_cast<const unsigned *>(sT)[t];
}

// ovf_points = [count, ids (<= S.n) ..., saved count]: the assembly kernels append the points they cannot take
// (neighbour lists longer than a warp) to the list of the index build; every call starts from the saved count.
static int overflow_list_save(fg_ctx *ctx, GaussSet &S) {
    FG_CUDA(ctx, cudaMemcpyAsync(S.cl.ovf_points.p + S.n + 2, S.cl.ovf_points.p, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    return FG_OK;
}
int fg_overflow_list_reset(fg_ctx *ctx, GaussSet &S) {
    FG_CUDA(ctx, cudaMemcpyAsync(S.cl.ovf_points.p, S.cl.ovf_points.p + S.n + 2, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    return FG_OK;
}

int fg_impl_cluster_tables(fg_ctx *ctx, GaussSet &S, bool upper_only) {
    Clusters &C = S.cl;
    Pattern &P = S.pat;
    if (!C.ready || !P.ready) return fg_fail(ctx, FG_E_STATE, "cluster tables need the clusters and the pattern");
    C.tab_ready = false;
    C.tab_upper = upper_only;
    if (C.n == 0) { C.tab_ready = true; return FG_OK; }
    FG_CUDA(ctx, C.col0.reserve(C.n * (int64_t)C.ucap));
    FG_CUDA(ctx, C.coln.reserve(C.n * (int64_t)C.ucap));
    FG_CUDA(ctx, C.tab.reserve(C.n * (int64_t)C.ucap * C.ucap));
    int rc = fg_overflow_list_reset(ctx, S);
    if (rc) return rc;
    const size_t smem = (size_t)C.ucap * C.ucap;
    FG_CUDA(ctx, cudaFuncSetAttribute(k_cluster_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cluster_tables<<<(unsigned)C.n, 128, smem, ctx->stream>>>(C.n, C.ucap, C.nodes.p, C.nU.p, P.colptr.p, P.rowval.p, C.col0.p, C.coln.p, C.tab.p,
                                                                C.grid.start.p, C.sorted.p, C.ovf_points.p, upper_only ? 1 : 0);
    FG_CHECK_LAUNCH(ctx);
    if ((rc = overflow_list_save(ctx, S))) return rc;
    C.tab_ready = true;
    return FG_OK;
}

// Sort the Gauss points of a set into the cells of the cluster grid (depends on (nodes, gs, ucap) only).
int fg_cluster_binning(fg_ctx *ctx, GaussSet &S, Clusters &C, int ucap, int cells_for) {
    NodeSet &N = ctx->nodes;
    if (ucap > 252) ucap = 252;
    ucap &= ~3;                      // rows of the flush table stay 4-byte, the table 16-byte aligned
    if (cells_for <= 0 || cells_for > ucap) cells_for = ucap;
    if (C.binned && C.ucap == ucap && C.cells_for == cells_for) return FG_OK;
    C.ready = false; C.tab_ready = false;
    C.ucap = ucap; C.cells_for = cells_for;
    if (S.n == 0) { C.n = 0; C.binned = true; return FG_OK; }
    double size[FG_MAXDIM] = {1, 1, 1}, shift[FG_MAXDIM] = {0, 0, 0}, inv[FG_MAXDIM] = {1, 1, 1}, origin[FG_MAXDIM] = {0, 0, 0};
    int nb[FG_MAXDIM] = {1, 1, 1};
    choose_cluster_size(N, cells_for, size, shift);
    for (int d = 0; d < N.dim; ++d) {
        inv[d] = 1.0 / size[d];
        origin[d] = N.bins.origin[d] + shift[d];
        nb[d] = (int)std::min(2048.0, floor((N.ext[d] - shift[d]) * inv[d]) + 1.0);
    }
    int rc = fg_build_bins(ctx, N.dim, S.n, S.gs.p, origin, inv, nb, C.grid, C.sorted);
    if (rc) return rc;
    for (int d = 0; d < FG_MAXDIM; ++d) { C.csize[d] = size[d]; C.cshift[d] = shift[d]; }
    C.n = C.grid.nbins;
    C.binned = true;
    return FG_OK;
}

// The same index from the SEARCH's own data, without hashing: when the cluster grid nests in the search grid, all points of a
// cluster drew their neighbours from one id-sorted candidate list and carry hit masks over it.  The union is the OR of the masks,
// the local index of candidate b is the number of set union bits below b, and a point's list is the ascending walk of its own mask.
// One warp per cluster; nothing but bit operations (k_cluster_build: 5.7 ms at 2.7e7 points, this kernel: see DESIGN.md).
__global__ void __launch_bounds__(128) k_cluster_build_masks(int64_t nclusters, const int *__restrict__ cstart, const int *__restrict__ csorted,
                                                             const int64_t *__restrict__ off, const double *__restrict__ gs, int64_t ngauss, int dim,
                                                             BinGeom sg, const int *__restrict__ cand, const unsigned *__restrict__ masks, int capc,
                                                             int ucap, int *__restrict__ nodes, int *__restrict__ nU, unsigned char *__restrict__ lidx,
                                                             int *__restrict__ ovf_points, int *__restrict__ max_u, unsigned *__restrict__ inv, int mode,
                                                             unsigned *__restrict__ pm, int *__restrict__ pm_flag) {
    // mode bit 0: the index proper (union, nU, overflow list, gather map); bit 1: lidx.  The point-batched kernels of 32-node blocks read
    // only the gather map, so the index build of such blocks leaves lidx to the first kernel that wants it (fg_ensure_lidx: mode 2 --
    // 5e8 scattered byte stores kept the L1 at 85 % of its throughput and were half of this kernel's time).
    const bool build = (mode & 1) != 0, want_lidx = (mode & 2) != 0;
    const int lane = threadIdx.x & 31;
    const int64_t c = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nclusters) return;
    const int p0 = cstart[c], p1 = cstart[c + 1];
    if (p0 == p1) { if (lane == 0 && build) nU[c] = 0; return; }
    if (!build && nU[c] <= 0) return;                 // lidx only: clusters outside the index have none
    // pm (optional): per local node the POINT mask of the sparsity build (bit q = the cluster's q-th point holds the node), the transpose
    // of the per-point membership words this kernel has in registers anyway
    const bool want_pm = build && pm != nullptr;
    if (want_pm && p1 - p0 > 32 && lane == 0) atomicExch(pm_flag, 1);                // more points than mask bits: the caller falls back
    // the search cluster of the first point (all points of this cluster share it: nested grids)
    int cs;
    {
        const int g0 = csorted[p0];
        int b[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < FG_MAXDIM; ++d) if (d < dim) b[d] = fg_bin_coord(gs[d * ngauss + g0], sg.origin[d], sg.inv_size[d], sg.nb[d]);
        cs = (b[2] * sg.nb[1] + b[1]) * sg.nb[0] + b[0];
    }
    const int *cl = cand + cs * (int64_t)(1 + 2 * capc);
    unsigned um[4] = {0u, 0u, 0u, 0u};
    bool stray = false;                 // a point of another search cluster (cannot happen with nested grids; the direct path catches it)
    for (int p = p0 + lane; p < p1; p += 32) {
        const int g = csorted[p];
        int b[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < FG_MAXDIM; ++d) if (d < dim) b[d] = fg_bin_coord(gs[d * ngauss + g], sg.origin[d], sg.inv_size[d], sg.nb[d]);
        stray |= ((b[2] * sg.nb[1] + b[1]) * sg.nb[0] + b[0]) != cs;
        const uint4 mk = __ldg(reinterpret_cast<const uint4 *>(masks + (int64_t)g * 4));
        um[0] |= mk.x; um[1] |= mk.y; um[2] |= mk.z; um[3] |= mk.w;
    }
    const bool any_stray = __any_sync(0xffffffffu, stray);
#pragma unroll
    for (int w = 0; w < 4; ++w) um[w] = __reduce_or_sync(0xffffffffu, um[w]);
    int pre[4];
    pre[0] = 0; pre[1] = __popc(um[0]); pre[2] = pre[1] + __popc(um[1]); pre[3] = pre[2] + __popc(um[2]);
    const int n = pre[3] + __popc(um[3]);
    if (n > ucap || any_stray) {     // too many distinct nodes: these points take the direct path
        if (!build) return;
        if (lane == 0) nU[c] = -1;
        for (int p = p0 + lane; p < p1; p += 32) { const int slot = atomicAdd(ovf_points, 1); ovf_points[1 + slot] = csorted[p]; }
        return;
    }
    if (build) {
#pragma unroll
        for (int w = 0; w < 4; ++w)
            if ((um[w] >> lane) & 1u) nodes[c * ucap + pre[w] + __popc(um[w] & ((1u << lane) - 1u))] = cl[1 + 32 * w + lane];
        if (lane == 0) { nU[c] = n; atomicMax(max_u, n); }
    }
    const bool want_inv = build && inv != nullptr;
    // every point: local index of each list entry, and (32-node blocks) the inverse map
    unsigned held0 = 0u;                 // local nodes held by my point of the first 32
    for (int p = p0 + lane; p < p1; p += 32) {
        const int g = csorted[p];
        const uint4 mk = __ldg(reinterpret_cast<const uint4 *>(masks + (int64_t)g * 4));
        const unsigned m[4] = {mk.x, mk.y, mk.z, mk.w};
        const int64_t o0 = off[g];
        unsigned held = 0u;
        unsigned wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = 0xFFFFFFFFu;
        int pos = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned mm = m[w];
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1;
                const int rank = pre[w] + __popc(um[w] & ((1u << b) - 1u));
                if (want_lidx) lidx[o0 + pos] = (unsigned char)rank;
                if (rank < 32) held |= 1u << rank;
                if (want_inv && rank < 32 && pos < 255) {
                    const int q = rank & 7, sh = 8 * (rank >> 3);
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (j == q) wv[j] = (wv[j] & ~(0xFFu << sh)) | ((unsigned)pos << sh);
                }
                ++pos;
            }
        }
        if (want_inv) {
            uint4 *dst = reinterpret_cast<uint4 *>(inv + (int64_t)g * 8);
            dst[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            dst[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        }
        if (p < p0 + 32) held0 = held;
    }
    if (want_pm) {
        unsigned P = 0u;
#pragma unroll
        for (int a = 0; a < 32; ++a) {
            const unsigned col = __ballot_sync(0xffffffffu, (held0 >> a) & 1u);
            if (a == lane) P = col;
        }
        pm[c * 32 + lane] = P;
    }
}

// Does every cell of grid `a` lie inside one cell of grid `s`?  (same origin up to the lattice shift, integer size ratio)
static bool grids_nest(const NodeSet &N, const Clusters &a, const Clusters &s) {
    for (int d = 0; d < N.dim; ++d) {
        if (!(a.csize[d] > 0.0) || !(s.csize[d] > 0.0)) return false;
        const double ratio = s.csize[d] / a.csize[d];
        if (fabs(ratio - floor(ratio + 0.5)) > 1e-6 || ratio < 0.999) return false;
        if (fabs(a.cshift[d] - s.cshift[d]) > 1e-9 * a.csize[d]) return false;
    }
    return true;
}

// Search clusters (cells of the neighbour search's grid, numbered slowest dimension last) that hold every Gauss point of the assembly
// clusters [0, c1): both grids number their cells the same way, so when the assembly cells nest in the search cells the answer is "all
// search cells up to the slab of the last assembly cell".  Grids that do not nest: everything.
int64_t fg_search_clusters_below(const fg_ctx *ctx, const GaussSet &S, int64_t c1) {
    const NodeSet &N = N_(ctx);
    const Clusters &a = S.cl, &s = S.nbcl;
    if (c1 <= 0) return 0;
    if (!a.binned || !s.binned || !grids_nest(N, a, s) || c1 >= a.n) return s.n;
    const int sd = N.dim - 1;                                   // slowest dimension of the cell numbering
    int64_t fast_a = 1, fast_s = 1;
    for (int d = 0; d < sd; ++d) { fast_a *= a.grid.nb[d]; fast_s *= s.grid.nb[d]; }
    const int64_t slab_a = (c1 - 1) / fast_a;                   // slab of the last assembly cell of the range
    const int ratio = (int)floor(s.csize[sd] / a.csize[sd] + 0.5);
    const int64_t slab_s = slab_a / std::max(ratio, 1);
    return std::min<int64_t>(s.n, (slab_s + 1) * fast_s);
}

// lidx of an index whose build left it out (32-node blocks from the search's masks): the same kernel, writing nothing else
int fg_ensure_lidx(fg_ctx *ctx, GaussSet &S) {
    Clusters &C = S.cl;
    if (!C.ready || C.lidx_ready || C.n == 0) return FG_OK;
    if (!C.from_masks) return fg_fail(ctx, FG_E_STATE, "cluster index without lidx");
    k_cluster_build_masks<<<(unsigned)((C.n + 3) / 4), 128, 0, ctx->stream>>>(C.n, C.grid.start.p, C.sorted.p, S.off.p, S.gs.p, S.n, ctx->nodes.dim,
        fg_geom(S.nbcl.grid), S.cand.p, S.masks.p, S.nb_capc, C.ucap, C.nodes.p, C.nU.p, C.lidx.p, C.ovf_points.p, ctx->flags.p + 8, nullptr, 2, nullptr, nullptr);
    FG_CHECK_LAUNCH(ctx);
    C.lidx_ready = true;
    return FG_OK;
}

int fg_impl_clusters(fg_ctx *ctx, GaussSet &S, int ucap, int cells_for) {
    Clusters &C = S.cl;
    C.ready = false; C.tab_ready = false;
    int rc = fg_cluster_binning(ctx, S, C, ucap, cells_for);
    if (rc) return rc;
    ucap = C.ucap;
    if (S.n == 0) { C.ready = true; return FG_OK; }
    FG_CUDA(ctx, C.nU.reserve(C.n));
    FG_CUDA(ctx, C.nodes.reserve(C.n * (int64_t)ucap));
    FG_CUDA(ctx, C.lidx.reserve(S.total_L));
    if (ucap == 32) FG_CUDA(ctx, C.inv.reserve((size_t)S.n * 8));
    FG_CUDA(ctx, C.ovf_points.reserve(S.n + 4));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    FG_CUDA(ctx, cudaMemsetAsync(C.ovf_points.p, 0, sizeof(int), ctx->stream));
    FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p + 8, 0, sizeof(int), ctx->stream));
    C.lidx_ready = true; C.from_masks = false; C.pmask_many = -1;
    if (ctx->pmask_last == &S) ctx->pmask_last = nullptr;
    if (S.nb_clustered && S.nb_capc == 128 && S.nbcl.binned && grids_nest(N_(ctx), C, S.nbcl) && fg_opt(ctx, "cluster_from_masks", 1) != 0) {
        // the search's candidate lists and hit masks describe DOM completely: no hashing, DOM itself is not even read.
        // 32-node blocks: lidx is left to the first kernel that reads it (fg_ensure_lidx); the point-batched kernels never do
        const bool lazy_lidx = ucap == 32 && fg_opt(ctx, "cluster_lidx_eager", 0) == 0;
        // 32-node blocks: the point masks of the sparsity build (fg_pattern.cu) come out of the same pass
        unsigned *pmp = nullptr;
        if (ucap == 32 && ctx->pat_pmask.reserve((size_t)C.n * 32) == cudaSuccess) pmp = ctx->pat_pmask.p; else cudaGetLastError();
        FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p + 10, 0, sizeof(int), ctx->stream));
        k_cluster_build_masks<<<(unsigned)((C.n + 3) / 4), 128, 0, ctx->stream>>>(C.n, C.grid.start.p, C.sorted.p, S.off.p, S.gs.p, S.n, ctx->nodes.dim,
            fg_geom(S.nbcl.grid), S.cand.p, S.masks.p, S.nb_capc, ucap, C.nodes.p, C.nU.p, C.lidx.p, C.ovf_points.p, ctx->flags.p + 8,
            ucap == 32 ? C.inv.p : nullptr, lazy_lidx ? 1 : 3, pmp, ctx->flags.p + 10);
        FG_CHECK_LAUNCH(ctx);
        ctx->pmask_last = pmp ? &S : nullptr; ctx->pmask_ptr = pmp;
        C.lidx_ready = !lazy_lidx; C.from_masks = true;
    } else {
        if ((rc = fg_ensure_dom(ctx, S))) return rc;
        int hbits = 7;                                   // >= 4 x ucap slots
        while ((1 << hbits) < 4 * ucap && (1 << hbits) < CL_HASH) ++hbits;
        k_cluster_build<<<(unsigned)C.n, 128, 0, ctx->stream>>>(C.n, C.grid.start.p, C.sorted.p, S.off.p, S.dom.p, ucap, hbits, C.nodes.p, C.nU.p,
                                                                C.lidx.p, C.ovf_points.p, ctx->flags.p + 8, ucap == 32 ? C.inv.p : nullptr);
        FG_CHECK_LAUNCH(ctx);
    }
    if ((rc = overflow_list_save(ctx, S))) return rc;
    C.stat_overflow_points = -1; C.stat_max_u = -1;
    C.stream_ready = false;
    C.ready = true;
    return FG_OK;
}

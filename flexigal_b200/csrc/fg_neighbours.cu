// fg_neighbours.cu -- kernel 1: cell-binned support-domain neighbour search.
//
// Replaces domain() (src/Shape_Functions.jl:3-27), which scans ALL nodes per Gauss point.
// Nodes are binned on a uniform grid whose cell is the largest support half-width, so the
// candidates of a Gauss point lie in the 3^dim bins around it; every candidate is then put
// through the reference's own FP64 predicate (fg_inside_1d / the circular form), so the
// neighbour lists are bit-identical to findall(inside): same members, ascending 1-based ids.
//
// One thread per Gauss point (neighbouring points scan the same bins, so the loads of a
// warp are mostly broadcasts).  Two passes: count -> exclusive scan -> fill.  The fill pass
// keeps each thread's hits in shared memory (slot-major, conflict free), insertion-sorts them
// by node id (bins are not id-ordered) and the block then writes its contiguous slice of DOM
// with coalesced stores.
#include "fg_internal.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <stdlib.h>

struct NbArgs {
    int64_t nnodes, ngauss;
    const double *xs, *dms;   // bin-ordered node data
    const int *sorted_id, *bin_start;
    const double *gs;
    BinGeom geom;
    double reach[FG_MAXDIM];
    const unsigned *mask;     // node subset of the set (bit per node id), or nullptr: every node is a candidate
};

__device__ __forceinline__ bool fg_mask_bit(const unsigned *mask, int id) { return (__ldg(mask + (id >> 5)) >> (id & 31)) & 1u; }

template <int DIM, int SHAPE, typename Hit>
__device__ __forceinline__ void scan_candidates(const NbArgs &A, const double (&g)[FG_MAXDIM], Hit &&hit) {
    int lo[FG_MAXDIM] = {0, 0, 0}, hi[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        lo[d] = fg_bin_coord(g[d] - A.reach[d], A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
        hi[d] = fg_bin_coord(g[d] + A.reach[d], A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
    }
    const int64_t n = A.nnodes;
    for (int bz = lo[2]; bz <= hi[2]; ++bz)
        for (int by = lo[1]; by <= hi[1]; ++by) {
            const int row = (bz * A.geom.nb[1] + by) * A.geom.nb[0];
            const int k0 = A.bin_start[row + lo[0]], k1 = A.bin_start[row + hi[0] + 1];
            for (int k = k0; k < k1; ++k) {
                bool in;
                if (SHAPE == FG_SHAPE_RECTANGULAR) {
                    in = fg_inside_1d(A.xs[k], g[0], A.dms[k]);
                    if (in) in = fg_inside_1d(A.xs[n + k], g[1], A.dms[n + k]);
                    if (DIM == 3) { if (in) in = fg_inside_1d(A.xs[2 * n + k], g[2], A.dms[2 * n + k]); }
                } else {
                    double dist_sq = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        const double t = __dsub_rn(A.xs[d * n + k], g[d]);
                        dist_sq = __dadd_rn(dist_sq, __dmul_rn(t, t));
                    }
                    const double r = A.dms[k];
                    in = dist_sq <= __dmul_rn(r, r);
                }
                if (in && A.mask) in = fg_mask_bit(A.mask, A.sorted_id[k]);
                if (in) hit(k);
            }
        }
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(128) k_nb_count(NbArgs A, int *__restrict__ cnt) {
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gi >= A.ngauss) return;
    double g[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) g[d] = A.gs[d * A.ngauss + gi];
    int c = 0;
    scan_candidates<DIM, SHAPE>(A, g, [&](int) { ++c; });
    cnt[gi] = c;
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(128) k_nb_fill(NbArgs A, const int64_t *__restrict__ off, int *__restrict__ dom, int cap) {
    extern __shared__ int smem[];
    const int bd = blockDim.x, tid = threadIdx.x;
    int *soff = smem;                 // bd + 1 offsets relative to the block's first entry
    int *slist = smem + bd + 1;       // cap x bd, slot-major
    const int64_t g0 = blockIdx.x * (int64_t)bd;
    const int64_t gi = g0 + tid;
    const int64_t base = off[g0];
    int L = 0;
    if (gi < A.ngauss) {
        double g[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < DIM; ++d) g[d] = A.gs[d * A.ngauss + gi];
        scan_candidates<DIM, SHAPE>(A, g, [&](int k) { if (L < cap) slist[L * bd + tid] = A.sorted_id[k]; ++L; });
        if (L > cap) L = cap;         // cannot happen: cap = max count of the count pass
        for (int k = 1; k < L; ++k) { // insertion sort (lists are short and nearly sorted)
            const int v = slist[k * bd + tid];
            int j = k - 1;
            while (j >= 0 && slist[j * bd + tid] > v) { slist[(j + 1) * bd + tid] = slist[j * bd + tid]; --j; }
            slist[(j + 1) * bd + tid] = v;
        }
        soff[tid] = (int)(off[gi] - base);
    }
    const int npts = (int)min((int64_t)bd, A.ngauss - g0);
    if (tid == 0) soff[npts] = (int)(off[g0 + npts] - base);
    __syncthreads();
    const int total = soff[npts];
    for (int e = tid; e < total; e += bd) {
        int lo = 0, hi = npts - 1;    // largest p with soff[p] <= e
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (soff[mid] <= e) lo = mid; else hi = mid - 1; }
        dom[base + e] = slist[(e - soff[lo]) * bd + lo];
    }
}

// Single-pass variant for a set whose neighbour lists were already computed once (buffers and the list
// capacity are known): every block keeps its lists in shared memory, obtains the offset of its slice with a
// decoupled look-back over the per-block totals (tickets are handed out in launch order, so a block only ever
// waits for blocks that are already running), and writes offsets and DOM in one go.
//   state[b] = (flag << 62) | value, flag 1 = block total, 2 = inclusive prefix
//   info[0] = ticket counter, info[1] = max L seen, info[2] = overflow (a list longer than cap / capacity)
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(128) k_nb_fused(NbArgs A, int64_t *__restrict__ off, int *__restrict__ dom, int cap, int64_t dom_cap,
                                                  unsigned long long *state, int *info) {
    extern __shared__ int smem[];
    const int bd = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int *soff = smem;                                   // bd + 1
    int *slist = smem + bd + 1;                         // cap x bd, slot-major
    unsigned char *sowner = (unsigned char *)(slist + (size_t)cap * bd);   // cap x bd
    __shared__ int s_ticket, s_wsum[4];
    __shared__ long long s_base;
    if (tid == 0) s_ticket = atomicAdd(&info[0], 1);
    __syncthreads();
    const int ticket = s_ticket;
    const int64_t g0 = ticket * (int64_t)bd;
    const int64_t gi = g0 + tid;
    const int npts = (int)min((int64_t)bd, A.ngauss - g0);
    int L = 0;
    if (gi < A.ngauss) {
        double g[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < DIM; ++d) g[d] = A.gs[d * A.ngauss + gi];
        scan_candidates<DIM, SHAPE>(A, g, [&](int k) { if (L < cap) slist[L * bd + tid] = A.sorted_id[k]; ++L; });
        if (L > cap) { atomicMax(&info[2], L); L = cap; }
        for (int k = 1; k < L; ++k) {
            const int v = slist[k * bd + tid];
            int j = k - 1;
            while (j >= 0 && slist[j * bd + tid] > v) { slist[(j + 1) * bd + tid] = slist[j * bd + tid]; --j; }
            slist[(j + 1) * bd + tid] = v;
        }
    }
    // block-level exclusive scan of L
    int incl = L;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
    if (lane == 31) s_wsum[wid] = incl;
    int mx = L;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    if (lane == 0 && mx > 0) atomicMax(&info[1], mx);
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_wsum[w];
    const int excl = wbase + incl - L;
    int total = 0;
    for (int w = 0; w < (bd >> 5); ++w) total += s_wsum[w];
    soff[tid] = excl;
    if (tid == 0) soff[bd] = total;
    for (int a = 0; a < L; ++a) sowner[excl + a] = (unsigned char)tid;
    // decoupled look-back (warp 0)
    if (wid == 0) {
        const unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, MASK = (1ull << 62) - 1;
        long long exclusive = 0;
        if (ticket == 0) {
            if (lane == 0) atomicExch(&state[0], FLAG_INC | (unsigned long long)total);
        } else {
            if (lane == 0) atomicExch(&state[ticket], FLAG_AGG | (unsigned long long)total);
            int j = ticket - 1;
            while (true) {
                const int idx = j - lane;
                unsigned long long st = FLAG_INC;                 // before the first block: inclusive 0
                if (idx >= 0) {
                    do { st = *((volatile unsigned long long *)&state[idx]); } while ((st >> 62) == 0);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                long long v = (long long)(st & MASK);
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    if (lane > first) v = 0;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
                    exclusive += v;
                    break;
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
                exclusive += v;
                j -= 32;
            }
            if (lane == 0) atomicExch(&state[ticket], FLAG_INC | (unsigned long long)(exclusive + total));
        }
        if (lane == 0) s_base = exclusive;
    }
    __syncthreads();
    const int64_t base = s_base;
    if (gi < A.ngauss) off[gi] = base + excl;
    if (g0 + npts == A.ngauss && tid == 0) off[A.ngauss] = base + total;
    if (base + total > dom_cap) { if (tid == 0) atomicMax(&info[2], 1 << 30); return; }
    for (int e = tid; e < total; e += bd) {
        const int p = sowner[e];
        dom[base + e] = slist[(e - soff[p]) * bd + p];
    }
}

// --------------------------------------------------------------------------------------------
// Cluster search (default): one warp per cluster of Gauss points (the cells of the cluster grid of
// fg_cluster.cu).  The nodes whose support box meets the cluster's bounding box are gathered from the node bins
// ONCE per cluster, rank-sorted by id and parked in shared memory with their coordinates and supports; every
// Gauss point of the cluster (one per lane) then runs the reference's predicate over that short list with
// broadcast shared-memory reads.  Because the list is id-sorted, the hits come out ascending: no per-point sort.
// Count pass -> exclusive scan -> fill pass (the fill pass reuses the sorted candidate lists written by the
// count pass).  A cluster with more than `capc` candidates raises a flag and the whole search is redone by the
// point-wise kernels above.
// --------------------------------------------------------------------------------------------
struct NbClArgs {
    const int *cstart, *csorted;     // cluster -> Gauss points
    int *cand;                       // per cluster: [0] n, [1..capc] sorted ids, [1+capc .. 1+2capc) bin positions
    unsigned *masks;                 // per Gauss point: capc/32 words, bit b = candidate b of its cluster is a neighbour
    int capc;
};

template <int DIM, int SHAPE, bool FILL>
__global__ void __launch_bounds__(32) k_nb_cluster(NbArgs A, NbClArgs C, int64_t nclusters, int *__restrict__ cnt,
                                                   const int64_t *__restrict__ off, int *__restrict__ dom, int *__restrict__ flag) {
    constexpr int NDM = SHAPE == FG_SHAPE_RECTANGULAR ? DIM : 1;
    constexpr int RS = DIM + NDM;                          // doubles per candidate record
    extern __shared__ double smd[];
    const int lane = threadIdx.x;
    const int capc = C.capc;
    double *srec = smd;                                    // capc x RS
    int *sid = (int *)(smd + (size_t)capc * RS);           // capc sorted ids
    int *spos = sid + capc;                                // capc bin positions (unsorted during the gather)
    int *stmp = spos + capc;                               // capc ids (unsorted during the gather)
    const int64_t n = A.nnodes;
    for (int64_t c = blockIdx.x; c < nclusters; c += gridDim.x) {
        const int p0 = C.cstart[c], p1 = C.cstart[c + 1];
        if (p0 == p1) continue;
        int *cand = C.cand + c * (int64_t)(1 + 2 * capc);
        int nc;
        __syncwarp();
        if (!FILL) {
            // ---- bounding box of the cluster's points ------------------------------------------------------
            double lo[DIM], hi[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) { lo[d] = 1e300; hi[d] = -1e300; }
            for (int p = p0 + lane; p < p1; p += 32) {
                const int g = C.csorted[p];
#pragma unroll
                for (int d = 0; d < DIM; ++d) { const double v = A.gs[d * A.ngauss + g]; lo[d] = fmin(lo[d], v); hi[d] = fmax(hi[d], v); }
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d)
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) { lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], sft)); hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], sft)); }
            // ---- gather: nodes whose support box meets the box (conservative; the exact predicate comes later) ---
            int blo[FG_MAXDIM] = {0, 0, 0}, bhi[FG_MAXDIM] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                blo[d] = fg_bin_coord(lo[d] - A.reach[d], A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
                bhi[d] = fg_bin_coord(hi[d] + A.reach[d], A.geom.origin[d], A.geom.inv_size[d], A.geom.nb[d]);
            }
            nc = 0;
            // the (y,z) bin rows of the box: their node ranges are fetched 32 rows at a time (one per lane), then the
            // concatenated ranges are walked 32 nodes at a time -- few dependent global-load rounds per cluster
            const int nry = bhi[1] - blo[1] + 1, nrows = nry * (bhi[2] - blo[2] + 1);
            for (int r0 = 0; r0 < nrows; r0 += 32) {
                const int r = r0 + lane;
                int k0 = 0, len = 0;
                if (r < nrows) {
                    const int by = blo[1] + r % nry, bz = blo[2] + r / nry;
                    const int row = (bz * A.geom.nb[1] + by) * A.geom.nb[0];
                    k0 = A.bin_start[row + blo[0]];
                    len = A.bin_start[row + bhi[0] + 1] - k0;
                }
                int incl = len;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - len;
                for (int i0 = 0; i0 < tot; i0 += 32) {
                    const int i = i0 + lane;
                    // row of flat index i: the last lane whose exclusive prefix is <= i
                    int rr = 0;
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
                            const double xa = A.xs[d * n + k], da = A.dms[(d < NDM ? d : 0) * n + k] * (1.0 + 1e-9);
                            keep = keep & (xa + da >= lo[d]) & (xa - da <= hi[d]);      // no short-circuit: the loads of all dimensions are issued together
                        }
                    }
                    int kid = 0;
                    if (keep) { kid = A.sorted_id[k]; if (A.mask) keep = fg_mask_bit(A.mask, kid); }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) { const int pos = nc + __popc(m & ((1u << lane) - 1)); if (pos < capc) { spos[pos] = k; stmp[pos] = kid; } }
                    nc += __popc(m);
                }
            }
            if (nc > capc) { if (lane == 0) atomicMax(flag, nc); if (lane == 0) cand[0] = -1; continue; }
            if (lane == 0) { atomicMax(flag + 1, p1 - p0); atomicMax(flag + 2, nc); }        // sizes the cluster-centric IMLS launch
            __syncwarp();
            // ---- rank sort by node id (ids are distinct), publish the sorted list for the fill pass --------------
            for (int e = lane; e < nc; e += 32) {
                const int v = stmp[e];
                int rank = 0;
                for (int t = 0; t < nc; ++t) rank += stmp[t] < v;
                sid[rank] = v;
                cand[1 + rank] = v;
                cand[1 + capc + rank] = spos[e];
            }
            if (lane == 0) cand[0] = nc;
            __syncwarp();
            for (int e = lane; e < nc; e += 32) spos[e] = cand[1 + capc + e];       // now sorted order (own writes, same lane mapping not guaranteed)
        } else {
            nc = cand[0];
            if (nc < 0) continue;
            for (int e = lane; e < nc; e += 32) { sid[e] = cand[1 + e]; spos[e] = cand[1 + capc + e]; }
        }
        __syncwarp();
        if (FILL) {
            // the count pass left one hit mask per point: expand it against the id-sorted candidate list
            // into a staging buffer (the record area, unused in this pass); the warp then copies the lists out with
            // full-line stores (a lane writing its own list issues one 4-byte store per entry: 5e8 partial-sector writes).
            constexpr int NW = 4;                                  // capc / 32
            int *stage = reinterpret_cast<int *>(srec);
            const int stage_cap = capc * RS * 2;
            // header of a chunk (point id, list offset, the four mask words as one 16-byte load); the next chunk's
            // header is requested before the current chunk is expanded
            auto header = [&](int pb, int64_t &o, uint4 &mk) {
                const int p = pb + lane;
                o = 0; mk = make_uint4(0u, 0u, 0u, 0u);
                if (p < p1) {
                    const int g = C.csorted[p];
                    o = off[g];
                    mk = *reinterpret_cast<const uint4 *>(C.masks + (int64_t)g * NW);
                }
            };
            int64_t o_next; uint4 mk_next;
            header(p0, o_next, mk_next);
            for (int pb = p0; pb < p1; pb += 32) {
                const bool live = pb + lane < p1;
                const int64_t o = o_next;
                const uint4 mk = mk_next;
                if (pb + 32 < p1) header(pb + 32, o_next, mk_next);
                unsigned m[NW] = {mk.x, mk.y, mk.z, mk.w};
                int L = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) L += __popc(m[w]);
                // compact staging: lane q's list sits at the exclusive prefix of the list lengths; runs of lanes whose
                // DOM ranges are adjacent (consecutive Gauss points) are copied out as one range
                int incl = L;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, sft); if (lane >= sft) incl += t; }
                const int tot = __shfl_sync(0xffffffffu, incl, 31), pre = incl - L;
                if (tot <= stage_cap) {
                    int k = pre;
#pragma unroll
                    for (int w = 0; w < NW; ++w) { unsigned mm = m[w]; while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1; stage[k++] = sid[32 * w + b]; } }
                    const int64_t prev_end = __shfl_up_sync(0xffffffffu, o + L, 1);
                    unsigned runs = __ballot_sync(0xffffffffu, live && (lane == 0 || o != prev_end));
                    __syncwarp();
                    while (runs) {
                        const int a = __ffs(runs) - 1;
                        runs &= runs - 1;
                        const int bnext = runs ? __ffs(runs) - 1 : 32;               // first lane of the next run
                        const int src0 = __shfl_sync(0xffffffffu, pre, a);
                        const int src1 = bnext < 32 ? __shfl_sync(0xffffffffu, pre, bnext) : tot;
                        const int64_t dst = __shfl_sync(0xffffffffu, o, a);
                        for (int i = lane; i < src1 - src0; i += 32) dom[dst + i] = stage[src0 + i];
                    }
                    __syncwarp();
                } else if (live) {
                    int64_t oo = o;
#pragma unroll
                    for (int w = 0; w < NW; ++w) { unsigned mm = m[w]; while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1; dom[oo++] = sid[32 * w + b]; } }
                }
            }
            continue;
        }
        for (int e = lane; e < nc; e += 32) {
            const int k = FILL ? spos[e] : cand[1 + capc + e];
#pragma unroll
            for (int d = 0; d < DIM; ++d) srec[e * RS + d] = A.xs[d * n + k];
#pragma unroll
            for (int d = 0; d < NDM; ++d) srec[e * RS + DIM + d] = A.dms[d * n + k];
        }
        __syncwarp();
        // ---- the reference predicate, one Gauss point per lane, candidates broadcast from shared memory ------
        for (int pb = p0; pb < p1; pb += 32) {
            const int p = pb + lane;
            const bool live = p < p1;
            const int g = live ? C.csorted[p] : 0;
            double gp[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) gp[d] = live ? A.gs[d * A.ngauss + g] : 0.0;
            int64_t o = 0;
            if (FILL && live) o = off[g];
            int total = 0;
            for (int w0 = 0; w0 < nc; w0 += 32) {
                unsigned mask = 0;
                const int wend = min(32, nc - w0);
#pragma unroll 8
                for (int b = 0; b < wend; ++b) {           // independent tests: unrolled so that their latencies overlap
                    const double *r = srec + (w0 + b) * RS;
                    bool in;
                    if (SHAPE == FG_SHAPE_RECTANGULAR) {
                        // all dimensions are tested unconditionally: the lanes (different points) disagree on where a
                        // short-circuit would stop, so a branch only adds divergence
                        const bool i0 = fg_inside_1d(r[0], gp[0], r[DIM]);
                        const bool i1 = fg_inside_1d(r[1], gp[1], r[DIM + 1]);
                        const bool i2 = DIM == 3 ? fg_inside_1d(r[DIM - 1], gp[DIM - 1], r[2 * DIM - 1]) : true;
                        in = i0 & i1 & i2;
                    } else {
                        double dist_sq = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) { const double t = __dsub_rn(r[d], gp[d]); dist_sq = __dadd_rn(dist_sq, __dmul_rn(t, t)); }
                        in = dist_sq <= __dmul_rn(r[DIM], r[DIM]);
                    }
                    mask |= (in ? 1u : 0u) << b;
                }
                total += __popc(mask);
                if (live) C.masks[(int64_t)g * (capc >> 5) + (w0 >> 5)] = mask;
            }
            if (live) for (int w = (nc + 31) >> 5; w < (capc >> 5); ++w) C.masks[(int64_t)g * (capc >> 5) + w] = 0u;
            if (!FILL && live) cnt[g] = total;
        }
    }
}

template <int DIM, int SHAPE>
static int run_cluster(fg_ctx *ctx, GaussSet &S, const NbArgs &A, bool *done) {
    *done = false;
    // the search wants coarser cells than the assembly (fewer candidate gathers per point): its own grid
    Clusters &Cl = S.nbcl;
    if (fg_opt(ctx, "nb_rebin", 0) != 0) Cl.binned = false;      // benchmark honesty: a search on NEW inputs pays for the binning too
    int rc = fg_cluster_binning(ctx, S, Cl, fg_opt(ctx, "nb_ucap", 64));
    if (rc) return rc;
    constexpr int NDM = SHAPE == FG_SHAPE_RECTANGULAR ? DIM : 1;
    const int capc = 128;                          // candidates per cluster held in shared memory (more -> point-wise fallback)
    const size_t smem = sizeof(double) * (size_t)capc * (DIM + NDM) + sizeof(int) * 3 * capc;
    FG_CUDA(ctx, S.cand.reserve((size_t)Cl.n * (1 + 2 * capc)));
    FG_CUDA(ctx, S.masks.reserve((size_t)S.n * (capc / 32)));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p + 12, 0, 3 * sizeof(int), ctx->stream));
    NbClArgs C;
    C.cstart = Cl.grid.start.p; C.csorted = Cl.sorted.p; C.cand = S.cand.p; C.masks = S.masks.p; C.capc = capc;
    const unsigned grid = (unsigned)std::min<int64_t>(Cl.n, (int64_t)ctx->sm_count * 32);
    FG_CUDA(ctx, cudaFuncSetAttribute(k_nb_cluster<DIM, SHAPE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FG_CUDA(ctx, cudaFuncSetAttribute(k_nb_cluster<DIM, SHAPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nb_cluster<DIM, SHAPE, false><<<grid, 32, smem, ctx->stream>>>(A, C, Cl.n, S.cnt.p, nullptr, nullptr, ctx->flags.p + 12);
    FG_CHECK_LAUNCH(ctx);
    rc = fg_exclusive_scan_i32_to_i64(ctx, S.cnt.p, S.off.p, S.n);
    if (rc) return rc;
    size_t bytes = 0;
    FG_CUDA(ctx, cub::DeviceReduce::Max(nullptr, bytes, S.cnt.p, ctx->flags.p, (int)S.n, ctx->stream));
    FG_CUDA(ctx, ctx->temp.reserve(bytes));
    FG_CUDA(ctx, cub::DeviceReduce::Max(ctx->temp.p, bytes, S.cnt.p, ctx->flags.p, (int)S.n, ctx->stream));
    int maxL = 0; int ovf3[3] = {0, 0, 0}; int64_t total = 0;
    FG_CUDA(ctx, cudaMemcpyAsync(&maxL, ctx->flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(ovf3, ctx->flags.p + 12, sizeof ovf3, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(&total, S.off.p + S.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int ovf = ovf3[0];
    if (ovf > 0) return FG_OK;                     // a cluster has too many candidates: the caller falls back
    S.total_L = total; S.max_L = maxL; S.nb_max_points = ovf3[1]; S.nb_max_cand = ovf3[2];
    FG_CUDA(ctx, S.dom.reserve(total));
    S.nb_clustered = true; S.nb_capc = capc;
    S.dom_pending = total > 0;
    // The fill pass (hit masks -> DOM) is deferred: the cluster-centric IMLS kernel writes DOM together with phi / dphi while it
    // walks the very same masks (fg_shape_functions), and every other consumer materialises it on demand (fg_ensure_dom).
    if (fg_opt(ctx, "nb_lazy_dom", 1) == 0) { int rc2 = fg_ensure_dom(ctx, S); if (rc2) return rc2; }
    *done = true;
    return FG_OK;
}

template <int DIM, int SHAPE>
static int fill_cluster(fg_ctx *ctx, GaussSet &S, const NbArgs &A) {
    Clusters &Cl = S.nbcl;
    constexpr int NDM = SHAPE == FG_SHAPE_RECTANGULAR ? DIM : 1;
    const int capc = S.nb_capc;
    const size_t smem = sizeof(double) * (size_t)capc * (DIM + NDM) + sizeof(int) * 3 * capc;
    NbClArgs C;
    C.cstart = Cl.grid.start.p; C.csorted = Cl.sorted.p; C.cand = S.cand.p; C.masks = S.masks.p; C.capc = capc;
    const unsigned grid = (unsigned)std::min<int64_t>(Cl.n, (int64_t)ctx->sm_count * 32);
    FG_CUDA(ctx, cudaFuncSetAttribute(k_nb_cluster<DIM, SHAPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nb_cluster<DIM, SHAPE, true><<<grid, 32, smem, ctx->stream>>>(A, C, Cl.n, nullptr, S.off.p, S.dom.p, ctx->flags.p + 12);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

template <int DIM, int SHAPE>
static int run(fg_ctx *ctx, GaussSet &S, const NbArgs &A) {
    const int bd = 128;
    const unsigned grid = (unsigned)((S.n + bd - 1) / bd);
    if (fg_opt(ctx, "nb_pointwise", 0) == 0) {              // default: cluster search; point-wise kernels when a cluster overflows
        bool done = false;
        int rcc = run_cluster<DIM, SHAPE>(ctx, S, A, &done);
        if (rcc) return rcc;
        if (done) return FG_OK;
    }
    // ---- fast path: same set computed before -> one pass with a decoupled look-back ------------------------
    const int known_L = S.max_L > 0 ? S.max_L : S.max_L_hint;
    if (known_L > 0 && S.dom.cap > 0 && fg_opt(ctx, "nb_fused", 0) != 0) {   // opt-in: measured 13.0 ms vs 11.3 ms two-pass at 2.7e7 points (round 1)
        const int cap = known_L + 4;
        const size_t smem_f = sizeof(int) * ((size_t)bd + 1 + (size_t)cap * bd) + (size_t)cap * bd;
        if (smem_f <= 200 * 1024) {
            FG_CUDA(ctx, S.lookback.reserve(grid + 8));
            FG_CUDA(ctx, cudaMemsetAsync(S.lookback.p, 0, sizeof(unsigned long long) * (grid + 8), ctx->stream));
            int *info = (int *)(S.lookback.p + grid + 2);
            FG_CUDA(ctx, cudaFuncSetAttribute(k_nb_fused<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
            k_nb_fused<DIM, SHAPE><<<grid, bd, smem_f, ctx->stream>>>(A, S.off.p, S.dom.p, cap, (int64_t)S.dom.cap, S.lookback.p, info);
            FG_CHECK_LAUNCH(ctx);
            int h[4] = {0, 0, 0, 0}; int64_t total = 0;
            FG_CUDA(ctx, cudaMemcpyAsync(h, info, sizeof(int) * 3, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaMemcpyAsync(&total, S.off.p + S.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (h[2] == 0) { S.total_L = total; S.max_L = h[1]; return FG_OK; }
            // a list outgrew the capacity assumed from the previous call: redo with the two-pass path below
        }
    }
    k_nb_count<DIM, SHAPE><<<grid, bd, 0, ctx->stream>>>(A, S.cnt.p);
    FG_CHECK_LAUNCH(ctx);
    int rc = fg_exclusive_scan_i32_to_i64(ctx, S.cnt.p, S.off.p, S.n);
    if (rc) return rc;
    // max L and total L
    FG_CUDA(ctx, ctx->flags.reserve(64));
    size_t bytes = 0;
    FG_CUDA(ctx, cub::DeviceReduce::Max(nullptr, bytes, S.cnt.p, ctx->flags.p, (int)S.n, ctx->stream));
    FG_CUDA(ctx, ctx->temp.reserve(bytes));
    FG_CUDA(ctx, cub::DeviceReduce::Max(ctx->temp.p, bytes, S.cnt.p, ctx->flags.p, (int)S.n, ctx->stream));
    int maxL = 0; int64_t total = 0;
    FG_CUDA(ctx, cudaMemcpyAsync(&maxL, ctx->flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(&total, S.off.p + S.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.total_L = total; S.max_L = maxL;
    FG_CUDA(ctx, S.dom.reserve(total));
    if (total == 0) return FG_OK;
    // shared memory: (bd+1) + cap*bd ints; shrink the block for very long lists
    int fbd = bd;
    size_t smem = sizeof(int) * ((size_t)fbd + 1 + (size_t)maxL * fbd);
    while (smem > 200 * 1024 && fbd > 32) { fbd >>= 1; smem = sizeof(int) * ((size_t)fbd + 1 + (size_t)maxL * fbd); }
    if (smem > 200 * 1024)
        return fg_fail(ctx, FG_E_CAPACITY, "fg_neighbours: a Gauss point has %d neighbours; the limit is %d", maxL, (200 * 1024 / 4 - 33) / 32);
    FG_CUDA(ctx, cudaFuncSetAttribute(k_nb_fill<DIM, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nb_fill<DIM, SHAPE><<<(unsigned)((S.n + fbd - 1) / fbd), fbd, smem, ctx->stream>>>(A, S.off.p, S.dom.p, maxL);
    FG_CHECK_LAUNCH(ctx);
    return FG_OK;
}

static NbArgs nb_args(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    NbArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.xs = N.xs.p; A.dms = N.dms.p; A.sorted_id = N.sorted_id.p; A.bin_start = N.bins.start.p;
    A.gs = S.gs.p; A.geom = fg_geom(N.bins);
    A.mask = S.nsub > 0 ? S.nmask.p : nullptr;
    for (int d = 0; d < FG_MAXDIM; ++d) A.reach[d] = N.reach[d];
    return A;
}

int fg_ensure_dom(fg_ctx *ctx, GaussSet &S) {
    if (!S.dom_pending) return FG_OK;
    if (!S.nb_clustered) return fg_fail(ctx, FG_E_STATE, "neighbour lists pending without the cluster search's masks");
    NodeSet &N = ctx->nodes;
    const NbArgs A = nb_args(ctx, S);
    int rc;
    if (N.dim == 2) rc = N.shape == FG_SHAPE_RECTANGULAR ? fill_cluster<2, FG_SHAPE_RECTANGULAR>(ctx, S, A) : fill_cluster<2, FG_SHAPE_CIRCULAR>(ctx, S, A);
    else rc = N.shape == FG_SHAPE_RECTANGULAR ? fill_cluster<3, FG_SHAPE_RECTANGULAR>(ctx, S, A) : fill_cluster<3, FG_SHAPE_CIRCULAR>(ctx, S, A);
    if (rc) return rc;
    S.dom_pending = false;
    return FG_OK;
}

int fg_impl_neighbours(fg_ctx *ctx, GaussSet &S) {
    NodeSet &N = ctx->nodes;
    if (!N.ready) return fg_fail(ctx, FG_E_STATE, "fg_neighbours: call fg_set_nodes first");
    S.have_nb = S.have_sf = false;   // the sparsity depends on (nodes, gs) only and stays valid
    S.sf_lazy = false;
    S.nb_clustered = false; S.dom_pending = false;
    // (the cluster index of kernel 3 also survives: DOM is a deterministic function of (nodes, gs))
    FG_CUDA(ctx, S.off.reserve(S.n + 1));
    FG_CUDA(ctx, S.cnt.reserve(S.n));
    if (S.n == 0) {
        FG_CUDA(ctx, cudaMemsetAsync(S.off.p, 0, sizeof(int64_t), ctx->stream));
        S.total_L = 0; S.max_L = 0; S.have_nb = true;
        return FG_OK;
    }
    NbArgs A;
    A.nnodes = N.n; A.ngauss = S.n; A.xs = N.xs.p; A.dms = N.dms.p; A.sorted_id = N.sorted_id.p; A.bin_start = N.bins.start.p;
    A.gs = S.gs.p; A.geom = fg_geom(N.bins);
    A.mask = S.nsub > 0 ? S.nmask.p : nullptr;
    if (S.nsub > 1) return fg_fail(ctx, FG_E_STATE, "fg_neighbours: set is a concatenation of sets with different node subsets");
    for (int d = 0; d < FG_MAXDIM; ++d) A.reach[d] = N.reach[d];
    int rc;
    if (N.dim == 2) rc = N.shape == FG_SHAPE_RECTANGULAR ? run<2, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<2, FG_SHAPE_CIRCULAR>(ctx, S, A);
    else rc = N.shape == FG_SHAPE_RECTANGULAR ? run<3, FG_SHAPE_RECTANGULAR>(ctx, S, A) : run<3, FG_SHAPE_CIRCULAR>(ctx, S, A);
    if (rc) return rc;
    S.have_nb = true;
    return FG_OK;
}

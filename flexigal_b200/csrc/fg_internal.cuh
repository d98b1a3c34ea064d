// fg_internal.cuh -- context, device containers and helpers shared by the kernels.
// Nothing here is part of the ABI (include/flexigal_gpu.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <map>
#include <vector>
#include "../../include/flexigal_gpu.h"

#define FG_MAXDIM 3

// Device buffer that only grows (re-used across calls: Newton steps re-assemble every iteration).
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) cap = n ? n : 1;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Uniform bin grid over a point cloud: bin b holds sorted[start[b] .. start[b+1]).
struct BinGrid {
    double origin[FG_MAXDIM] = {0, 0, 0};
    double inv_size[FG_MAXDIM] = {0, 0, 0};
    int nb[FG_MAXDIM] = {1, 1, 1};
    int64_t nbins = 0;
    DevBuf<int> start;   // nbins + 1
};

struct NodeSet {
    int dim = 0, shape = 0;
    int64_t n = 0;
    DevBuf<double> x, dm;      // as given (column-major): x[d*n+i]; dm[d*n+i] or dm[i]
    DevBuf<double> idm;        // 1/dm, same layout as dm
    DevBuf<double> nrec;       // AoS per node: [x_0..x_{dim-1}, 1/dm_0..1/dm_{dim-1}] (circular: the radius repeated)
    double reach[FG_MAXDIM] = {0, 0, 0};   // max support half-width per dim (slightly widened)
    double ext[FG_MAXDIM] = {0, 0, 0};     // bounding-box extent
    BinGrid bins;
    DevBuf<int> sorted_id;     // node ids ordered by bin (ascending id inside a bin)
    DevBuf<double> xs, dms;    // x / dm gathered in bin order: xs[d*n+k]
    bool ready = false;
};

struct Pattern {
    int64_t nnz = 0;           // node-level
    DevBuf<int64_t> colptr;    // nnodes+1, 0-based
    DevBuf<int> rowval;        // nnz, 0-based, ascending per column
    DevBuf<double> nzval;      // nnz*D*D (last assembly)
    int D = 0;
    int max_col = 0;
    bool ready = false;
    DevBuf<int> tmap;          // for every entry below the diagonal the position of its transpose, else -1 (built on first use)
    bool tmap_ready = false;
    uint64_t version = 0;                     // bumped by every rebuild (peer mappings of the column exchange carry offsets into it)
    std::map<int64_t, int64_t> colptr_host;   // colptr values already fetched by the column-range entry points (cleared with the pattern)
};


// Spatial clusters of Gauss points (kernel 3 reduces a cluster's contributions in shared memory
// before touching global memory): cluster c owns points sorted[start[c]..start[c+1]) and the
// sorted union `nodes[c*ucap .. +nU[c])` of their neighbour lists; lidx[e] is the position of
// DOM entry e in that union.  nU[c] < 0 marks a cluster whose union did not fit (its points go
// through the direct-atomics kernel instead).
struct Clusters {
    bool ready = false;      // union / lidx valid for the current DOM
    bool binned = false;     // grid / sorted valid for the current (nodes, gs)
    int ucap = 0;
    int cells_for = 0;       // the cell size was chosen so that a REGULAR lattice fills at most this many nodes (<= ucap: headroom for irregular clouds)
    int64_t n = 0;
    BinGrid grid;
    double csize[FG_MAXDIM] = {0, 0, 0}, cshift[FG_MAXDIM] = {0, 0, 0};   // cell size and lattice shift the grid was built with
    DevBuf<int> sorted;
    DevBuf<int> nU;
    DevBuf<int> nodes;
    DevBuf<unsigned char> lidx;
    DevBuf<unsigned> inv;        // 32-node blocks only: per Gauss point 8 words; byte r of word gid = position of local node 8r+gid in the
                                 // point's neighbour list, 255 = not a neighbour (the gather map of k_bilinear_pts)
    DevBuf<int> ovf_points;      // [0] = count, then point ids; the slot after the ids keeps the count of the index build
                                 // (an assembly call appends its own long-list points and is reset to that count first)
    DevBuf<int64_t> col0;        // n*ucap: CSC start of every block column (flush table, needs the pattern)
    DevBuf<int> coln;            // n*ucap: CSC length of every block column
    DevBuf<unsigned char> tab;   // n*ucap*ucap: position of block row i inside block column j (255 = absent)
    bool tab_ready = false;
    bool lidx_ready = false;     // lidx written (32-node blocks built from the search's masks leave it to fg_ensure_lidx)
    bool from_masks = false;     // built by k_cluster_build_masks (the search's candidate lists / hit masks are the source)
    int pmask_many = -1;         // 1: some cluster has more than 32 points (masks unusable), 0: fine, -1: not read back yet
    int want = 0;            // block size the index was asked for (ucap is 64 when a 32-node index overflowed too often)
    bool tab_upper = false;  // tables hold rows <= column only (enough for the symmetric kernels)
    int64_t stat_overflow_points = -1;
    int stat_max_u = -1;
    // streamed download (fg_assemble_bilinear_async): the clusters are assembled in FG_STREAM_RANGES launches; after launch i every column
    // below stream_cols[i+1] is final (no later cluster holds the node), so it is mirrored and copied out while the next launch runs
    std::vector<int64_t> stream_cols, stream_ent;      // FG_STREAM_RANGES+1 column bounds and their colptr values
    bool stream_ready = false;
    int stream_novf = 0;         // points the index build left to the direct kernel (streaming needs 0)
    uint64_t stream_pattern = 0; // pattern version the bounds belong to
};
#define FG_STREAM_RANGES 8

struct GaussSet {
    int64_t n = 0;
    DevBuf<double> gs;         // gs[c*n+g], c < dim+2
    DevBuf<int64_t> off;       // n+1
    DevBuf<int> cnt;           // n (scratch: counts)
    DevBuf<unsigned long long> lookback;   // per-block scan state of the single-pass neighbour kernel
    DevBuf<unsigned> masks;    // per Gauss point: hit mask over its cluster's candidate list (count pass -> fill pass)
    DevBuf<int> cand;          // per cluster: candidate nodes of the cluster search ([0] count, then sorted ids, then bin positions)
    DevBuf<int> dom;           // total_L, 0-based ascending per point
    DevBuf<double> phi, dphi;  // total_L, total_L*dim
    int64_t total_L = 0;
    int max_L = 0;
    int max_L_hint = 0;        // longest list seen on the previous Gauss table of this set (capacity hint)
    bool have_nb = false, have_sf = false;
    bool sf_lazy = false;       // option "lazy_shapes": fg_shape_functions(IMLS) only set the buffers up; search clusters [0, sf_done) are computed,
    int64_t sf_done = 0;        // the rest follows on demand (fg_ensure_shapes) or range by range inside the streamed assembly
    bool nb_clustered = false;  // DOM came from the cluster search: cand / masks / nbcl describe it (the cluster-centric IMLS kernel reads them)
    int nb_capc = 0;
    int nb_max_points = 0, nb_max_cand = 0;   // largest search cluster (points / candidate nodes), from the count pass
    bool dom_pending = false;   // offsets are final, DOM not yet written: the cluster-centric IMLS kernel writes it with phi / dphi (its phase B
                                // walks the same hit masks), any other consumer calls fg_ensure_dom first (the search's fill pass, on demand)
    int64_t singular = 0;      // host copy, valid after fg_singular_count
    DevBuf<int> singular_dev;  // device counter written by the shape-function kernels
    DevBuf<int> imls_perm;     // Gauss points of every 1024-point tile ordered by neighbour count (k_imls_order)
    BinGrid gbins;             // Gauss points binned on the node grid geometry (pattern build)
    DevBuf<int> gsorted;       // Gauss ids in bin order
    bool gbins_ready = false;
    Pattern pat;
    Clusters cl;               // cluster index of the assembly kernels
    Clusters nbcl;             // cluster grid of the neighbour search (binning only; coarser cells: fewer gathers per point)
    DevBuf<double> fvec; int fvec_D = 0;
    // Node subset(s) of the set (fg_set_node_subset = build_space's x[global_node_ids,:], src/FlexiGal.jl:250-252):
    // nsub bit masks of mask_words words each, back to back; bit i of mask k = node i may be a neighbour of the points
    // of subset k.  A plain set has nsub <= 1 and psub empty; a concatenation of sets with different subsets keeps one
    // mask per source and psub[g] = the mask of Gauss point g.
    DevBuf<unsigned> nmask;
    DevBuf<unsigned char> psub;
    int nsub = 0;
    int64_t mask_words = 0;
    // FEM sets (fg_generate_fem): runs of Gauss points whose lists are the connectivity of their background cell (npts points per cell,
    // DOM in vertex order); empty for EFG sets.  The sparsity of such a set comes from DOM (fg_impl_pattern_fem).
    struct FemSeg { int64_t first = 0, count = 0; int npts = 1; };
    std::vector<FemSeg> fem;
    // multi-GPU column exchange (fg_exchange_columns): receive staging of the NCCL path, peer mappings of the direct path
    DevBuf<double> xbuf;
    struct PeerMap { int peer = -1; void *mapped = nullptr; int64_t s0 = 0, s1 = 0; cudaIpcMemHandle_t h; };     // the peer's buffer in my address space; its send range for me
    struct XchgState { const void *my_base = nullptr; uint64_t version = 0; std::vector<PeerMap> peers; } xchg[2];   // [0] nzval, [1] fvec
};

struct fg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // downloads that overlap the next kernels (fg_get_values_async); joined by fg_sync
    cudaEvent_t copy_event = nullptr;
    bool copy_pending = false;
    double *stream_to = nullptr;     // fg_assemble_bilinear_async: host destination of the streamed download (consumed by fg_impl_bilinear)
    bool streamed = false;           // ... and whether that call did stream
    std::string err;
    int64_t launches = 0;            // kernels of this library launched so far (cub helpers not counted)
    NodeSet nodes;
    std::map<int, GaussSet *> sets;
    DevBuf<unsigned char> temp;      // cub scratch
    DevBuf<int> flags;               // small device scratch for counters / flags
    DevBuf<double> coef;             // operator coefficients
    DevBuf<unsigned char> stage;     // conversion staging for getters
    // grow-only work buffers of the binning and the sparsity build: kept for the life of the context, because a
    // cudaMalloc/cudaFree pair of several hundred MB per call costs tens of milliseconds and synchronises the device
    DevBuf<int> bin_key_in, bin_key_out, bin_val_in, pat_cnt, pat_rows;
    DevBuf<unsigned> pat_pmask;      // per cluster and local node: which of the cluster's points hold the node (sparsity from the cluster index)
    const void *pmask_last = nullptr; // the set whose index build (k_cluster_build_masks) wrote pat_pmask last; nullptr: nobody's
    const void *pmask_ptr = nullptr;  // ... and the allocation it wrote (a grown buffer is a new, empty one)
    DevBuf<double> pat_gxs;          // Gauss-point coordinates in Gauss-bin order (witness search of the sparsity build)
    DevBuf<unsigned char> pat_gsub;  // subset index of the Gauss points in the same order (merged sets with node subsets)
    DevBuf<int64_t> ids64;           // upload staging of fg_set_node_subset
    DevBuf<double> field_u, field_val, field_grad;   // fg_evaluate_field staging (grow-only)
    struct SolveWork { DevBuf<double> v[7]; DevBuf<int64_t> cp; DevBuf<int> ri; } solve;   // PCG work vectors (grow-only)
    std::map<std::string, int> opts; // fg_set_option (kernel-selection switches; tuning builds also seed them from FG_* environment variables)
    void *comm = nullptr;            // ncclComm_t of fg_comm_init (multi-GPU exchange inside the library)
    int comm_rank = 0, comm_world = 1;
    DevBuf<unsigned long long> fem_keys_in, fem_keys_out;   // (column,row) pairs of the FEM sparsity build
    DevBuf<unsigned char> xmsg;      // device staging of the mapping messages / handshake tokens of the direct exchange
};

int fg_fail(fg_ctx *ctx, int code, const char *fmt, ...);
// value of a kernel-selection switch (fg_set_option), or `dflt` when it was never set
inline int fg_opt(const fg_ctx *ctx, const char *key, int dflt) {
    auto it = ctx->opts.find(key);
    return it == ctx->opts.end() ? dflt : it->second;
}
// profiling switches inside kernels exist in tuning builds only (make TUNING=1): release kernels carry no such branches
#ifdef FG_TUNING
#define FG_DBG(x) (x)
#else
#define FG_DBG(x) 0
#endif

#define FG_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fg_fail(ctx, e__ == cudaErrorMemoryAllocation ? FG_E_OOM : FG_E_CUDA,     \
                           "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define FG_CHECK_LAUNCH(ctx) do { ++(ctx)->launches; FG_CUDA(ctx, cudaGetLastError()); } while (0)

GaussSet *fg_find_set(fg_ctx *ctx, int set_id);

// Host-side helpers implemented in fg_api.cu
int fg_build_bins(fg_ctx *ctx, int dim, int64_t n, const double *coords /*device, [d*n+i]*/,
                  const double *origin, const double *inv_size, const int *nb,
                  BinGrid &grid, DevBuf<int> &sorted);
int fg_exclusive_scan_i32_to_i64(fg_ctx *ctx, const int *in, int64_t *out, int64_t n);  // out[n] = total
int fg_colptr_at(fg_ctx *ctx, Pattern &P, const int64_t *cols, int n, int64_t *out);          // colptr[cols[k]] on the host (cached)

// implemented in the kernel translation units
int fg_impl_neighbours(fg_ctx *ctx, GaussSet &gs);
int fg_ensure_dom(fg_ctx *ctx, GaussSet &gs);               // materialise DOM if the search left it pending
int fg_impl_imls(fg_ctx *ctx, GaussSet &gs);
bool fg_imls_rangeable(const fg_ctx *ctx, const GaussSet &gs);                 // the cluster IMLS kernel applies (ranged launches possible)
int fg_impl_imls_range(fg_ctx *ctx, GaussSet &gs, int64_t c_lo, int64_t c_hi);  // IMLS of the search clusters [c_lo, c_hi)
int fg_defer_imls(fg_ctx *ctx, GaussSet &gs);                                  // option "lazy_shapes": set up, compute later
int fg_ensure_shapes(fg_ctx *ctx, GaussSet &gs);                               // compute whatever a deferred fg_shape_functions left
int64_t fg_search_clusters_below(const fg_ctx *ctx, const GaussSet &gs, int64_t c1);   // search clusters that cover the assembly clusters [0, c1)
int fg_impl_mk(fg_ctx *ctx, GaussSet &gs);
int fg_impl_pattern(fg_ctx *ctx, GaussSet &gs, GaussSet &witness_points);
int fg_impl_pattern_fem(fg_ctx *ctx, GaussSet &gs);
int fg_impl_bilinear(fg_ctx *ctx, GaussSet &gs, const fg_operator *op);
int fg_impl_linear(fg_ctx *ctx, GaussSet &gs, const fg_operator *op);
int fg_default_ucap(const fg_ctx *ctx, const NodeSet &nodes);                  // cluster block size (32 or 64) for this node set
int fg_ensure_clusters(fg_ctx *ctx, GaussSet &gs);                             // the assembly's cluster index, built on first use (with the irregular-cloud fall-backs)
int fg_impl_clusters(fg_ctx *ctx, GaussSet &gs, int ucap, int cells_for = 0);   // cells_for: block size the cell size is chosen for (0 = ucap)
int fg_impl_cluster_tables(fg_ctx *ctx, GaussSet &gs, bool upper_only);
int fg_ensure_lidx(fg_ctx *ctx, GaussSet &gs);                                 // lidx on demand (kernels other than the point-batched ones)
int fg_overflow_list_reset(fg_ctx *ctx, GaussSet &gs);     // overflow list := the points of the index build only
int fg_cluster_binning(fg_ctx *ctx, GaussSet &gs, Clusters &cl, int ucap, int cells_for = 0);

// --------------------------------------------------------------------------------------------
// Device helpers
// --------------------------------------------------------------------------------------------
struct BinGeom {
    double origin[FG_MAXDIM];
    double inv_size[FG_MAXDIM];
    int nb[FG_MAXDIM];
};

static inline BinGeom fg_geom(const BinGrid &g) {
    BinGeom r;
    for (int d = 0; d < FG_MAXDIM; ++d) { r.origin[d] = g.origin[d]; r.inv_size[d] = g.inv_size[d]; r.nb[d] = g.nb[d]; }
    return r;
}

__device__ __forceinline__ int fg_bin_coord(double v, double origin, double inv_size, int nb) {
    double t = floor((v - origin) * inv_size);
    int b = t < 0.0 ? 0 : (t >= (double)nb ? nb - 1 : (int)t);
    return b;
}

// The reference's inclusion predicates, evaluated exactly as written there (no FMA contraction):
//   rectangular  abs(x[i,j] - g[j]) <= dm[i,j]  for all j      src/Shape_Functions.jl:12
//   circular     sum_j (x[i,j]-g[j])^2 <= dm[i]^2                src/Shape_Functions.jl:17-21
__device__ __forceinline__ bool fg_inside_1d(double x, double g, double dm) {
    return fabs(__dsub_rn(x, g)) <= dm;
}

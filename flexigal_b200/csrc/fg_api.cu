// fg_api.cu -- context, inputs, binning, getters: the host side of the C ABI.
// Kernels live in fg_neighbours.cu / fg_mls.cu / fg_mk.cu / fg_pattern.cu / fg_assemble.cu.
#include "fg_internal.cuh"
#include <cub/cub.cuh>
#include <stdarg.h>
#include <math.h>
#include <algorithm>
#include <stdlib.h>

static std::string g_create_error;

int fg_fail(fg_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

GaussSet *fg_find_set(fg_ctx *ctx, int set_id) {
    auto it = ctx->sets.find(set_id);
    return it == ctx->sets.end() ? nullptr : it->second;
}

static void free_set(GaussSet *s) {
    s->gs.release(); s->off.release(); s->cnt.release(); s->lookback.release(); s->cand.release(); s->masks.release(); s->dom.release(); s->phi.release(); s->dphi.release();
    s->gbins.start.release(); s->gsorted.release(); s->singular_dev.release(); s->imls_perm.release();
    s->nbcl.grid.start.release(); s->nbcl.sorted.release(); s->cl.grid.start.release(); s->cl.sorted.release(); s->cl.nU.release(); s->cl.nodes.release(); s->cl.lidx.release(); s->cl.inv.release(); s->cl.ovf_points.release(); s->cl.col0.release(); s->cl.coln.release(); s->cl.tab.release();
    s->pat.colptr.release(); s->pat.rowval.release(); s->pat.nzval.release(); s->pat.tmap.release(); s->fvec.release();
    s->nmask.release(); s->psub.release(); s->xbuf.release();
    for (auto &x : s->xchg) for (auto &pm : x.peers) if (pm.mapped) cudaIpcCloseMemHandle(pm.mapped);
    delete s;
}

extern "C" int fg_version(int *abi, int *sm) {
    if (abi) *abi = 2;
    if (sm) *sm = 100;
    return FG_OK;
}

extern "C" const char *fg_last_error(const fg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int fg_create(fg_ctx **out, int device) {
    if (!out) return fg_fail(nullptr, FG_E_ARG, "fg_create: ctx is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: no CUDA device (%s); this library has no CPU path",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fg_fail(nullptr, FG_E_ARG, "fg_create: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: cudaSetDevice: %s", cudaGetErrorString(e));
    fg_ctx *c = new fg_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(c->own_stream); delete c;
        return fg_fail(nullptr, FG_E_CUDA, "fg_create: copy stream / event");
    }
#ifdef FG_TUNING
    {   // tuning builds: the sweep scripts under tools/ drive the switches through the environment
        static const char *const env[][2] = {{"FG_ASSEMBLY_DIRECT", "assembly_direct"}, {"FG_ASSEMBLY_SYM", "assembly_sym"}, {"FG_ASSEMBLY_REG", "assembly_reg"},
            {"FG_ASSEMBLY_MMA", "assembly_mma"}, {"FG_ASSEMBLY_PTS", "assembly_pts"}, {"FG_IMLS_WARP", "imls_warp"}, {"FG_IMLS_CLUSTER", "imls_cluster"}, {"FG_CLUSTER_UCAP", "cluster_ucap"}, {"FG_NB_UCAP", "nb_ucap"}, {"FG_NB_POINTWISE", "nb_pointwise"},
            {"FG_NB_FUSED", "nb_fused"}, {"FG_DEBUG_SKIP", "debug_skip"}, {"FG_FUSED", "fused"}};
        for (auto &e : env) { const char *v = getenv(e[0]); if (v) c->opts[e[1]] = atoi(v); }
    }
#endif
    *out = c;
    return FG_OK;
}

extern "C" int fg_set_option(fg_ctx *ctx, const char *key, int value) {
    if (!ctx || !key) return FG_E_ARG;
    static const char *const known[] = {"assembly_direct", "assembly_sym", "assembly_reg", "assembly_mma", "assembly_pts", "imls_warp", "imls_cluster", "imls_pipe", "imls_grid", "pattern_from_clusters", "linear_pts", "lazy_shapes", "mk_rows", "cluster_ucap", "nb_ucap", "nb_pointwise", "nb_fused", "nb_rebin", "nb_lazy_dom", "xchg_direct", "cluster_from_masks", "cluster_lidx_eager", "pattern_tables", "coef_on_device", "fused",
#ifdef FG_TUNING
                                        "debug_skip",
#endif
                                        nullptr};
    for (int k = 0; known[k]; ++k)
        if (std::string(known[k]) == key) { ctx->opts[key] = value; return FG_OK; }
    return fg_fail(ctx, FG_E_ARG, "fg_set_option: unknown option '%s'", key);
}

extern "C" int fg_destroy(fg_ctx *ctx) {
    if (!ctx) return FG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    fg_comm_destroy(ctx);
    for (auto &kv : ctx->sets) free_set(kv.second);
    NodeSet &n = ctx->nodes;
    n.x.release(); n.dm.release(); n.idm.release(); n.nrec.release(); n.bins.start.release(); n.sorted_id.release(); n.xs.release(); n.dms.release();
    ctx->temp.release(); ctx->flags.release(); ctx->coef.release(); ctx->stage.release();
    ctx->fem_keys_in.release(); ctx->fem_keys_out.release(); ctx->xmsg.release(); ctx->pat_gsub.release(); ctx->ids64.release(); ctx->field_u.release(); ctx->field_val.release(); ctx->field_grad.release();
    for (auto &v : ctx->solve.v) v.release();
    ctx->solve.cp.release(); ctx->solve.ri.release();
    ctx->bin_key_in.release(); ctx->bin_key_out.release(); ctx->bin_val_in.release(); ctx->pat_cnt.release(); ctx->pat_rows.release(); ctx->pat_pmask.release(); ctx->pat_gxs.release();
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return FG_OK;
}

extern "C" int fg_set_stream(fg_ctx *ctx, void *s) {
    if (!ctx) return FG_E_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return FG_OK;
}

extern "C" int fg_launch_count(fg_ctx *ctx, int64_t *count) {
    if (!ctx || !count) return FG_E_ARG;
    *count = ctx->launches;
    return FG_OK;
}

extern "C" int fg_sync(fg_ctx *ctx) {
    if (!ctx) return FG_E_ARG;
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_pending) { FG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// scans / reductions (cub = plumbing, not the hot path)
// --------------------------------------------------------------------------------------------
struct I32ToI64 { __host__ __device__ int64_t operator()(int v) const { return (int64_t)v; } };

int fg_exclusive_scan_i32_to_i64(fg_ctx *ctx, const int *in, int64_t *out, int64_t n) {
    // out[0..n] : exclusive prefix sums with the grand total in out[n]
    cub::TransformInputIterator<int64_t, I32ToI64, const int *> it(in, I32ToI64());
    size_t bytes = 0;
    FG_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, bytes, it, out + 1, n, ctx->stream));
    FG_CUDA(ctx, ctx->temp.reserve(bytes));
    FG_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(int64_t), ctx->stream));
    if (n > 0) FG_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->temp.p, bytes, it, out + 1, n, ctx->stream));
    return FG_OK;
}

static int reduce_minmax(fg_ctx *ctx, const double *d, int64_t n, double *mn, double *mx) {
    FG_CUDA(ctx, ctx->flags.reserve(64));
    double *res = (double *)ctx->flags.p;
    size_t b1 = 0, b2 = 0;
    FG_CUDA(ctx, cub::DeviceReduce::Min(nullptr, b1, d, res, n, ctx->stream));
    FG_CUDA(ctx, cub::DeviceReduce::Max(nullptr, b2, d, res + 1, n, ctx->stream));
    FG_CUDA(ctx, ctx->temp.reserve(std::max(b1, b2)));
    FG_CUDA(ctx, cub::DeviceReduce::Min(ctx->temp.p, b1, d, res, n, ctx->stream));
    FG_CUDA(ctx, cub::DeviceReduce::Max(ctx->temp.p, b2, d, res + 1, n, ctx->stream));
    double h[2];
    FG_CUDA(ctx, cudaMemcpyAsync(h, res, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *mn = h[0]; *mx = h[1];
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// binning
// --------------------------------------------------------------------------------------------
__global__ void k_bin_index(int dim, int64_t n, const double *__restrict__ c, BinGeom g, int *__restrict__ key, int *__restrict__ val) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b[FG_MAXDIM] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) b[d] = fg_bin_coord(c[d * n + i], g.origin[d], g.inv_size[d], g.nb[d]);
    key[i] = (b[2] * g.nb[1] + b[1]) * g.nb[0] + b[0];
    val[i] = (int)i;
}

__global__ void k_bin_starts(int64_t n, int64_t nbins, const int *__restrict__ key, int *__restrict__ start) {
    // key sorted ascending; start[b] = first k with key[k] >= b; start[nbins] = n
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n) return;
    int prev = k == 0 ? -1 : key[k - 1];
    int cur = k == n ? (int)nbins : key[k];
    for (int b = prev + 1; b <= cur; ++b) start[b] = (int)k;
}

int fg_build_bins(fg_ctx *ctx, int dim, int64_t n, const double *coords, const double *origin, const double *inv_size,
                  const int *nb, BinGrid &grid, DevBuf<int> &sorted) {
    for (int d = 0; d < FG_MAXDIM; ++d) { grid.origin[d] = origin[d]; grid.inv_size[d] = inv_size[d]; grid.nb[d] = nb[d]; }
    grid.nbins = (int64_t)nb[0] * nb[1] * nb[2];
    FG_CUDA(ctx, grid.start.reserve(grid.nbins + 1));
    FG_CUDA(ctx, sorted.reserve(n));
    DevBuf<int> &key_in = ctx->bin_key_in, &key_out = ctx->bin_key_out, &val_in = ctx->bin_val_in;
    FG_CUDA(ctx, key_in.reserve(n)); FG_CUDA(ctx, key_out.reserve(n)); FG_CUDA(ctx, val_in.reserve(n));
    int rc = FG_OK;
    do {
        if (n > 0) k_bin_index<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dim, n, coords, fg_geom(grid), key_in.p, val_in.p);
        int bits = 1; while ((1ll << bits) < grid.nbins) ++bits;
        size_t bytes = 0;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.p, key_out.p, val_in.p, sorted.p, (int)n, 0, bits, ctx->stream);
        if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "radix sort size: %s", cudaGetErrorString(e)); break; }
        if ((e = ctx->temp.reserve(bytes)) != cudaSuccess) { rc = fg_fail(ctx, FG_E_OOM, "radix sort temp: %s", cudaGetErrorString(e)); break; }
        if (n > 0) {
            e = cub::DeviceRadixSort::SortPairs(ctx->temp.p, bytes, key_in.p, key_out.p, val_in.p, sorted.p, (int)n, 0, bits, ctx->stream);
            if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "radix sort: %s", cudaGetErrorString(e)); break; }
        }
        k_bin_starts<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(n, grid.nbins, key_out.p, grid.start.p);
        e = cudaGetLastError();          // no stream synchronisation here: every consumer of the bins runs on the same stream
        if (e != cudaSuccess) { rc = fg_fail(ctx, FG_E_CUDA, "binning: %s", cudaGetErrorString(e)); break; }
    } while (0);
    return rc;
}

__global__ void k_gather_sorted(int ncomp, int64_t n, const int *__restrict__ sorted, const double *__restrict__ src, double *__restrict__ dst) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = sorted[k];
    for (int d = 0; d < ncomp; ++d) dst[d * n + k] = src[d * n + i];
}

__global__ void k_node_records(int dim, int ndm, int64_t n, const double *__restrict__ x, const double *__restrict__ idm, double *__restrict__ rec) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int d = 0; d < dim; ++d) { rec[i * 2 * dim + d] = x[d * n + i]; rec[i * 2 * dim + dim + d] = idm[(d < ndm ? d : 0) * n + i]; }
}

__global__ void k_reciprocal(int64_t n, const double *__restrict__ a, double *__restrict__ r) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) r[k] = 1.0 / a[k];
}

extern "C" int fg_set_nodes(fg_ctx *ctx, int dim, int64_t nnodes, const double *x, const double *dm, int shape) {
    if (!ctx) return FG_E_ARG;
    if (dim != 2 && dim != 3) return fg_fail(ctx, FG_E_ARG, "Only 2D or 3D problems are supported. Detected dimension: %d.", dim);
    if (shape != FG_SHAPE_RECTANGULAR && shape != FG_SHAPE_CIRCULAR)
        return fg_fail(ctx, FG_E_ARG, "Unknown shape of influence domain: %d. Shapes supported are: rectangular, circular, spherical.", shape);
    if (nnodes <= 0 || nnodes >= (1ll << 31) || !x || !dm) return fg_fail(ctx, FG_E_ARG, "fg_set_nodes: bad nnodes/x/dm");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    NodeSet &N = ctx->nodes;
    N.ready = false;
    N.dim = dim; N.shape = shape; N.n = nnodes;
    const int ndm = shape == FG_SHAPE_RECTANGULAR ? dim : 1;
    FG_CUDA(ctx, N.x.reserve(nnodes * dim)); FG_CUDA(ctx, N.dm.reserve(nnodes * ndm)); FG_CUDA(ctx, N.idm.reserve(nnodes * ndm));
    FG_CUDA(ctx, N.xs.reserve(nnodes * dim)); FG_CUDA(ctx, N.dms.reserve(nnodes * ndm));
    FG_CUDA(ctx, cudaMemcpyAsync(N.x.p, x, sizeof(double) * nnodes * dim, cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(N.dm.p, dm, sizeof(double) * nnodes * ndm, cudaMemcpyHostToDevice, ctx->stream));
    k_reciprocal<<<(unsigned)((nnodes * ndm + 255) / 256), 256, 0, ctx->stream>>>(nnodes * ndm, N.dm.p, N.idm.p);
    FG_CHECK_LAUNCH(ctx);
    FG_CUDA(ctx, N.nrec.reserve(nnodes * 2 * dim));
    k_node_records<<<(unsigned)((nnodes + 255) / 256), 256, 0, ctx->stream>>>(dim, ndm, nnodes, N.x.p, N.idm.p, N.nrec.p);
    FG_CHECK_LAUNCH(ctx);
    // bounding box, reach, bin geometry
    double origin[FG_MAXDIM] = {0, 0, 0}, inv_size[FG_MAXDIM] = {1, 1, 1}, ext[FG_MAXDIM] = {0, 0, 0};
    int nb[FG_MAXDIM] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        double mn, mx, dmn, dmx;
        int rc = reduce_minmax(ctx, N.x.p + d * nnodes, nnodes, &mn, &mx);
        if (rc) return rc;
        rc = reduce_minmax(ctx, N.dm.p + (shape == FG_SHAPE_RECTANGULAR ? d : 0) * nnodes, nnodes, &dmn, &dmx);
        if (rc) return rc;
        if (!(dmn > 0.0) || !std::isfinite(dmx) || !std::isfinite(mn) || !std::isfinite(mx))
            return fg_fail(ctx, FG_E_ARG, "fg_set_nodes: non-finite coordinates or non-positive support size in dimension %d", d + 1);
        origin[d] = mn; ext[d] = mx - mn; N.ext[d] = ext[d];
        // widened a little so that candidate ranges are conservative w.r.t. the FP64 predicate
        N.reach[d] = dmx * (1.0 + 1e-9) + 1e-300;
    }
    double size[FG_MAXDIM];
    double total = 1.0;
    for (int d = 0; d < dim; ++d) { size[d] = N.reach[d]; nb[d] = (int)std::min(4096.0, floor(ext[d] / size[d]) + 1.0); total *= nb[d]; }
    const double max_bins = std::max(4096.0, 8.0 * (double)nnodes);
    while (total > max_bins) {   // very small supports relative to the box: coarsen (still conservative)
        total = 1.0;
        for (int d = 0; d < dim; ++d) { size[d] *= 1.26; nb[d] = (int)std::min(4096.0, floor(ext[d] / size[d]) + 1.0); total *= nb[d]; }
    }
    for (int d = 0; d < dim; ++d) inv_size[d] = 1.0 / size[d];
    int rc = fg_build_bins(ctx, dim, nnodes, N.x.p, origin, inv_size, nb, N.bins, N.sorted_id);
    if (rc) return rc;
    k_gather_sorted<<<(unsigned)((nnodes + 255) / 256), 256, 0, ctx->stream>>>(dim, nnodes, N.sorted_id.p, N.x.p, N.xs.p);
    k_gather_sorted<<<(unsigned)((nnodes + 255) / 256), 256, 0, ctx->stream>>>(ndm, nnodes, N.sorted_id.p, N.dm.p, N.dms.p);
    FG_CHECK_LAUNCH(ctx);
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // every Gauss set refers to the old node cloud
    for (auto &kv : ctx->sets) { kv.second->have_nb = kv.second->have_sf = false; kv.second->pat.ready = false; kv.second->gbins_ready = false; kv.second->cl.ready = kv.second->cl.binned = false; kv.second->nbcl.binned = false; kv.second->nsub = 0; }
    N.ready = true;
    return FG_OK;
}

extern "C" int fg_set_gauss(fg_ctx *ctx, int set_id, int64_t ngauss, const double *gs) {
    if (!ctx) return FG_E_ARG;
    if (!ctx->nodes.ready) return fg_fail(ctx, FG_E_STATE, "fg_set_gauss: call fg_set_nodes first");
    if (set_id < 0 || ngauss < 0 || ngauss >= (1ll << 31) || (ngauss > 0 && !gs)) return fg_fail(ctx, FG_E_ARG, "fg_set_gauss: bad set_id/ngauss/gs");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) { S = new GaussSet(); ctx->sets[set_id] = S; }
    if (S->max_L > 0) S->max_L_hint = S->max_L;
    S->n = ngauss; S->have_nb = S->have_sf = false; S->pat.ready = false; S->gbins_ready = false; S->cl.ready = S->cl.binned = false; S->nbcl.binned = false; S->total_L = 0; S->max_L = 0;
    S->nsub = 0;                     // a new table searches all nodes until fg_set_node_subset says otherwise
    S->nb_clustered = false; S->fem.clear();
    const int nc = ctx->nodes.dim + 2;
    FG_CUDA(ctx, S->gs.reserve(ngauss * nc));
    if (ngauss > 0) FG_CUDA(ctx, cudaMemcpyAsync(S->gs.p, gs, sizeof(double) * ngauss * nc, cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

// bit id-1 of the mask for every listed (1-based) id; ids outside [1, nnodes] raise the flag
__global__ void k_mask_from_ids(int64_t n, const int64_t *__restrict__ ids, int64_t nnodes, unsigned *__restrict__ mask, int *__restrict__ bad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t id = ids[k] - 1;
    if (id < 0 || id >= nnodes) { atomicAdd(bad, 1); return; }
    atomicOr(mask + (id >> 5), 1u << (id & 31));
}

extern "C" int fg_set_node_subset(fg_ctx *ctx, int set_id, const int64_t *ids, int64_t n) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_STATE, "fg_set_node_subset: unknown set %d (call fg_set_gauss first)", set_id);
    if (n < 0 || (n > 0 && !ids)) return fg_fail(ctx, FG_E_ARG, "fg_set_node_subset: bad ids/n");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    S->have_nb = S->have_sf = false; S->pat.ready = false; S->cl.ready = false;
    if (!ids) { S->nsub = 0; return FG_OK; }
    const int64_t nn = ctx->nodes.n, words = (nn + 31) / 32;
    FG_CUDA(ctx, S->nmask.reserve(words));
    FG_CUDA(ctx, ctx->ids64.reserve(n));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    FG_CUDA(ctx, cudaMemsetAsync(S->nmask.p, 0, sizeof(unsigned) * words, ctx->stream));
    FG_CUDA(ctx, cudaMemsetAsync(ctx->flags.p + 16, 0, sizeof(int), ctx->stream));
    if (n > 0) {
        FG_CUDA(ctx, cudaMemcpyAsync(ctx->ids64.p, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        k_mask_from_ids<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->ids64.p, nn, S->nmask.p, ctx->flags.p + 16);
        FG_CHECK_LAUNCH(ctx);
    }
    int bad = 0;
    FG_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flags.p + 16, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // also: ids is the caller's again
    if (bad) return fg_fail(ctx, FG_E_ARG, "fg_set_node_subset: %d node ids outside [1, %lld]", bad, (long long)nn);
    S->nsub = 1; S->mask_words = words;
    return FG_OK;
}

extern "C" int fg_drop_set(fg_ctx *ctx, int set_id) {
    if (!ctx) return FG_E_ARG;
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return FG_OK;
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_set(it->second);
    ctx->sets.erase(it);
    return FG_OK;
}

extern "C" int fg_neighbours(fg_ctx *ctx, int set_id, int64_t *total_L, int32_t *max_L) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_STATE, "fg_neighbours: unknown set %d", set_id);
    if (!S->fem.empty()) return fg_fail(ctx, FG_E_STATE, "fg_neighbours: set %d holds FEM shape functions (fg_generate_fem)", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = fg_impl_neighbours(ctx, *S);
    if (rc) return rc;
    if (total_L) *total_L = S->total_L;
    if (max_L) *max_L = S->max_L;
    return FG_OK;
}

extern "C" int fg_shape_functions(fg_ctx *ctx, int set_id, int technique) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_STATE, "fg_shape_functions: unknown set %d", set_id);
    if (technique != FG_TECH_IMLS && technique != FG_TECH_MK)
        return fg_fail(ctx, FG_E_ARG, "Unknown technique: %d. Techniques supported are: IMLS, MK.", technique);
    if (!S->fem.empty()) return fg_fail(ctx, FG_E_STATE, "fg_shape_functions: set %d holds FEM shape functions (fg_generate_fem)", set_id);
    if (!S->have_nb) return fg_fail(ctx, FG_E_STATE, "fg_shape_functions: call fg_neighbours first");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    FG_CUDA(ctx, S->phi.reserve(S->total_L));
    FG_CUDA(ctx, S->dphi.reserve(S->total_L * ctx->nodes.dim));
    S->sf_lazy = false;
    int rc;
    // option "lazy_shapes": the IMLS kernels run when a consumer needs their output -- the streamed assembly (fg_assemble_bilinear_async)
    // then computes them range by range between its assembly launches, so that the download of finished columns starts after one
    // eighth of the shape functions instead of after all of them; every other consumer (getters, linear forms, field evaluation,
    // concatenation, plain assembly) materialises them first (fg_ensure_shapes)
    if (technique == FG_TECH_IMLS && fg_opt(ctx, "lazy_shapes", 0) != 0 && S->n > 0 && S->total_L > 0 && fg_imls_rangeable(ctx, *S)) rc = fg_defer_imls(ctx, *S);
    else rc = technique == FG_TECH_IMLS ? fg_impl_imls(ctx, *S) : fg_impl_mk(ctx, *S);
    if (rc) return rc;
    S->have_sf = true;
    return FG_OK;
}

extern "C" int fg_singular_count(fg_ctx *ctx, int set_id, int64_t *count) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !count) return fg_fail(ctx, FG_E_ARG, "fg_singular_count: bad set/pointer");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (S->have_sf) { const int rcs = fg_ensure_shapes(ctx, *S); if (rcs) return rcs; }
    if (S->singular_dev.p) {
        int sing = 0;
        FG_CUDA(ctx, cudaMemcpyAsync(&sing, S->singular_dev.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        S->singular = sing;
    }
    *count = S->singular;
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// getters (device int32 0-based -> Julia Int64 1-based)
// --------------------------------------------------------------------------------------------
__global__ void k_to_i64_plus(int64_t n, const int *__restrict__ in, int64_t add, int64_t *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (int64_t)in[k] + add;
}

static int fetch_i32_as_i64(fg_ctx *ctx, const int *dev, int64_t n, int64_t add, int64_t *host) {
    const int64_t chunk = 1ll << 26;
    FG_CUDA(ctx, ctx->stage.reserve(sizeof(int64_t) * std::min(n, chunk)));
    for (int64_t s = 0; s < n; s += chunk) {
        int64_t m = std::min(chunk, n - s);
        k_to_i64_plus<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(m, dev + s, add, (int64_t *)ctx->stage.p);
        FG_CHECK_LAUNCH(ctx);
        FG_CUDA(ctx, cudaMemcpyAsync(host + s, ctx->stage.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FG_OK;
}

// out[(e - e0)*dim + d] = comp[d*tl + e] for e0 <= e < e0 + ne
__global__ void k_rows_from_components(int dim, int64_t tl, int64_t e0, int64_t ne, const double *__restrict__ comp, double *__restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= ne * dim) return;
    const int64_t e = k / dim; const int d = (int)(k - e * dim);
    out[k] = comp[d * tl + e0 + e];
}

extern "C" int fg_get_shapes(fg_ctx *ctx, int set_id, int64_t *offsets, int64_t *dom, double *phi, double *dphi) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->have_nb) return fg_fail(ctx, FG_E_STATE, "fg_get_shapes: neighbours of set %d not computed", set_id);
    if ((phi || dphi) && !S->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_get_shapes: shape functions of set %d not computed", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (phi || dphi) { const int rcs = fg_ensure_shapes(ctx, *S); if (rcs) return rcs; }
    if (offsets) FG_CUDA(ctx, cudaMemcpyAsync(offsets, S->off.p, sizeof(int64_t) * (S->n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (phi && S->total_L) FG_CUDA(ctx, cudaMemcpyAsync(phi, S->phi.p, sizeof(double) * S->total_L, cudaMemcpyDeviceToHost, ctx->stream));
    if (dphi && S->total_L) {
        // the device keeps the gradient component-major (dphi[d*total_L + e]); the caller gets [e][d] rows: transposed
        // through a bounded staging buffer, chunk by chunk
        const int dim = ctx->nodes.dim;
        const int64_t chunk = std::min<int64_t>(S->total_L, (int64_t)4 << 20);
        FG_CUDA(ctx, ctx->stage.reserve(sizeof(double) * chunk * dim));
        for (int64_t e0 = 0; e0 < S->total_L; e0 += chunk) {
            const int64_t ne = std::min(chunk, S->total_L - e0);
            k_rows_from_components<<<(unsigned)((ne * dim + 255) / 256), 256, 0, ctx->stream>>>(dim, S->total_L, e0, ne, S->dphi.p, (double *)ctx->stage.p);
            FG_CHECK_LAUNCH(ctx);
            FG_CUDA(ctx, cudaMemcpyAsync(dphi + e0 * dim, ctx->stage.p, sizeof(double) * ne * dim, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the staging buffer is reused by the next chunk
        }
    }
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (dom && S->total_L) { int rc = fg_ensure_dom(ctx, *S); if (rc) return rc; rc = fetch_i32_as_i64(ctx, S->dom.p, S->total_L, 1, dom); if (rc) return rc; }
    return FG_OK;
}

// (PHI, DPHI, DOM) of the Gauss points [g_lo, g_hi) only: offsets relative to the first fetched entry
extern "C" int fg_get_shapes_range(fg_ctx *ctx, int set_id, int64_t g_lo, int64_t g_hi, int64_t *offsets, int64_t *dom, double *phi, double *dphi) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->have_nb) return fg_fail(ctx, FG_E_STATE, "fg_get_shapes_range: neighbours of set %d not computed", set_id);
    if ((phi || dphi) && !S->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_get_shapes_range: shape functions of set %d not computed", set_id);
    if (g_lo < 0 || g_hi < g_lo || g_hi > S->n || !offsets) return fg_fail(ctx, FG_E_ARG, "fg_get_shapes_range: bad range / offsets pointer");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (phi || dphi) { const int rcs = fg_ensure_shapes(ctx, *S); if (rcs) return rcs; }
    const int64_t ng = g_hi - g_lo;
    FG_CUDA(ctx, cudaMemcpyAsync(offsets, S->off.p + g_lo, sizeof(int64_t) * (ng + 1), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t e0 = offsets[0], ne = offsets[ng] - e0;
    for (int64_t k = 0; k <= ng; ++k) offsets[k] -= e0;
    if (ne == 0) return FG_OK;
    if (phi) FG_CUDA(ctx, cudaMemcpyAsync(phi, S->phi.p + e0, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    if (dphi) {
        const int dim = ctx->nodes.dim;
        const int64_t chunk = std::min<int64_t>(ne, (int64_t)4 << 20);
        FG_CUDA(ctx, ctx->stage.reserve(sizeof(double) * chunk * dim));
        for (int64_t c0 = 0; c0 < ne; c0 += chunk) {
            const int64_t m = std::min(chunk, ne - c0);
            k_rows_from_components<<<(unsigned)((m * dim + 255) / 256), 256, 0, ctx->stream>>>(dim, S->total_L, e0 + c0, m, S->dphi.p, (double *)ctx->stage.p);
            FG_CHECK_LAUNCH(ctx);
            FG_CUDA(ctx, cudaMemcpyAsync(dphi + c0 * dim, ctx->stage.p, sizeof(double) * m * dim, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (dom) { int rc = fg_ensure_dom(ctx, *S); if (rc) return rc; rc = fetch_i32_as_i64(ctx, S->dom.p + e0, ne, 1, dom); if (rc) return rc; }
    return FG_OK;
}

__global__ void k_shift_offsets(int64_t n, const int64_t *__restrict__ in, int64_t add, int64_t *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k] + add;
}

extern "C" int fg_concat_sets(fg_ctx *ctx, const int *set_ids, int nsets, int dst_id) {
    if (!ctx) return FG_E_ARG;
    if (!set_ids || nsets <= 0 || dst_id < 0) return fg_fail(ctx, FG_E_ARG, "fg_concat_sets: bad arguments");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int dim = ctx->nodes.dim, nc = dim + 2;
    int64_t n = 0, tl = 0; int maxL = 0; int64_t sing = 0;
    std::vector<GaussSet *> src;
    for (int i = 0; i < nsets; ++i) {
        GaussSet *s = fg_find_set(ctx, set_ids[i]);
        if (!s || !s->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_concat_sets: set %d has no shape functions", set_ids[i]);
        if (set_ids[i] == dst_id) return fg_fail(ctx, FG_E_ARG, "fg_concat_sets: dst_id among the sources");
        { int rcd = fg_ensure_shapes(ctx, *s); if (rcd) return rcd; }
        { int rcd = fg_ensure_dom(ctx, *s); if (rcd) return rcd; }
        int64_t cnt = 0;
        int rcs = fg_singular_count(ctx, set_ids[i], &cnt);      // refreshes the host copy from the source's device counter
        if (rcs) return rcs;
        src.push_back(s); n += s->n; tl += s->total_L; maxL = std::max(maxL, s->max_L); sing += cnt;
    }
    int nfem = 0;
    for (GaussSet *s : src) nfem += s->fem.empty() ? 0 : 1;
    if (nfem != 0 && nfem != nsets)
        return fg_fail(ctx, FG_E_ARG, "fg_concat_sets: FEM and EFG sets in one concatenation (assemble the two measures separately and add the operators, "
                                       "as Linear_Problem does for the terms of a weak form)");
    bool any_mask = false;
    for (GaussSet *s : src) { if (s->nsub > 1) return fg_fail(ctx, FG_E_ARG, "fg_concat_sets: a source is itself a concatenation with node subsets"); any_mask |= s->nsub == 1; }
    if (any_mask && nsets > 255) return fg_fail(ctx, FG_E_CAPACITY, "fg_concat_sets: more than 255 sets with node subsets");
    GaussSet *D = fg_find_set(ctx, dst_id);
    if (!D) { D = new GaussSet(); ctx->sets[dst_id] = D; }
    D->n = n; D->total_L = tl; D->max_L = maxL; D->singular = sing; D->pat.ready = false; D->gbins_ready = false; D->cl.ready = D->cl.binned = false; D->nbcl.binned = false;
    FG_CUDA(ctx, D->gs.reserve(n * nc)); FG_CUDA(ctx, D->off.reserve(n + 1)); FG_CUDA(ctx, D->cnt.reserve(n));
    FG_CUDA(ctx, D->dom.reserve(tl)); FG_CUDA(ctx, D->phi.reserve(tl)); FG_CUDA(ctx, D->dphi.reserve(tl * dim));
    int64_t g0 = 0, l0 = 0;
    for (GaussSet *s : src) {
        for (int c = 0; c < nc; ++c)
            if (s->n) FG_CUDA(ctx, cudaMemcpyAsync(D->gs.p + c * n + g0, s->gs.p + c * s->n, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, ctx->stream));
        if (s->n) k_shift_offsets<<<(unsigned)((s->n + 1 + 255) / 256), 256, 0, ctx->stream>>>(s->n + 1, s->off.p, l0, D->off.p + g0);
        if (s->total_L) {
            FG_CUDA(ctx, cudaMemcpyAsync(D->dom.p + l0, s->dom.p, sizeof(int) * s->total_L, cudaMemcpyDeviceToDevice, ctx->stream));
            FG_CUDA(ctx, cudaMemcpyAsync(D->phi.p + l0, s->phi.p, sizeof(double) * s->total_L, cudaMemcpyDeviceToDevice, ctx->stream));
            for (int d = 0; d < dim; ++d)
                FG_CUDA(ctx, cudaMemcpyAsync(D->dphi.p + d * tl + l0, s->dphi.p + d * s->total_L, sizeof(double) * s->total_L, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        g0 += s->n; l0 += s->total_L;
    }
    if (n == 0) FG_CUDA(ctx, cudaMemsetAsync(D->off.p, 0, sizeof(int64_t), ctx->stream));
    FG_CHECK_LAUNCH(ctx);
    // node subsets: one mask per source (all ones for a source that searched every node), subset index per Gauss point
    D->nsub = 0;
    if (any_mask) {
        const int64_t words = (ctx->nodes.n + 31) / 32;
        FG_CUDA(ctx, D->nmask.reserve(words * nsets)); FG_CUDA(ctx, D->psub.reserve(n));
        int64_t g1 = 0;
        for (int i = 0; i < nsets; ++i) {
            GaussSet *s = src[i];
            if (s->nsub == 1) FG_CUDA(ctx, cudaMemcpyAsync(D->nmask.p + words * i, s->nmask.p, sizeof(unsigned) * words, cudaMemcpyDeviceToDevice, ctx->stream));
            else FG_CUDA(ctx, cudaMemsetAsync(D->nmask.p + words * i, 0xFF, sizeof(unsigned) * words, ctx->stream));
            if (s->n) FG_CUDA(ctx, cudaMemsetAsync(D->psub.p + g1, i, s->n, ctx->stream));
            g1 += s->n;
        }
        D->nsub = nsets; D->mask_words = words;
    }
    // the merged set's own singular counter: the sum of its sources' (fg_singular_count(dst) reads the device copy)
    FG_CUDA(ctx, D->singular_dev.reserve(1));
    { const int sing32 = (int)std::min<int64_t>(sing, 0x7fffffff);
      FG_CUDA(ctx, cudaMemcpyAsync(D->singular_dev.p, &sing32, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    D->fem.clear();
    { int64_t g1 = 0; for (GaussSet *s : src) { for (auto seg : s->fem) { seg.first += g1; D->fem.push_back(seg); } g1 += s->n; } }
    D->have_nb = D->have_sf = true; D->nb_clustered = false; D->dom_pending = false;
    return FG_OK;
}

// --------------------------------------------------------------------------------------------
// pattern / values / vectors
// --------------------------------------------------------------------------------------------
extern "C" int fg_pattern(fg_ctx *ctx, int set_id, int64_t *nnz_nodes) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_STATE, "fg_pattern: unknown set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!S->pat.ready) { int rc = S->fem.empty() ? fg_impl_pattern(ctx, *S, *S) : fg_impl_pattern_fem(ctx, *S); if (rc) return rc; }
    if (nnz_nodes) *nnz_nodes = S->pat.nnz;
    return FG_OK;
}

extern "C" int fg_pattern_from(fg_ctx *ctx, int set_id, int points_set_id, int64_t *nnz_nodes) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id), *W = fg_find_set(ctx, points_set_id);
    if (!S || !W) return fg_fail(ctx, FG_E_STATE, "fg_pattern_from: unknown set %d or %d", set_id, points_set_id);
    if (!S->fem.empty()) return fg_fail(ctx, FG_E_STATE, "fg_pattern_from: set %d is a FEM set (its sparsity comes from the connectivity)", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = fg_impl_pattern(ctx, *S, *W);
    if (rc) return rc;
    if (nnz_nodes) *nnz_nodes = S->pat.nnz;
    return FG_OK;
}

// DOF-level expansion (global dof = (node-1)*D + i, src/Assembling_Operators.jl:65-68):
// column (j, jc) lists, for each row node k of column j, the D dofs of that node.
// Column range [lo, hi) of the node-level pattern: colptr is relative to the range's first entry (1-based), row ids are
// shifted by row_off nodes (a slab's local numbering -> the global one).
__global__ void k_dof_colptr(int64_t lo, int64_t hi, int D, const int64_t *__restrict__ colptr, int64_t *__restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t ncol = (hi - lo) * D;
    if (t > ncol) return;
    const int64_t e0 = colptr[lo];
    if (t == ncol) { out[t] = (colptr[hi] - e0) * D * D + 1; return; }
    int64_t j = lo + t / D; int jc = (int)(t % D);
    int64_t nj = colptr[j + 1] - colptr[j];
    out[t] = (colptr[j] - e0) * D * D + jc * nj * D + 1;
}

__global__ void k_dof_rowval(int64_t lo, int64_t hi, int D, const int64_t *__restrict__ colptr, const int *__restrict__ rowval,
                             int64_t base, int64_t count, int64_t row_off, int64_t *__restrict__ out) {
    // one thread per DOF-level entry p in [base, base+count), p counted from the start of the whole pattern
    int64_t p = base + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= base + count) return;
    // node column j: colptr[j]*D*D <= p < colptr[j+1]*D*D
    int64_t a = lo, b = hi;
    while (a < b) { int64_t mid = (a + b + 1) >> 1; if (colptr[mid] * (int64_t)D * D <= p) a = mid; else b = mid - 1; }
    int64_t j = a;
    int64_t c0 = colptr[j], nj = colptr[j + 1] - c0;
    int64_t r = p - c0 * D * D;
    int64_t within = r % (nj * D);        // k*D + ir
    int64_t k = within / D; int ir = (int)(within % D);
    out[p - base] = ((int64_t)rowval[c0 + k] + row_off) * D + ir + 1;
}

int fg_colptr_at(fg_ctx *ctx, Pattern &P, const int64_t *cols, int n, int64_t *out) {
    bool fetched = false;
    for (int k = 0; k < n; ++k) {
        auto it = P.colptr_host.find(cols[k]);
        if (it == P.colptr_host.end()) {
            FG_CUDA(ctx, cudaMemcpyAsync(&P.colptr_host[cols[k]], P.colptr.p + cols[k], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            fetched = true;
        }
    }
    if (fetched) FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n; ++k) out[k] = P.colptr_host[cols[k]];
    return FG_OK;
}

static int get_pattern_cols(fg_ctx *ctx, GaussSet *S, int D, int64_t lo, int64_t hi, int64_t row_off, int64_t *colptr, int64_t *rowval) {
    Pattern &P = S->pat;
    if (colptr) {
        const int64_t ncol = (hi - lo) * D;
        FG_CUDA(ctx, ctx->stage.reserve(sizeof(int64_t) * (ncol + 1)));
        k_dof_colptr<<<(unsigned)((ncol + 1 + 255) / 256), 256, 0, ctx->stream>>>(lo, hi, D, P.colptr.p, (int64_t *)ctx->stage.p);
        FG_CHECK_LAUNCH(ctx);
        FG_CUDA(ctx, cudaMemcpyAsync(colptr, ctx->stage.p, sizeof(int64_t) * (ncol + 1), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (rowval) {
        const int64_t ends[2] = {lo, hi}; int64_t e[2];
        int rc = fg_colptr_at(ctx, P, ends, 2, e); if (rc) return rc;
        const int64_t first = e[0] * D * D, total = (e[1] - e[0]) * D * D, chunk = 1ll << 26;
        FG_CUDA(ctx, ctx->stage.reserve(sizeof(int64_t) * std::max<int64_t>(1, std::min(total, chunk))));
        for (int64_t s = 0; s < total; s += chunk) {
            int64_t m = std::min(chunk, total - s);
            k_dof_rowval<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(lo, hi, D, P.colptr.p, P.rowval.p, first + s, m, row_off, (int64_t *)ctx->stage.p);
            FG_CHECK_LAUNCH(ctx);
            FG_CUDA(ctx, cudaMemcpyAsync(rowval + s, ctx->stage.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return FG_OK;
}

extern "C" int fg_get_pattern(fg_ctx *ctx, int set_id, int D, int64_t *colptr, int64_t *rowval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready) return fg_fail(ctx, FG_E_STATE, "fg_get_pattern: call fg_pattern first");
    if (D < 1 || D > 3) return fg_fail(ctx, FG_E_ARG, "fg_get_pattern: D must be 1, 2 or 3");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    return get_pattern_cols(ctx, S, D, 0, ctx->nodes.n, 0, colptr, rowval);
}

extern "C" int fg_get_pattern_cols(fg_ctx *ctx, int set_id, int D, int64_t col_lo, int64_t col_hi, int64_t row_offset, int64_t *nnz_dof,
                                   int64_t *colptr, int64_t *rowval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready) return fg_fail(ctx, FG_E_STATE, "fg_get_pattern_cols: call fg_pattern first");
    if (D < 1 || D > 3 || col_lo < 0 || col_hi < col_lo || col_hi > ctx->nodes.n) return fg_fail(ctx, FG_E_ARG, "fg_get_pattern_cols: bad D / column range");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t ends[2] = {col_lo, col_hi}; int64_t e[2];
    int rc = fg_colptr_at(ctx, S->pat, ends, 2, e); if (rc) return rc;
    if (nnz_dof) *nnz_dof = (e[1] - e[0]) * D * D;
    return get_pattern_cols(ctx, S, D, col_lo, col_hi, row_offset, colptr, rowval);
}

extern "C" int fg_get_values_cols(fg_ctx *ctx, int set_id, int64_t col_lo, int64_t col_hi, double *nzval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready || S->pat.D == 0 || !nzval) return fg_fail(ctx, FG_E_STATE, "fg_get_values_cols: nothing assembled for set %d", set_id);
    if (col_lo < 0 || col_hi < col_lo || col_hi > ctx->nodes.n) return fg_fail(ctx, FG_E_ARG, "fg_get_values_cols: bad column range");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t ends[2] = {col_lo, col_hi}; int64_t e[2];
    int rc = fg_colptr_at(ctx, S->pat, ends, 2, e); if (rc) return rc;
    const int64_t dd = (int64_t)S->pat.D * S->pat.D;
    if (e[1] > e[0]) FG_CUDA(ctx, cudaMemcpyAsync(nzval, S->pat.nzval.p + e[0] * dd, sizeof(double) * (e[1] - e[0]) * dd, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

static int check_op(fg_ctx *ctx, const fg_operator *op, bool bilinear) {
    if (!op) return fg_fail(ctx, FG_E_ARG, "operator is NULL");
    if (op->D < 1 || op->D > 3) return fg_fail(ctx, FG_E_ARG, "operator: D must be 1, 2 or 3 (got %d)", op->D);
    if (bilinear) {
        if (op->kind < FG_OP_GENERIC || op->kind > FG_OP_ELASTICITY) return fg_fail(ctx, FG_E_ARG, "unknown bilinear operator kind %d", op->kind);
        if (op->kind == FG_OP_ADVECTION && op->D != 1) return fg_fail(ctx, FG_E_ARG, "FG_OP_ADVECTION is a scalar-field operator");
        if (op->kind == FG_OP_ELASTICITY && op->D != ctx->nodes.dim) return fg_fail(ctx, FG_E_ARG, "FG_OP_ELASTICITY needs D == dim");
    } else if (op->kind != FG_LIN_GENERIC && op->kind != FG_LIN_SOURCE)
        return fg_fail(ctx, FG_E_ARG, "unknown linear operator kind %d", op->kind);
    if (op->kind == 0 && !op->coef) return fg_fail(ctx, FG_E_ARG, "generic operator without coefficients");
    if (op->coef_stride < 0) return fg_fail(ctx, FG_E_ARG, "negative coef_stride");
    return FG_OK;
}

extern "C" int fg_assemble_bilinear(fg_ctx *ctx, int set_id, const fg_operator *op, double *nzval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_assemble_bilinear: shape functions of set %d not computed", set_id);
    int rc = check_op(ctx, op, true);
    if (rc) return rc;
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!S->pat.ready) { rc = S->fem.empty() ? fg_impl_pattern(ctx, *S, *S) : fg_impl_pattern_fem(ctx, *S); if (rc) return rc; }
    rc = fg_impl_bilinear(ctx, *S, op);
    if (rc) return rc;
    if (nzval) return fg_get_values(ctx, set_id, nzval);
    return FG_OK;
}

// Assembly and download as one pipeline (what Bilinear_Assembler's caller wants: the SparseMatrixCSC values on the host).  On the
// point-batched scalar path the clusters are assembled in ranges and every finished column block is copied out on the copy stream under
// the next range; elsewhere this is fg_assemble_bilinear + fg_get_values_async.  nzval is complete after the next fg_sync.
extern "C" int fg_assemble_bilinear_async(fg_ctx *ctx, int set_id, const fg_operator *op, double *nzval) {
    if (!ctx) return FG_E_ARG;
    if (!nzval) return fg_fail(ctx, FG_E_ARG, "fg_assemble_bilinear_async: nzval is NULL");
    ctx->stream_to = nzval; ctx->streamed = false;
    const int rc = fg_assemble_bilinear(ctx, set_id, op, nullptr);
    ctx->stream_to = nullptr;
    if (rc) return rc;
    return ctx->streamed ? FG_OK : fg_get_values_async(ctx, set_id, nzval);
}

extern "C" int fg_assembly_stats(fg_ctx *ctx, int set_id, int64_t *nclusters, int64_t *direct_points, int32_t *max_union) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_ARG, "fg_assembly_stats: unknown set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t ncl = 0, ovf = S->n; int mu = 0;
    if (S->cl.ready && S->cl.n > 0) {
        int h[2] = {0, 0};
        FG_CUDA(ctx, cudaMemcpyAsync(&h[0], S->cl.ovf_points.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaMemcpyAsync(&h[1], ctx->flags.p + 8, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ncl = S->cl.n; ovf = h[0]; mu = h[1];
    }
    if (nclusters) *nclusters = ncl;
    if (direct_points) *direct_points = ovf;
    if (max_union) *max_union = mu;
    return FG_OK;
}

extern "C" int fg_get_values(fg_ctx *ctx, int set_id, double *nzval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready || S->pat.D == 0 || !nzval) return fg_fail(ctx, FG_E_STATE, "fg_get_values: nothing assembled for set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = S->pat.nnz * S->pat.D * S->pat.D;
    if (total) FG_CUDA(ctx, cudaMemcpyAsync(nzval, S->pat.nzval.p, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

// The download runs on the context's copy stream behind an event on the main stream: kernels enqueued afterwards (the linear forms, the
// next set) overlap it.  nzval must stay valid until fg_sync returns (pinned host memory makes the copy truly asynchronous).  A later
// assembly on the SAME set must not start before fg_sync.
extern "C" int fg_get_values_async(fg_ctx *ctx, int set_id, double *nzval) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->pat.ready || S->pat.D == 0 || !nzval) return fg_fail(ctx, FG_E_STATE, "fg_get_values_async: nothing assembled for set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = S->pat.nnz * S->pat.D * S->pat.D;
    FG_CUDA(ctx, cudaEventRecord(ctx->copy_event, ctx->stream));
    FG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
    if (total) FG_CUDA(ctx, cudaMemcpyAsync(nzval, S->pat.nzval.p, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->copy_stream));
    ctx->copy_pending = true;
    return FG_OK;
}

extern "C" int fg_assemble_linear(fg_ctx *ctx, int set_id, const fg_operator *op, double *F) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !S->have_sf) return fg_fail(ctx, FG_E_STATE, "fg_assemble_linear: shape functions of set %d not computed", set_id);
    int rc = check_op(ctx, op, false);
    if (rc) return rc;
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((rc = fg_ensure_shapes(ctx, *S))) return rc;
    rc = fg_impl_linear(ctx, *S, op);
    if (rc) return rc;
    if (F) return fg_get_vector(ctx, set_id, F);
    return FG_OK;
}

extern "C" int fg_get_vector(fg_ctx *ctx, int set_id, double *F) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || S->fvec_D == 0 || !F) return fg_fail(ctx, FG_E_STATE, "fg_get_vector: nothing assembled for set %d", set_id);
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    FG_CUDA(ctx, cudaMemcpyAsync(F, S->fvec.p, sizeof(double) * ctx->nodes.n * S->fvec_D, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

extern "C" int fg_device_buffer(fg_ctx *ctx, int set_id, int which, void **dev_ptr, int64_t *count) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S || !dev_ptr || !count) return fg_fail(ctx, FG_E_ARG, "fg_device_buffer: bad set/pointers");
    const int dim = ctx->nodes.dim;
    switch (which) {
    case FG_BUF_COLPTR: *dev_ptr = S->pat.colptr.p; *count = S->pat.ready ? ctx->nodes.n + 1 : 0; break;
    case FG_BUF_ROWVAL: *dev_ptr = S->pat.rowval.p; *count = S->pat.ready ? S->pat.nnz : 0; break;
    case FG_BUF_NZVAL: *dev_ptr = S->pat.nzval.p; *count = S->pat.ready ? S->pat.nnz * S->pat.D * S->pat.D : 0; break;
    case FG_BUF_FVEC: *dev_ptr = S->fvec.p; *count = ctx->nodes.n * S->fvec_D; break;
    case FG_BUF_OFFSETS: *dev_ptr = S->off.p; *count = S->have_nb ? S->n + 1 : 0; break;
    case FG_BUF_DOM: { int rc = fg_ensure_dom(ctx, *S); if (rc) return rc; } *dev_ptr = S->dom.p; *count = S->have_nb ? S->total_L : 0; break;
    case FG_BUF_PHI: { int rc = fg_ensure_shapes(ctx, *S); if (rc) return rc; } *dev_ptr = S->phi.p; *count = S->have_sf ? S->total_L : 0; break;
    case FG_BUF_DPHI: { int rc = fg_ensure_shapes(ctx, *S); if (rc) return rc; } *dev_ptr = S->dphi.p; *count = S->have_sf ? S->total_L * dim : 0; break;
    default: return fg_fail(ctx, FG_E_ARG, "fg_device_buffer: unknown buffer %d", which);
    }
    return FG_OK;
}

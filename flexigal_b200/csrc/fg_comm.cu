// fg_comm.cu -- multi-GPU exchange inside the library (SURVEY 8b "Multi-GPU is internal to the library", 8e).
//
// One process per GPU, one fg_ctx per process.  The assembly of a slab leaves partial sums in the columns of the halo
// nodes (owned by the neighbouring rank); fg_exchange_columns sends those column slices to their owner and adds what
// arrives to the owned columns -- ncclSend / ncclRecv inside one group on the context's stream, then one add kernel, no
// host synchronisation.  Values only: both ranks build the sparsity of shared columns from the same Gauss positions
// (fg_pattern_from), so a column slice has the same layout on both sides.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): in a process that already carries an NCCL (torch's bundled one) the
// loader hands back that very library, in a Julia process the system library is loaded.  The caller moves the 128-byte
// unique id from rank 0 to the other ranks with whatever transport it has (MPI.jl, torch.distributed, a file).
#include "fg_internal.cuh"
#include <dlfcn.h>
#include <string.h>
#include <vector>
#include <algorithm>

namespace {
typedef struct { char internal[128]; } nccl_uid;          // ncclUniqueId (nccl.h:37-38)
typedef void *nccl_comm;
typedef int (*pfn_get_uid)(nccl_uid *);
typedef int (*pfn_init_rank)(nccl_comm *, int, nccl_uid, int);
typedef int (*pfn_destroy)(nccl_comm);
typedef int (*pfn_sendrecv)(void *, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*pfn_group)(void);
typedef const char *(*pfn_errstr)(int);
typedef int (*pfn_allreduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);

struct Nccl {
    void *h = nullptr;
    pfn_get_uid get_uid = nullptr; pfn_init_rank init_rank = nullptr; pfn_destroy destroy = nullptr;
    pfn_sendrecv send = nullptr, recv = nullptr; pfn_group group_start = nullptr, group_end = nullptr; pfn_errstr errstr = nullptr;
    pfn_allreduce allreduce = nullptr;
} g_nccl;

const int NCCL_DOUBLE = 8;   // ncclFloat64 (nccl.h ncclDataType_t)
const int NCCL_MAX = 2;      // ncclMax (ncclRedOp_t: sum 0, prod 1, max 2, min 3)

int bind_nccl(fg_ctx *ctx) {
    if (g_nccl.h) return FG_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int k = 0; names[k] && !h; ++k) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fg_fail(ctx, FG_E_STATE, "fg_comm: cannot load libnccl.so.2 (%s)", dlerror());
    Nccl n; n.h = h;
    n.get_uid = (pfn_get_uid)dlsym(h, "ncclGetUniqueId"); n.init_rank = (pfn_init_rank)dlsym(h, "ncclCommInitRank");
    n.destroy = (pfn_destroy)dlsym(h, "ncclCommDestroy"); n.send = (pfn_sendrecv)dlsym(h, "ncclSend"); n.recv = (pfn_sendrecv)dlsym(h, "ncclRecv");
    n.group_start = (pfn_group)dlsym(h, "ncclGroupStart"); n.group_end = (pfn_group)dlsym(h, "ncclGroupEnd");
    n.errstr = (pfn_errstr)dlsym(h, "ncclGetErrorString"); n.allreduce = (pfn_allreduce)dlsym(h, "ncclAllReduce");
    if (!n.get_uid || !n.init_rank || !n.destroy || !n.send || !n.recv || !n.group_start || !n.group_end || !n.allreduce)
        return fg_fail(ctx, FG_E_STATE, "fg_comm: libnccl.so.2 lacks a required symbol");
    g_nccl = n;
    return FG_OK;
}

#define FG_NCCL(ctx, call)                                                                                                  \
    do {                                                                                                                    \
        const int r__ = (call);                                                                                             \
        if (r__ != 0) return fg_fail(ctx, FG_E_CUDA, "%s:%d %s: NCCL error %d (%s)", __FILE__, __LINE__, #call, r__,       \
                                     g_nccl.errstr ? g_nccl.errstr(r__) : "?");                                             \
    } while (0)
}   // namespace

extern "C" int fg_comm_unique_id(fg_ctx *ctx, void *id128) {
    if (!ctx || !id128) return FG_E_ARG;
    int rc = bind_nccl(ctx); if (rc) return rc;
    nccl_uid id;
    FG_NCCL(ctx, g_nccl.get_uid(&id));
    memcpy(id128, &id, sizeof id);
    return FG_OK;
}

extern "C" int fg_comm_init(fg_ctx *ctx, const void *id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fg_fail(ctx, FG_E_ARG, "fg_comm_init: bad id/rank/world");
    if (ctx->comm) return fg_fail(ctx, FG_E_STATE, "fg_comm_init: the context already has a communicator");
    int rc = bind_nccl(ctx); if (rc) return rc;
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, id128, sizeof id);
    nccl_comm comm = nullptr;
    FG_NCCL(ctx, g_nccl.init_rank(&comm, world, id, rank));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
    return FG_OK;
}

extern "C" int fg_comm_destroy(fg_ctx *ctx) {
    if (!ctx) return FG_E_ARG;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.destroy((nccl_comm)ctx->comm);
        ctx->comm = nullptr; ctx->comm_world = 1; ctx->comm_rank = 0;
    }
    return FG_OK;
}

// dst[i] += src[i] for up to four (dst, src, n) segments in one launch
struct AddSegs { double *dst[4]; const double *src[4]; int64_t n[4]; int count; };
__global__ void k_add_segments(AddSegs A) {
    for (int s = 0; s < A.count; ++s)
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.n[s]; i += (int64_t)gridDim.x * blockDim.x) A.dst[s][i] += A.src[s][i];
}

extern "C" int fg_exchange_columns(fg_ctx *ctx, int set_id, int which, int npeers, const int *peers, const int64_t *send_lo, const int64_t *send_hi,
                                   const int64_t *recv_lo, const int64_t *recv_hi) {
    if (!ctx) return FG_E_ARG;
    GaussSet *S = fg_find_set(ctx, set_id);
    if (!S) return fg_fail(ctx, FG_E_ARG, "fg_exchange_columns: unknown set %d", set_id);
    if (npeers == 0) return FG_OK;
    if (npeers < 0 || npeers > 4 || !peers || !send_lo || !send_hi || !recv_lo || !recv_hi) return fg_fail(ctx, FG_E_ARG, "fg_exchange_columns: bad peer lists (at most 4 peers)");
    if (!ctx->comm) return fg_fail(ctx, FG_E_STATE, "fg_exchange_columns: call fg_comm_init first");
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nn = ctx->nodes.n;
    double *base = nullptr;
    std::vector<int64_t> s0(npeers), s1(npeers), r0(npeers), r1(npeers);
    if (which == FG_BUF_NZVAL) {
        Pattern &P = S->pat;
        if (!P.ready || P.D == 0) return fg_fail(ctx, FG_E_STATE, "fg_exchange_columns: nothing assembled for set %d", set_id);
        std::vector<int64_t> cols, vals(4 * npeers);
        for (int k = 0; k < npeers; ++k) {
            if (send_lo[k] < 0 || send_hi[k] < send_lo[k] || send_hi[k] > nn || recv_lo[k] < 0 || recv_hi[k] < recv_lo[k] || recv_hi[k] > nn)
                return fg_fail(ctx, FG_E_ARG, "fg_exchange_columns: column range of peer %d outside [0, %lld]", peers[k], (long long)nn);
            cols.push_back(send_lo[k]); cols.push_back(send_hi[k]); cols.push_back(recv_lo[k]); cols.push_back(recv_hi[k]);
        }
        int rc = fg_colptr_at(ctx, P, cols.data(), (int)cols.size(), vals.data()); if (rc) return rc;
        const int64_t dd = (int64_t)P.D * P.D;
        for (int k = 0; k < npeers; ++k) { s0[k] = vals[4 * k] * dd; s1[k] = vals[4 * k + 1] * dd; r0[k] = vals[4 * k + 2] * dd; r1[k] = vals[4 * k + 3] * dd; }
        base = P.nzval.p;
    } else if (which == FG_BUF_FVEC) {
        if (S->fvec_D == 0) return fg_fail(ctx, FG_E_STATE, "fg_exchange_columns: no vector assembled for set %d", set_id);
        for (int k = 0; k < npeers; ++k) { s0[k] = send_lo[k] * S->fvec_D; s1[k] = send_hi[k] * S->fvec_D; r0[k] = recv_lo[k] * S->fvec_D; r1[k] = recv_hi[k] * S->fvec_D; }
        base = S->fvec.p;
    } else
        return fg_fail(ctx, FG_E_ARG, "fg_exchange_columns: which must be FG_BUF_NZVAL or FG_BUF_FVEC");
    nccl_comm comm = (nccl_comm)ctx->comm;
    // ---- direct path (default): the add kernel reads the peer's halo columns straight from the peer's HBM over NVLink ------------
    // (SURVEY 8e "accumulating straight from the peer buffer").  The peers' buffers are mapped once with CUDA IPC; per call NCCL only
    // carries two one-integer handshakes per peer (stream-ordered rendezvous: "my columns are final" before the reads, "I am done
    // reading" after them), no payload and no staging buffer.  Measured at N = 8: 0.5 ms of ncclSend/ncclRecv for 4 x 20 MB -> see DESIGN.
    if (fg_opt(ctx, "xchg_direct", 1) != 0) {
        GaussSet::XchgState &X = S->xchg[which == FG_BUF_NZVAL ? 0 : 1];
        const uint64_t version = which == FG_BUF_NZVAL ? S->pat.version * 16 + (uint64_t)S->pat.D : (uint64_t)S->fvec_D;   // offsets depend on the sparsity and on D
        bool remap = X.my_base != (const void *)base || X.version != version || (int)X.peers.size() != npeers;
        for (int k = 0; k < npeers && !remap; ++k) remap = X.peers[k].peer != peers[k];
        struct Msg { cudaIpcMemHandle_t h; long long s0, s1; };
        FG_CUDA(ctx, ctx->xmsg.reserve(sizeof(Msg) * 9 + 64));
        Msg *dmsg = reinterpret_cast<Msg *>(ctx->xmsg.p);                 // [0] mine per peer (4), [4..8) theirs
        int *tok = reinterpret_cast<int *>(ctx->xmsg.p + sizeof(Msg) * 9);   // handshake tokens
        if (remap) {
            std::vector<GaussSet::PeerMap> old_maps;
            old_maps.swap(X.peers);                 // a peer whose allocation did not move keeps its mapping (opening one costs milliseconds)
            X.peers.assign(npeers, GaussSet::PeerMap());
            Msg mine[4], theirs[4];
            for (int k = 0; k < npeers; ++k) {
                FG_CUDA(ctx, cudaIpcGetMemHandle(&mine[k].h, base));
                mine[k].s0 = s0[k]; mine[k].s1 = s1[k];
            }
            FG_CUDA(ctx, cudaMemcpyAsync(dmsg, mine, sizeof(Msg) * npeers, cudaMemcpyHostToDevice, ctx->stream));
            FG_NCCL(ctx, g_nccl.group_start());
            for (int k = 0; k < npeers; ++k) {
                FG_NCCL(ctx, g_nccl.send(dmsg + k, sizeof(Msg), 0 /* ncclInt8 */, peers[k], comm, ctx->stream));
                FG_NCCL(ctx, g_nccl.recv(dmsg + 4 + k, sizeof(Msg), 0, peers[k], comm, ctx->stream));
            }
            FG_NCCL(ctx, g_nccl.group_end());
            FG_CUDA(ctx, cudaMemcpyAsync(theirs, dmsg + 4, sizeof(Msg) * npeers, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int k = 0; k < npeers; ++k) {
                if (theirs[k].s1 - theirs[k].s0 != r1[k] - r0[k])
                    return fg_fail(ctx, FG_E_STATE, "fg_exchange_columns: peer %d sends %lld values, this rank expects %lld (column layouts differ)",
                                   peers[k], (long long)(theirs[k].s1 - theirs[k].s0), (long long)(r1[k] - r0[k]));
                X.peers[k].peer = peers[k]; X.peers[k].s0 = theirs[k].s0; X.peers[k].s1 = theirs[k].s1; X.peers[k].h = theirs[k].h;
                for (auto &om : old_maps)
                    if (om.mapped && om.peer == peers[k] && memcmp(&om.h, &theirs[k].h, sizeof(cudaIpcMemHandle_t)) == 0) { X.peers[k].mapped = om.mapped; om.mapped = nullptr; }
                if (!X.peers[k].mapped) FG_CUDA(ctx, cudaIpcOpenMemHandle(&X.peers[k].mapped, theirs[k].h, cudaIpcMemLazyEnablePeerAccess));
            }
            for (auto &om : old_maps) if (om.mapped) cudaIpcCloseMemHandle(om.mapped);
            X.my_base = base; X.version = version;
        }
        auto handshake = [&]() -> int {
            FG_NCCL(ctx, g_nccl.group_start());
            for (int k = 0; k < npeers; ++k) {
                FG_NCCL(ctx, g_nccl.send(tok, 1, 2 /* ncclInt32 */, peers[k], comm, ctx->stream));
                FG_NCCL(ctx, g_nccl.recv(tok + 1 + k, 1, 2, peers[k], comm, ctx->stream));
            }
            FG_NCCL(ctx, g_nccl.group_end());
            return FG_OK;
        };
        int rc = handshake(); if (rc) return rc;                       // every neighbour's columns are final
        AddSegs A; A.count = npeers;
        int64_t most = 0;
        for (int k = 0; k < npeers; ++k) {
            A.dst[k] = base + r0[k]; A.src[k] = reinterpret_cast<const double *>(X.peers[k].mapped) + X.peers[k].s0; A.n[k] = r1[k] - r0[k];
            most = std::max(most, A.n[k]);
        }
        if (most > 0) {
            k_add_segments<<<(unsigned)std::min<int64_t>(ctx->sm_count * 16, (most + 255) / 256), 256, 0, ctx->stream>>>(A);
            FG_CHECK_LAUNCH(ctx);
        }
        return handshake();                                            // every neighbour is done reading mine: the next assembly may overwrite
    }
    int64_t rtotal = 0;
    for (int k = 0; k < npeers; ++k) rtotal += r1[k] - r0[k];
    FG_CUDA(ctx, S->xbuf.reserve(rtotal));
    AddSegs A; A.count = npeers;
    FG_NCCL(ctx, g_nccl.group_start());
    int64_t ro = 0;
    for (int k = 0; k < npeers; ++k) {
        if (s1[k] > s0[k]) FG_NCCL(ctx, g_nccl.send(base + s0[k], (size_t)(s1[k] - s0[k]), NCCL_DOUBLE, peers[k], comm, ctx->stream));
        if (r1[k] > r0[k]) FG_NCCL(ctx, g_nccl.recv(S->xbuf.p + ro, (size_t)(r1[k] - r0[k]), NCCL_DOUBLE, peers[k], comm, ctx->stream));
        A.dst[k] = base + r0[k]; A.src[k] = S->xbuf.p + ro; A.n[k] = r1[k] - r0[k];
        ro += r1[k] - r0[k];
    }
    FG_NCCL(ctx, g_nccl.group_end());
    if (rtotal > 0) {
        k_add_segments<<<(unsigned)std::min<int64_t>(ctx->sm_count * 8, (rtotal + 255) / 256), 256, 0, ctx->stream>>>(A);
        FG_CHECK_LAUNCH(ctx);
    }
    return FG_OK;
}

// max over ranks of a host double (the reference's alpha = maximum(A)*100 needs the global maximum, Assembling_Operators.jl:238)
extern "C" int fg_comm_allreduce_max(fg_ctx *ctx, double *value) {
    if (!ctx || !value) return FG_E_ARG;
    if (!ctx->comm) return FG_OK;             // single rank
    FG_CUDA(ctx, cudaSetDevice(ctx->device));
    FG_CUDA(ctx, ctx->flags.reserve(64));
    double *slot = (double *)(ctx->flags.p + 24);
    FG_CUDA(ctx, cudaMemcpyAsync(slot, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FG_NCCL(ctx, g_nccl.allreduce(slot, slot, 1, NCCL_DOUBLE, NCCL_MAX, (nccl_comm)ctx->comm, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyAsync(value, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

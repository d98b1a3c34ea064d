"""Multi-GPU sharding of the hot path: one process per GPU, z-slabs of background cells.

SURVEY 8(e): neighbour search and MLS are independent per Gauss point; assembly is a sum over
Gauss points into columns owned by nodes.  Rank r owns the cells of z-layers [c0, c1) (hence their
Gauss points) and the node layers [c0, c1) (the last rank also the top layer).  Its Gauss points
also touch a few node layers of the neighbouring slabs (halo columns); after local assembly the
partial sums of those columns are sent to their owner (one send/recv pair per interface over
torch.distributed: NCCL on the GPUs, gloo in the CPU tests) and added there.  Values only: both
sides build the sparsity of the shared columns from the same Gauss points (own + a few halo cell
layers, positions only -- no shape functions are evaluated for them), so the column slices have
identical layout on both ranks.

Everything here is host-side index arithmetic (NumPy) plus the exchange; the numbers come from
the CUDA kernels behind the C ABI.
"""
from __future__ import annotations

import math

import numpy as np

from . import geometry as G


class SlabPartition:
    """Rank's share of a (nx, ny, nz) cell grid over `domain`, split along z into `world` slabs."""

    def __init__(self, domain, divisions, dmax, world, rank, degree=3):
        self.domain, self.divisions, self.world, self.rank, self.degree = tuple(domain), tuple(int(v) for v in divisions), world, rank, degree
        nx, ny, nz = self.divisions
        if not 0 <= rank < world or world > nz:
            raise ValueError("need 0 <= rank < world <= z divisions")
        self.dmax = [float(dmax)] * 3 if np.isscalar(dmax) else [float(v) for v in dmax]
        # balanced slabs: the first nz % world ranks get one layer more (100 layers over 8 ranks = 4 x 13 + 4 x 12)
        base, extra = divmod(nz, world)
        bounds = [r * base + min(r, extra) for r in range(world + 1)]
        self.bounds = bounds
        self.c0, self.c1 = bounds[rank], bounds[rank + 1]
        self.cz = self.c1 - self.c0
        xi = G.pgauss(degree)[0]
        xi_max = 0.5 * (1.0 + float(xi.max()))                      # outermost Gauss abscissa inside a cell, in cells
        dz = self.dmax[2]
        self.reach = int(math.floor(xi_max + dz))                   # own Gauss points touch node layers [c0+1-reach, c1-1+reach]
        if world > 1 and base < self.reach:
            # a thinner slab's Gauss points would touch node layers owned by rank +-2: the one-hop plan below would drop them
            raise ValueError(f"slabs of {base} cell layers are thinner than the support reach ({self.reach} layers); use fewer ranks")
        self.halo = int(math.ceil(self.reach + 2.0 * dz)) if world > 1 else 0   # node / Gauss-cell layers kept beyond the slab
        self.lay = (nx + 1) * (ny + 1)
        self.n_lo = max(0, self.c0 - self.halo)                     # first local node layer (global layer index)
        self.n_hi = min(nz, self.c1 + self.halo)                    # last local node layer, inclusive
        self.g_lo = max(0, self.c0 - self.halo)                     # pattern Gauss cells [g_lo, g_hi)
        self.g_hi = min(nz, self.c1 + self.halo)
        self.node_offset = self.n_lo * self.lay                     # global id = local id + node_offset

    @classmethod
    def window(cls, domain, divisions, dmax, lay_lo, lay_hi, degree=3):
        """A stand-alone single-rank problem made of the cell layers [lay_lo, lay_hi) of the global grid and the node layers
        [lay_lo, lay_hi] (global coordinates, node_offset = lay_lo * layer size): used to cross-check the columns around a
        slab interface against an independent single-GPU assembly."""
        p = cls(domain, divisions, dmax, 1, 0, degree)
        p.c0, p.c1, p.cz = lay_lo, lay_hi, lay_hi - lay_lo
        p.n_lo, p.n_hi, p.g_lo, p.g_hi = lay_lo, lay_hi, lay_lo, lay_hi
        p.node_offset = lay_lo * p.lay
        return p

    # ---- arrays that cross the C ABI (identical bits to the global mesh of generate_cartesian_mesh) ----
    def _local_mesh(self, cell_lo, cell_hi):
        """Nodes of layers [n_lo, n_hi] and the cells of layers [cell_lo, cell_hi) in local numbering."""
        nx, ny, nz = self.divisions
        step = [self.domain[d] / self.divisions[d] for d in range(3)]
        ax = [step[0] * np.arange(nx + 1, dtype=np.float64), step[1] * np.arange(ny + 1, dtype=np.float64),
              step[2] * np.arange(self.n_lo, self.n_hi + 1, dtype=np.float64)]
        nl = self.n_hi - self.n_lo + 1
        X = np.empty((self.lay * nl, 3), order="F")
        X[:, 0] = np.tile(ax[0], (ny + 1) * nl)
        X[:, 1] = np.tile(np.repeat(ax[1], nx + 1), nl)
        X[:, 2] = np.repeat(ax[2], self.lay)
        k, j, i = np.meshgrid(np.arange(cell_lo - self.n_lo, cell_hi - self.n_lo), np.arange(ny), np.arange(nx), indexing="ij")
        n0 = (i + 1 + j * (nx + 1) + k * self.lay).reshape(-1)
        n1, n3 = n0 + 1, n0 + nx + 1
        n2 = n1 + nx + 1
        conn = np.stack([n0, n1, n2, n3, n0 + self.lay, n1 + self.lay, n2 + self.lay, n3 + self.lay], axis=1).astype(np.int64)
        return X, conn

    def local_nodes(self):
        if not hasattr(self, "_x"):
            self._x, self._conn = self._local_mesh(self.c0, self.c1)
        return self._x

    def local_dm(self):
        x = self.local_nodes()
        dm = np.empty((x.shape[0], 3), order="F")
        for d in range(3):
            dm[:, d] = self.dmax[d] * self.domain[d] / self.divisions[d]
        return dm

    def local_gauss(self):
        """Gauss table of the rank's own cells (the rows the reference's egauss gives for them)."""
        x = self.local_nodes()
        return G.egauss(x, self._conn, G.pgauss(self.degree))

    def pattern_gauss(self):
        """Gauss points of own + halo cell layers: positions for the sparsity build only."""
        if not hasattr(self, "_pgs"):
            x = self.local_nodes()
            self._pgs = G.egauss(x, self.pattern_conn(), G.pgauss(self.degree))
        return self._pgs

    def pattern_conn(self):
        """Connectivity (local numbering) of own + halo cell layers: fg_generate_gauss builds the same table on the device
        from these 64 bytes per cell instead of an upload of 27 x 40 bytes per cell."""
        if not hasattr(self, "_pconn"):
            self.local_nodes()
            _, self._pconn = self._local_mesh(self.g_lo, self.g_hi)
        return self._pconn

    def local_conn(self):
        self.local_nodes()
        return self._conn

    # ---- ownership and the exchange plan --------------------------------------------------------------
    def owned_layers(self):
        nz = self.divisions[2]
        return self.c0, (self.c1 if self.rank < self.world - 1 else nz + 1)      # [lo, hi) global node layers

    def local_node_range(self, lay_lo, lay_hi):
        """Local node-id range [a, b) of global node layers [lay_lo, lay_hi)."""
        return (lay_lo - self.n_lo) * self.lay, (lay_hi - self.n_lo) * self.lay

    def plan(self):
        """[(peer, send_layers, recv_layers)]: columns whose partial sums go to / come from each neighbour.
        send = halo node layers this rank touches but the peer owns; recv = owned layers the peer touches."""
        out = []
        r = self.reach
        if self.rank > 0:                      # interface at node layer c0 (owned by this rank)
            out.append((self.rank - 1, (self.c0 + 1 - r, self.c0), (self.c0, self.c0 + r)))
        if self.rank < self.world - 1:         # interface at node layer c1 (owned by the next rank)
            out.append((self.rank + 1, (self.c1, self.c1 + r), (self.c1 + 1 - r, self.c1)))
        return out

    # ---- helpers used by bench.py / the GPU tests -----------------------------------------------------
    def pattern_set(self, ctx, sid, gs=None, on_device=False, ps=None):
        """The halo-extended Gauss positions as their own set (None on a single rank): uploaded, or with ``on_device``
        generated by the library from the connectivity (fg_generate_gauss, bit-identical to the host table)."""
        if self.world == 1:
            return None
        ps = ctx.new_set_id() if ps is None else ps
        if on_device:
            ctx.generate_gauss(ps, self.pattern_conn(), G.pgauss(self.degree))
        else:
            ctx.set_gauss(ps, self.pattern_gauss() if gs is None else gs)
        return ps

    def local_plan(self):
        """plan() in local 0-based node-column ranges: what fg_exchange_columns takes."""
        return [(peer, self.local_node_range(*send_l), self.local_node_range(*recv_l)) for peer, send_l, recv_l in self.plan()]

    def owned_columns(self):
        """Local 0-based node-column range [a, b) of the columns this rank owns."""
        return self.local_node_range(*self.owned_layers())

    def build_pattern(self, ctx, sid, pattern_sid):
        return ctx.pattern_from(sid, pattern_sid)


class _DeviceArray:
    """Raw device memory of the library exposed through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def exchange_slices(part: SlabPartition, colptr, values, dist, torch_mod, block=1):
    """The exchange step on any torch tensors (device tensors under NCCL, CPU tensors under gloo):
    send the column slices of `values` named by part.plan(), add what arrives.  `colptr` is the
    0-based column pointer (host numpy int64) of the local pattern; `block` = D*D."""
    ops, recvs = [], []
    for peer, send_l, recv_l in part.plan():
        a, b = part.local_node_range(*send_l)
        sa, sb = int(colptr[a]) * block, int(colptr[b]) * block
        a, b = part.local_node_range(*recv_l)
        ra, rb = int(colptr[a]) * block, int(colptr[b]) * block
        buf = torch_mod.empty(rb - ra, dtype=values.dtype, device=values.device)
        ops.append(dist.P2POp(dist.isend, values[sa:sb], peer))
        ops.append(dist.P2POp(dist.irecv, buf, peer))
        recvs.append((ra, rb, buf))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for ra, rb, buf in recvs:
        values[ra:rb] += buf
    return values


class HaloExchange:
    """Exchange of halo-column partial sums for one assembled set, inside the library (fg_exchange_columns: ncclSend /
    ncclRecv + add on the context's own stream, ordered after the assembly kernels -- no torch tensors, no stream handshake).
    The communicator is created once per context: rank 0's id travels over torch.distributed (any backend)."""

    def __init__(self, part: SlabPartition, ctx, sid, torch_mod=None, dist=None):
        self.part, self.ctx, self.sid = part, ctx, sid
        if not getattr(ctx, "_comm_ready", False):
            box = [ctx.comm_unique_id() if part.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0], part.rank, part.world)
            ctx._comm_ready = True
        self.plan = part.local_plan()

    def exchange_values(self, D=1):
        from . import _lib
        self.ctx.exchange_columns(self.sid, _lib.BUF_NZVAL, self.plan)

    def exchange_vector(self, D=1):
        from . import _lib
        self.ctx.exchange_columns(self.sid, _lib.BUF_FVEC, self.plan)


def gather_global_csc(part: SlabPartition, ctx, sid, D, dist):
    """SURVEY 8(e) "Return to Julia": every rank fetches the column block it owns (global row numbering) and rank 0
    concatenates the blocks into one 1-based SparseMatrixCSC triple (colptr, rowval, nzval); other ranks get None."""
    a, b = part.owned_columns()
    colptr, rowval = ctx.get_pattern_cols(sid, D, a, b, part.node_offset)
    nz = np.empty(rowval.size)
    ctx.get_values_cols(sid, a, b, nz)
    blocks = [None] * part.world if part.rank == 0 else None
    dist.gather_object((colptr, rowval, nz), blocks, dst=0)
    if part.rank != 0:
        return None
    cps, shift = [], 0
    for cp, _, _ in blocks:
        cps.append(cp[:-1] + shift)
        shift += cp[-1] - 1
    colptr_g = np.concatenate(cps + [np.array([shift + 1], dtype=np.int64)])
    return colptr_g, np.concatenate([blk[1] for blk in blocks]), np.concatenate([blk[2] for blk in blocks])

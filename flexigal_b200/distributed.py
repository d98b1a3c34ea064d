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
        if nz % world:
            raise ValueError("the z divisions must be a multiple of the number of ranks")
        self.dmax = [float(dmax)] * 3 if np.isscalar(dmax) else [float(v) for v in dmax]
        self.cz = nz // world
        self.c0, self.c1 = rank * self.cz, (rank + 1) * self.cz
        xi = G.pgauss(degree)[0]
        xi_max = 0.5 * (1.0 + float(xi.max()))                      # outermost Gauss abscissa inside a cell, in cells
        dz = self.dmax[2]
        self.reach = int(math.floor(xi_max + dz))                   # own Gauss points touch node layers [c0+1-reach, c1-1+reach]
        self.halo = int(math.ceil(self.reach + 2.0 * dz)) if world > 1 else 0   # node / Gauss-cell layers kept beyond the slab
        self.lay = (nx + 1) * (ny + 1)
        self.n_lo = max(0, self.c0 - self.halo)                     # first local node layer (global layer index)
        self.n_hi = min(nz, self.c1 + self.halo)                    # last local node layer, inclusive
        self.g_lo = max(0, self.c0 - self.halo)                     # pattern Gauss cells [g_lo, g_hi)
        self.g_hi = min(nz, self.c1 + self.halo)
        self.node_offset = self.n_lo * self.lay                     # global id = local id + node_offset

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
            _, conn = self._local_mesh(self.g_lo, self.g_hi)
            self._pgs = G.egauss(x, conn, G.pgauss(self.degree))
        return self._pgs

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
    def pattern_set(self, ctx, sid, gs=None):
        """Upload the halo-extended Gauss positions as their own set (None on a single rank)."""
        if self.world == 1:
            return None
        ps = ctx.new_set_id()
        ctx.set_gauss(ps, self.pattern_gauss() if gs is None else gs)
        return ps

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
    """Exchange of halo-column partial sums for one assembled set on the GPU (values stay in HBM)."""

    def __init__(self, part: SlabPartition, ctx, sid, torch_mod, dist):
        from . import _lib
        self.part, self.ctx, self.sid, self.torch, self.dist = part, ctx, sid, torch_mod, dist
        p, n = ctx.device_buffer(sid, _lib.BUF_COLPTR)
        cp = torch_mod.as_tensor(_DeviceArray(p, n, "<i8"), device="cuda")
        self.colptr = cp.cpu().numpy()

    def exchange_values(self, D=1):
        from . import _lib
        p, n = self.ctx.device_buffer(self.sid, _lib.BUF_NZVAL)
        vals = self.torch.as_tensor(_DeviceArray(p, n), device="cuda")
        exchange_slices(self.part, self.colptr, vals, self.dist, self.torch, D * D)
        return vals

    def exchange_vector(self, D=1):
        from . import _lib
        p, n = self.ctx.device_buffer(self.sid, _lib.BUF_FVEC)
        vec = self.torch.as_tensor(_DeviceArray(p, n), device="cuda")
        ident = np.arange(self.part.local_nodes().shape[0] + 1, dtype=np.int64)
        exchange_slices(self.part, ident, vec, self.dist, self.torch, D)
        return vec

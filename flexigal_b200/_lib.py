"""ctypes binding of include/flexigal_gpu.h -- the same symbols FlexiGal's Julia shim `ccall`s.

There is no CPU path behind these calls: if the shared library is missing, or no sm_100 GPU is
usable, every entry point raises.  Nothing under ``oracle/`` is imported here.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLEXIGAL_GPU_LIB") or os.path.join(_HERE, "lib", "libflexigal_gpu.so")   # the variable INTEGRATION.md's shim reads too

FG_OK, FG_E_ARG, FG_E_CUDA, FG_E_OOM, FG_E_STATE, FG_E_CAPACITY = 0, -1, -2, -3, -4, -5
SHAPE_RECTANGULAR, SHAPE_CIRCULAR = 0, 1
TECH_IMLS, TECH_MK = 0, 1
OP_GENERIC, OP_GRADGRAD, OP_MASS, OP_ADVECTION, OP_ELASTICITY = 0, 1, 2, 3, 4
LIN_GENERIC, LIN_SOURCE = 0, 1
BUF_COLPTR, BUF_ROWVAL, BUF_NZVAL, BUF_FVEC, BUF_OFFSETS, BUF_DOM, BUF_PHI, BUF_DPHI = range(8)

SHAPES = {"rectangular": SHAPE_RECTANGULAR, "circular": SHAPE_CIRCULAR, "spherical": SHAPE_CIRCULAR}
TECHNIQUES = {"IMLS": TECH_IMLS, "MK": TECH_MK}

# every symbol include/flexigal_gpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fg_create", "fg_destroy", "fg_last_error", "fg_set_stream", "fg_sync", "fg_set_nodes", "fg_set_gauss",
    "fg_drop_set", "fg_neighbours", "fg_shape_functions", "fg_singular_count", "fg_get_shapes", "fg_concat_sets",
    "fg_pattern", "fg_get_pattern", "fg_assemble_bilinear", "fg_get_values", "fg_assemble_linear", "fg_get_vector",
    "fg_device_buffer", "fg_version", "fg_launch_count", "fg_pattern_from", "fg_assembly_stats", "fg_evaluate_field", "fg_solve_spd", "fg_generate_gauss", "fg_get_gauss",
    "fg_set_option", "fg_set_node_subset", "fg_get_pattern_cols", "fg_get_values_cols", "fg_values_max", "fg_solve_spd_sets",
    "fg_get_shapes_range", "fg_generate_fem", "fg_get_values_async", "fg_assemble_bilinear_async", "fg_comm_unique_id", "fg_comm_init", "fg_comm_destroy", "fg_exchange_columns", "fg_comm_allreduce_max",
]


class FlexiGalGPUError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"flexigal_gpu error {code}: {msg}")
        self.code = code


class fg_operator(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("D", ctypes.c_int32), ("c", ctypes.c_double * 4),
                ("coef", ctypes.c_void_p), ("coef_stride", ctypes.c_int64)]


_lib = None


def load():
    """Load libflexigal_gpu.so (built by ``__graft_entry__.build()`` / ``make -C flexigal_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "flexigal_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    P = ctypes.POINTER
    sig = {
        "fg_create": [P(vp), i32], "fg_destroy": [vp], "fg_set_stream": [vp, vp], "fg_sync": [vp],
        "fg_set_nodes": [vp, i32, i64, vp, vp, i32], "fg_set_gauss": [vp, i32, i64, vp], "fg_drop_set": [vp, i32],
        "fg_neighbours": [vp, i32, P(i64), P(ctypes.c_int32)], "fg_shape_functions": [vp, i32, i32],
        "fg_singular_count": [vp, i32, P(i64)], "fg_get_shapes": [vp, i32, vp, vp, vp, vp],
        "fg_concat_sets": [vp, P(i32), i32, i32], "fg_pattern": [vp, i32, P(i64)], "fg_get_pattern": [vp, i32, i32, vp, vp],
        "fg_assemble_bilinear": [vp, i32, P(fg_operator), vp], "fg_get_values": [vp, i32, vp], "fg_get_values_async": [vp, i32, vp], "fg_assemble_bilinear_async": [vp, i32, P(fg_operator), vp],
        "fg_assemble_linear": [vp, i32, P(fg_operator), vp], "fg_get_vector": [vp, i32, vp],
        "fg_device_buffer": [vp, i32, i32, P(vp), P(i64)], "fg_launch_count": [vp, P(i64)], "fg_pattern_from": [vp, i32, i32, P(i64)], "fg_assembly_stats": [vp, i32, P(i64), P(i64), P(ctypes.c_int32)],
        "fg_evaluate_field": [vp, i32, i32, vp, vp, vp],
        "fg_generate_gauss": [vp, i32, i64, i32, vp, i32, vp], "fg_get_gauss": [vp, i32, vp], "fg_generate_fem": [vp, i32, i64, i32, vp, i32, vp],
        "fg_solve_spd": [vp, i64, vp, vp, vp, vp, vp, ctypes.c_double, i32, P(i32), P(ctypes.c_double)], "fg_version": [P(i32), P(i32)],
        "fg_set_option": [vp, ctypes.c_char_p, i32], "fg_set_node_subset": [vp, i32, vp, i64],
        "fg_get_pattern_cols": [vp, i32, i32, i64, i64, i64, P(i64), vp, vp], "fg_get_values_cols": [vp, i32, i64, i64, vp],
        "fg_values_max": [vp, i32, P(ctypes.c_double)], "fg_get_shapes_range": [vp, i32, i64, i64, vp, vp, vp, vp],
        "fg_solve_spd_sets": [vp, P(i32), P(ctypes.c_double), i32, i32, vp, vp, ctypes.c_double, i32, P(i32), P(ctypes.c_double)],
        "fg_comm_unique_id": [vp, vp], "fg_comm_init": [vp, vp, i32, i32], "fg_comm_destroy": [vp],
        "fg_exchange_columns": [vp, i32, i32, i32, P(i32), P(i64), P(i64), P(i64), P(i64)], "fg_comm_allreduce_max": [vp, P(ctypes.c_double)],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = ctypes.c_int
    lib.fg_last_error.argtypes = [vp]
    lib.fg_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Context:
    """One fg_ctx = one GPU.  Thin, explicit wrapper: method names are the C symbols minus ``fg_``."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.fg_create(ctypes.byref(h), device)
        if rc != FG_OK:
            raise FlexiGalGPUError(rc, self.lib.fg_last_error(None).decode())
        self.h = h
        self.device = device
        self.dim = 0
        self.nnodes = 0
        self._next_set = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.fg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != FG_OK:
            raise FlexiGalGPUError(rc, self.lib.fg_last_error(self.h).decode())

    def new_set_id(self) -> int:
        self._next_set += 1
        return self._next_set - 1

    def set_stream(self, cuda_stream: int | None):
        self._ck(self.lib.fg_set_stream(self.h, ctypes.c_void_p(cuda_stream or 0)))

    def launch_count(self) -> int:
        c = ctypes.c_int64()
        self._ck(self.lib.fg_launch_count(self.h, ctypes.byref(c)))
        return c.value

    def sync(self):
        self._ck(self.lib.fg_sync(self.h))

    def set_nodes(self, x, dm, shape="rectangular"):
        if shape not in SHAPES:
            raise ValueError(f"Unknown shape of influence domain: {shape}. Shapes supported are: rectangular, circular, spherical.")
        x = np.asfortranarray(x, dtype=np.float64)
        dm = np.asfortranarray(dm, dtype=np.float64)
        nn, dim = x.shape
        want = (nn, dim) if SHAPES[shape] == SHAPE_RECTANGULAR else (nn,)
        if dm.shape != want:
            raise ValueError(f"dm has shape {dm.shape}, expected {want} for shape={shape}")
        self._ck(self.lib.fg_set_nodes(self.h, dim, nn, ptr(x), ptr(dm), SHAPES[shape]))
        self.dim, self.nnodes = dim, nn

    def set_gauss(self, set_id, gs):
        gs = np.asfortranarray(gs, dtype=np.float64)
        if gs.ndim != 2 or gs.shape[1] != self.dim + 2:
            raise ValueError(f"gs must be ngauss x {self.dim + 2}")
        self._ck(self.lib.fg_set_gauss(self.h, set_id, gs.shape[0], ptr(gs)))
        return gs.shape[0]

    def generate_gauss(self, set_id, conn, gauss, fetch=False):
        """egauss / egauss_bound on the device from the connectivity (ncells x nverts, 1-based) and pgauss(k)."""
        conn = np.asfortranarray(conn, dtype=np.int64)
        gauss = np.asfortranarray(gauss, dtype=np.float64)
        self._ck(self.lib.fg_generate_gauss(self.h, set_id, conn.shape[0], conn.shape[1], ptr(conn), gauss.shape[1], ptr(gauss)))
        npts = gauss.shape[1] ** (self.dim if conn.shape[1] == 2 ** self.dim else self.dim - 1)
        ng = conn.shape[0] * npts
        if not fetch:
            return ng
        gs = np.empty((ng, self.dim + 2), order="F")
        self._ck(self.lib.fg_get_gauss(self.h, set_id, ptr(gs)))
        return gs

    def generate_fem(self, set_id, conn, gauss):
        """egauss_fem / egauss_bound_fem on the device: Gauss table + element shape functions; returns (ngauss, total_L, gs)."""
        conn = np.asfortranarray(conn, dtype=np.int64)
        gauss = np.asfortranarray(gauss, dtype=np.float64)
        self._ck(self.lib.fg_generate_fem(self.h, set_id, conn.shape[0], conn.shape[1], ptr(conn), gauss.shape[1], ptr(gauss)))
        npts = gauss.shape[1] ** (self.dim if conn.shape[1] == 2 ** self.dim else self.dim - 1)
        ng = conn.shape[0] * npts
        gs = np.empty((ng, self.dim + 2), order="F")
        self._ck(self.lib.fg_get_gauss(self.h, set_id, ptr(gs)))
        return ng, ng * conn.shape[1], gs

    def set_option(self, key: str, value: int):
        self._ck(self.lib.fg_set_option(self.h, key.encode(), int(value)))

    def set_node_subset(self, set_id, ids):
        """Restrict the set's neighbour search to the nodes ``ids`` (1-based); None removes the restriction."""
        if ids is None:
            self._ck(self.lib.fg_set_node_subset(self.h, set_id, None, 0))
            return
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._ck(self.lib.fg_set_node_subset(self.h, set_id, ptr(ids), ids.size))

    def drop_set(self, set_id):
        self._ck(self.lib.fg_drop_set(self.h, set_id))

    def neighbours(self, set_id):
        tot, mx = ctypes.c_int64(), ctypes.c_int32()
        self._ck(self.lib.fg_neighbours(self.h, set_id, ctypes.byref(tot), ctypes.byref(mx)))
        return tot.value, mx.value

    def shape_functions(self, set_id, technique="IMLS"):
        if technique not in TECHNIQUES:
            raise ValueError(f"Unknown technique: {technique}. Techniques supported are: IMLS, MK. If you wrote MLS, "
                             "please write IMLS instead, which is more efficient than standard MLS providing the same results.")
        self._ck(self.lib.fg_shape_functions(self.h, set_id, TECHNIQUES[technique]))

    def singular_count(self, set_id):
        c = ctypes.c_int64()
        self._ck(self.lib.fg_singular_count(self.h, set_id, ctypes.byref(c)))
        return c.value

    def get_shapes(self, set_id, ngauss, total_L, want_values=True):
        off = np.empty(ngauss + 1, dtype=np.int64)
        dom = np.empty(total_L, dtype=np.int64)
        phi = np.empty(total_L) if want_values else None
        dphi = np.empty((total_L, self.dim)) if want_values else None
        self._ck(self.lib.fg_get_shapes(self.h, set_id, ptr(off), ptr(dom), ptr(phi), ptr(dphi)))
        return off, dom, phi, dphi

    def get_shapes_range(self, set_id, g_lo, g_hi, max_L=64, want_values=True):
        """(offsets relative, DOM 1-based, PHI, DPHI) of the Gauss points [g_lo, g_hi)."""
        ng = g_hi - g_lo
        off = np.empty(ng + 1, dtype=np.int64)
        self._ck(self.lib.fg_get_shapes_range(self.h, set_id, g_lo, g_hi, ptr(off), None, None, None))
        tot = int(off[-1])
        dom = np.empty(tot, dtype=np.int64)
        phi = np.empty(tot) if want_values else None
        dphi = np.empty((tot, self.dim)) if want_values else None
        self._ck(self.lib.fg_get_shapes_range(self.h, set_id, g_lo, g_hi, ptr(off), ptr(dom), ptr(phi), ptr(dphi)))
        return off, dom, phi, dphi

    def concat_sets(self, set_ids, dst_id):
        arr = (ctypes.c_int * len(set_ids))(*set_ids)
        self._ck(self.lib.fg_concat_sets(self.h, arr, len(set_ids), dst_id))

    def pattern(self, set_id):
        nnz = ctypes.c_int64()
        self._ck(self.lib.fg_pattern(self.h, set_id, ctypes.byref(nnz)))
        return nnz.value

    def pattern_from(self, set_id, points_set_id):
        nnz = ctypes.c_int64()
        self._ck(self.lib.fg_pattern_from(self.h, set_id, points_set_id, ctypes.byref(nnz)))
        return nnz.value

    def get_pattern(self, set_id, D, nnz):
        colptr = np.empty(self.nnodes * D + 1, dtype=np.int64)
        rowval = np.empty(nnz * D * D, dtype=np.int64)
        self._ck(self.lib.fg_get_pattern(self.h, set_id, D, ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def assemble_bilinear(self, set_id, op: fg_operator, out=None):
        self._ck(self.lib.fg_assemble_bilinear(self.h, set_id, ctypes.byref(op), ptr(out)))
        return out

    def assembly_stats(self, set_id):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        self._ck(self.lib.fg_assembly_stats(self.h, set_id, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"clusters": a.value, "direct_points": b.value, "max_union": c.value}

    def evaluate_field(self, set_id, ngauss, u, D=1, values=True, gradients=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        if u.size != self.nnodes * D:
            raise ValueError(f"nodal vector has {u.size} entries, expected {self.nnodes * D}")
        val = np.empty((ngauss, D)) if values else None
        grad = np.empty((ngauss, self.dim, D)) if gradients else None      # [g][d][i] = SMatrix{D,dim} memory order
        self._ck(self.lib.fg_evaluate_field(self.h, set_id, D, ptr(u), ptr(val), ptr(grad)))
        return val, grad

    def solve_spd(self, A, b, rtol=1e-13, max_iter=20000):
        """A: scipy CSC/CSR symmetric positive definite matrix -> (x, iterations, true relative residual)."""
        A = A.tocsc()
        A.sort_indices()
        colptr = (A.indptr.astype(np.int64) + 1); rowval = (A.indices.astype(np.int64) + 1)
        nz = np.ascontiguousarray(A.data, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(A.shape[0]); it = ctypes.c_int(); res = ctypes.c_double()
        self._ck(self.lib.fg_solve_spd(self.h, A.shape[0], ptr(colptr), ptr(rowval), ptr(nz), ptr(b), ptr(x), rtol, max_iter,
                                       ctypes.byref(it), ctypes.byref(res)))
        return x, it.value, res.value

    def solve_spd_sets(self, set_ids, rhs, D=1, scales=None, rtol=1e-13, max_iter=20000):
        """u = (sum_k scales[k] A_k) \\ rhs with the A_k resident on the device (the matrices last assembled on the sets)."""
        ids = (ctypes.c_int * len(set_ids))(*set_ids)
        sc = None if scales is None else (ctypes.c_double * len(set_ids))(*[float(v) for v in scales])
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        if rhs.size != self.nnodes * D:
            raise ValueError(f"rhs has {rhs.size} entries, expected {self.nnodes * D}")
        x = np.empty(rhs.size); it = ctypes.c_int(); res = ctypes.c_double()
        self._ck(self.lib.fg_solve_spd_sets(self.h, ids, sc, len(set_ids), D, ptr(rhs), ptr(x), rtol, max_iter, ctypes.byref(it), ctypes.byref(res)))
        return x, it.value, res.value

    def values_max(self, set_id):
        m = ctypes.c_double()
        self._ck(self.lib.fg_values_max(self.h, set_id, ctypes.byref(m)))
        return m.value

    def get_pattern_cols(self, set_id, D, col_lo, col_hi, row_offset=0):
        """Column block [col_lo, col_hi) of the pattern: (colptr relative 1-based, rowval shifted by row_offset nodes)."""
        nnz = ctypes.c_int64()
        self._ck(self.lib.fg_get_pattern_cols(self.h, set_id, D, col_lo, col_hi, row_offset, ctypes.byref(nnz), None, None))
        colptr = np.empty((col_hi - col_lo) * D + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        self._ck(self.lib.fg_get_pattern_cols(self.h, set_id, D, col_lo, col_hi, row_offset, ctypes.byref(nnz), ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def get_values_cols(self, set_id, col_lo, col_hi, out):
        self._ck(self.lib.fg_get_values_cols(self.h, set_id, col_lo, col_hi, ptr(out)))
        return out

    # ---- multi-GPU inside the library ------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self._ck(self.lib.fg_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = ctypes.create_string_buffer(bytes(uid), 128)
        self._ck(self.lib.fg_comm_init(self.h, buf, rank, world))

    def comm_destroy(self):
        self._ck(self.lib.fg_comm_destroy(self.h))

    def exchange_columns(self, set_id, which, plan):
        """plan: [(peer, (send_lo, send_hi), (recv_lo, recv_hi))] in local 0-based node columns."""
        n = len(plan)
        if n == 0:
            return
        I, L = ctypes.c_int * n, ctypes.c_int64 * n
        self._ck(self.lib.fg_exchange_columns(self.h, set_id, which, n, I(*[p[0] for p in plan]), L(*[p[1][0] for p in plan]), L(*[p[1][1] for p in plan]),
                                              L(*[p[2][0] for p in plan]), L(*[p[2][1] for p in plan])))

    def comm_allreduce_max(self, value: float) -> float:
        v = ctypes.c_double(value)
        self._ck(self.lib.fg_comm_allreduce_max(self.h, ctypes.byref(v)))
        return v.value

    def get_values(self, set_id, out):
        self._ck(self.lib.fg_get_values(self.h, set_id, ptr(out)))
        return out

    def get_values_async(self, set_id, out):
        """Download on the copy stream; `out` (pinned for a real overlap) is complete after the next sync()."""
        self._ck(self.lib.fg_get_values_async(self.h, set_id, ptr(out)))
        return out

    def assemble_bilinear_async(self, set_id, op: fg_operator, out):
        """Assembly with the download pipelined behind it (finished column blocks leave the device while the remaining clusters
        are assembled); `out` (pinned) is complete after the next sync()."""
        self._ck(self.lib.fg_assemble_bilinear_async(self.h, set_id, ctypes.byref(op), ptr(out)))
        return out

    def assemble_linear(self, set_id, op: fg_operator, out=None):
        self._ck(self.lib.fg_assemble_linear(self.h, set_id, ctypes.byref(op), ptr(out)))
        return out

    def get_vector(self, set_id, out):
        self._ck(self.lib.fg_get_vector(self.h, set_id, ptr(out)))
        return out

    def device_buffer(self, set_id, which):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        self._ck(self.lib.fg_device_buffer(self.h, set_id, which, ctypes.byref(p), ctypes.byref(n)))
        return p.value or 0, n.value

#!/usr/bin/env python
"""Correctness of the N > 1 path on real GPUs (run under torchrun, one rank per GPU; writes one JSON line on rank 0).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_multi_gpu.py [--cells 24]

Every rank assembles its z-slab of the SAME global (n, n, n) Poisson problem through the C ABI, the halo columns are
exchanged inside the library (fg_exchange_columns over NCCL), every rank fetches the column block it owns in global numbering
(fg_get_pattern_cols / fg_get_values_cols), rank 0 concatenates the blocks into one SparseMatrixCSC triple (SURVEY 8e "Return
to Julia") and compares it with (a) a single-GPU assembly of the global grid on its own device: colptr / rowval bit-exact,
nzval <= 1e-12; (b) the oracle's CPU assembly (K from the __float128 shape functions) when the grid is small enough; the load
vector goes through the same exchange and check.  The oracle is the checker here, never the product."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, default=24)
    ap.add_argument("--D", type=int, default=1)
    ap.add_argument("--xchg", default="direct", choices=["direct", "nccl"], help="direct = peer memory over CUDA IPC + NCCL handshakes (default); nccl = payload through ncclSend/ncclRecv")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    sys.stdout.flush(); fd = os.dup(1); os.dup2(2, 1)          # NCCL banners off stdout
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import flexigal_b200 as fg
    from flexigal_b200 import _lib
    from flexigal_b200.distributed import SlabPartition, HaloExchange, gather_global_csc
    n, dmax, deg = args.n, 1.35, 3
    part = SlabPartition((1.0, 1.0, 1.0), (n, n, n), dmax, world, rank, degree=deg)
    cloud = fg.NodeCloud(part.local_nodes(), part.local_dm(), "rectangular", device=local)
    ctx = cloud.ctx
    ctx.set_option("xchg_direct", 1 if args.xchg == "direct" else 0)
    sid, ps = ctx.new_set_id(), ctx.new_set_id()
    ng = ctx.set_gauss(sid, part.local_gauss())
    ctx.neighbours(sid); ctx.shape_functions(sid, "IMLS")
    ctx.generate_gauss(ps, part.pattern_conn(), fg.pgauss(deg))
    ctx.pattern_from(sid, ps)
    halo = HaloExchange(part, ctx, sid, torch, dist)
    ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), None)
    halo.exchange_values()
    ctx.assemble_linear(sid, fg.Source(5000.0).c_struct(3, ng), None)
    halo.exchange_vector()
    ctx.sync()
    glob = gather_global_csc(part, ctx, sid, 1, dist)
    F = np.empty(part.local_nodes().shape[0]); ctx.get_vector(sid, F)
    a, b = part.owned_columns()
    Fs = [None] * world if rank == 0 else None
    dist.gather_object(F[a:b], Fs, dst=0)
    # a second exchange on the same buffers must not double count (fresh assembly each time)
    ctx.assemble_bilinear(sid, fg.GradGrad(1.0).c_struct(3, ng), None); halo.exchange_values(); ctx.sync()
    glob2 = gather_global_csc(part, ctx, sid, 1, dist)
    os.dup2(fd, 1); os.close(fd)
    if rank == 0:
        model = fg.create_model((1.0, 1.0, 1.0), (n, n, n), boundary=False)
        one = fg.NodeCloud(model.x, fg.Influence_Domains(model, dmax), "rectangular", device=local)
        c1 = one.ctx
        s1 = c1.new_set_id()
        gs = fg.egauss(model.x, model.conn, fg.pgauss(deg))
        ng1 = c1.set_gauss(s1, gs)
        c1.neighbours(s1); c1.shape_functions(s1, "IMLS")
        nnz = c1.pattern(s1)
        cp, rv = c1.get_pattern(s1, 1, nnz)
        nz = np.empty(nnz); c1.assemble_bilinear(s1, fg.GradGrad(1.0).c_struct(3, ng1), nz)
        F1 = np.empty(model.nnodes); c1.assemble_linear(s1, fg.Source(5000.0).c_struct(3, ng1), F1)
        colptr, rowval, nzval = glob
        rep = {"world": world, "n": n, "exchange": args.xchg, "nodes": int(model.nnodes), "nnz": int(nnz),
               "colptr_bit_exact": bool(np.array_equal(colptr, cp)), "rowval_bit_exact": bool(np.array_equal(rowval, rv)),
               "nzval_rel_err_vs_single_gpu": float(np.abs(nzval - nz).max() / np.abs(nz).max()),
               "F_rel_err_vs_single_gpu": float(np.abs(np.concatenate(Fs) - F1).max() / np.abs(F1).max()),
               "repeat_rel_diff": float(np.abs(glob2[2] - nzval).max() / np.abs(nz).max()),
               "slabs": [int(p) for p in part.bounds]}
        if model.nnodes <= 40000:
            import efg_oracle as o
            o.build()
            dm = fg.Influence_Domains(model, dmax)
            off, dom, phiq, dphiq = o.shape_fun(gs, model.x, dm, nthreads=os.cpu_count() or 1, quad=True)
            ocp, orv = o.pattern(model.nnodes, off, dom)
            Kq = o.assemble_bilinear(gs, off, dom, phiq, dphiq, o.coef_gradgrad(3), 1, ocp, orv, quad=True)
            rep["pattern_bit_exact_vs_oracle"] = bool(np.array_equal(colptr, ocp) and np.array_equal(rowval, orv))
            rep["nzval_rel_err_vs_oracle_truth"] = float(np.abs(nzval - Kq).max() / np.abs(Kq).max())
        rep["ok"] = bool(rep["colptr_bit_exact"] and rep["rowval_bit_exact"] and rep["nzval_rel_err_vs_single_gpu"] <= 1e-12
                         and rep["F_rel_err_vs_single_gpu"] <= 1e-12 and rep["repeat_rel_diff"] <= 1e-13
                         and rep.get("pattern_bit_exact_vs_oracle", True) and rep.get("nzval_rel_err_vs_oracle_truth", 0.0) <= 1e-12)
        print(json.dumps(rep))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""The N>1 path on CPU: two gloo ranks, z-slab partition, halo-column exchange.  The numbers come from the
oracle here (no GPU in this container); what is tested is the sharding logic that bench.py / the GPU path use
unchanged -- local clouds, ownership, the exchange plan, identical column layout on both sides -- against the
single-domain result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import efg_oracle as o
    from flexigal_b200.distributed import SlabPartition, exchange_slices
    domain, div, dmax, deg = (1.0, 1.0, 2.25), (4, 4, 9), 1.35, 3          # uneven slabs: 5 + 4 cell layers
    part = SlabPartition(domain, div, dmax, world, rank, degree=deg)
    x, dm, gs = part.local_nodes(), part.local_dm(), part.local_gauss()
    off, dom, phi, dphi = o.shape_fun(gs, x, dm)
    offp, domp, _, _ = o.shape_fun(part.pattern_gauss(), x, dm)
    cp, rv = o.pattern(x.shape[0], offp, domp)                       # sparsity from own + halo Gauss positions
    K = o.assemble_bilinear(gs, off, dom, phi, dphi, o.coef_gradgrad(3), 1, cp, rv)
    F = o.assemble_linear(gs, off, dom, phi, dphi, o.coef_source(3, 7.0), 1, x.shape[0])
    Kt = exchange_slices(part, cp - 1, torch.from_numpy(K), dist, torch)
    Ft = exchange_slices(part, np.arange(x.shape[0] + 1, dtype=np.int64), torch.from_numpy(F), dist, torch)
    # single-domain reference
    model = o.create_model(domain, div)
    gdm = o.influence_domains(model, dmax)
    ggs, _ = o.background_integration(model, "Domain", deg)
    goff, gdom, gphi, gdphi = o.shape_fun(ggs, model.x, gdm)
    gcp, grv = o.pattern(model.nnodes, goff, gdom)
    gK = o.assemble_bilinear(ggs, goff, gdom, gphi, gdphi, o.coef_gradgrad(3), 1, gcp, grv)
    gF = o.assemble_linear(ggs, goff, gdom, gphi, gdphi, o.coef_source(3, 7.0), 1, model.nnodes)
    lo, hi = part.owned_layers()
    a, b = part.local_node_range(lo, hi)
    ok = bool(np.array_equal(x, model.x[part.node_offset:part.node_offset + x.shape[0]]))
    worst = 0.0
    for jl in range(a, b):
        jg = jl + part.node_offset
        rows_l = rv[cp[jl] - 1:cp[jl + 1] - 1] + part.node_offset
        rows_g = grv[gcp[jg] - 1:gcp[jg + 1] - 1]
        ok = ok and np.array_equal(rows_l, rows_g)
        worst = max(worst, float(np.abs(Kt.numpy()[cp[jl] - 1:cp[jl + 1] - 1] - gK[gcp[jg] - 1:gcp[jg + 1] - 1]).max()))
    ferr = float(np.abs(Ft.numpy()[a:b] - gF[a + part.node_offset:b + part.node_offset]).max())
    out[rank] = (ok, worst / float(np.abs(gK).max()), ferr / float(np.abs(gF).max()), b - a)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_assembly_equals_single_domain():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    owned = 0
    for r in range(world):
        ok, kerr, ferr, n = out[r]
        assert ok, f"rank {r}: local cloud / owned column patterns differ from the single-domain ones"
        assert kerr < 1e-13 and ferr < 1e-13, (r, kerr, ferr)
        owned += n
    assert owned == 5 * 5 * 10                                     # every node is owned exactly once


def test_partition_bookkeeping():
    sys.path.insert(0, ROOT)
    from flexigal_b200.distributed import SlabPartition
    parts = [SlabPartition((1.0, 1.0, 4.0), (6, 6, 24), 1.35, 4, r) for r in range(4)]
    assert [p.owned_layers() for p in parts] == [(0, 6), (6, 12), (12, 18), (18, 25)]
    assert all(p.reach == 2 and p.halo == 5 for p in parts)
    assert sum(p.local_gauss().shape[0] for p in parts) == 6 * 6 * 24 * 27
    # plans are mutually consistent: what r sends to r+1 is what r+1 expects from r
    for r in range(3):
        send = [pl for pl in parts[r].plan() if pl[0] == r + 1][0]
        recv = [pl for pl in parts[r + 1].plan() if pl[0] == r][0]
        assert send[1] == recv[2] and send[2] == recv[1]
    one = SlabPartition((1.0, 1.0, 1.0), (6, 6, 6), 1.35, 1, 0)
    assert one.plan() == [] and one.halo == 0 and one.local_nodes().shape[0] == 7 ** 3
    # uneven slabs (the 10^6-node grid over 8 ranks: 100 = 4 x 13 + 4 x 12), every layer owned once
    p8 = [SlabPartition((1.0, 1.0, 1.0), (4, 4, 100), 1.35, 8, r) for r in range(8)]
    assert [p.cz for p in p8] == [13, 13, 13, 13, 12, 12, 12, 12]
    assert [p.owned_layers()[0] for p in p8] == [0, 13, 26, 39, 52, 64, 76, 88] and p8[-1].owned_layers()[1] == 101
    for r in range(7):
        send = [pl for pl in p8[r].plan() if pl[0] == r + 1][0]
        recv = [pl for pl in p8[r + 1].plan() if pl[0] == r][0]
        assert send[1] == recv[2] and send[2] == recv[1]
    lp = p8[3].local_plan()
    assert all(0 <= a <= b <= p8[3].local_nodes().shape[0] for _, (a, b), _ in lp)
    # slabs thinner than the support reach would need a multi-hop exchange: refused
    with pytest.raises(ValueError, match="thinner"):
        SlabPartition((1.0, 1.0, 1.0), (6, 6, 7), 1.35, 7, 0)
    with pytest.raises(ValueError):
        SlabPartition((1.0, 1.0, 1.0), (6, 6, 4), 1.35, 5, 0)


def test_slab_plan_is_consistent_for_every_world_size():
    """Pure index logic of the sharding (no devices): owned node layers tile the cloud, every send range of a rank is
    the matching receive range of its neighbour, halos cover what the rank's own Gauss points can touch."""
    from flexigal_b200.distributed import SlabPartition
    for world in (1, 2, 4, 8):
        nz = 8 * world
        parts = [SlabPartition((1.0, 1.0, 1.0 * world), (4, 4, nz), 1.35, world, r) for r in range(world)]
        covered = []
        for p in parts:
            lo, hi = p.owned_layers()
            covered += list(range(lo, hi))
            # own Gauss points touch node layers [c0 + 1 - reach, c1 - 1 + reach]: inside the local cloud
            assert p.n_lo <= max(0, p.c0 + 1 - p.reach) and min(nz, p.c1 - 1 + p.reach) <= p.n_hi
            assert p.node_offset == p.n_lo * p.lay
        assert covered == list(range(nz + 1))
        plans = [dict((peer, (s, r)) for peer, s, r in p.plan()) for p in parts]
        for r in range(world):
            assert set(plans[r]) == {q for q in (r - 1, r + 1) if 0 <= q < world}
            for peer, (send, recv) in plans[r].items():
                assert plans[peer][r][1] == send and plans[peer][r][0] == recv        # my send = their receive
                lo, hi = parts[peer].owned_layers()
                assert lo <= send[0] and send[1] <= hi                               # I send columns the peer owns
                mine = parts[r].owned_layers()
                assert mine[0] <= recv[0] and recv[1] <= mine[1]                     # and receive columns I own

"""world_size-2 gloo test (CPU) of the N>1 host path: ring neighbours, slab ownership, the
staple_geometry_plan halo offsets driven over torch.distributed, and the sum-allreduce of local
reductions -- against the reference's own two-rank output committed in tests/golden/ref_4x4x4x4_r2.npz.
The per-rank arithmetic here is the CPU oracle (this is a test); the exchange and the sharding
arithmetic are the product's host logic."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import openstaple_b200 as osb
        from openstaple_b200.sharding import communicate_fermion_borders_hostonly, owned_d3_range, ring_neighbours
        from oracle.pyoracle import Restatement
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_4x4x4x4_r2.npz")))
        loc = tuple(int(x) for x in g["loc_n"])
        assert world == int(g["nranks"])
        p = osb.geometry_plan(loc, world)
        S = Restatement(*loc, nr=world)
        assert p["sizeh"] == S.sizeh
        assert ring_neighbours(rank, world) == ((rank - 1) % world, (rank + 1) % world)
        lo, hi = owned_d3_range(rank, loc[3])
        assert [int(x) for x in g["gl_snum_r%d" % rank]][p["d3_halo"]] == lo * p["vol3h"]
        # local box of this rank (scatter includes the halos), local operator, then the exchange under test
        lu, lv = S.scatter_conf(rank, g["u"]), S.scatter_vec(rank, g["v"])
        ph = S.phases(rank, tuple(g["eb"]), float(g["mu"]), float(g["charge"]))
        out = S.dslash("doe", lu, lv, ph)
        assert np.abs(out - g["doe_unsafe_r%d" % rank]).max() < 1e-15
        t = torch.from_numpy(out)
        communicate_fermion_borders_hostonly(dist, t, loc)
        ok_halo = bool(np.array_equal(t.numpy(), g["doe_exchanged_r%d" % rank]))
        # global reduction = allreduce(sum) of the local interior reductions (fermionic_utilities.c:97-118)
        nrm = torch.tensor([S.l2norm2(lv)], dtype=torch.float64)
        dist.all_reduce(nrm)
        ok_norm = abs(float(nrm) / float(g["l2norm2_global"]) - 1) < 1e-14
        # max-over-ranks timing reduction used by bench.py
        tm = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ok_max = float(tm) == float(world)
        dist.destroy_process_group()
        q.put((rank, ok_halo, ok_norm, ok_max, ""))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, False, False, False, traceback.format_exc()))


def test_two_rank_halo_exchange_and_reductions_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port(); world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_halo, ok_norm, ok_max, err in sorted(res):
        assert err == "", err
        assert ok_halo, "rank %d: halo slices differ from the reference's exchange" % rank
        assert ok_norm and ok_max


def _borders_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from make_golden import border_fields
        from openstaple_b200 import sharding as sh
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_borders_4x4x4x4_r2.npz")))
        loc = tuple(int(x) for x in g["loc_n"]); sizeh = int(g["sizeh"])
        bad = []
        for name, fn, which, thickness in (("su3_t2", sh.communicate_su3_borders_hostonly, 0, 2), ("su3_t1", sh.communicate_su3_borders_hostonly, 0, 1),
                                           ("gl3_t1", sh.communicate_gl3_borders, 0, 1), ("tamat_t1", sh.communicate_tamat_soa_borders, 1, 1),
                                           ("thmat_t1", sh.communicate_thmat_soa_borders, 1, 1)):
            t = torch.from_numpy(border_fields(sizeh, rank)[which].copy())
            fn(dist, t, loc, thickness)
            if not np.array_equal(t.numpy(), g["%s_r%d" % (name, rank)]):
                bad.append(name)
        dist.destroy_process_group()
        q.put((rank, bad, ""))
    except Exception:      # pragma: no cover
        import traceback
        q.put((rank, ["exception"], traceback.format_exc()))


def test_two_rank_link_and_force_border_exchanges_gloo():
    """host-side su3 (thickness 2 and 1), gl3, tamat and thmat border exchanges over gloo land byte-identical to the reference's
    own exchanges between two ranks (tests/golden/make_golden.py:borders_multi, mailbox MPI)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port(); world = 2
    procs = [ctx.Process(target=_borders_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, bad, err in sorted(res):
        assert err == "", err
        assert bad == [], (rank, bad)

"""bench.py's in-run parity machinery, checked on the CPU: the per-slice field generator, the slab <-> global-slice mapping and the
window comparison must accept a correct result (here: the oracle's own global Doe / Deo.Doe scattered into a rank's local+halo box)
and reject a result with a corrupted halo slice."""
import types

import numpy as np
import pytest
import torch

import bench
from oracle.pyoracle import Restatement


class _Lat:
    def __init__(self, loc, world):
        self.loc_n = loc; self.nranks = world
        self.d3_halo = 2 if world > 1 else 0
        self.nd = (loc[0], loc[1], loc[2], loc[3] + 2 * self.d3_halo)
        self.vol3h = loc[0] * loc[1] * loc[2] // 2
        self.sizeh = self.vol3h * self.nd[3]
        self.device = torch.device("cpu")

    def new_conf(self):
        return torch.zeros((8, 3, 3, self.sizeh), dtype=torch.complex128)

    def new_vec(self):
        return torch.zeros((3, self.sizeh), dtype=torch.complex128)


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1), (4, 0)])
def test_windows_accept_the_oracle_and_reject_a_bad_halo(world, rank):
    gl = (4, 4, 4, 16)
    loc = (4, 4, 4, 16 // world)
    lat = _Lat(loc, world)
    V = lat.vol3h
    # global fields from the same per-slice generator
    ug = torch.zeros((8, 3, 3, V * gl[3]), dtype=torch.complex128); vg = torch.zeros((3, V * gl[3]), dtype=torch.complex128)
    for g3 in range(gl[3]):
        us, vs = bench.slice_fields(torch, lat.device, V, g3)
        ug[..., g3 * V:(g3 + 1) * V] = us; vg[:, g3 * V:(g3 + 1) * V] = vs
    G = Restatement(*gl)
    phg = bench.staggered_phase_slices(np, gl[0], gl[1], gl[2], list(range(gl[3])), gl[3])
    assert np.array_equal(phg, G.phases(0))
    doe = G.dslash("doe", ug.numpy(), vg.numpy(), phg)
    deo = G.dslash("deo", ug.numpy(), doe, phg)
    # this rank's box of the generated fields equals the scatter of the global ones
    u, v = bench.make_fields(torch, lat, rank)
    for d3 in range(lat.nd[3]):
        g3 = bench.global_slice(lat, rank, d3)
        assert torch.equal(u[..., d3 * V:(d3 + 1) * V], ug[..., g3 * V:(g3 + 1) * V])

    def box(a):
        o = np.zeros((3, lat.sizeh), np.complex128)
        for d3 in range(lat.nd[3]):
            g3 = bench.global_slice(lat, rank, d3)
            o[:, d3 * V:(d3 + 1) * V] = a[:, g3 * V:(g3 + 1) * V]
        return torch.from_numpy(o)

    job = types.SimpleNamespace(lat=lat, torch=torch, gl=gl, loc=loc, rank=rank, world=world, dev=lat.device)
    err, ncmp, nwin = bench.parity_windows(job, box(doe), box(deo))
    assert err == 0.0 and ncmp >= 8 and nwin >= 2
    if world > 1:
        bad = box(deo)
        lo = (lat.d3_halo - 1) * V
        bad[:, lo:lo + V] = 0                       # the halo slice received from rank L
        err, _, _ = bench.parity_windows(job, box(doe), bad)
        assert err > 0.5

"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU, D3 slabs,
NCCL / peer-memory halo exchange inside acc_Deo/acc_Doe, all-reduced CG scalars -- against the
single-rank CPU oracle on the GLOBAL lattice and the reference's committed two-rank outputs.

Run directly on a multi-GPU box:  python -m pytest tests/test_gpu_multirank.py -m gpu -x -q
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _worker(rank, world, port, loc, mode, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import openstaple_b200 as osb
        from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf
        torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", rank)))
        lat = osb.Lattice(loc, nranks_d3=world, device=rank)
        lat.init_multidev(dist, async_comm_fermion=mode["async"], p2p=mode["p2p"])
        S = Restatement(*loc, nr=world)
        G = Restatement(loc[0], loc[1], loc[2], loc[3] * world)
        u = random_su3_conf(G.sizeh, 3); v = gaussian_vec(G.sizeh, 4)
        phg = G.phases(0, EB, 1.0, 2.0)
        lu, lv, ph = S.scatter_conf(rank, u), S.scatter_vec(rank, v), S.phases(rank, EB, 1.0, 2.0)
        errs = {}
        # ---- link halos: scatter WITHOUT valid halos, then communicate_su3_borders must restore them
        lu_nohalo = lu.copy()
        h = S.d3_halo * S.vol3h
        lu_nohalo[..., :h] = 0; lu_nohalo[..., S.sizeh - h:] = 0
        du = lat.to_device(lu_nohalo)
        lat.communicate_su3_borders(du, 2)
        got = du.cpu().numpy()
        errs["su3_borders_rows01"] = float(np.abs(got[:, :2] - lu[:, :2]).max())        # rows r0,r1 only are exchanged
        du = lat.to_device(lu)
        # ---- fermion halos
        lv_nohalo = lv.copy(); lv_nohalo[:, :h] = 0; lv_nohalo[:, S.sizeh - h:] = 0
        dv = lat.to_device(lv_nohalo)
        lat.communicate_fermion_borders(dv)
        r1lo, r1hi = S.g.r1_lo, S.g.r1_hi
        errs["fermion_borders"] = float(np.abs(dv.cpu().numpy()[:, r1lo:r1hi] - lv[:, r1lo:r1hi]).max())
        dv = lat.to_device(lv); dph = lat.to_device(ph)
        # ---- operator with exchange: compare interior + 1 halo slice with the global oracle result
        for name, which in (("acc_Doe", "doe"), ("acc_Deo", "deo")):
            want = S.scatter_vec(rank, G.dslash(which, u, v, phg))
            out = lat.new_vec()
            for _ in range(3):          # repeated: exercises the double-buffered peer staging
                getattr(lat, name)(du, out, dv, dph)
            o = out.cpu().numpy()
            errs[name] = _relerr(o[:, r1lo:r1hi], want[:, r1lo:r1hi])
        # ---- dirac_times (fermion_matrix.c:196-205): rank 0 counts and times the BLOCKING exchanges, nothing else does
        import ctypes as C

        class _DT(C.Structure):
            _fields_ = [("totTransferTime", C.c_double), ("count", C.c_uint)]
        dt = _DT.in_dll(lat.L, "dirac_times")
        if mode["async"] == 0 and mode["p2p"] == 0 and rank == 0:
            errs["dirac_times"] = 0.0 if (dt.count == 6 and dt.totTransferTime > 0) else 1.0
        else:
            errs["dirac_times"] = 0.0 if dt.count == 0 else 1.0
        # ---- M^+M and reductions
        pars = lat.ferm_param(0.0507, dph)
        out, tmp = lat.new_vec(), lat.new_vec()
        lat.fermion_matrix_multiplication(du, out, dv, tmp, pars)
        want = S.scatter_vec(rank, G.mdagm(u, v, phg, 0.0507))
        errs["mdagm"] = _relerr(out.cpu().numpy()[:, r1lo:r1hi], want[:, r1lo:r1hi])
        errs["l2norm2"] = abs(lat.l2norm2_global(dv) / G.l2norm2(v) - 1)
        # ---- CG-M against the global oracle
        shifts = np.array([1e-3, 1e-2, 0.1, 1.0])
        wantx, cg_ref, ok, _ = G.multishift_invert(u, phg, 0.0507, shifts, v, 1e-8, 5000)
        approx = osb.RationalApprox.make(1.0, np.ones(4), shifts)
        sol, ps = lat.new_vec(4), lat.new_vec(4)
        r, hh, s, p = (lat.new_vec() for _ in range(4))
        st, cg = lat.multishift_invert(du, pars, approx, sol, dv, 1e-8, r, hh, s, p, ps, 5000)
        got = sol.cpu().numpy()
        r0lo, r0hi = S.g.r0_lo, S.g.r0_hi
        e = 0.0
        for i in range(4):
            w = S.scatter_vec(rank, wantx[i])
            e = max(e, _relerr(got[i][:, r0lo:r0hi], w[:, r0lo:r0hi]))
        errs["cgm_sol"] = e
        errs["cgm_iters"] = abs(cg - cg_ref) / cg_ref
        errs["cgm_status"] = 0.0 if st == 1 else 1.0
        # ---- fermion-force outer products (row N2): acc_Doe with its exchange + outer products over the interior
        shg = gaussian_vec(G.sizeh, 5, n=3); ra = np.array([0.4, -1.1, 0.7])
        auxg = np.zeros((8, 3, 3, G.sizeh), np.complex128)
        G.compute_fermion_force(u, auxg, shg, phg, ra)
        dsh = lat.to_device(np.stack([S.scatter_vec(rank, shg[i]) for i in range(3)]))
        daux = lat.new_conf()
        fp = lat.ferm_param(0.0507, dph); fp.approx_md.approx_order = 3
        for i in range(3):
            fp.approx_md.RA_a[i] = ra[i]
        lat.ker_openacc_compute_fermion_force(du, daux, dsh, lat.new_vec(), lat.new_vec(), fp)
        wantaux = S.scatter_conf(rank, auxg)
        ilo, ihi = S.d3_halo * S.vol3h, (S.d3_halo + loc[3]) * S.vol3h
        errs["force"] = _relerr(daux.cpu().numpy()[..., ilo:ihi], wantaux[..., ilo:ihi])
        # ---- stout smearing (row N4): two levels, link halos (thickness 2) exchanged after each (stouting.c:27-72)
        wantst = G.stout_wrapper(u, 0.15, 2)
        arr = torch.zeros((2, 8, 3, 3, S.sizeh), dtype=torch.complex128, device=lat.device)
        lat.set_stout(0.15, 2, lat.new_conf(), lat.new_conf(), lat.new_tamat())
        lat.stout_wrapper(du, arr, 0)
        got = arr.cpu().numpy()
        errs["stout"] = max(_relerr(got[l][:, :2], S.scatter_conf(rank, wantst[l])[:, :2]) for l in range(2))   # halos included
        # ---- stouted force chain Sigma' -> Sigma with its four border exchanges (fermion_force.c:52-163)
        spg = gaussian_vec(G.sizeh, 6, n=24).reshape(8, 3, 3, G.sizeh).copy()
        tag = G.stout_isotropic(u, 0.15)[3]
        lamg, _ = G.compute_lambda(spg, u, tag)
        sgg = spg.copy(); G.compute_sigma(lamg, u, sgg, tag, 0.15)
        lat.set_stout(0.15, 1)
        dsg = lat.to_device(S.scatter_conf(rank, spg))
        lat.compute_sigma_from_sigma_prime_backinto_sigma_prime(dsg, lat.new_tamat(), lat.new_tamat(), du, lat.new_conf(), 0)
        wsg = S.scatter_conf(rank, sgg)
        flo, fhi = (S.d3_halo - 1) * S.vol3h, (S.d3_halo + loc[3] + 1) * S.vol3h      # interior + the exchanged halo slice
        errs["stout_force"] = _relerr(dsg.cpu().numpy()[..., flo:fhi], wsg[..., flo:fhi])
        # ---- operator with a field (field_times_fermion_matrix.c): exchange through communicate_fermion_borders
        def scat(a):        # any [..., gl_sizeh] field -> this rank's local+halo box: whole d3 slices (communications.c:1104-1257)
            a2 = a.reshape(-1, G.sizeh); o = np.zeros((a2.shape[0], S.sizeh), a.dtype)
            for d3 in range(S.nd[3]):
                g3 = (d3 + loc[3] * rank - S.d3_halo) % (loc[3] * world)
                o[:, d3 * S.vol3h:(d3 + 1) * S.vol3h] = a2[:, g3 * S.vol3h:(g3 + 1) * S.vol3h]
            return o.reshape(a.shape[:-1] + (S.sizeh,))
        rng = np.random.default_rng(7); freg, fimg = rng.standard_normal((8, G.sizeh)), rng.standard_normal((8, G.sizeh))
        out = lat.new_vec()
        lat.acc_Deo_wf(du, out, dv, dph, lat.to_device(scat(freg)), lat.to_device(scat(fimg)))
        want = S.scatter_vec(rank, G.dslash_wf("deo", u, v, phg, freg, fimg))
        errs["deo_wf"] = _relerr(out.cpu().numpy()[:, r1lo:r1hi], want[:, r1lo:r1hi])
        # ---- the whole MD fermion force (fermion_force.c:166-357): smearing, CG-M, outer products, Sigma' -> Sigma, TA
        fin = gaussian_vec(G.sizeh, 8, n=1)
        flg = [dict(mass=0.08, ph=phg, number_of_ps=1, first_ps=0, ra_a=[0.4, 0.1], ra_b=[0.02, 0.3])]
        want_ipdot = G.fermion_force(u, flg, fin, 1e-10, 5000, 0.12, 1)[0]
        lat.set_stout(0.12, 1, lat.new_conf(), lat.new_conf(), lat.new_tamat())
        lat.set_force_globals(aux_th=lat.new_tamat(), aux_ta=lat.new_tamat())
        fpars = lat.ferm_param_array([dict(mass=0.08, phases=dph, number_of_ps=1, first_ps=0, ra_a=[0.4, 0.1], ra_b=[0.02, 0.3])])
        ip = osb.InverterPackage()
        lat.setup_inverter_package_dp(ip, du, lat.new_vec(2), 2, lat.new_vec(), lat.new_vec(), lat.new_vec(), lat.new_vec())
        ipdot = lat.new_tamat()
        lat.fermion_force_soloopenacc(du, torch.zeros((1, 8, 3, 3, S.sizeh), dtype=torch.complex128, device=lat.device), lat.new_conf(),
                                      ipdot, fpars, 1, lat.to_device(np.stack([S.scatter_vec(rank, fin[0])])), 1e-10, lat.new_conf(),
                                      lat.new_vec(2), ip, 5000)
        # tamat_soa packs three COMPLEX arrays and two real ones per link: slice by site field by field
        errs["full_force"] = max(_relerr(a[..., ilo:ihi], scat(np.ascontiguousarray(b))[..., ilo:ihi])
                                 for a, b in zip(osb.tamat_fields(ipdot), osb.tamat_fields(want_ipdot)))
        lat.shutdown_multidev()
        dist.destroy_process_group()
        q.put((rank, errs, ""))
    except Exception:      # pragma: no cover
        import traceback
        q.put((rank, {}, traceback.format_exc()))


def _run(world, loc, mode):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, loc, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, errs, tb in sorted(res):
        assert tb == "", tb
        assert errs["su3_borders_rows01"] == 0.0 and errs["fermion_borders"] == 0.0, (rank, errs)
        for k in ("acc_Doe", "acc_Deo", "mdagm", "force", "stout", "stout_force"):
            assert errs[k] < 1e-13, (rank, k, errs)
        assert errs["l2norm2"] < 1e-13 and errs["cgm_status"] == 0.0 and errs["dirac_times"] == 0.0, (rank, errs)
        assert errs["cgm_iters"] <= 0.02 and errs["cgm_sol"] < 1e-6, (rank, errs)
        assert errs["deo_wf"] < 1e-13 and errs["full_force"] < 1e-7, (rank, errs)      # force: iterative solves to 1e-10 inside
    print("multi-rank errors:", sorted(res)[0][1])


@pytest.mark.parametrize("mode", [dict(**{"async": a, "p2p": p}) for a, p in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4))],
                         ids=["sync-nccl", "async-nccl", "p2p-one-launch-staged-halos", "p2p-three-queues", "p2p-launch+unpack", "p2p-one-launch-eager"])
@pytest.mark.parametrize("world,loc", [(2, (8, 8, 8, 8)), (2, (8, 4, 6, 2))])
def test_two_gpus(world, loc, mode):
    _run(world, loc, mode)


@pytest.mark.parametrize("mode", [dict(**{"async": 1, "p2p": 1}), dict(**{"async": 1, "p2p": 0})], ids=["p2p", "nccl"])
def test_four_gpus(mode):
    _run(4, (8, 8, 8, 4), mode)


def test_eight_gpus_all_surface():
    """LOC_N3 = 2: no bulk at all, every site is on a face (the 64^3 x 16 strong-scaling end point)."""
    _run(8, (8, 8, 8, 2), {"async": 1, "p2p": 1})

"""The D3-slab code path on ONE GPU (staple_init_loopback): local+halo box, segmented operator launch with face / bulk / unpack
blocks, per-chunk flags, staged halos inside M^+M and CG-M, the pipelined host round trip -- every kernel of the peer-memory
transport, with this rank as its own L and R neighbour.  The lattice is then the single-rank LOC lattice (periodic in d3) stored
with halos, so the oracle is the plain single-rank restatement and the halo slices must hold the periodic images.  These tests
run on the single-GPU box; tests/test_gpu_multirank.py repeats them over NVLink."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)
MODES = {"staged-halos": 1, "three-queues": 2, "launch+unpack": 3, "one-launch-eager": 4, "copies": 0}


def _box(a, loc3, V, halo=2):
    """single-rank field [..., loc3*V] -> local+halo box [..., (loc3+2*halo)*V] with periodic images in the halo slices"""
    out = np.zeros(a.shape[:-1] + ((loc3 + 2 * halo) * V,), a.dtype)
    for d3 in range(loc3 + 2 * halo):
        g3 = (d3 - halo) % loc3
        out[..., d3 * V:(d3 + 1) * V] = a[..., g3 * V:(g3 + 1) * V]
    return out


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module", params=[(8, 8, 8, 8), (8, 4, 6, 2), (6, 6, 6, 4)], ids=["8x8x8x8", "8x4x6x2-no-bulk", "6x6x6x4-ragged-chunks"])
def setup(request):
    import openstaple_b200 as osb
    from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf
    loc = request.param
    S = Restatement(*loc)
    u = random_su3_conf(S.sizeh, 3); v = gaussian_vec(S.sizeh, 4)
    ph = S.phases(0, EB, 1.0, 2.0)
    return osb, loc, S, u, v, ph


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
def test_loopback_operator_and_solver(setup, mode):
    osb, loc, S, u, v, ph = setup
    lat = osb.Lattice(loc, nranks_d3=2, device=0)
    lat.init_loopback(MODES[mode])
    try:
        V = S.vol3h; L3 = loc[3]
        box = lambda a: _box(a, L3, V)
        du, dph = lat.to_device(box(u)), lat.to_device(box(ph))
        r1lo, r1hi = lat.ranges[2], lat.ranges[3]
        # fermion and link halos from an array without halos
        nh = box(v); nh[:, :2 * V] = 0; nh[:, (L3 + 2) * V:] = 0
        dv = lat.to_device(nh)
        lat.communicate_fermion_borders(dv)
        assert np.array_equal(dv.cpu().numpy()[:, r1lo:r1hi], box(v)[:, r1lo:r1hi])
        un = box(u); un[..., :2 * V] = 0; un[..., (L3 + 2) * V:] = 0
        dun = lat.to_device(un)
        lat.communicate_su3_borders(dun, 2)
        assert np.array_equal(dun.cpu().numpy()[:, :2], box(u)[:, :2])
        # operator with exchange, repeated (double-buffered staging), interior + halo slices
        for name, which in (("acc_Doe", "doe"), ("acc_Deo", "deo")):
            want = box(S.dslash(which, u, v, ph))
            out = lat.new_vec()
            for _ in range(3):
                getattr(lat, name)(du, out, dv, dph)
            assert _relerr(out.cpu().numpy()[:, r1lo:r1hi], want[:, r1lo:r1hi]) < 1e-13, name
        # M^+M (API: out AND temp1 leave with valid halos) and the reductions
        pars = lat.ferm_param(0.0507, dph)
        out, tmp = lat.new_vec(), lat.new_vec()
        lat.fermion_matrix_multiplication(du, out, dv, tmp, pars)
        assert _relerr(out.cpu().numpy()[:, r1lo:r1hi], box(S.mdagm(u, v, ph, 0.0507))[:, r1lo:r1hi]) < 1e-13
        assert _relerr(tmp.cpu().numpy()[:, r1lo:r1hi], box(S.dslash("doe", u, v, ph))[:, r1lo:r1hi]) < 1e-13
        assert abs(lat.l2norm2_global(dv) / S.l2norm2(v) - 1) < 1e-13
        # CG-M: staged halos of h and s inside the iteration (mode 1), eager otherwise
        shifts = np.array([1e-3, 1e-2, 0.1, 1.0])
        wantx, cg_ref, ok, _ = S.multishift_invert(u, ph, 0.0507, shifts, v, 1e-8, 5000)
        approx = osb.RationalApprox.make(1.0, np.ones(4), shifts)
        sol, ps = lat.new_vec(4), lat.new_vec(4)
        r, hh, s, p = (lat.new_vec() for _ in range(4))
        st, cg = lat.multishift_invert(du, pars, approx, sol, dv, 1e-8, r, hh, s, p, ps, 5000)
        assert st == 1 and abs(cg - cg_ref) <= 0.02 * cg_ref, (cg, cg_ref)
        got = sol.cpu().numpy()
        for i in range(4):        # update range R1: the halo slices of the solutions are kept consistent by the BLAS over R1
            assert _relerr(got[i][:, r1lo:r1hi], box(wantx[i])[:, r1lo:r1hi]) < 1e-6
        # restarted CG and the FP32-inner mixed-precision CG through the wrappers
        ip = osb.InverterPackage()
        lat.setup_inverter_package_dp(ip, du, ps, 4, r, hh, s, p)
        x = lat.new_vec()
        lat.set_inverter_tricks(0, 0, 0.1, 10000)
        its = lat.inverter_wrapper(ip, pars, x, dv, 1e-9, 5000, 0.01, osb.CONVERGENCE_NONCRITICAL)
        wx, it_ref, _ = S.cg(u, ph, 0.0507, v, 1e-9, 5000, 0.01)[:3]
        assert abs(its - it_ref) <= 0.02 * it_ref + 1 and _relerr(x.cpu().numpy()[:, r1lo:r1hi], box(wx)[:, r1lo:r1hi]) < 1e-7
        # host round trip in one call (pipelined over d3 chunks when the peer-memory single-launch transport is on)
        h_in = lat.host_array((3, lat.sizeh), np.complex128); h_out = lat.host_array((3, lat.sizeh), np.complex128)
        h_in.np[...] = box(v)
        d_tmp = lat.new_vec()
        for chunk in (0, 1, 3):
            h_out.np[...] = 0
            lat.acc_Doe_Deo_streamed(du, h_out, h_in, d_tmp, dph, chunk)
            want = box(S.dslash("deo", u, S.dslash("doe", u, v, ph), ph))
            assert _relerr(h_out.np[:, r1lo:r1hi], want[:, r1lo:r1hi]) < 1e-13, chunk
        h_in.free(); h_out.free()
    finally:
        lat.shutdown_multidev()


def test_loopback_fp32(setup):
    osb, loc, S, u, v, ph = setup
    import torch
    lat = osb.Lattice(loc, nranks_d3=2, device=0)
    lat.init_loopback(1)
    try:
        V = S.vol3h; L3 = loc[3]
        box = lambda a: _box(a, L3, V)
        uf, vf = u.astype(np.complex64), v.astype(np.complex64)
        phf = S.phases(0, EB, 1.0, 2.0, single=True)
        du, dv, dph = lat.to_device(box(uf)), lat.to_device(box(vf)), lat.to_device(box(phf))
        r1lo, r1hi = lat.ranges[2], lat.ranges[3]
        pars = lat.ferm_param(0.0507, None, dph)
        out, tmp = lat.new_vec(single=True), lat.new_vec(single=True)
        lat.fermion_matrix_multiplication(du, out, dv, tmp, pars)
        assert _relerr(out.cpu().numpy()[:, r1lo:r1hi], box(S.mdagm(uf, vf, phf, 0.0507))[:, r1lo:r1hi]) < 2e-6
        shifts = np.array([1e-2, 0.1, 1.0])
        wantx, cg_ref, ok, _ = S.multishift_invert(uf, phf, 0.0507, shifts, vf, 1e-4, 5000)
        approx = osb.RationalApprox.make(1.0, np.ones(3), shifts)
        sol, ps = lat.new_vec(3, single=True), lat.new_vec(3, single=True)
        r, hh, s, p = (lat.new_vec(single=True) for _ in range(4))
        st, cg = lat.multishift_invert(du, pars, approx, sol, dv, 1e-4, r, hh, s, p, ps, 5000)
        assert st == 1 and abs(cg - cg_ref) <= 0.02 * cg_ref + 1, (cg, cg_ref)
    finally:
        lat.shutdown_multidev()


def test_reinit_with_a_larger_lattice_rebuilds_the_mailbox():
    """staging slots are sized by vol3h; a second staple_init_geometry with a larger lattice must not reuse them"""
    import openstaple_b200 as osb
    from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf
    for loc in ((4, 4, 4, 4), (8, 8, 8, 4), (4, 4, 4, 4)):
        lat = osb.Lattice(loc, nranks_d3=2, device=0)
        lat.init_loopback(1)
        S = Restatement(*loc)
        u = random_su3_conf(S.sizeh, 3); v = gaussian_vec(S.sizeh, 4); ph = S.phases(0)
        box = lambda a: _box(a, loc[3], S.vol3h)
        out = lat.new_vec()
        lat.acc_Doe(lat.to_device(box(u)), out, lat.to_device(box(v)), lat.to_device(box(ph)))
        r1lo, r1hi = lat.ranges[2], lat.ranges[3]
        assert _relerr(out.cpu().numpy()[:, r1lo:r1hi], box(S.dslash("doe", u, v, ph))[:, r1lo:r1hi]) < 1e-13
    lat.shutdown_multidev()

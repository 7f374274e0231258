"""Parity of the callers of the path through the C ABI: the whole MD fermion force (OpenAcc/fermion_force.c:166-357, both
precisions), eo_inversion (Meas/ferm_meas.c:50-72) and the operator with a field (OpenAcc/field_times_fermion_matrix.c)
against the CPU oracle and the committed outputs of the reference's own build (tests/golden/make_golden.py:callers_single).

Tolerances.  The operator with a field is a single kernel: FP64 relative 1e-13.  The force and eo_inversion contain
ITERATIVE solves to a requested residual (1e-10 here), so two correct implementations agree to residual x condition
number (~1e3 for the smallest shift here), not to rounding: 1e-7 is asserted and the measured values are printed;
iteration counts are compared through multishift_invert_iterations (+-2%).  FP32 force: solves to 1e-5 in float, 5e-3
on the force."""
import ctypes as C

import numpy as np
import pytest

from conftest import callers_flavours, relerr
from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf

pytestmark = pytest.mark.gpu
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)


@pytest.fixture(scope="module")
def osb():
    import openstaple_b200
    return openstaple_b200


def _force(osb, lat, u, flavours, ferm_in, res, rho, steps, single=False, sp_accel=False):
    """allocate what alloc_vars.c allocates for the force and call fermion_force_soloopenacc[_f] -> (ipdot, gl3_aux, iterations)"""
    cd = np.complex64 if single else np.complex128
    du = lat.to_device(u.astype(cd))
    nsh = max(len(f["ra_b"]) for f in flavours)
    stout = lat.torch.zeros((max(steps, 1), 8, 3, 3, lat.sizeh), dtype=du.dtype, device=du.device)
    gl3, taux = lat.new_conf(single), lat.new_conf(single)
    ipdot, th, ta = lat.new_tamat(single), lat.new_tamat(single), lat.new_tamat(single)
    lat.set_stout(rho, steps, auxbis=lat.new_conf(single), staples=lat.new_conf(single), ipdot=lat.new_tamat(single), single=single)
    conf_f = lat.new_conf(True) if sp_accel else None
    lat.set_force_globals(aux_th=th, aux_ta=ta, conf_acc_f=conf_f, single=single)
    fl = []
    for f in flavours:
        ph = lat.to_device(np.ascontiguousarray(f["ph"], np.float32 if single else np.float64))
        phf = lat.to_device(np.ascontiguousarray(f["ph"], np.float32)) if sp_accel else None
        fl.append(dict(mass=f["mass"], phases=None if single else ph, phases_f=ph if single else phf,
                       number_of_ps=f["number_of_ps"], first_ps=f["first_ps"], ra_a=f["ra_a"], ra_b=f["ra_b"]))
    pars = lat.ferm_param_array(fl)
    ip = osb.InverterPackage()
    vec = lambda s=single, n=None: lat.new_vec(n, single=s)
    if single:
        lat.setup_inverter_package_sp(ip, du, vec(n=nsh), nsh, vec(), vec(), vec(), vec(), vec())
    else:
        lat.setup_inverter_package_dp(ip, du, vec(n=nsh), nsh, vec(), vec(), vec(), vec())
        if sp_accel:
            lat.setup_inverter_package_sp(ip, conf_f, vec(True, nsh), nsh, vec(True), vec(True), vec(True), vec(True), vec(True))
            lat.set_sp_globals(vec(True), vec(True, nsh))
    lat.set_inverter_tricks(singlePInvAccelMultiInv=1 if sp_accel else 0)
    shiftmulti = vec(n=nsh)
    it = C.c_int.in_dll(lat.L, "multishift_invert_iterations"); it0 = it.value
    md0 = C.c_int.in_dll(lat.L, "nMdInversionPerformed").value
    lat.fermion_force_soloopenacc(du, stout, gl3, ipdot, pars, len(fl), lat.to_device(np.ascontiguousarray(ferm_in, cd)), res,
                                  taux, shiftmulti, ip, 5000)
    lat.set_inverter_tricks()
    assert C.c_int.in_dll(lat.L, "nMdInversionPerformed").value == md0 + 1
    return ipdot.cpu().numpy(), gl3.cpu().numpy(), it.value - it0


def test_fermion_force_vs_golden(osb, golden_r1, golden_callers):
    g, gc = golden_r1, golden_callers
    lat = osb.Lattice((4, 4, 4, 4))
    for steps in (2, 0):
        ipdot, gl3, _ = _force(osb, lat, g["u"], callers_flavours(gc), gc["ferm_in"], float(gc["res"]), float(gc["rho"]), steps)
        e1, e2 = relerr(ipdot, gc["ipdot_s%d" % steps]), relerr(gl3, gc["gl3_s%d" % steps])
        print("force vs reference, %d stout levels: ipdot %.1e gl3 %.1e" % (steps, e1, e2))
        assert e1 < 1e-7 and e2 < 1e-7
    ipdot_f, _, _ = _force(osb, lat, g["u"], callers_flavours(gc, True), gc["ferm_in"], float(gc["res_f"]), float(gc["rho"]), 2, single=True)
    e = relerr(ipdot_f, gc["ipdot_s2_f"])
    print("force_f vs reference: %.1e" % e)
    assert e < 5e-3 and relerr(ipdot_f, gc["ipdot_s2"]) < 5e-3


@pytest.mark.parametrize("loc_n,steps", [((8, 8, 8, 8), 1), ((8, 4, 6, 10), 2)])
def test_fermion_force_vs_oracle(osb, loc_n, steps):
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n); n = S.sizeh
    u = random_su3_conf(n, 71); fin = gaussian_vec(n, 72, n=3)
    fl = [dict(mass=0.08, ph=S.phases(0, EB, 1.0, 2.0), number_of_ps=2, first_ps=1, ra_a=[0.4, 0.1, -0.3], ra_b=[0.02, 0.3, 1.5]),
          dict(mass=0.2, ph=S.phases(0, EB, 1.0, -1.0), number_of_ps=1, first_ps=0, ra_a=[0.25], ra_b=[0.05])]
    want_ipdot, want_gl3, _, cgs = S.fermion_force(u, fl, fin, 1e-10, 5000, 0.12, steps)
    ipdot, gl3, its = _force(osb, lat, u, fl, fin, 1e-10, 0.12, steps)
    e1, e2 = relerr(ipdot, want_ipdot), relerr(gl3, want_gl3)
    print("force vs oracle %s: ipdot %.1e gl3 %.1e, iterations %d (oracle %d)" % (loc_n, e1, e2, its, sum(cgs)))
    assert e1 < 1e-7 and e2 < 1e-7
    assert abs(its - sum(cgs)) <= 0.02 * sum(cgs)
    # the force is anti-hermitian traceless by storage; its diagonal parts are real numbers of ordinary size
    assert np.isfinite(ipdot).all() and np.abs(ipdot).max() > 1e-3
    # singlePInvAccelMultiInv: FP32 multishift + per-shift FP64 refinement (inverter_wrappers.c:45-115) gives the same force
    ipdot_a, _, _ = _force(osb, lat, u, fl, fin, 1e-10, 0.12, steps, sp_accel=True)
    e3 = relerr(ipdot_a, want_ipdot)
    print("  FP32-accelerated solves: %.1e" % e3)
    assert e3 < 1e-6


def test_eo_inversion(osb, golden_r1, golden_callers):
    g, gc = golden_r1, golden_callers
    lat = osb.Lattice((4, 4, 4, 4)); S = Restatement(4, 4, 4, 4)
    m = float(g["mass"])
    du, dph = lat.to_device(g["u"]), lat.to_device(gc["ph0"])
    pars = lat.ferm_param(m, dph)
    ip = osb.InverterPackage()
    lat.setup_inverter_package_dp(ip, du, lat.new_vec(1), 1, lat.new_vec(), lat.new_vec(), lat.new_vec(), lat.new_vec())
    ie, io = lat.to_device(g["v"]), lat.to_device(g["w"])
    oe, oo, pe, po = (lat.new_vec() for _ in range(4))
    lat.eo_inversion(ip, pars, float(gc["res"]), 5000, ie, io, oe, oo, pe, po)
    e1, e2 = relerr(oe.cpu().numpy(), gc["eo_out_e"]), relerr(oo.cpu().numpy(), gc["eo_out_o"])
    print("eo_inversion vs reference: %.1e %.1e" % (e1, e2))
    assert e1 < 1e-7 and e2 < 1e-7
    # (D + m) x = b on the full lattice, checked with the library's own operator
    t = lat.new_vec()
    lat.acc_Deo(du, t, oo, dph); assert relerr((t + m * oe).cpu().numpy(), g["v"]) < 1e-8
    lat.acc_Doe(du, t, oe, dph); assert relerr((t + m * oo).cpu().numpy(), g["w"]) < 1e-8


@pytest.mark.parametrize("loc_n", [(4, 4, 4, 4), (8, 4, 6, 10), (16, 16, 16, 16)])
def test_dslash_wf(osb, golden_r1, golden_callers, loc_n):
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n); n = S.sizeh
    if loc_n == (4, 4, 4, 4):
        u, v, ph, fre, fim = golden_r1["u"], golden_r1["v"], golden_callers["ph0"], golden_callers["field_re"], golden_callers["field_im"]
    else:
        rng = np.random.default_rng(17)
        u, v, ph = random_su3_conf(n, 18), gaussian_vec(n, 19), S.phases(0, EB, 1.0, 2.0)
        fre, fim = rng.standard_normal((8, n)), rng.standard_normal((8, n))
    du, dv, dph, dre, dim = (lat.to_device(x) for x in (u, v, ph, fre, fim))
    out = lat.new_vec()
    for which, fn, fn_unsafe in (("deo", lat.acc_Deo_wf, lat.acc_Deo_wf_unsafe), ("doe", lat.acc_Doe_wf, lat.acc_Doe_wf_unsafe)):
        want = S.dslash_wf(which, u, v, ph, fre, fim)
        fn(du, out, dv, dph, dre, dim)
        assert relerr(out.cpu().numpy(), want) < 1e-13
        if loc_n == (4, 4, 4, 4):
            assert relerr(out.cpu().numpy(), golden_callers[which + "_wf"]) < 1e-13
        out2 = lat.new_vec(); fn_unsafe(du, out2, dv, dph, dre, dim)
        assert lat.torch.equal(out, out2)
    # field = 1 reduces to the plain operator; the operator is linear in the field
    one, zero = lat.torch.ones_like(dre), lat.torch.zeros_like(dre)
    plain = lat.new_vec(); lat.acc_Deo(du, plain, dv, dph); lat.acc_Deo_wf(du, out, dv, dph, one, zero)
    assert relerr(out.cpu().numpy(), plain.cpu().numpy()) < 1e-14
    a, b = lat.new_vec(), lat.new_vec()
    lat.acc_Doe_wf(du, a, dv, dph, dre, dim); lat.acc_Doe_wf(du, b, dv, dph, 2 * dre, 2 * dim)
    assert relerr(b.cpu().numpy(), 2 * a.cpu().numpy()) < 1e-14


def test_convert_su3_covers_all_links(osb):
    """float_double_conv.c:94-150: one call converts the eight su3_soa of a configuration, rows r0, r1, r2"""
    lat = osb.Lattice((4, 4, 6, 8))
    u = random_su3_conf(lat.sizeh, 4); du = lat.to_device(u)
    uf = lat.new_conf(single=True)
    lat.convert_double_to_float_su3_soa(du, uf)
    assert np.array_equal(uf.cpu().numpy(), u.astype(np.complex64))
    back = lat.new_conf()
    lat.convert_float_to_double_su3_soa(uf, back)
    assert np.array_equal(back.cpu().numpy(), u.astype(np.complex64).astype(np.complex128))


def test_tamat_thmat_conversions(osb):
    lat = osb.Lattice((4, 4, 6, 8))
    rng = np.random.default_rng(5)
    ta = rng.standard_normal((8, 8, lat.sizeh))
    d = lat.to_device(ta); f = lat.new_tamat(single=True); back = lat.new_tamat()
    lat.convert_double_to_float_tamat_soa(d, f)
    assert np.array_equal(f.cpu().numpy(), ta.astype(np.float32))
    lat.convert_float_to_double_tamat_soa(f, back)
    assert np.array_equal(back.cpu().numpy(), ta.astype(np.float32).astype(np.float64))
    f.zero_(); lat.convert_double_to_float_thmat_soa(d, f)
    assert np.array_equal(f.cpu().numpy(), ta.astype(np.float32))
    back.zero_(); lat.convert_float_to_double_thmat_soa(f, back)
    assert np.array_equal(back.cpu().numpy(), ta.astype(np.float32).astype(np.float64))
    # host-side single colour vector (vec3 by value)
    v = np.array([1.5 + 2j, -3.25 + 0.125j, 1 / 3 + 1j / 7]); vf = np.zeros(3, np.complex64); vd = np.zeros(3, np.complex128)
    lat.L.convert_double_to_float_vec3(v.ctypes.data, vf.ctypes.data); lat.L.convert_float_to_double_vec3(vf.ctypes.data, vd.ctypes.data)
    assert np.array_equal(vf, v.astype(np.complex64)) and np.array_equal(vd, vf.astype(np.complex128))


def test_min_eigenvalue(osb):
    """ker_find_min_eigenvalue_openacc (find_min_max.c:62-98) against the restatement: a converged power iteration,
    stopped by a 1e-5 relative criterion on both sides (one iteration more or less moves the value by up to that) -> 3e-5"""
    loc_n = (8, 8, 8, 8)
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n)
    u = random_su3_conf(S.sizeh, 96); w = gaussian_vec(S.sizeh, 97); ph = S.phases(0, EB, 1.0, 2.0)
    mx = S.max_eigenvalue(u, ph, 0.3, w)
    want = S.min_eigenvalue(u, ph, 0.3, w, mx * 1.1)
    du, dph = lat.to_device(u), lat.to_device(ph)
    pars = lat.ferm_param(0.3, dph)
    got_max = lat.ker_find_max_eigenvalue_openacc(du, pars, lat.new_vec(), lat.new_vec(), lat.to_device(w))
    got = lat.ker_find_min_eigenvalue_openacc(du, pars, lat.new_vec(), lat.new_vec(), lat.to_device(w), mx * 1.1)
    print("eigenvalues: max %.12g (oracle %.12g)  min %.12g (oracle %.12g)" % (got_max, mx, got, want))
    assert abs(got_max / mx - 1) < 3e-5 and abs(got / want - 1) < 3e-5


def test_shutdown_releases_and_reinit_works(osb):
    """staple_shutdown frees the library's streams, scratch and cached graphs; staple_init_geometry brings everything back"""
    loc_n = (8, 8, 8, 8)
    S = Restatement(*loc_n)
    u = random_su3_conf(S.sizeh, 31); v = gaussian_vec(S.sizeh, 32); ph = S.phases(0, EB, 1.0, 2.0)
    want = S.mdagm(u, v, ph, 0.0507)
    for _ in range(2):
        lat = osb.Lattice(loc_n)
        du, dv, dph = lat.to_device(u), lat.to_device(v), lat.to_device(ph)
        out, tmp = lat.new_vec(), lat.new_vec()
        lat.fermion_matrix_multiplication(du, out, dv, tmp, lat.ferm_param(0.0507, dph))
        assert relerr(out.cpu().numpy(), want) < 1e-13
        sol, ps = lat.new_vec(2), lat.new_vec(2)
        r, h, s, p = (lat.new_vec() for _ in range(4))
        st, cg = lat.multishift_invert(du, lat.ferm_param(0.0507, dph), osb.RationalApprox.make(1.0, np.ones(2), np.array([0.01, 0.5])),
                                       sol, dv, 1e-8, r, h, s, p, ps, 5000)
        assert st == 1 and cg > 10
        lat.synchronize()
        lat.L.staple_shutdown()
    # a compute call after shutdown fails loudly (subprocess: the library exits)
    import subprocess
    import sys
    code = ("import openstaple_b200 as o, torch\nlat = o.Lattice((4, 4, 4, 4)); v = lat.new_vec(); lat.L.staple_shutdown()\n"
            "lat.L.l2norm2_global(v.data_ptr())\nprint('SURVIVED')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       cwd=__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    assert r.returncode != 0 and "SURVIVED" not in r.stdout and "staple_init_geometry" in r.stderr

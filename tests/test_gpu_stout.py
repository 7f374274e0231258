"""Parity of isotropic stout smearing (stouting.c, plaquettes.c:196-255, su3_utilities.c:210-237,
cayley_hamilton.h) through the C ABI against the CPU oracle and the committed outputs of the reference's own build.
FP64 relative 1e-13, FP32 relative 1e-6 (max-norm)."""
import numpy as np
import pytest

from conftest import relerr
from oracle.pyoracle import Restatement, random_su3_conf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def osb():
    import openstaple_b200
    return openstaple_b200


@pytest.mark.parametrize("rho", [0.15, 2e-3])
@pytest.mark.parametrize("loc_n", [(8, 8, 8, 8), (8, 4, 6, 10), (4, 2, 2, 6)])
def test_stout_isotropic_vs_oracle(osb, loc_n, rho):
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n)
    u = random_su3_conf(S.sizeh, 61)
    wup, wstap, waux, wta = S.stout_isotropic(u, rho)
    lat.set_stout(rho, 1)
    du = lat.to_device(u)
    up, stap, aux, ta = lat.new_conf(), lat.new_conf(), lat.new_conf(), lat.new_tamat()
    stap.fill_(7.0)                                  # stout_isotropic zeroes the parking field itself
    lat.stout_isotropic(du, up, stap, aux, ta, 0)
    assert relerr(stap.cpu().numpy(), wstap) < 1e-13
    assert relerr(ta.cpu().numpy(), wta) < 1e-13
    assert relerr(aux.cpu().numpy(), waux) < 1e-13
    assert relerr(up.cpu().numpy(), wup) < 1e-13
    # the three steps through their own entry points give the same thing
    stap2, ta2, up2, aux2 = lat.new_conf(), lat.new_tamat(), lat.new_conf(), lat.new_conf()
    lat.calc_loc_staples_nnptrick_all_onlyferms(du, stap2)
    lat.RHO_times_conf_times_staples_ta_part(du, stap2, ta2, 0)
    lat.exp_minus_QA_times_conf(du, ta2, up2, aux2)
    assert relerr(stap2.cpu().numpy(), wstap) < 1e-13 and relerr(ta2.cpu().numpy(), wta) < 1e-13
    assert relerr(up2.cpu().numpy(), wup) < 1e-13
    # calc_loc_staples ACCUMULATES (mat4 += ..., su3_utilities.h:739-750)
    lat.calc_loc_staples_nnptrick_all_onlyferms(du, stap2)
    assert relerr(stap2.cpu().numpy(), 2 * wstap) < 1e-13


def test_stout_vs_golden(osb, golden_r1, golden_stout):
    g, gs = golden_r1, golden_stout
    lat = osb.Lattice((4, 4, 4, 4))
    rho = float(gs["rho"])
    du = lat.to_device(g["u"])
    up, stap, aux, ta = lat.new_conf(), lat.new_conf(), lat.new_conf(), lat.new_tamat()
    lat.set_stout(rho, 2, stap, aux, ta)
    lat.stout_isotropic(du, up, stap, aux, ta, 0)
    for got, key in ((up, "uprime"), (stap, "staples"), (aux, "exp_aux"), (ta, "tipdot")):
        assert relerr(got.cpu().numpy(), gs[key]) < 1e-13, key
    arr = lat.torch.zeros((2, 8, 3, 3, lat.sizeh), dtype=lat.torch.complex128, device=lat.device)
    lat.stout_wrapper(du, arr, 0)                     # two levels, parking arrays from the library globals
    assert relerr(arr.cpu().numpy(), gs["wrapper2"]) < 1e-13
    lat.set_stout(float(gs["rho_small"]), 1)
    lat.stout_isotropic(du, up, stap, aux, ta, 0)
    assert relerr(up.cpu().numpy(), gs["uprime_small"]) < 1e-13
    # FP32 twin
    duf = lat.to_device(g["u"].astype(np.complex64))
    upf, stapf, auxf = (lat.new_conf(single=True) for _ in range(3))
    taf = lat.new_tamat(single=True)
    lat.set_stout(rho, 1)
    lat.stout_isotropic(duf, upf, stapf, auxf, taf, 0)
    assert relerr(upf.cpu().numpy(), gs["uprime_f"]) < 1e-6 and relerr(taf.cpu().numpy(), gs["tipdot_f"]) < 1e-6


def test_stout_properties_32(osb):
    """32^4: rho = 0 is the identity map bit for bit (Q = 0 -> exp = 1), smeared links stay in SU(3), and the
    smeared operator differs from the thin one (the links really changed)."""
    import torch
    import bench
    lat = osb.Lattice((32, 32, 32, 32))
    u, v = bench.make_fields(torch, lat, 9)
    up, stap, aux, ta = lat.new_conf(), lat.new_conf(), lat.new_conf(), lat.new_tamat()
    lat.set_stout(0.0, 1)
    lat.stout_isotropic(u, up, stap, aux, ta, 0)
    assert torch.equal(up[:, :2], u[:, :2])
    lat.set_stout(0.15, 1)
    lat.stout_isotropic(u, up, stap, aux, ta, 0)
    r0, r1 = up[:, 0], up[:, 1]
    assert float(((r0.abs() ** 2).sum(1) - 1).abs().max()) < 1e-13
    assert float((r0.conj() * r1).sum(1).abs().max()) < 1e-13
    ph = lat.to_device(bench.staggered_phases(lat, 0))
    o1, o2 = lat.new_vec(), lat.new_vec()
    lat.acc_Deo(u, o1, v, ph); lat.acc_Deo(up, o2, v, ph)
    assert float((o1 - o2).abs().max()) > 1e-2


# ----------------------------------------------------------------------------- force side: Sigma' -> Sigma
def _thmat(lat, single=False):
    return lat.new_tamat(single=single)          # thmat_soa[8] has the tamat packing (struct_c_def.h:45-58)


@pytest.mark.parametrize("rho", [0.15, 2e-3])
@pytest.mark.parametrize("loc_n", [(8, 8, 8, 8), (8, 4, 6, 10)])
def test_stout_force_vs_oracle(osb, loc_n, rho):
    from oracle.pyoracle import gaussian_vec
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n); n = S.sizeh
    u = random_su3_conf(n, 62)
    sp = gaussian_vec(n, 63, n=24).reshape(8, 3, 3, n).copy()
    ta = S.stout_isotropic(u, rho)[3]
    wlam, wtmp = S.compute_lambda(sp, u, ta)
    wsg = sp.copy(); wtmp2 = S.compute_sigma(wlam, u, wsg, ta, rho)
    lat.set_stout(rho, 1)
    du, dsp, dta = lat.to_device(u), lat.to_device(sp), lat.to_device(ta)
    lam, tmp = _thmat(lat), lat.new_conf()
    lat.compute_lambda(lam, dsp, du, dta, tmp)
    assert relerr(lam.cpu().numpy(), wlam) < 1e-13 and relerr(tmp.cpu().numpy(), wtmp) < 1e-13
    lat.compute_sigma(lat.to_device(wlam), du, dsp, dta, tmp, 0)
    assert relerr(dsp.cpu().numpy(), wsg) < 1e-13 and relerr(tmp.cpu().numpy(), wtmp2) < 1e-13
    # the whole chain of fermion_force.c:52-163 from U and Sigma' alone
    sg, lam2, qa, tmp3 = lat.to_device(sp), _thmat(lat), lat.new_tamat(), lat.new_conf()
    lat.compute_sigma_from_sigma_prime_backinto_sigma_prime(sg, lam2, qa, du, tmp3, 0)
    assert relerr(qa.cpu().numpy(), ta) < 1e-13 and relerr(lam2.cpu().numpy(), wlam) < 1e-13
    assert relerr(sg.cpu().numpy(), wsg) < 1e-13


def test_stout_force_vs_golden(osb, golden_r1, golden_stout, golden_stoutforce):
    g, gs, gf = golden_r1, golden_stout, golden_stoutforce
    lat = osb.Lattice((4, 4, 4, 4))
    lat.set_stout(float(gf["rho"]), 1)
    du, dta = lat.to_device(g["u"]), lat.to_device(gs["tipdot"])
    lam, tmp = _thmat(lat), lat.new_conf()
    lat.compute_lambda(lam, lat.to_device(gf["sigma_prime"]), du, dta, tmp)
    assert relerr(lam.cpu().numpy(), gf["lambda"]) < 1e-13 and relerr(tmp.cpu().numpy(), gf["lambda_tmp"]) < 1e-13
    sg = lat.to_device(gf["sigma_prime"])
    lat.compute_sigma(lat.to_device(gf["lambda"]), du, sg, dta, tmp, 0)
    assert relerr(sg.cpu().numpy(), gf["sigma"]) < 1e-13 and relerr(tmp.cpu().numpy(), gf["sigma_tmp"]) < 1e-13
    # FP32 twin (5e-6: the b_ij coefficients are ill-conditioned in float, see tests/test_oracle_vs_reference.py)
    duf, dtaf = lat.to_device(g["u"].astype(np.complex64)), lat.to_device(gs["tipdot_f"])
    lamf, tmpf = _thmat(lat, single=True), lat.new_conf(single=True)
    lat.compute_lambda(lamf, lat.to_device(gf["sigma_prime"].astype(np.complex64)), duf, dtaf, tmpf)
    assert relerr(lamf.cpu().numpy(), gf["lambda_f"]) < 5e-6
    sgf = lat.to_device(gf["sigma_prime"].astype(np.complex64))
    lat.compute_sigma(lat.to_device(gf["lambda_f"]), duf, sgf, dtaf, tmpf, 0)
    assert relerr(sgf.cpu().numpy(), gf["sigma_f"]) < 5e-6

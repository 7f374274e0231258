"""Parity of the fermion-force outer products (OpenAcc/fermion_force_utilities.[ch]) through the
C ABI against the CPU oracle and the committed outputs of the reference's own build.
FP64 relative 1e-13, FP32 relative 1e-6 (max-norm, as everywhere)."""
import numpy as np
import pytest

from conftest import relerr
from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf

pytestmark = pytest.mark.gpu
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)


@pytest.fixture(scope="module")
def osb():
    import openstaple_b200
    return openstaple_b200


def _pars(lat, ph, phf, ra_a):
    p = lat.ferm_param(0.0507, ph, phf)
    p.approx_md.approx_order = len(ra_a)
    for i, a in enumerate(ra_a):
        p.approx_md.RA_a[i] = a
    return p


@pytest.mark.parametrize("order", [1, 2, 5, 6])
@pytest.mark.parametrize("loc_n", [(8, 8, 8, 8), (8, 4, 6, 10), (2, 2, 2, 2)])
def test_compute_fermion_force_vs_oracle(osb, loc_n, order):
    """odd and even numbers of shifts: the CUDA path takes shifts in pairs (one aux_u pass per pair)."""
    lat = osb.Lattice(loc_n); S = Restatement(*loc_n); n = S.sizeh
    u = random_su3_conf(n, 41); sh = gaussian_vec(n, 42, n=order); ph = S.phases(0, EB, 1.0, 2.0)
    ra = np.linspace(-0.7, 1.3, order)
    aux0 = gaussian_vec(n, 43, n=24).reshape(8, 3, 3, n).copy()
    want = aux0.copy(); ws, wh = S.compute_fermion_force(u, want, sh, ph, ra)
    d_u, d_sh, d_ph, d_aux = lat.to_device(u), lat.to_device(sh), lat.to_device(ph), lat.to_device(aux0)
    s, h = lat.new_vec(), lat.new_vec()
    lat.ker_openacc_compute_fermion_force(d_u, d_aux, d_sh, s, h, _pars(lat, d_ph, None, ra))
    assert relerr(d_aux.cpu().numpy(), want) < 1e-13
    assert np.array_equal(s.cpu().numpy()[:, S.g.r1_lo:S.g.r1_hi], ws[:, S.g.r1_lo:S.g.r1_hi])     # loc_s = last shift
    assert relerr(h.cpu().numpy(), wh) < 1e-13                                                    # loc_h = Doe(loc_s)
    # one shift at a time through direct_product_of_fermions_into_auxmat: same additions in the same order
    d_one = lat.to_device(aux0); approx = osb.RationalApprox.make(1.0, ra, np.zeros(order))
    for i in range(order):
        lat.acc_Doe(d_u, h, d_sh[i], d_ph)
        lat.direct_product_of_fermions_into_auxmat(d_sh[i], h, d_one, approx, i)
    assert np.array_equal(d_one.cpu().numpy(), d_aux.cpu().numpy())


def test_force_chain_vs_golden(osb, golden_r1, golden_force):
    """the whole post-solve chain on the reference's own inputs and outputs (tests/golden/make_golden.py:force_single)"""
    g, gf = golden_r1, golden_force
    lat = osb.Lattice((4, 4, 4, 4))
    # FP32: the force is a SUM of outer products whose terms (|a_i| |s_i| |Doe s_i| ~ 1e4 here, the s_i being solutions
    # of (M^+M + b_i) x = phi with small b_i) are two orders of magnitude larger than the accumulated result, so the
    # FP32 bar of 1e-6 is taken relative to the magnitude of the terms; FP64 keeps 1e-13 relative to the result.
    term = float(np.abs(gf["ra_a"]).max() * np.abs(g["ms_out"]).max() ** 2)

    def relerr32(a, b):
        return float(np.abs(np.asarray(a) - np.asarray(b)).max() / term)

    for single in (False, True):
        cd, rd, tol, sfx = (np.complex64, np.float32, 1e-6, "_f") if single else (np.complex128, np.float64, 1e-13, "")
        relerr = relerr32 if single else globals()["relerr"]
        u, sh = lat.to_device(g["u"].astype(cd)), lat.to_device(g["ms_out"].astype(cd))
        ph, phf = lat.to_device(g["ph_bf"]), lat.to_device(g["phf_bf"])
        pars = _pars(lat, ph, phf, gf["ra_a"])
        aux = lat.to_device(gf["aux0"].astype(cd))
        s, h = lat.new_vec(single=single), lat.new_vec(single=single)
        lat.ker_openacc_compute_fermion_force(u, aux, sh, s, h, pars)
        assert relerr(aux.cpu().numpy(), gf["force_aux" + sfx]) < tol
        pseudo = lat.to_device(np.ascontiguousarray(gf["aux0"][::-1]).astype(cd))
        lat.multiply_backfield_times_force(pars, aux, pseudo)
        assert relerr(pseudo.cpu().numpy(), gf["backfield" + sfx]) < tol
        if not single:
            assert relerr(h.cpu().numpy(), gf["force_loc_h"]) < tol
            one = lat.to_device(gf["aux0"]); approx = osb.RationalApprox.make(1.0, [0.37], [0.0])
            lat.direct_product_of_fermions_into_auxmat(lat.to_device(g["v"]), lat.to_device(g["w"]), one, approx, 0)
            assert relerr(one.cpu().numpy(), gf["direct_product"]) < tol
            lat.accumulate_gl3soa_into_gl3soa(aux, pseudo)
            assert relerr(pseudo.cpu().numpy(), gf["accumulated"]) < tol
        ta = lat.to_device(gf["ta0"].astype(rd))
        lat.multiply_conf_times_force_and_take_ta_nophase(u, pseudo, ta)
        assert relerr(ta.cpu().numpy(), gf["ta" + sfx]) < tol
    # zero initialisers
    lat.set_tamat_soa_to_zero(ta); lat.set_su3_soa_to_zero(aux)
    assert float(ta.abs().max()) == 0.0 and float(aux.abs().max()) == 0.0


def test_force_properties_32(osb):
    """full-size (32^4) properties: linearity in the residues, TA output traceless anti-hermitian by construction
    (ic00 + ic11 + ic22 = 0 is implied by storage), U(1) phases with theta = 0 make backfield == accumulate."""
    import torch
    import bench
    lat = osb.Lattice((32, 32, 32, 32))
    u, v = bench.make_fields(torch, lat, 5)
    ph = lat.to_device(bench.staggered_phases(lat, 0))
    sh = torch.stack([v, 0.5 * v.flip(1), v.roll(7, 1)])
    s, h = lat.new_vec(), lat.new_vec()
    a1, a2 = lat.new_conf(), lat.new_conf()
    lat.ker_openacc_compute_fermion_force(u, a1, sh, s, h, _pars(lat, ph, None, [0.3, -0.2, 0.9]))
    lat.ker_openacc_compute_fermion_force(u, a2, sh, s, h, _pars(lat, ph, None, [0.6, -0.4, 1.8]))
    assert float((a2 - 2 * a1).abs().max()) < 1e-12 * float(a1.abs().max())
    zero_ph = torch.zeros_like(ph)
    p1, p2 = lat.new_conf(), lat.new_conf()
    lat.multiply_backfield_times_force(_pars(lat, zero_ph, None, [1.0]), a1, p1)
    lat.accumulate_gl3soa_into_gl3soa(a1, p2)
    assert torch.equal(p1, p2) and torch.equal(p1, a1)

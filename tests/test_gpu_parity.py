"""Parity tests proper: the CUDA path, called through the C ABI (libstaple_b200.so), against the
CPU oracle on identical seeded inputs and against the committed reference outputs (tests/golden).

Tolerances (BASELINE.json north_star): FP64 relative 1e-13, FP32 relative 1e-6, CG iteration counts
within +-2% of the reference, indexing bit-exact.
"""
import numpy as np
import pytest

from conftest import relerr
from oracle.pyoracle import Restatement, gaussian_vec, random_su3_conf

pytestmark = pytest.mark.gpu

TOL64 = 1e-13
TOL32 = 1e-6
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)


@pytest.fixture(scope="module")
def osb():
    import openstaple_b200
    return openstaple_b200


def make_case(osb, loc_n, seed=1, eb=EB, mu=1.0, charge=2.0):
    lat = osb.Lattice(loc_n)
    S = Restatement(*loc_n)
    assert S.sizeh == lat.sizeh
    c = dict(lat=lat, S=S)
    c["u"] = random_su3_conf(S.sizeh, seed); c["v"] = gaussian_vec(S.sizeh, seed + 1); c["w"] = gaussian_vec(S.sizeh, seed + 2)
    c["ph"] = S.phases(0, eb, mu, charge); c["phf"] = S.phases(0, eb, mu, charge, single=True)
    c["uf"] = c["u"].astype(np.complex64); c["vf"] = c["v"].astype(np.complex64); c["wf"] = c["w"].astype(np.complex64)
    for k in ("u", "v", "w", "ph", "phf", "uf", "vf", "wf"):
        c["d_" + k] = lat.to_device(c[k])
    return c


@pytest.mark.parametrize("loc_n", [(8, 8, 8, 8), (8, 4, 6, 10), (2, 2, 2, 2), (4, 2, 2, 6)])
def test_deo_doe_fp64(osb, loc_n):
    c = make_case(osb, loc_n)
    lat, S = c["lat"], c["S"]
    for name, which in (("acc_Deo", "deo"), ("acc_Doe", "doe"), ("acc_Deo_unsafe", "deo"), ("acc_Doe_unsafe", "doe")):
        out = lat.new_vec()
        getattr(lat, name)(c["d_u"], out, c["d_v"], c["d_ph"])
        ref = S.dslash(which, c["u"], c["v"], c["ph"])
        assert relerr(out.cpu().numpy(), ref) < TOL64, name


@pytest.mark.parametrize("loc_n", [(8, 8, 8, 8), (8, 4, 6, 10)])
def test_deo_doe_fp32(osb, loc_n):
    c = make_case(osb, loc_n)
    lat, S = c["lat"], c["S"]
    for name, which in (("acc_Deo", "deo"), ("acc_Doe", "doe")):
        out = lat.new_vec(single=True)
        getattr(lat, name)(c["d_uf"], out, c["d_vf"], c["d_phf"])
        ref = S.dslash(which, c["uf"], c["vf"], c["phf"])
        assert relerr(out.cpu().numpy(), ref) < TOL32, name


def test_zero_field_phases_exact_pi(osb):
    """theta in {0, pi}: the staggered signs and the antiperiodic boundary only."""
    c = make_case(osb, (8, 8, 8, 8), eb=(0,) * 6, mu=0.0, charge=0.0)
    lat, S = c["lat"], c["S"]
    out = lat.new_vec()
    lat.acc_Deo(c["d_u"], out, c["d_v"], c["d_ph"])
    assert relerr(out.cpu().numpy(), S.dslash("deo", c["u"], c["v"], c["ph"])) < TOL64


def test_mdagm(osb):
    c = make_case(osb, (8, 8, 8, 8))
    lat, S = c["lat"], c["S"]
    pars = lat.ferm_param(0.0507, c["d_ph"], c["d_phf"])
    out, tmp = lat.new_vec(), lat.new_vec()
    lat.fermion_matrix_multiplication(c["d_u"], out, c["d_v"], tmp, pars)
    assert relerr(out.cpu().numpy(), S.mdagm(c["u"], c["v"], c["ph"], 0.0507)) < TOL64
    assert relerr(tmp.cpu().numpy(), S.dslash("doe", c["u"], c["v"], c["ph"])) < TOL64
    lat.fermion_matrix_multiplication_shifted(c["d_u"], out, c["d_v"], tmp, pars, 0.37)
    assert relerr(out.cpu().numpy(), S.mdagm(c["u"], c["v"], c["ph"], 0.0507, 0.37)) < TOL64
    outf, tmpf = lat.new_vec(single=True), lat.new_vec(single=True)
    lat.fermion_matrix_multiplication(c["d_uf"], outf, c["d_vf"], tmpf, pars)
    assert relerr(outf.cpu().numpy(), S.mdagm(c["uf"], c["vf"], c["phf"], 0.0507)) < TOL32


def test_golden_operator(osb, golden_r1):
    """CUDA output against the reference's own output on the committed 4^4 fixture."""
    g = golden_r1
    lat = osb.Lattice(tuple(int(x) for x in g["loc_n"]))
    u, v = lat.to_device(g["u"]), lat.to_device(g["v"])
    uf, vf = lat.to_device(g["u"].astype(np.complex64)), lat.to_device(g["v"].astype(np.complex64))
    for tag in ("p0", "bf"):
        ph, phf = lat.to_device(g["ph_" + tag]), lat.to_device(g["phf_" + tag])
        pars = lat.ferm_param(float(g["mass"]), ph, phf)
        out, tmp = lat.new_vec(), lat.new_vec()
        lat.acc_Deo(u, out, v, ph); assert relerr(out.cpu().numpy(), g["deo_" + tag]) < TOL64
        lat.acc_Doe(u, out, v, ph); assert relerr(out.cpu().numpy(), g["doe_" + tag]) < TOL64
        lat.fermion_matrix_multiplication(u, out, v, tmp, pars)
        assert relerr(out.cpu().numpy(), g["mdagm_" + tag]) < TOL64
        lat.fermion_matrix_multiplication_shifted(u, out, v, tmp, pars, 0.37)
        assert relerr(out.cpu().numpy(), g["mdagm_sh_" + tag]) < TOL64
        outf, tmpf = lat.new_vec(single=True), lat.new_vec(single=True)
        lat.acc_Deo(uf, outf, vf, phf); assert relerr(outf.cpu().numpy(), g["deo_f_" + tag]) < TOL32
        lat.acc_Doe(uf, outf, vf, phf); assert relerr(outf.cpu().numpy(), g["doe_f_" + tag]) < TOL32
        lat.fermion_matrix_multiplication(uf, outf, vf, tmpf, pars)
        assert relerr(outf.cpu().numpy(), g["mdagm_f_" + tag]) < TOL32


def test_reductions(osb, golden_r1):
    c = make_case(osb, (8, 4, 6, 10))
    lat, S = c["lat"], c["S"]
    assert abs(lat.l2norm2_global(c["d_v"]) / S.l2norm2(c["v"]) - 1) < 1e-13
    assert abs(lat.real_scal_prod_global(c["d_v"], c["d_w"]) - S.real_scal_prod(c["v"], c["w"])) < 1e-11
    z = lat.scal_prod_global(c["d_v"], c["d_w"])
    assert abs(z - np.vdot(c["v"], c["w"])) < 1e-11
    assert abs(lat.l2norm2_global(c["d_vf"]) / S.l2norm2(c["vf"]) - 1) < 1e-13   # double accumulators in _f too
    assert abs(lat.real_scal_prod_global(c["d_vf"], c["d_wf"]) - S.real_scal_prod(c["vf"], c["wf"])) < 1e-10
    # determinism: bit-identical on repetition
    assert lat.l2norm2_global(c["d_v"]) == lat.l2norm2_global(c["d_v"])
    g = golden_r1
    lat = osb.Lattice((4, 4, 4, 4))
    assert abs(lat.l2norm2_global(lat.to_device(g["v"])) / float(g["l2norm2"]) - 1) < 1e-13
    assert abs(lat.real_scal_prod_global(lat.to_device(g["v"]), lat.to_device(g["w"])) - float(g["real_scal_prod"])) < 1e-11


BLAS_CASES = [
    ("combine_in1xfactor_plus_in2", lambda L, a, b, c, o: L.combine_in1xfactor_plus_in2(a, 0.37, b, o), lambda a, b, c, o: a * 0.37 + b),
    ("multiply_fermion_x_doublefactor", lambda L, a, b, c, o: L.multiply_fermion_x_doublefactor(o, 1.7), lambda a, b, c, o: 1.7 * o),
    ("combine_add_factor_x_in2_to_in1", lambda L, a, b, c, o: L.combine_add_factor_x_in2_to_in1(o, a, -0.3), lambda a, b, c, o: o - 0.3 * a),
    ("combine_in1xferm_mass2_minus_in2_minus_in3", lambda L, a, b, c, o: L.combine_in1xferm_mass2_minus_in2_minus_in3(a, 0.2, b, c, o), lambda a, b, c, o: a * 0.2 - b - c),
    ("combine_in1xferm_mass_minus_in2", lambda L, a, b, c, o: L.combine_in1xferm_mass_minus_in2(a, 0.2, o), lambda a, b, c, o: a * 0.2 - o),
    ("combine_in1_minus_in2", lambda L, a, b, c, o: L.combine_in1_minus_in2(a, b, o), lambda a, b, c, o: a - b),
    ("assign_in_to_out", lambda L, a, b, c, o: L.assign_in_to_out(a, o), lambda a, b, c, o: a + 0 * o),
    ("combine_in1_x_fact1_minus_in2_back_into_in2", lambda L, a, b, c, o: L.combine_in1_x_fact1_minus_in2_back_into_in2(a, 0.9, o), lambda a, b, c, o: 0.9 * a - o),
    ("combine_in1_minus_in2_allxfact", lambda L, a, b, c, o: L.combine_in1_minus_in2_allxfact(a, b, 1.1, o), lambda a, b, c, o: 1.1 * (a - b)),
]


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("case", BLAS_CASES, ids=[c[0] for c in BLAS_CASES])
def test_blas1(osb, case, single):
    name, run, ref = case
    lat = osb.Lattice((8, 4, 6, 10))
    dt = np.complex64 if single else np.complex128
    a, b, c, o = (gaussian_vec(lat.sizeh, s, dtype=dt) for s in (1, 2, 3, 4))
    da, db, dc, do = (lat.to_device(x) for x in (a, b, c, o))
    run(lat, da, db, dc, do)
    want = ref(a.astype(np.complex128), b.astype(np.complex128), c.astype(np.complex128), o.astype(np.complex128)).astype(dt)
    assert relerr(do.cpu().numpy(), want) < (2e-7 if single else 1e-15), name


def test_blas1_aliasing_and_multi(osb):
    """p = r + g*p (in == out) as the solvers call it, zero, inside_loop, and the multi-vector updates."""
    lat = osb.Lattice((8, 4, 6, 10))
    S = lat.sizeh
    p, r, s = (gaussian_vec(S, k) for k in (5, 6, 7))
    dp, dr, ds = lat.to_device(p), lat.to_device(r), lat.to_device(s)
    lat.combine_in1xfactor_plus_in2(dp, 0.25, dr, dp)
    assert relerr(dp.cpu().numpy(), p * 0.25 + r) < 1e-15
    out = lat.to_device(p)
    lat.combine_inside_loop(out, dr, ds, dp, 0.5)
    assert relerr(out.cpu().numpy(), p + 0.5 * (p * 0.25 + r)) < 1e-15
    assert relerr(dr.cpu().numpy(), r - 0.5 * s) < 1e-15
    lat.set_vec3_soa_to_zero(out); assert float(out.abs().max()) == 0.0
    n = 5
    x = gaussian_vec(S, 8, n=n); y = gaussian_vec(S, 9, n=n)
    dx, dy = lat.to_device(x), lat.to_device(y)
    flag = [1, 0, 1, 1, 1]; om = [0.1, 0.2, 0.3, 0.4, 0.5]
    lat.multiple_combine_in1_minus_in2x_factor_back_into_in1(dx, dy, 4, flag, om)
    want = x.copy()
    for i in range(4):
        if flag[i]:
            want[i] -= om[i] * y[i]
    assert relerr(dx.cpu().numpy(), want) < 1e-15
    gm = [1.1, 1.2, 1.3, 1.4, 1.5]
    lat.multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1(dx, 5, flag, gm, dr, om)
    r_now = dr.cpu().numpy()
    for i in range(5):
        if flag[i]:
            want[i] = gm[i] * want[i] + om[i] * r_now
    assert relerr(dx.cpu().numpy(), want) < 1e-15
    z = gaussian_vec(S, 10, n=4); dz = lat.to_device(z)
    lat.calc_new_trialsol_for_inversion_in_force(2, dz, 1)
    want = z.copy(); want[2:] = 2 * z[:2] - z[2:]
    assert relerr(dz.cpu().numpy(), want) < 1e-15
    lat.calc_new_trialsol_for_inversion_in_force(2, dz, 2)
    want2 = want.copy(); want2[:2] = 2 * want[2:] - want[:2]
    assert relerr(dz.cpu().numpy(), want2) < 1e-15


def test_conversions(osb):
    lat = osb.Lattice((8, 4, 6, 10))
    v = gaussian_vec(lat.sizeh, 3); dv = lat.to_device(v)
    f = lat.new_vec(single=True)
    lat.convert_double_to_float_vec3_soa(dv, f)
    assert np.array_equal(f.cpu().numpy(), v.astype(np.complex64))
    d = lat.new_vec()
    lat.convert_float_to_double_vec3_soa(f, d)
    assert np.array_equal(d.cpu().numpy(), v.astype(np.complex64).astype(np.complex128))
    u = random_su3_conf(lat.sizeh, 4); du = lat.to_device(u); uf = lat.new_conf(single=True)
    lat.convert_double_to_float_su3_soa(du, uf)
    assert np.array_equal(uf.cpu().numpy(), u.astype(np.complex64))


def test_multishift_vs_golden(osb, golden_r1):
    g = golden_r1
    lat = osb.Lattice((4, 4, 4, 4))
    u, v, ph = lat.to_device(g["u"]), lat.to_device(g["v"]), lat.to_device(g["ph_bf"])
    pars = lat.ferm_param(float(g["mass"]), ph)
    approx = osb.RationalApprox.make(float(g["ra_a0"]), g["ra_a"], g["shifts"])
    n = len(g["shifts"])
    out, ps = lat.new_vec(n), lat.new_vec(n)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    st, cg = lat.multishift_invert(u, pars, approx, out, v, 1e-9, r, h, s, p, ps, 5000)
    assert st == osb.INVERTER_SUCCESS
    assert abs(cg - int(g["ms_cg"])) <= max(1, 0.02 * int(g["ms_cg"])), (cg, int(g["ms_cg"]))
    assert relerr(out.cpu().numpy(), g["ms_out"]) < 1e-7      # both are solutions to residual 1e-9
    rec = lat.new_vec()
    lat.recombine_shifted_vec3_to_vec3(lat.to_device(g["ms_out"]), v, rec, approx)
    assert relerr(rec.cpu().numpy(), g["ms_recombined"]) < 1e-14
    # FP32 twin
    uf, vf, phf = (lat.to_device(g[k].astype(t)) for k, t in (("u", np.complex64), ("v", np.complex64), ("phf_bf", np.float32)))
    parsf = lat.ferm_param(float(g["mass"]), None, phf)
    outf, psf = lat.new_vec(n, single=True), lat.new_vec(n, single=True)
    rf, hf, sf, pf = (lat.new_vec(single=True) for _ in range(4))
    st, cgf = lat.multishift_invert(uf, parsf, approx, outf, vf, 1e-4, rf, hf, sf, pf, psf, 5000)
    assert st == osb.INVERTER_SUCCESS
    assert abs(cgf - int(g["ms_f_cg"])) <= max(1, 0.02 * int(g["ms_f_cg"])), (cgf, int(g["ms_f_cg"]))
    assert relerr(outf.cpu().numpy(), g["ms_f_out"]) < 1e-3


def test_multishift_vs_oracle_8(osb):
    """8^4, 8 shifts spanning 1e-5..10: iteration count within 2% of the oracle's, true residuals reached,
    early-converged shifts frozen exactly like the reference does."""
    c = make_case(osb, (8, 8, 8, 8))
    lat, S = c["lat"], c["S"]
    shifts = np.array([1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.5, 2.0, 10.0])
    mass, res = 0.0507, 1e-8
    want, cg_ref, ok, rel = S.multishift_invert(c["u"], c["ph"], mass, shifts, c["v"], res, 10000)
    pars = lat.ferm_param(mass, c["d_ph"])
    approx = osb.RationalApprox.make(1.0, np.ones(8), shifts)
    out, ps = lat.new_vec(8), lat.new_vec(8)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    st, cg = lat.multishift_invert(c["d_u"], pars, approx, out, c["d_v"], res, r, h, s, p, ps, 10000)
    assert st == osb.INVERTER_SUCCESS and ok == 1
    assert abs(cg - cg_ref) <= 0.02 * cg_ref, (cg, cg_ref)
    got = out.cpu().numpy()
    for i, b in enumerate(shifts):
        resid = c["v"] - S.mdagm(c["u"], got[i], c["ph"], mass, b)
        assert np.linalg.norm(resid) / np.linalg.norm(c["v"]) < 1.5 * res, i
        assert relerr(got[i], want[i]) < 1e-6
    it, act, ms = lat.last_solve_stats()
    assert it == cg and 0 < act <= 8 * cg
    # max_cg cut-off is honoured exactly
    st, cg2 = lat.multishift_invert(c["d_u"], pars, approx, out, c["d_v"], res, r, h, s, p, ps, 37)
    assert cg2 == 37


@pytest.mark.parametrize("order", [1, 25])
@pytest.mark.parametrize("single", [False, True])
def test_multishift_extreme_orders(osb, order, single):
    """approx_order = 1 and = MAX_APPROX_ORDER (rationalapprox.h:8), both precisions, against the oracle: iteration counts, every
    shifted solution, and the recombination with 25 terms"""
    c = make_case(osb, (4, 4, 4, 4), seed=5)
    lat, S = c["lat"], c["S"]
    shifts = np.array([0.05]) if order == 1 else np.geomspace(1e-4, 5.0, 25)
    ra = np.linspace(0.2, 1.4, order)
    mass, res = 0.0507, (1e-5 if single else 1e-9)
    u, v, ph = (c["uf"], c["vf"], c["phf"]) if single else (c["u"], c["v"], c["ph"])
    want, cg_ref, ok, _ = S.multishift_invert(u, ph, mass, shifts, v, res, 10000)
    pars = lat.ferm_param(mass, c["d_ph"], c["d_phf"])
    approx = osb.RationalApprox.make(0.3, ra, shifts)
    out, ps = lat.new_vec(order, single=single), lat.new_vec(order, single=single)
    r, h, s_, p = (lat.new_vec(single=single) for _ in range(4))
    d_u, d_v = (c["d_uf"], c["d_vf"]) if single else (c["d_u"], c["d_v"])
    st, cg = lat.multishift_invert(d_u, pars, approx, out, d_v, res, r, h, s_, p, ps, 10000)
    # +-2 % of the oracle's count; a 69-iteration FP32 solve is shorter than 2 % resolves, so at least two iterations of slack
    assert st == osb.INVERTER_SUCCESS and abs(cg - cg_ref) <= max(2, 0.02 * cg_ref), (cg, cg_ref)
    got = out.cpu().numpy()
    tol = 2e-3 if single else 1e-6          # iterative solutions: residual x condition number, not rounding
    assert max(relerr(got[i], want[i]) for i in range(order)) < tol
    rec = lat.new_vec(single=single)
    lat.recombine_shifted_vec3_to_vec3(out, d_v, rec, approx)
    assert relerr(rec.cpu().numpy(), S.recombine(got, v, 0.3, ra)) < (TOL32 if single else TOL64)


def test_cg_and_mixed(osb, golden_r1):
    g = golden_r1
    lat = osb.Lattice((4, 4, 4, 4))
    u, v, ph, phf = lat.to_device(g["u"]), lat.to_device(g["v"]), lat.to_device(g["ph_bf"]), lat.to_device(g["phf_bf"])
    uf = lat.to_device(g["u"].astype(np.complex64))
    pars = lat.ferm_param(float(g["mass"]), ph, phf)
    sol = lat.new_vec(); r, h, s, p = (lat.new_vec() for _ in range(4))
    lat.set_inverter_tricks(0, 0, 0.1, 10000)
    st, cg = lat.ker_invert_openacc(u, pars, sol, v, 1e-10, r, h, s, p, 5000, 0.01)
    assert st == osb.INVERTER_SUCCESS
    assert abs(cg - int(g["cg_cg"])) <= max(1, 0.02 * int(g["cg_cg"]))
    assert relerr(sol.cpu().numpy(), g["cg_sol"]) < 1e-8
    # mixed precision (inverter_package by value)
    ip = osb.InverterPackage()
    st_d = lat.new_vec(1); st_f = lat.new_vec(1, single=True)
    rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5))
    lat.setup_inverter_package_dp(ip, u, st_d, 1, r, h, s, p)
    lat.setup_inverter_package_sp(ip, uf, st_f, 1, rf, hf, sf, pf, of)
    lat.set_inverter_tricks(0, 1, 0.1, 10000)
    sol2 = lat.new_vec()
    st, cgm = lat.inverter_mixed_precision(ip, pars, sol2, v, 1e-10, 5000, 0.01)
    assert st == osb.INVERTER_SUCCESS
    assert abs(cgm - int(g["mixed_cg"])) <= max(1, 0.02 * int(g["mixed_cg"])), (cgm, int(g["mixed_cg"]))       # north star: +-2 %
    assert relerr(sol2.cpu().numpy(), g["mixed_sol"]) < 1e-7
    # wrapper dispatch (inverter_wrappers.c:117-159)
    sol3 = lat.new_vec()
    its = lat.inverter_wrapper(ip, pars, sol3, v, 1e-10, 5000, 0.01, osb.CONVERGENCE_NONCRITICAL)
    assert its == cgm
    lat.set_inverter_tricks(0, 0, 0.1, 10000)
    # power iteration (find_min_max.c)
    w = lat.to_device(g["w"])
    mx = lat.ker_find_max_eigenvalue_openacc(u, pars, r, h, w)
    assert abs(mx / float(g["max_eig"]) - 1) < 1e-9


@pytest.mark.parametrize("loc_n", [(4, 4, 4, 4), (8, 8, 8, 8)])
def test_device_resident_cg_equals_host_driven_cg(osb, loc_n):
    """ker_invert_openacc and inverter_mixed_precision with the iteration loop on the device (control block, recurrences in the
    kernels' tails, CUDA-graph batches) against the same solvers reading their scalars back every iteration: same iteration
    counts, same number of magic touches (same solution to rounding), and both against the oracle"""
    c = make_case(osb, loc_n)
    lat, S = c["lat"], c["S"]
    mass, res, shift = 0.0507, 1e-10, 0.003
    pars = lat.ferm_param(mass, c["d_ph"], c["d_phf"])
    ip = osb.InverterPackage()
    r, h, s, p = (lat.new_vec() for _ in range(4)); st_d = lat.new_vec(1)
    rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5)); st_f = lat.new_vec(1, single=True)
    lat.setup_inverter_package_dp(ip, c["d_u"], st_d, 1, r, h, s, p)
    lat.setup_inverter_package_sp(ip, c["d_uf"], st_f, 1, rf, hf, sf, pf, of)
    want, it_ref, _ = S.cg(c["u"], c["ph"], mass, c["v"], res, 5000, shift)
    wantm, itm_ref, _, touches_ref = S.mixed_cg(c["u"], c["u"].astype(np.complex64), c["ph"], c["phf"], mass, c["v"], res, 5000, shift)
    got = {}
    for dev in (1, 0):
        lat.L.staple_set_cg_device_loops(dev)
        for restart in (10000, 7):                        # 7: the restart branch (inverter_full.c:66-77) is taken many times
            lat.set_inverter_tricks(0, 0, 0.1, restart)
            x = lat.new_vec()
            st, cg = lat.ker_invert_openacc(c["d_u"], pars, x, c["d_v"], res, r, h, s, p, 5000, shift)
            got[("cg", dev, restart)] = (st, cg, x.cpu().numpy())
        lat.set_inverter_tricks(0, 1, 0.1, 10000)
        x = lat.new_vec()
        st, cg = lat.inverter_mixed_precision(ip, pars, x, c["d_v"], res, 5000, shift)
        got[("mixed", dev)] = (st, cg, x.cpu().numpy())
    lat.L.staple_set_cg_device_loops(1)
    lat.set_inverter_tricks(0, 0, 0.1, 10000)
    for restart in (10000, 7):
        a, b = got[("cg", 1, restart)], got[("cg", 0, restart)]
        assert a[0] == b[0] == osb.INVERTER_SUCCESS and abs(a[1] - b[1]) <= 1, (restart, a[1], b[1])     # reductions sum in different orders
        assert relerr(a[2], b[2]) < 1e-9
    assert abs(got[("cg", 1, 10000)][1] - it_ref) <= max(1, 0.02 * it_ref) and relerr(got[("cg", 1, 10000)][2], want) < 1e-8
    wr, itr, _ = S.cg(c["u"], c["ph"], mass, c["v"], res, 5000, shift, restarting_every=7)
    assert abs(got[("cg", 1, 7)][1] - itr) <= max(1, 0.02 * itr)
    a, b = got[("mixed", 1)], got[("mixed", 0)]
    assert a[0] == b[0] == osb.INVERTER_SUCCESS and abs(a[1] - b[1]) <= max(1, 0.02 * b[1]), (a[1], b[1])
    assert abs(a[1] - itm_ref) <= max(1, 0.02 * itm_ref) and relerr(a[2], wantm) < 1e-7


def test_multishift_wrapper_sp_accelerated(osb):
    c = make_case(osb, (4, 4, 4, 4))
    lat, S = c["lat"], c["S"]
    shifts = np.array([1e-3, 1e-2, 0.3]); mass, res = 0.0507, 1e-9
    pars = lat.ferm_param(mass, c["d_ph"], c["d_phf"])
    approx = osb.RationalApprox.make(1.0, np.ones(3), shifts)
    ip = osb.InverterPackage()
    r, h, s, p = (lat.new_vec() for _ in range(4)); st_d = lat.new_vec(3)
    rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5)); st_f = lat.new_vec(3, single=True)
    lat.setup_inverter_package_dp(ip, c["d_u"], st_d, 3, r, h, s, p)
    lat.setup_inverter_package_sp(ip, c["d_uf"], st_f, 3, rf, hf, sf, pf, of)
    lat.set_sp_globals(lat.new_vec(single=True), lat.new_vec(3, single=True))
    for accel in (0, 1):
        lat.set_inverter_tricks(accel, 1, 0.1, 10000)
        out = lat.new_vec(3)
        its = lat.inverter_multishift_wrapper(ip, pars, approx, out, c["d_v"], res, 5000, osb.CONVERGENCE_NONCRITICAL)
        assert its > 0
        if accel:
            # the reference's accounting, literally (inverter_wrappers.c:76,95): the FP32 multishift count once, plus once more per shift
            assert its == lat.last_solve_stats()[0] * (len(shifts) + 1)
            assert lat.last_refinement_iterations() > 0
        got = out.cpu().numpy()
        for i, b in enumerate(shifts):
            resid = c["v"] - S.mdagm(c["u"], got[i], c["ph"], mass, b)
            assert np.linalg.norm(resid) / np.linalg.norm(c["v"]) < 2 * res
    lat.set_inverter_tricks(0, 0, 0.1, 10000)


def test_host_pointer_boundary(osb):
    """Host arrays made present (posix_memalign_wrapper + enter data) behave like the reference's
    OpenACC host pointers: update device -> operator -> update host."""
    c = make_case(osb, (8, 8, 8, 8))
    lat, S = c["lat"], c["S"]
    hu = lat.host_array((8, 3, 3, lat.sizeh), np.complex128); hu.np[...] = c["u"]; hu.update_device()
    hph = lat.host_array((8, lat.sizeh), np.float64); hph.np[...] = c["ph"]; hph.update_device()
    hin = lat.host_array((3, lat.sizeh), np.complex128); hin.np[...] = c["v"]; hin.update_device()
    hout = lat.host_array((3, lat.sizeh), np.complex128)
    lat.acc_Deo(hu, hout, hin, hph)
    hout.update_host()
    assert relerr(hout.np, S.dslash("deo", c["u"], c["v"], c["ph"])) < TOL64
    # interior pointers (&out[i]) resolve too
    hmany = lat.host_array((2, 3, lat.sizeh), np.complex128)
    lat.acc_Doe(hu, hmany.ptr + 3 * lat.sizeh * 16, hin, hph)
    hmany.update_host()
    assert relerr(hmany.np[1], S.dslash("doe", c["u"], c["v"], c["ph"])) < TOL64
    for a in (hu, hph, hin, hout, hmany):
        a.free()


@pytest.mark.parametrize("on_stream", [False, True], ids=["legacy-stream-direct", "own-stream-graph"])
@pytest.mark.parametrize("loc_n,chunk", [((8, 8, 8, 8), 0), ((8, 8, 8, 8), 1), ((8, 8, 8, 8), 2), ((8, 8, 8, 8), 8),
                                         ((8, 4, 6, 10), 5), ((4, 4, 4, 2), 1), ((8, 8, 8, 48), 0), ((8, 8, 8, 12), 3)])
def test_streamed_host_round_trip(osb, loc_n, chunk, on_stream):
    """staple_acc_Doe_Deo_streamed (update device / acc_Doe / acc_Deo / update host pipelined over d3 chunks)
    is BIT-identical to the plain sequence and matches the oracle; repeated calls reuse the buffers (and, on a
    capturable stream, the cached CUDA graph of the schedule) safely."""
    import contextlib
    import torch
    ctx = torch.cuda.stream(torch.cuda.Stream()) if on_stream else contextlib.nullcontext()
    with ctx:
        c = make_case(osb, loc_n)
        lat, S = c["lat"], c["S"]
        hin = lat.host_array((3, lat.sizeh), np.complex128)
        hout = lat.host_array((3, lat.sizeh), np.complex128)
        tmp, plain = lat.new_vec(), lat.new_vec()
        for rep, src in enumerate((c["v"], c["w"], c["v"], c["w"])):
            lat.L.staple_set_streamed_mode(rep % 2)      # 0: copy-engine downloads, 1: Deo kernels store to the host
            hin.np[...] = src
            hout.np[...] = 0
            lat.acc_Doe_Deo_streamed(c["d_u"], hout, hin, tmp, c["d_ph"], chunk)
            d_src = lat.to_device(src)
            lat.acc_Doe(c["d_u"], tmp, d_src, c["d_ph"])
            lat.acc_Deo(c["d_u"], plain, tmp, c["d_ph"])
            assert np.array_equal(hout.np, plain.cpu().numpy()), rep
            ref = S.dslash("deo", c["u"], S.dslash("doe", c["u"], src, c["ph"]), c["ph"])
            assert relerr(hout.np, ref) < TOL64
        # device pointers fall through to the plain sequence
        dout = lat.new_vec()
        lat.acc_Doe_Deo_streamed(c["d_u"], dout, c["d_v"], tmp, c["d_ph"], chunk)
        lat.acc_Doe(c["d_u"], tmp, c["d_v"], c["d_ph"]); lat.acc_Deo(c["d_u"], plain, tmp, c["d_ph"])
        assert np.array_equal(dout.cpu().numpy(), plain.cpu().numpy())
        hin.free(); hout.free()
        lat.L.staple_set_streamed_mode(0)
        torch.cuda.synchronize()
    lat.use_torch_stream()


def test_not_present_pointer_aborts():
    """A plain host pointer is a fatal error (no silent CPU path), like an OpenACC `present` miss."""
    import subprocess, sys, os
    code = (
        "import numpy as np, openstaple_b200 as o\n"
        "lat = o.Lattice((4,4,4,4))\n"
        "a = np.zeros((3, lat.sizeh), np.complex128)\n"
        "lat.L.l2norm2_global(a.ctypes.data)\n"
        "print('SURVIVED')\n")
    root = os.path.join(os.path.dirname(__file__), "..")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout
    assert "not present on the device" in r.stderr


@pytest.mark.parametrize("nr,loc_n", [(2, (8, 8, 8, 4)), (4, (4, 4, 4, 2))])
def test_slab_kernels_single_process(osb, nr, loc_n):
    """Multi-rank geometry (halos of width 2) exercised rank by rank in one process: unsafe = bulk+d3p+d3m
    = d3c pieces, all equal to the oracle on the same local box; update/reduction ranges R1/R0."""
    lat = osb.Lattice(loc_n, nranks_d3=nr)
    S = Restatement(*loc_n, nr=nr)
    assert lat.sizeh == S.sizeh and lat.ranges == (S.g.r0_lo, S.g.r0_hi, S.g.r1_lo, S.g.r1_hi)
    gl = (loc_n[0], loc_n[1], loc_n[2], loc_n[3] * nr)
    G = Restatement(*gl)
    u = random_su3_conf(G.sizeh, 3); v = gaussian_vec(G.sizeh, 4)
    for rank in range(nr):
        lu, lv, ph = S.scatter_conf(rank, u), S.scatter_vec(rank, v), S.phases(rank, EB, 1.0, 2.0)
        du, dv, dph = lat.to_device(lu), lat.to_device(lv), lat.to_device(ph)
        for par, which in (("Deo", "deo"), ("Doe", "doe")):
            want = S.dslash(which, lu, lv, ph)
            o1 = lat.new_vec(); getattr(lat, "acc_%s_unsafe" % par)(du, o1, dv, dph)
            assert relerr(o1.cpu().numpy(), want) < TOL64
            o2 = lat.new_vec()
            for piece in ("bulk", "d3p", "d3m"):
                getattr(lat, "acc_%s_%s" % (par, piece))(du, o2, dv, dph)
            assert np.array_equal(o2.cpu().numpy(), o1.cpu().numpy())
            o3 = lat.new_vec()
            getattr(lat, "acc_%s_d3c" % par)(du, o3, dv, dph, lat.d3_halo, 1)
            getattr(lat, "acc_%s_d3c" % par)(du, o3, dv, dph, lat.d3_halo + 1, loc_n[3] - 1)
            assert np.array_equal(o3.cpu().numpy(), o1.cpu().numpy())
        # reductions cover the interior only, updates the interior + 1 halo slice
        assert abs(lat.l2norm2_global(dv) / S.l2norm2(lv) - 1) < 1e-13
        o = lat.new_vec(); lat.assign_in_to_out(dv, o)
        want = np.zeros_like(lv); want[:, S.g.r1_lo:S.g.r1_hi] = lv[:, S.g.r1_lo:S.g.r1_hi]
        assert np.array_equal(o.cpu().numpy(), want)


def test_properties_32(osb):
    """Size-independent properties at the BASELINE 32^4 size (too big for the CPU oracle to be quick):
    D is anti-Hermitian between parities, M^+M is Hermitian positive, everything is linear."""
    lat = osb.Lattice((32, 32, 32, 32))
    S = Restatement(32, 32, 32, 32)
    u = lat.to_device(random_su3_conf(lat.sizeh, 7))
    ph = lat.to_device(S.phases(0, EB, 1.0, 2.0))
    a, b = lat.to_device(gaussian_vec(lat.sizeh, 8)), lat.to_device(gaussian_vec(lat.sizeh, 9))
    Da, Db, t = lat.new_vec(), lat.new_vec(), lat.new_vec()
    lat.acc_Deo(u, Db, b, ph)      # b taken as an odd-site vector -> even
    lat.acc_Doe(u, Da, a, ph)      # a taken as an even-site vector -> odd
    lhs = lat.scal_prod_global(a, Db); rhs = lat.scal_prod_global(Da, b)
    assert abs(lhs + rhs) < 1e-12 * abs(lhs)            # <a, Deo b> = -<Doe a, b>
    pars = lat.ferm_param(0.0018, ph)
    Ma, Mb = lat.new_vec(), lat.new_vec()
    lat.fermion_matrix_multiplication(u, Ma, a, t, pars)
    lat.fermion_matrix_multiplication(u, Mb, b, t, pars)
    x, y = lat.scal_prod_global(a, Mb), lat.scal_prod_global(Ma, b)
    assert abs(x - y) < 1e-12 * abs(x)
    assert lat.real_scal_prod_global(a, Ma) > 0
    ab = a * 0.3 + b * (-1.2)
    Mab = lat.new_vec(); lat.fermion_matrix_multiplication(u, Mab, ab, t, pars)
    assert float((Mab - (Ma * 0.3 + Mb * (-1.2))).abs().max()) < 1e-12 * float(Mab.abs().max())
    # one oracle-checked d3 slab of the 32^4 output (a few thousand sites; seconds on the CPU)
    un, an, phn = u.cpu().numpy(), a.cpu().numpy(), ph.cpu().numpy()
    want = S.dslash("doe", un, an, phn, d3lo=31, d3hi=32)
    lo = 31 * S.vol3h
    assert relerr(Da.cpu().numpy()[:, lo:], want[:, lo:]) < TOL64


def test_multishift_graph_replay_matches_direct_launches(osb):
    """On a non-default stream CG-M replays its iteration batches as a CUDA graph: same iteration count and
    bit-identical solutions as direct launches (every dependence lives in device memory)."""
    import torch
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        c = make_case(osb, (8, 8, 8, 8))
        lat, S = c["lat"], c["S"]
        shifts = np.array([1e-4, 1e-3, 1e-2, 0.1, 1.0, 5.0])
        pars = lat.ferm_param(0.0507, c["d_ph"])
        approx = osb.RationalApprox.make(1.0, np.ones(6), shifts)
        res = []
        for graphs in (1, 0):
            lat.L.staple_set_use_graphs(graphs)
            out, ps = lat.new_vec(6), lat.new_vec(6)
            r, h, s_, p = (lat.new_vec() for _ in range(4))
            st, cg = lat.multishift_invert(c["d_u"], pars, approx, out, c["d_v"], 1e-9, r, h, s_, p, ps, 10000)
            assert st == osb.INVERTER_SUCCESS
            res.append((cg, out.cpu().numpy()))
        lat.L.staple_set_use_graphs(1)
        assert res[0][0] == res[1][0]
        assert np.array_equal(res[0][1], res[1][1])
        want, cg_ref, ok, _ = S.multishift_invert(c["u"], c["ph"], 0.0507, shifts, c["v"], 1e-9, 10000)
        assert abs(res[0][0] - cg_ref) <= 0.02 * cg_ref
        assert relerr(res[0][1], want) < 1e-7
    torch.cuda.synchronize()


@pytest.mark.parametrize("single", [False, True])
def test_multishift_fused_tail_matches_separate_kernels(osb, single):
    """The CG-M scalar recurrences run in the tail of the Deo kernel (alpha) and of the shifted pass (lambda);
    with the one-warp kernels of their own instead, iteration count and solutions are bit-identical."""
    c = make_case(osb, (8, 8, 8, 8))
    lat = c["lat"]
    shifts = np.array([1e-4, 1e-3, 1e-2, 0.1, 1.0, 5.0])
    pars = lat.ferm_param(0.0507, c["d_ph"], c["d_phf"])
    approx = osb.RationalApprox.make(1.0, np.ones(6), shifts)
    u, v = (c["d_uf"], c["d_vf"]) if single else (c["d_u"], c["d_v"])
    res = []
    for fuse in (1, 0):
        lat.L.staple_set_cgm_fuse_tail(fuse)
        l0 = lat.kernel_launches()
        out, ps = lat.new_vec(6, single=single), lat.new_vec(6, single=single)
        r, h, s_, p = (lat.new_vec(single=single) for _ in range(4))
        st, cg = lat.multishift_invert(u, pars, approx, out, v, 1e-5 if single else 1e-9, r, h, s_, p, ps, 10000)
        assert st == osb.INVERTER_SUCCESS
        res.append((cg, out.cpu().numpy(), lat.kernel_launches() - l0))
    lat.L.staple_set_cgm_fuse_tail(1)
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][2] < res[1][2]          # 4 instead of 6 launches per iteration


def test_load_configuration_from_ildg_and_ascii(osb, tmp_path):
    """N5: a configuration file goes straight into the su3_soa[8] layout the kernels read"""
    from openstaple_b200 import io as sio
    c = make_case(osb, (4, 4, 4, 8))
    lat, S = c["lat"], c["S"]
    want = S.dslash("deo", c["u"], c["v"], c["ph"])
    for fmt, writer in (("ildg", lambda p: sio.print_su3_soa_ildg_binary(c["u"], p, (4, 4, 4, 8), 12)),
                        ("ascii", lambda p: sio.print_su3_soa_ASCII(c["u"], p, (4, 4, 4, 8), 12))):
        p = str(tmp_path / ("conf." + fmt)); writer(p)
        du, cid = lat.load_configuration(p, fmt)
        assert cid == 12
        out = lat.new_vec()
        lat.acc_Deo(du, out, c["d_v"], c["d_ph"])
        assert relerr(out.cpu().numpy(), want) < (TOL64 if fmt == "ildg" else 1e-12)    # ASCII keeps 18 decimals

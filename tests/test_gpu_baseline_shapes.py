"""Parity on the BASELINE.json shapes themselves: the full 32^4 operators, M^+M and 20 CG-M iterations
against the REFERENCE's own gcc build (oracle/_ref/libref_32x32x32x32_r1.so, built by oracle/build_ref.sh from the unmodified
sources), iterate for iterate; and window-sampled comparisons with the oracle at 48^3 x 96 and 64^3 x 16, FP64 and FP32 (the
same machinery bench.py runs inside every benchmark: first / middle / last windows of 8 d3 slices)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_32x4_against_the_reference_build():
    import openstaple_b200 as osb
    from oracle.pyoracle import RefLib, gaussian_vec, random_su3_conf, ref_lib_path
    import os
    n = (32, 32, 32, 32)
    if not os.path.exists(ref_lib_path(*n)):
        pytest.skip("oracle/_ref/libref_32x32x32x32_r1.so not built (needs /root/reference at build time)")
    R = RefLib(*n)
    lat = osb.Lattice(n)
    u = random_su3_conf(R.sizeh, 11); v = gaussian_vec(R.sizeh, 12)
    ph = R.phases((5.0, -5.0, 1.0, -5.0, 5.0, 3.0), 1.0, 2.0)           # general angles: the tools/test background field
    du, dv, dph = lat.to_device(u), lat.to_device(v), lat.to_device(ph)
    out, tmp = lat.new_vec(), lat.new_vec()
    lat.acc_Doe(du, out, dv, dph)
    want_doe = R.dslash("acc_Doe", u, v, ph)
    assert relerr(out.cpu().numpy(), want_doe) < 1e-13
    lat.acc_Deo(du, out, dv, dph)
    assert relerr(out.cpu().numpy(), R.dslash("acc_Deo", u, v, ph)) < 1e-13
    mass = 0.0507
    pars = lat.ferm_param(mass, dph)
    lat.fermion_matrix_multiplication(du, out, dv, tmp, pars)
    assert relerr(out.cpu().numpy(), R.mdagm(u, v, ph, mass)) < 1e-13
    assert relerr(tmp.cpu().numpy(), want_doe) < 1e-13
    # 20 CG-M iterations, iterate for iterate (max_cg stops both solvers at the same point of the recurrences)
    shifts = np.array([2e-5, 1e-3, 3e-2, 0.8])
    wantx, cg_ref, _ = R.multishift_invert(u, ph, mass, (1.0, np.ones(4), shifts), v, 1e-30, 20)
    approx = osb.RationalApprox.make(1.0, np.ones(4), shifts)
    sol, ps = lat.new_vec(4), lat.new_vec(4)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    st, cg = lat.multishift_invert(du, pars, approx, sol, dv, 1e-30, r, h, s, p, ps, 20)
    assert cg == cg_ref == 20
    got = sol.cpu().numpy()
    for i in range(4):
        assert relerr(got[i], wantx[i]) < 1e-11, i


@pytest.mark.parametrize("single", [False, True], ids=["fp64", "fp32"])
@pytest.mark.parametrize("gl", [(48, 48, 48, 96), (64, 64, 64, 16)], ids=["48x48x48x96", "64x64x64x16"])
def test_window_sampled_parity_at_the_baseline_shapes(gl, single):
    import torch
    import openstaple_b200 as osb
    import bench
    torch.cuda.set_device(0)
    job = bench.Job(torch, None, osb, gl, 1, 0, 0)
    lat = job.lat
    if single:
        u = lat.new_conf(single=True); lat.convert_double_to_float_su3_soa(job.u, u)
        v, ph = job.v.to(torch.complex64), job.ph.to(torch.float32)
    else:
        u, v, ph = job.u, job.v, job.ph
    b, o = lat.new_vec(single=single), lat.new_vec(single=single)
    lat.acc_Doe(u, b, v, ph)
    lat.acc_Deo(u, o, b, ph)
    err, ncmp, nwin = bench.parity_windows(job, b, o, single=single)
    assert ncmp >= 16 and nwin >= 2
    assert err < (2e-6 if single else 1e-13), err
    # M^+M = m^2 - Deo Doe on the same windows: the fused epilogue against the two operators
    pars = lat.ferm_param(bench.MASS, None if single else ph, ph if single else None)
    mm, tmp = lat.new_vec(single=single), lat.new_vec(single=single)
    lat.fermion_matrix_multiplication(u, mm, v, tmp, pars)
    want = (bench.MASS ** 2) * v.to(torch.complex128) - o.to(torch.complex128)
    assert float((mm.to(torch.complex128) - want).abs().max() / want.abs().max()) < (1e-6 if single else 1e-14)
    job.close()

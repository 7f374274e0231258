#!/usr/bin/env python
"""Timings of the callers of the path on one GPU (secondary figures for DESIGN.md, not the driver's bench line):
the operator with a field (acc_Deo_wf) against its 1056 B/site, and the whole MD fermion force
(fermion_force_soloopenacc: stout smearing -> CG-M on approx_md -> outer products -> Sigma' -> Sigma -> TA) with its stages
timed one by one.

  python scripts/bench_callers.py [--global-lattice 32x32x32x32] [--stout-steps 2] [--order 9]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-lattice", default="32x32x32x32")
    ap.add_argument("--mass", type=float, default=0.0507)
    ap.add_argument("--residue", type=float, default=1e-6)
    ap.add_argument("--order", type=int, default=9)
    ap.add_argument("--stout-steps", type=int, default=2)
    ap.add_argument("--rho", type=float, default=0.15)
    args = ap.parse_args()
    import torch
    import openstaple_b200 as osb
    torch.cuda.set_device(0)
    real_stdout = os.dup(1); os.dup2(2, 1)
    loc = tuple(int(x) for x in args.global_lattice.split("x"))
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", 0)))
    lat = osb.Lattice(loc, device=0)
    u, v = bench.make_fields(torch, lat, 0)
    ph = lat.to_device(bench.staggered_phases(lat, 0))
    n = lat.sizeh
    peak, _ = bench.peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"global_lattice": args.global_lattice, "mass": args.mass}
    # ---- operator with a field: 8 x (96 + 8 + 16) + 48 + 48 = 1056 B per site
    fre, fim = torch.randn((8, n), dtype=torch.float64, device=lat.device), torch.randn((8, n), dtype=torch.float64, device=lat.device)
    a, b = v.clone(), lat.new_vec()
    ms = timeit(lambda: lat.acc_Deo_wf(u, b, a, ph, fre, fim), 50)
    ms0 = timeit(lambda: lat.acc_Deo(u, b, a, ph), 50)
    out["acc_Deo_wf"] = {"us": ms * 1e3, "hbm_GBps": 1056.0 * n / ms / 1e6, "frac_of_measured_peak": 1056.0 * n / ms / 1e6 / peak,
                         "acc_Deo_us": ms0 * 1e3}
    # ---- whole fermion force, one flavour, one pseudofermion, shipped x^(-1/4) approximation of the given order as approx_md
    steps = args.stout_steps
    lat.set_stout(args.rho, steps, lat.new_conf(), lat.new_conf(), lat.new_tamat())
    stout = torch.zeros((max(steps, 1), 8, 3, 3, n), dtype=torch.complex128, device=lat.device)
    lat.stout_wrapper(u, stout, 0)
    smeared = stout[steps - 1] if steps > 0 else u
    pars1 = lat.ferm_param(args.mass, ph)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    lmax = lat.ker_find_max_eigenvalue_openacc(smeared, pars1, r, h, v.clone())
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_abi_approx.npz")))
    tag = "m14" if args.order == 19 else "m14o9"
    mother = osb.RationalApprox.make(float(g[tag + "_a0"]), g[tag + "_a"], g[tag + "_b"], int(g[tag + "_num"]), int(g[tag + "_den"]))
    mother.lambda_min, mother.lambda_max = float(g[tag + "_lmin"]), float(g[tag + "_lmax"])
    try:
        approx = mother.rescaled((args.mass ** 2, lmax))
    except ValueError:
        approx = mother.rescaled((1e9, lmax))
    nsh = approx.approx_order
    fl = [dict(mass=args.mass, phases=ph, number_of_ps=1, first_ps=0, ra_a=list(approx.RA_a[:nsh]), ra_b=list(approx.RA_b[:nsh]))]
    fpars = lat.ferm_param_array(fl)
    th, ta = lat.new_tamat(), lat.new_tamat()
    lat.set_force_globals(aux_th=th, aux_ta=ta)
    ip = osb.InverterPackage()
    st = lat.new_vec(nsh)
    lat.setup_inverter_package_dp(ip, u, st, nsh, r, h, s, p)
    gl3, taux, ipdot, shiftmulti = lat.new_conf(), lat.new_conf(), lat.new_tamat(), lat.new_vec(nsh)
    it = __import__("ctypes").c_int.in_dll(lat.L, "multishift_invert_iterations")

    def force():
        lat.fermion_force_soloopenacc(u, stout, gl3, ipdot, fpars, 1, v, args.residue, taux, shiftmulti, ip, 20000)

    force(); torch.cuda.synchronize()
    it0 = it.value; t0 = time.perf_counter()
    force(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out["fermion_force_soloopenacc"] = {"s": wall, "stout_steps": steps, "shifts": nsh, "cg_iterations": it.value - it0,
                                        "residue": args.residue, "lambda_max": lmax}
    # stages, timed alone on the same data
    stages = {}
    stages["stout_wrapper_ms"] = timeit(lambda: lat.stout_wrapper(u, stout, 0), 3, 1)
    fp1 = fpars[0]
    t0 = time.perf_counter()
    lat.inverter_multishift_wrapper(ip_with(ip, smeared), fp1, fp1.approx_md, shiftmulti, v, args.residue, 20000, osb.CONVERGENCE_NONCRITICAL)
    torch.cuda.synchronize(); stages["multishift_s"] = time.perf_counter() - t0
    stages["outer_products_ms"] = timeit(lambda: lat.ker_openacc_compute_fermion_force(smeared, taux, shiftmulti, s, h, fp1), 3, 1)
    stages["backfield_ms"] = timeit(lambda: lat.multiply_backfield_times_force(fp1, taux, gl3), 3, 1)
    stages["sigma_chain_per_level_ms"] = timeit(lambda: lat.compute_sigma_from_sigma_prime_backinto_sigma_prime(gl3, th, ta, u, taux, 0), 3, 1)
    stages["take_ta_ms"] = timeit(lambda: lat.multiply_conf_times_force_and_take_ta_nophase(u, gl3, ipdot), 3, 1)
    out["stages"] = stages
    os.dup2(real_stdout, 1)
    line = json.dumps(out)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "bench_callers.jsonl"), "a").write(line + "\n")


def ip_with(ip, u):
    """by-value copy of the package with another gauge field (fermion_force.c:210 `ipt.u = conf_to_use`)"""
    q = type(ip).from_buffer_copy(ip)
    q.u = u.data_ptr()
    return q


if __name__ == "__main__":
    main()

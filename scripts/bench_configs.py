#!/usr/bin/env python
"""Secondary benchmarks on the BASELINE.json configs that bench.py (the driver's contract: 32^4 Deo/Doe)
does not cover: multishift CG with the shipped order-19 rational approximation (config 3), FP32 and
FP32-accelerated solves, and strong/weak scaling of Deo/Doe + CG-M on bigger lattices (configs 4, 5).

  python scripts/bench_configs.py --global-lattice 48x48x48x96 [--order19] [--mass 0.0018]
  torchrun --nproc-per-node N ... scripts/bench_configs.py --global-lattice 64x64x64x128

Prints one JSON object per line (rank 0) and appends it to gpurun_out/bench_configs.jsonl.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic field generators, clock sampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-lattice", default="48x48x48x96")
    ap.add_argument("--mass", type=float, default=0.0507)
    ap.add_argument("--residue", type=float, default=1e-8)
    ap.add_argument("--order", type=int, default=19, help="19/9: shipped x^(-1/4) approximation of that order; else geomspace shifts")
    ap.add_argument("--max-cg", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--skip-fp32", action="store_true")
    ap.add_argument("--p2p", type=int, default=int(os.environ.get("STAPLE_P2P", "1")))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import openstaple_b200 as osb

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    real_stdout = os.dup(1); os.dup2(2, 1)
    gl = tuple(int(x) for x in args.global_lattice.split("x"))
    assert gl[3] % world == 0
    loc = (gl[0], gl[1], gl[2], gl[3] // world)
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", local_rank)))
    lat = osb.Lattice(loc, nranks_d3=world, device=local_rank)
    if world > 1:
        lat.init_multidev(dist, async_comm_fermion=1, p2p=args.p2p)
    u, v = bench.make_fields(torch, lat, rank)
    ph_host = bench.staggered_phases(lat, rank)
    ph = lat.to_device(ph_host); phf = lat.to_device(ph_host.astype(np.float32))
    if world > 1:
        lat.communicate_su3_borders(u, 2); lat.communicate_fermion_borders(v)
    pars = lat.ferm_param(args.mass, ph, phf)
    interior = lat.vol3h * loc[3]
    peak, _ = bench.peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world > 1:
            t = torch.tensor([x], device=lat.device, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
        return x

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        barrier(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); barrier()
        return maxr(e0.elapsed_time(e1) / reps)

    out = {"global_lattice": args.global_lattice, "n_gpus": world, "local_lattice": "x".join(map(str, loc)),
           "halo": ("p2p" if getattr(lat, "p2p", False) else "nccl") if world > 1 else "none", "mass": args.mass}
    a, b, tmp = v.clone(), lat.new_vec(), lat.new_vec()
    # ---- operator, FP64 and FP32
    ms = timeit(lambda: (lat.acc_Doe(u, b, a, ph), lat.acc_Deo(u, tmp, b, ph)), args.reps)
    out["deo_doe_fp64"] = {"ms_per_pair": ms, "gflops": 570.0 * 2 * interior * world / ms / 1e6,
                           "hbm_GBps_per_gpu": 928.0 * 2 * interior / ms / 1e6, "frac_of_measured_peak": 928.0 * 2 * interior / ms / 1e6 / peak}
    ms = timeit(lambda: lat.fermion_matrix_multiplication(u, tmp, a, b, pars), args.reps)
    out["mdagm_fp64"] = {"ms": ms, "hbm_GBps_per_gpu": 1904.0 * interior / ms / 1e6}
    uf = None
    if not args.skip_fp32:
        uf = lat.new_conf(single=True); lat.convert_double_to_float_su3_soa(u, uf)
        af, bf, tf = a.to(torch.complex64), lat.new_vec(single=True), lat.new_vec(single=True)
        ms = timeit(lambda: (lat.acc_Doe(uf, bf, af, phf), lat.acc_Deo(uf, tf, bf, phf)), args.reps)
        out["deo_doe_fp32"] = {"ms_per_pair": ms, "gflops": 570.0 * 2 * interior * world / ms / 1e6,
                               "hbm_GBps_per_gpu": 464.0 * 2 * interior / ms / 1e6, "frac_of_measured_peak": 464.0 * 2 * interior / ms / 1e6 / peak}
    # ---- rational approximation: shipped order-19 x^(-1/4), rescaled with the measured lambda_max (update_versatile.c:189-193)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    start = v.clone()
    t0 = time.perf_counter()
    lmax = lat.ker_find_max_eigenvalue_openacc(u, pars, r, h, start)
    out["lambda_max"] = lmax; out["find_max_eig_s"] = time.perf_counter() - t0
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_abi_approx.npz")))
    if args.order in (19, 9):
        tag = "m14" if args.order == 19 else "m14o9"
        mother = osb.RationalApprox.make(float(g[tag + "_a0"]), g[tag + "_a"], g[tag + "_b"], int(g[tag + "_num"]), int(g[tag + "_den"]))
        mother.lambda_min, mother.lambda_max = float(g[tag + "_lmin"]), float(g[tag + "_lmax"])
        try:
            approx = mother.rescaled((args.mass ** 2, lmax))
        except ValueError:
            approx = mother.rescaled((1e9, lmax))       # range check is advisory for a throughput run
        out["approx"] = "shipped x^(-1/4) order %d rescaled to lambda_max*1.05" % args.order
    else:
        n = args.order
        approx = osb.RationalApprox.make(1.0, np.ones(n), np.geomspace(1e-4, 2.0, n))
        out["approx"] = "%d shifts geomspace(1e-4,2)" % n
    n = approx.approx_order
    sol, ps = lat.new_vec(n), lat.new_vec(n)
    src = v.clone()
    if world > 1:
        lat.communicate_fermion_borders(src)
    lat.multishift_invert(u, pars, approx, sol, src, args.residue, r, h, s, p, ps, 30)
    barrier(); t0 = time.perf_counter()
    st, cg = lat.multishift_invert(u, pars, approx, sol, src, args.residue, r, h, s, p, ps, args.max_cg)
    barrier(); wall = maxr(time.perf_counter() - t0)
    it, act, loop_ms = lat.last_solve_stats()
    bytes_ = (2192.0 * it + 192.0 * act) * interior
    out["multishift_fp64"] = {"s_per_solve": wall, "iterations": cg, "status": st, "shifts": n, "residue": args.residue,
                              "ms_per_iteration": loop_ms / max(it, 1), "mean_active_shifts": act / max(it, 1),
                              "hbm_GBps_per_gpu": bytes_ / loop_ms / 1e6, "frac_of_measured_peak": bytes_ / loop_ms / 1e6 / peak}
    rec = lat.new_vec()
    t = timeit(lambda: lat.recombine_shifted_vec3_to_vec3(sol, src, rec, approx), 10)
    out["recombine_ms"] = t
    # ---- fermion force outer products (row N2) on the solutions just computed: ker_openacc_compute_fermion_force
    # (N acc_Doe + outer products into aux_u), multiply_backfield_times_force, ..._take_ta_nophase.
    # Algorithmic bytes per half-lattice index (FP64): Doe 928 per shift; outer products 2304 (aux_u read+write, 8 links
    # x 9 x 16 B x 2) per PAIR of shifts + 192 per shift (s, h at the site and at the +mu neighbours, each vector once)
    # + 96 for the final loc_s copy; the reference's structure moves 96 + 928 + 2304 + 192 per shift.
    aux, pseudo, ta = lat.new_conf(), lat.new_conf(), lat.new_tamat()
    fpars = lat.ferm_param(args.mass, ph, phf)
    fpars.approx_md.approx_order = n
    for i in range(n):
        fpars.approx_md.RA_a[i] = approx.RA_a[i]
    t = timeit(lambda: lat.ker_openacc_compute_fermion_force(u, aux, sol, s, h, fpars), 5)
    fb = (928.0 * n + 2304.0 * ((n + 1) // 2) + 192.0 * n + 96.0) * interior
    out["fermion_force"] = {"ms": t, "shifts": n, "hbm_GBps_per_gpu": fb / t / 1e6, "frac_of_measured_peak": fb / t / 1e6 / peak,
                            "reference_structure_bytes_ratio": (3520.0 * n) / (fb / interior)}
    t = timeit(lambda: lat.multiply_backfield_times_force(fpars, aux, pseudo), 10)
    out["backfield_times_force"] = {"ms": t, "hbm_GBps_per_gpu": 3520.0 * interior / t / 1e6}
    t = timeit(lambda: lat.multiply_conf_times_force_and_take_ta_nophase(u, pseudo, ta), 10)
    out["take_ta"] = {"ms": t, "hbm_GBps_per_gpu": 2944.0 * interior / t / 1e6}
    # ---- isotropic stout smearing (row N4): one level.  Bytes per half-lattice index as the two kernels move them:
    # staples+Q kernel reads the 8 links once (768) and writes staples (1152) + Q (512); exp kernel reads Q (512) + links
    # (768), writes exp_aux rows 0,1 (768) + smeared links rows 0,1 (768) = 5248 B.  Arithmetic: ~2.7 kflop per link for the
    # six staples + ~0.7 kflop for TA(U S), exp and exp*U => ~27 kflop per index: FP64-bound, not HBM-bound.
    lat.set_stout(0.15, 1)
    t = timeit(lambda: lat.stout_isotropic(u, aux, pseudo, rec_conf, ta, 0), 5) if (rec_conf := lat.new_conf()) is not None else 0
    out["stout_isotropic"] = {"ms": t, "hbm_GBps_per_gpu": 5248.0 * interior / t / 1e6, "approx_fp64_tflops": 27.0e3 * interior / t / 1e9}
    del aux, pseudo, ta, rec_conf
    if not args.skip_fp32:
        # FP32 CG-M (multishift_invert_f) to the reference's single-precision target (inverter_wrappers.c:62-64)
        solf, psf = lat.new_vec(n, single=True), lat.new_vec(n, single=True)
        rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5))
        srcf = src.to(torch.complex64)
        resf = max(args.residue, 8e-7 * np.sqrt(lat.sizeh))
        lat.multishift_invert(uf, pars, approx, solf, srcf, resf, rf, hf, sf, pf, psf, 30)
        barrier(); t0 = time.perf_counter()
        st, cgf = lat.multishift_invert(uf, pars, approx, solf, srcf, resf, rf, hf, sf, pf, psf, args.max_cg)
        barrier(); wall = maxr(time.perf_counter() - t0)
        it, act, loop_ms = lat.last_solve_stats()
        out["multishift_fp32"] = {"s_per_solve": wall, "iterations": cgf, "status": st, "target_res": resf,
                                  "ms_per_iteration": loop_ms / max(it, 1), "mean_active_shifts": act / max(it, 1),
                                  "hbm_GBps_per_gpu": (1096.0 * it + 96.0 * act) * interior / loop_ms / 1e6}
        # single-shift solves: FP64 CG vs FP32-inner mixed precision on the smallest shift
        ip = osb.InverterPackage()
        st_d, st_f = lat.new_vec(1), lat.new_vec(1, single=True)
        lat.setup_inverter_package_dp(ip, u, st_d, 1, r, h, s, p)
        lat.setup_inverter_package_sp(ip, uf, st_f, 1, rf, hf, sf, pf, of)
        shift0 = float(approx.RA_b[0])
        for name, mixed in (("cg_fp64", 0), ("cg_mixed", 1)):
            lat.set_inverter_tricks(0, mixed, 0.1, 10000)
            x = lat.new_vec()
            barrier(); t0 = time.perf_counter()
            its = lat.inverter_wrapper(ip, pars, x, src, args.residue, args.max_cg, shift0, osb.CONVERGENCE_NONCRITICAL)
            barrier(); wall = maxr(time.perf_counter() - t0)
            out[name] = {"s_per_solve": wall, "iterations": its, "shift": shift0, "ms_per_iteration": wall * 1e3 / max(its, 1)}
        lat.set_inverter_tricks(0, 0, 0.1, 10000)
    if rank == 0:
        os.dup2(real_stdout, 1)
        line = json.dumps(out)
        print(line, flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "bench_configs.jsonl"), "a").write(line + "\n")
    if world > 1:
        lat.shutdown_multidev(); dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the staggered fermion-solver hot path (BASELINE.json):
Deo/Doe GFLOP/s (570 flop/site) and HBM GB/s against the roofline, plus M^+M and multishift CG.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lattice 32x32x32x32]

One "step" = one acc_Doe followed by one acc_Deo on the SAME gauge field (the stock deo_doe_test
loop, src/tests_and_benchmarks/deo_doe_test.c:236-269), on synthetic Haar-random SU(3) links and a
Gaussian source.  N=1 workload: BASELINE configs[1], 32^4 FP64.  N>1: weak scaling -- every GPU owns
a 32^3 x 32 slab of a 32^3 x (32 N) lattice (the reference's D3 "salamino"), with the halo exchange of
acc_Deo/acc_Doe inside the step.  `value` is timed with inputs resident in HBM; `e2e` is the same step
driven through the C ABI with HOST buffers (pinned host memory made present, staple_acc_update_device
of the source before and staple_acc_update_host of the result after, both inside the timed region; the
gauge field stays resident, as it does across operator calls in the reference).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SITE = 570.0            # SURVEY 8d / BASELINE.json convention
BYTES_PER_SITE_FP64 = 928.0      # 8 links x 96 B + 8 phases x 8 B + 48 B spinor in + 48 B out
EB = (0.0,) * 6                  # throughput runs: zero background field (theta in {0, pi})


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", default="32x32x32x32", help="per-GPU local lattice LOC_N0xLOC_N1xLOC_N2xLOC_N3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solver", action="store_true")
    ap.add_argument("--shifts", type=int, default=15)
    ap.add_argument("--stream-mode", type=int, default=0, help="0: copy-engine downloads, 1: Deo chunk kernels store to the pinned host buffer")
    ap.add_argument("--stream-chunk", type=int, default=0, help="d3 slices per chunk of the pipelined host round trip (0 = library default)")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []; self.proc = None; self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def haar_su3_torch(torch, n, device, seed):
    """n Haar-random SU(3) matrices on the GPU -> complex128[3,3,n] (row, col, site).  Gram-Schmidt of two
    complex Gaussian 3-vectors gives Haar-distributed rows 0,1; row 2 = conj(r0 x r1) makes det = 1
    (pure elementwise torch: batched torch.linalg.qr launches thousands of tiny kernels)."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    z = torch.complex(torch.randn((2, 3, n), generator=g, device=device, dtype=torch.float64),
                      torch.randn((2, 3, n), generator=g, device=device, dtype=torch.float64))
    r0 = z[0] / torch.linalg.vector_norm(z[0], dim=0, keepdim=True)
    r1 = z[1] - (r0.conj() * z[1]).sum(0, keepdim=True) * r0
    r1 = r1 / torch.linalg.vector_norm(r1, dim=0, keepdim=True)
    r2 = torch.stack((r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0])).conj()
    return torch.stack((r0, r1, r2))


def make_fields(torch, lat, seed):
    """synthetic inputs in the reference layouts: u[8,3,3,sizeh], source[3,sizeh]."""
    S = lat.sizeh
    u = lat.new_conf()
    for k in range(8):
        u[k] = haar_su3_torch(torch, S, lat.device, seed * 1000 + k * 17)
    g = torch.Generator(device=lat.device); g.manual_seed(seed + 99)
    v = torch.complex(torch.randn((3, S), generator=g, device=lat.device, dtype=torch.float64),
                      torch.randn((3, S), generator=g, device=lat.device, dtype=torch.float64)) / np.sqrt(2.0)
    return u, v


def staggered_phases(lat, rank):
    """calc_u1_phases with zero EM field and zero chemical potential (backfield.c:20-187): staggered
    eta_mu and the antiperiodic time boundary only, theta in {0, pi}.  Host-side input producer."""
    nd0, nd1, nd2, nd3 = lat.nd
    gl_t = lat.loc_n[3] * lat.nranks
    d0, d1, d2, d3 = np.meshgrid(np.arange(nd0), np.arange(nd1), np.arange(nd2), np.arange(nd3), indexing="ij")
    t = (d3 + rank * lat.loc_n[3] - lat.d3_halo) % gl_t
    x, y, z = d0, d1, d2
    idxh = (d0 + nd0 * (d1 + nd1 * (d2 + nd2 * d3))) // 2
    par = (x + y + z + t) % 2
    ph = np.zeros((8, lat.sizeh))
    twopi = 2 * 3.14159265358979323846
    args = [np.zeros_like(x, dtype=float), 0.5 * (x & 1), 0.5 * ((x + y) & 1), 0.5 * ((x + y + z) & 1) + 0.5 * (t == gl_t - 1)]
    for mu in range(4):
        a = args[mu].astype(float)
        a = np.where(a > 0.5, a - 1.0, a)
        ph[2 * mu + par, idxh] = a * twopi
    return ph


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation (gcc build of the unmodified sources,
    oracle/_ref; single-threaded because OpenACC pragmas are ignored by gcc) on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.pyoracle import RefLib, Restatement, gaussian_vec, random_su3_conf, ref_lib_path
    loc = tuple(int(x) for x in args.lattice.split("x"))
    kind = "reference"
    try:
        R = RefLib(*loc)
        run = lambda u, a, b, ph: (R.dslash("acc_Doe", u, a, ph, out=b), R.dslash("acc_Deo", u, b, ph, out=a))
        ph = R.phases()
    except Exception:
        kind = "port"
        R = Restatement(*loc)
        run = lambda u, a, b, ph: (R.dslash("doe", u, a, ph, out=b), R.dslash("deo", u, b, ph, out=a))
        ph = R.phases(0)
    u = random_su3_conf(R.sizeh, 1); a = gaussian_vec(R.sizeh, 2); b = np.zeros_like(a)
    # Bounded sample: the reference's build is single-threaded (OpenACC pragmas ignored, no OpenMP anywhere in the tree), one
    # Doe+Deo pair on 32^4 takes 0.4-1 s, so at most 12 timed pairs (2 warm-up) of ONE rank's slab are run whatever K, W and N
    # are; its throughput is size-independent (bandwidth-bound stencil), so the figure stands for the whole workload.
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    steps, warm = max(1, min(args.steps, 12)), max(1, min(args.warmup, 2))
    for _ in range(warm):
        run(u, a, b, ph)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(u, a, b, ph)
    dt = (time.perf_counter() - t0) / steps
    gflops = 2 * FLOP_PER_SITE * R.sizeh / dt / 1e9
    line = {"impl": "reference", "metric": "deo_doe_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3 * world, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.lattice, loc, world, "none (single host process)"),
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                             "sample": "%d Doe+Deo pairs (of the %d steps asked for) on one %s slab, timed on 1 host thread -- the gcc "
                                       "build of the reference is single-threaded; ms_per_step = that time x %d slab(s)"
                                       % (steps, args.steps, args.lattice, world)},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(lattice, loc, world, halo):
    """the `config` object of both arms: the workload BASELINE.json's metric is quoted on (configs[1]) and how it is laid out"""
    return {"workload": "deo_doe %s per GPU, FP64, one acc_Doe + one acc_Deo per step, D3 slabs over %d GPU(s)" % (lattice, world),
            "global_lattice": "%dx%dx%dx%d" % (loc[0], loc[1], loc[2], loc[3] * world), "halo": halo,
            "l2_policy": "inputs larger than L2 (links 384 MiB read per application at 32^4 vs 126 MB L2)",
            "flop_per_site": FLOP_PER_SITE, "bytes_per_site": BYTES_PER_SITE_FP64}


def cpu_baseline(args, loc, u_host, v_host, ph_host):
    """rank 0, N=1: the reference (oracle/_ref) or, failing that, the oracle port on a bounded sample."""
    from oracle.pyoracle import RefLib, Restatement
    sizeh = v_host.shape[1]
    try:
        R = RefLib(*loc); kind = "reference"
        f = lambda a, b: (R.dslash("acc_Doe", u_host, a, ph_host, out=b), R.dslash("acc_Deo", u_host, b, ph_host, out=a))
    except Exception:
        R = Restatement(*loc); kind = "port"
        f = lambda a, b: (R.dslash("doe", u_host, a, ph_host, out=b), R.dslash("deo", u_host, b, ph_host, out=a))
    a = v_host.copy(); b = np.zeros_like(a)
    f(a, b)
    n = 0; t0 = time.perf_counter()
    while True:
        f(a, b); n += 1
        if time.perf_counter() - t0 > 10.0 or n >= 64:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": 2 * FLOP_PER_SITE * sizeh / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
            "sample": "%d Doe+Deo pairs on the full %s lattice (%.2f s each), 1 host thread; gcc -O3 build of the "
                      "reference is single-threaded (OpenACC pragmas ignored)" % (n, "x".join(map(str, loc)), dt)}


def main():
    args = parse()
    # the driver reads ONE JSON line from stdout: everything else that C or Python code prints while the
    # benchmark runs (the library keeps the reference's printf messages) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        return run_reference(args, emit)
    import torch
    import torch.distributed as dist
    import openstaple_b200 as osb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    loc = tuple(int(x) for x in args.lattice.split("x"))
    torch.cuda.set_device(local_rank)
    # everything (torch fills/copies, library kernels, timing events) on ONE non-default stream
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", local_rank)))
    lat = osb.Lattice(loc, nranks_d3=world, halo_width=2, device=local_rank)
    if world > 1:
        lat.init_multidev(dist, async_comm_fermion=1, p2p=int(os.environ.get("STAPLE_P2P", "1")))
    dev = lat.device
    u, v = make_fields(torch, lat, seed=1 + rank)
    ph_host = staggered_phases(lat, rank)
    ph = lat.to_device(ph_host)
    if world > 1:
        lat.communicate_su3_borders(u, 2)
        lat.communicate_fermion_borders(v)
    a, b = v.clone(), lat.new_vec()
    interior = lat.vol3h * loc[3]
    pars = lat.ferm_param(0.0507, ph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def step():
        lat.acc_Doe(u, b, a, ph)
        lat.acc_Deo(u, a, b, ph)

    def renorm():
        # keep the ping-pong vector O(1) without touching the timed region
        nrm = lat.l2norm2_global(a)
        lat.multiply_fermion_x_doublefactor(a, 1.0 / np.sqrt(nrm / (3 * interior * world)))
        if world > 1:
            lat.communicate_fermion_borders(a)

    # ---------------- device-resident timing (value)
    for _ in range(max(3, args.warmup)):
        step()
    renorm()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = lat.kernel_launches()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = lat.kernel_launches() - l0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    sites_per_step = 2 * interior * world                      # Doe + Deo outputs, all ranks
    gflops = FLOP_PER_SITE * sites_per_step / (ms_step * 1e-3) / 1e9

    # ---------------- dominant kernel alone (roofline): Deo launches back to back on this stream
    renorm()
    nk = max(20, args.steps // 2)
    barrier()
    e0.record()
    for _ in range(nk):
        lat.acc_Deo_unsafe(u, b, a, ph)
    e1.record()
    barrier()
    ms_kernel = e0.elapsed_time(e1) / nk
    peak, peak_src = peaks()
    achieved = BYTES_PER_SITE_FP64 * interior / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "dslash_kernel<double,0,EPI_NONE> (acc_Deo_unsafe)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "bytes_per_launch": BYTES_PER_SITE_FP64 * interior, "us_per_launch": ms_kernel * 1e3,
                "traffic": None}
    tf = os.path.join(ROOT, "profiles", "dslash_traffic.json")
    if os.path.exists(tf):
        try:
            t = json.load(open(tf))
            if t.get("lattice") == args.lattice:
                roofline["traffic"] = t.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---------------- M^+M (fused mass epilogue)
    tmp, out = lat.new_vec(), lat.new_vec()
    for _ in range(3):
        lat.fermion_matrix_multiplication(u, out, a, tmp, pars)
    barrier(); e0.record()
    for _ in range(nk):
        lat.fermion_matrix_multiplication(u, out, a, tmp, pars)
    e1.record(); barrier()
    ms_mdagm = max_over_ranks(e0.elapsed_time(e1) / nk)

    # ---------------- end to end through the C ABI with host buffers
    h_in = lat.host_array((3, lat.sizeh), np.complex128)
    h_out = lat.host_array((3, lat.sizeh), np.complex128)
    h_in.np[...] = a.cpu().numpy()
    d_tmp = lat.new_vec()
    vec_bytes = 48 * lat.sizeh

    def e2e_plain():
        h_in.update_device()
        lat.acc_Doe(u, d_tmp, h_in, ph)
        lat.acc_Deo(u, h_out, d_tmp, ph)
        h_out.update_host()

    def e2e_streamed():
        # the same round trip as ONE C-ABI call, software-pipelined over d3 chunks (single rank; with D3 slabs
        # the library runs the plain sequence incl. the halo exchanges)
        lat.acc_Doe_Deo_streamed(u, h_out, h_in, d_tmp, ph, args.stream_chunk)

    lat.L.staple_set_streamed_mode(args.stream_mode)

    def time_e2e(fn):
        for _ in range(3):
            fn()
        ne = max(10, args.steps // 4)
        barrier(); t0 = time.perf_counter(); e0.record()
        for _ in range(ne):
            fn()
        e1.record(); barrier()
        return max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / ne)

    ms_plain = time_e2e(e2e_plain)
    plain_result = h_out.np.copy()
    ms_e2e = time_e2e(e2e_streamed)
    if not np.array_equal(plain_result, h_out.np):
        raise SystemExit("bench: pipelined host round trip differs from the plain sequence")
    e2e = {"value": FLOP_PER_SITE * sites_per_step / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
           "h2d_bytes_per_step": vec_bytes, "d2h_bytes_per_step": vec_bytes, "ms_per_step": ms_e2e,
           "unpipelined_ms_per_step": ms_plain,
           "unpipelined_value": FLOP_PER_SITE * sites_per_step / (ms_plain * 1e-3) / 1e9,
           "note": "source uploaded from and result downloaded to pinned host memory every step through ONE C-ABI call "
                   "(staple_acc_Doe_Deo_streamed: copies and Doe/Deo d3-chunk launches software-pipelined on three "
                   "streams, result checked bit-identical to the unpipelined update-device/acc_Doe/acc_Deo/update-host "
                   "sequence); gauge field resident"}

    # ---------------- multishift CG (secondary metric: s/solve)
    solver = None
    if not args.no_solver:
        n = args.shifts
        shifts = np.geomspace(1e-4, 2.0, n)
        approx = osb.RationalApprox.make(1.0, np.ones(n), shifts)
        sol, ps = lat.new_vec(n), lat.new_vec(n)
        r, h, s, p = (lat.new_vec() for _ in range(4))
        src = v.clone()
        if world > 1:
            lat.communicate_fermion_borders(src)
        lat.multishift_invert(u, pars, approx, sol, src, 1e-8, r, h, s, p, ps, 40)       # warm-up
        barrier(); t0 = time.perf_counter()
        st, cg = lat.multishift_invert(u, pars, approx, sol, src, 1e-8, r, h, s, p, ps, 20000)
        barrier(); wall = time.perf_counter() - t0
        it, act, loop_ms = lat.last_solve_stats()
        fused_bytes = (2192.0 * it + 192.0 * act) * interior            # M^+M 1904 + r-update 144 + p-update 144 + 192 per active shift
        solver = {"s_per_solve": wall, "iterations": cg, "status": st, "shifts": n, "residue": 1e-8,
                  "ms_per_iteration": loop_ms / max(it, 1), "active_shift_iterations": act,
                  "algorithmic_GBps": fused_bytes / (loop_ms * 1e-3) / 1e9 if loop_ms > 0 else None,
                  "roofline_frac": fused_bytes / (loop_ms * 1e-3) / 1e9 / peak if loop_ms > 0 else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, loc, u.cpu().numpy(), v.cpu().numpy(), ph_host)

    if rank == 0:
        line = {"metric": "deo_doe_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args.lattice, loc, world, ("nvlink peer stores fused in the surface kernels" if getattr(lat, "p2p", False)
                                                                      else "nccl send/recv") if world > 1 else "none"),
                "hbm_GBps": BYTES_PER_SITE_FP64 * sites_per_step / world / (ms_step * 1e-3) / 1e9,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "mdagm": {"ms": ms_mdagm, "algorithmic_GBps": 1904.0 * interior / (ms_mdagm * 1e-3) / 1e9,
                          "gflops": 1140.0 * interior * world / (ms_mdagm * 1e-3) / 1e9},
                "multishift": solver}
        emit(line)
    h_in.free(); h_out.free()
    if world > 1:
        lat.shutdown_multidev()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the staggered fermion-solver hot path (BASELINE.json):
Deo/Doe GFLOP/s (570 flop/site) and HBM GB/s against the roofline, multishift CG s/solve, at 1/2/4/8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sections headline,config1,config5,config3]

HEADLINE (every N): STRONG scaling of one acc_Doe + one acc_Deo ("step", the stock deo_doe_test loop,
src/tests_and_benchmarks/deo_doe_test.c:236-269) on the 64^3 x 128 lattice of BASELINE configs[3], FP64, sharded in D3 slabs
(the reference's "salamino") over the N GPUs with the halo exchange of acc_Deo/acc_Doe inside the step.  All N work on the SAME
global gauge field and source (generated per global d3 slice from counter-based seeds, so every rank can rebuild any slice),
which is what makes the in-run parity check possible: every rank compares three windows of Doe and Deo.Doe of ITS slab
(lower boundary incl. the received halo, middle, upper boundary incl. the received halo) with the CPU oracle
(oracle/, the restatement pinned to the reference's gcc build) run on the same global slices, and the CG-M iteration counts with
the committed single-GPU counts.  The line carries `parity`; a violated tolerance makes the run exit non-zero.

`value` is timed with inputs resident in HBM; `e2e` is the same step through ONE C-ABI call with HOST buffers
(staple_acc_Doe_Deo_streamed: source up from, result down to pinned host memory inside the timed region; gauge field resident,
as it is across operator calls in the reference).  Secondary objects in the same line: `config1` (32^4 on one GPU, N=1 only),
`config5` (64^3 x 16 strong scaling: Deo/Doe and CG-M), `config3` (48^3 x 96 CG-M with the shipped order-19 approximation:
FP64, FP32, and the FP32-accelerated wrapper with FP64 refinement), each with its own roofline figures.
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
HEADLINE = (64, 64, 64, 128)     # BASELINE configs[3]
CONFIG1 = (32, 32, 32, 32)       # configs[1]
CONFIG3 = (48, 48, 48, 96)       # configs[2]
CONFIG5 = (64, 64, 64, 16)       # configs[4]
MASS = 0.0507                    # strange quark of tools/test/fermion_parameters.set
RESIDUE = 1e-8                   # residue_metro
SEED = 20261017
CPU_SAMPLE = (64, 64, 64, 4)     # one 64^3 x 4 slab (1/32 of the headline lattice): the reference arm's bounded sample
ITER_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_cgm_iterations.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", default="x".join(map(str, HEADLINE)), help="GLOBAL lattice of the headline workload")
    ap.add_argument("--sections", default="headline,config1,config5,config3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="development only: skip the in-run oracle comparison")
    ap.add_argument("--write-fixture", action="store_true", help="N=1: record the CG-M iteration counts as the committed expectation")
    ap.add_argument("--stream-chunk", type=int, default=0, help="d3 slices per chunk of the pipelined host round trip (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time bound of each cpu_baseline sample")
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


# ------------------------------------------------------------------------------------------------ synthetic inputs
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


def slice_fields(torch, device, vol3h, g3, seed=SEED):
    """gauge links [8,3,3,vol3h] and source [3,vol3h] of GLOBAL d3 slice g3: a pure function of (seed, g3, vol3h), the same
    on every rank and for every number of ranks (Philox counter-based generator)."""
    u = haar_su3_torch(torch, 8 * vol3h, device, seed * 4099 + 2 * g3).reshape(3, 3, 8, vol3h).permute(2, 0, 1, 3)
    g = torch.Generator(device=device); g.manual_seed(seed * 4099 + 2 * g3 + 1)
    v = torch.complex(torch.randn((3, vol3h), generator=g, device=device, dtype=torch.float64),
                      torch.randn((3, vol3h), generator=g, device=device, dtype=torch.float64)) / np.sqrt(2.0)
    return u, v


def global_slice(lat, rank, d3):
    """global d3 slice behind local slice d3 of this rank's local+halo box (Mpi/geometry_multidev.h:300-328)"""
    return (d3 - lat.d3_halo + rank * lat.loc_n[3]) % (lat.loc_n[3] * lat.nranks)


def make_fields(torch, lat, rank, halos=True):
    """this rank's local+halo box in the reference layouts: u[8,3,3,sizeh], source[3,sizeh].  halos=False leaves the link halo
    slices zero (communicate_su3_borders must then restore them)."""
    u, v = lat.new_conf(), lat.new_vec()
    V = lat.vol3h
    for d3 in range(lat.nd[3]):
        halo = lat.nranks > 1 and not (lat.d3_halo <= d3 < lat.d3_halo + lat.loc_n[3])
        us, vs = slice_fields(torch, lat.device, V, global_slice(lat, rank, d3))
        if halos or not halo:
            u[..., d3 * V:(d3 + 1) * V] = us
        if not halo or lat.d3_halo - 1 <= d3 <= lat.d3_halo + lat.loc_n[3]:
            v[:, d3 * V:(d3 + 1) * V] = vs
    return u, v


def staggered_phase_slices(xp, nd0, nd1, nd2, g3_of_slice, gl_t, **kw):
    """calc_u1_phases with zero EM field and zero chemical potential (backfield.c:20-187): staggered eta_mu and the antiperiodic
    time boundary only, theta in {0, pi}, for the given list of GLOBAL d3 coordinates -> [8, len*vol3h].  xp = numpy or torch."""
    ns = len(g3_of_slice)
    d0, d1, d2, sl = xp.meshgrid(xp.arange(nd0, **kw), xp.arange(nd1, **kw), xp.arange(nd2, **kw), xp.arange(ns, **kw), indexing="ij")
    t = (xp.asarray(g3_of_slice, **kw) if xp is np else xp.as_tensor(g3_of_slice, **kw))[sl]
    idxh = (d0 + nd0 * (d1 + nd1 * (d2 + nd2 * sl))) // 2
    par = (d0 + d1 + d2 + t) % 2
    f64 = dict(dtype=xp.float64)
    ph = xp.zeros((8, nd0 * nd1 * nd2 * ns // 2), **f64, **({k: v for k, v in kw.items() if k == "device"}))
    twopi = 2 * 3.14159265358979323846
    args = [0 * d0, d0 & 1, (d0 + d1) & 1, ((d0 + d1 + d2) & 1) + (t == gl_t - 1)]
    for mu in range(4):
        a = args[mu] * 0.5 if xp is np else args[mu].to(xp.float64) * 0.5
        a = xp.where(a > 0.5, a - 1.0, a)
        ph[2 * mu + par.reshape(-1), idxh.reshape(-1)] = (a * twopi).reshape(-1)
    return ph


def staggered_phases(lat, rank, torch=None):
    g3 = [global_slice(lat, rank, d3) for d3 in range(lat.nd[3])]
    gl_t = lat.loc_n[3] * lat.nranks
    if torch is None:
        return staggered_phase_slices(np, lat.nd[0], lat.nd[1], lat.nd[2], g3, gl_t)
    return staggered_phase_slices(torch, lat.nd[0], lat.nd[1], lat.nd[2], g3, gl_t, device=lat.device, dtype=torch.int64)


# ------------------------------------------------------------------------------------------------ CPU legs (oracle / reference)
def cpu_operator_throughput(seconds, max_steps, warm, threads=None):
    """The reference's own Doe+Deo (oracle/_ref: its unmodified sources built by gcc; single-threaded by construction, OpenACC
    pragmas ignored, no OpenMP in the tree) on a bounded sample of the headline workload: `threads` host threads each apply it
    to their own source on one 64^3 x 4 slab (shared gauge field), the way an MPI run of the reference would use the cores.
    -> dict(value GFLOP/s, cores, kind, steps, seconds, single_thread_value, sample)."""
    from oracle.pyoracle import RefLib, Restatement, gaussian_vec, random_su3_conf
    loc = CPU_SAMPLE
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
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    T = max(1, min(threads or ncpu, ncpu, 64))
    u = random_su3_conf(R.sizeh, 1)
    vecs = [(gaussian_vec(R.sizeh, 2 + i), np.zeros((3, R.sizeh), np.complex128)) for i in range(T)]
    # single-thread figure first (one pair, after one warm-up pair)
    run(u, vecs[0][0], vecs[0][1], ph)
    t0 = time.perf_counter(); run(u, vecs[0][0], vecs[0][1], ph); single = time.perf_counter() - t0
    done = [0] * T
    stop = threading.Event()

    def worker(i, n):
        a, b = vecs[i]
        for _ in range(n):
            if stop.is_set():
                break
            run(u, a, b, ph); done[i] += 1

    def batch(n):
        th = [threading.Thread(target=worker, args=(i, n)) for i in range(T)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for i in range(T):
        done[i] = 0
    if warm > 0:
        batch(min(warm, 1) if seconds < 30 else warm)
    steps = 0; total = 0.0
    while steps < max_steps and total < seconds:
        for i in range(T):
            done[i] = 0
        total += batch(1); steps += 1
    sites = 2 * R.sizeh * T * steps                       # Doe + Deo output half-sites processed
    value = FLOP_PER_SITE * sites / total / 1e9
    return {"value": value, "unit": "GFLOP/s", "cores": T, "kind": kind, "steps": steps, "seconds": total,
            "single_thread_value": FLOP_PER_SITE * 2 * R.sizeh / single / 1e9,
            "sample": "%d step(s), each = %d host threads applying Doe+Deo to their own source on one %s slab (1/%d of the %s lattice, "
                      "shared gauge field); the reference's gcc build is single-threaded, threads stand in for its MPI ranks"
                      % (steps, T, "x".join(map(str, loc)), HEADLINE[3] // loc[3], "x".join(map(str, HEADLINE)))}


def cpu_cgm_per_site_iteration(approx_b, seconds):
    """seconds per half-site and CG-M iteration of the reference (oracle/_ref, 1 thread): a few iterations with the SAME shifts on
    a 32^4 lattice, max_cg bounded (inverter_multishift_full.c:108-191 timed by the wall clock)."""
    from oracle.pyoracle import RefLib, Restatement, gaussian_vec, random_su3_conf
    loc = CONFIG1
    n = len(approx_b)
    try:
        R = RefLib(*loc); kind = "reference"; ph = R.phases()
        solve = lambda u, v, it: R.multishift_invert(u, ph, MASS, (1.0, np.ones(n), np.asarray(approx_b)), v, 1e-30, it)
    except Exception:
        R = Restatement(*loc); kind = "port"; ph = R.phases(0)
        solve = lambda u, v, it: R.multishift_invert(u, ph, MASS, np.asarray(approx_b), v, 1e-30, it)
    u = random_su3_conf(R.sizeh, 1); v = gaussian_vec(R.sizeh, 2)
    t0 = time.perf_counter(); solve(u, v, 1); t1 = time.perf_counter() - t0        # set-up + post-loop checks + 1 iteration
    its = max(2, min(12, int(seconds / max(t1 / (n + 2), 0.2))))
    t0 = time.perf_counter(); solve(u, v, 1 + its); t2 = time.perf_counter() - t0
    per_it = (t2 - t1) / its
    return per_it / R.sizeh, kind, "%d CG-M iterations with the same %d shifts on a 32^4 lattice, 1 host thread (%.2f s per iteration); " \
                                   "scaled by half-sites x iterations of the GPU solve" % (its, n, per_it)


def workload_config(lattice, world):
    """the `config` object of BOTH arms (byte-identical for the same --gpus)"""
    gl = tuple(int(x) for x in lattice.split("x"))
    return {"workload": "deo_doe %s FP64 (BASELINE configs[3]): one acc_Doe + one acc_Deo per step on the global lattice, D3 slabs over %d GPU(s)"
                        % (lattice, world),
            "global_lattice": lattice, "local_lattice": "%dx%dx%dx%d" % (gl[0], gl[1], gl[2], gl[3] // world),
            "l2_policy": "inputs larger than L2 (links: 768 B x %d half-sites per GPU and application vs 126 MB L2)" % (gl[0] * gl[1] * gl[2] * gl[3] // 2 // world),
            "flop_per_site": FLOP_PER_SITE, "bytes_per_site": BYTES_PER_SITE_FP64}


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation on bounded samples of the same workload.  Honours --steps and
    --warmup up to a time cap of 90 s of timed work (each step is ~0.5-1 s of wall clock on all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    r = cpu_operator_throughput(90.0, max(1, args.steps), max(1, args.warmup))
    gl = tuple(int(x) for x in args.lattice.split("x"))
    sites_per_step = gl[0] * gl[1] * gl[2] * gl[3]            # Doe + Deo output half-sites of one step on the global lattice
    ms_per_step = FLOP_PER_SITE * sites_per_step / (r["value"] * 1e9) * 1e3
    line = {"impl": "reference", "metric": "deo_doe_gflops", "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": max(1, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.lattice, world),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_value")},
            "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "steps_requested": args.steps,
            "note": "time cap 90 s: fewer than --steps steps are timed when they do not fit; ms_per_step = one step of the global lattice at "
                    "the measured throughput"}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
class Job:
    """one global lattice sharded over the ranks: library geometry, fields, phases, helpers"""

    def __init__(self, torch, dist, osb, gl, world, rank, local_rank, check_links=True):
        self.torch, self.dist, self.osb = torch, dist, osb
        self.gl, self.world, self.rank = gl, world, rank
        assert gl[3] % world == 0 and (gl[3] // world) % 2 == 0 or world == 1, "global d3 extent must split into even slabs"
        self.loc = (gl[0], gl[1], gl[2], gl[3] // world)
        self.lat = lat = osb.Lattice(self.loc, nranks_d3=world, halo_width=2, device=local_rank)
        if world > 1:
            lat.init_multidev(dist, async_comm_fermion=1, p2p=int(os.environ.get("STAPLE_P2P", "1")))
        self.u, self.v = make_fields(torch, lat, rank, halos=True)
        self.link_halo_exact = None
        if world > 1 and check_links:
            # links: exchange the halos of a copy whose halo slices were zeroed and compare with the generated ones (rows r0, r1 travel)
            w = self.u.clone()
            h = lat.d3_halo * lat.vol3h
            w[..., :h] = 0; w[..., lat.sizeh - h:] = 0
            lat.communicate_su3_borders(w, 2)
            self.link_halo_exact = bool(torch.equal(w[:, :2], self.u[:, :2]))
            del w
        self.ph = staggered_phases(lat, rank, torch)
        self.interior = lat.vol3h * self.loc[3]
        self.dev = lat.device

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, x):
        if self.world > 1:
            t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def timeit(self, fn, reps, warm=3):
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        for _ in range(warm):
            fn()
        self.barrier(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); self.barrier()
        return self.maxr(e0.elapsed_time(e1) / reps)

    def close(self):
        if self.world > 1:
            self.lat.shutdown_multidev()


def parity_windows(job, doe_out, deo_out, single=False):
    """Every rank: Doe(v) and Deo(Doe(v)) of its slab against the CPU oracle on three 8-slice windows of the GLOBAL lattice --
    around its lower boundary (covers the halo slice received from rank L), its middle, and its upper boundary (halo from R).
    The oracle lattice is the window itself (periodic wrap never reaches the compared slices: Doe is valid on window slices
    1..6, Deo.Doe on 2..5).  -> max relative error over everything compared."""
    from oracle.pyoracle import Restatement
    lat, torch = job.lat, job.torch
    V = lat.vol3h; gl3 = job.gl[3]; loc3 = job.loc[3]; r = job.rank
    NW = 8
    S = Restatement(job.gl[0], job.gl[1], job.gl[2], NW)
    own_lo = r * loc3
    starts = sorted({(own_lo - 4) % gl3, (own_lo + (loc3 // 2 // 2) * 2 - 4) % gl3, (own_lo + loc3 - 4) % gl3})
    worst = 0.0; compared = 0
    dh, do = doe_out, deo_out
    for w0 in starts:
        g3s = [(w0 + k) % gl3 for k in range(NW)]
        uw = torch.zeros((8, 3, 3, NW * V), dtype=torch.complex128, device=job.dev)
        vw = torch.zeros((3, NW * V), dtype=torch.complex128, device=job.dev)
        for k, g3 in enumerate(g3s):
            us, vs = slice_fields(torch, job.dev, V, g3)
            uw[..., k * V:(k + 1) * V] = us; vw[:, k * V:(k + 1) * V] = vs
        phw = staggered_phase_slices(np, job.gl[0], job.gl[1], job.gl[2], g3s, gl3)
        uh, vh = uw.cpu().numpy(), vw.cpu().numpy()
        del uw, vw
        if single:          # FP32 twin: inputs rounded to float, the oracle's float restatement (phases: theta in {0, pi} rounds exactly as the float code computes it)
            uh, vh, phw = uh.astype(np.complex64), vh.astype(np.complex64), phw.astype(np.float32)
        want_doe = S.dslash("doe", uh, vh, phw, 0, NW)
        want_deo = S.dslash("deo", uh, want_doe, phw, 0, NW)
        for k, g3 in enumerate(g3s):
            # local slice that holds global slice g3 on this rank: interior, or one of the two fermion halo slices
            rel = (g3 - own_lo) % gl3
            if rel < loc3:
                d3 = lat.d3_halo + rel
            elif job.world > 1 and rel == gl3 - 1:
                d3 = lat.d3_halo - 1
            elif job.world > 1 and rel == loc3:
                d3 = lat.d3_halo + loc3
            else:
                continue
            for lo_ok, hi_ok, got, want in ((1, NW - 2, dh, want_doe), (2, NW - 3, do, want_deo)):
                if lo_ok <= k <= hi_ok:
                    g = got[:, d3 * V:(d3 + 1) * V].cpu().numpy(); w = want[:, k * V:(k + 1) * V]
                    worst = max(worst, float(np.abs(g - w).max() / np.abs(w).max())); compared += 1
    return worst, compared, len(starts)


def shipped_order19(osb, lmax, mass):
    """the shipped order-19 x^(-1/4) approximation (saved_approxs/approx_-1_over_4_order_19_mloglm_6.4.REMEZ, committed as
    tests/golden/ref_abi_approx.npz by tests/golden/make_golden.py), rescaled with the measured lambda_max as
    update_versatile.c:189-193 does"""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_abi_approx.npz")))
    mother = osb.RationalApprox.make(float(g["m14_a0"]), g["m14_a"], g["m14_b"], int(g["m14_num"]), int(g["m14_den"]))
    mother.lambda_min, mother.lambda_max = float(g["m14_lmin"]), float(g["m14_lmax"])
    try:
        return mother.rescaled((mass ** 2, lmax))
    except ValueError:
        return mother.rescaled((1e9, lmax))       # the range check is advisory for a throughput run


def cgm_bytes(it, act, interior, single=False):
    """SURVEY 8d fused figure with the ACTUAL number of active shifts: per iteration M^+M 1904 + r update 144 + p update 144,
    + 192 per active shift (FP32: half)"""
    return (2192.0 * it + 192.0 * act) * interior * (0.5 if single else 1.0)


def solver_section(job, peak, want_fp32, want_accel, cpu_seconds, tag, fixture, parity):
    """CG-M with the shipped order-19 approximation on job's lattice: FP64, FP32, FP32-accelerated wrapper"""
    torch, lat, osb = job.torch, job.lat, job.osb
    u, v, ph = job.u, job.v, job.ph
    phf = ph.to(torch.float32)
    pars = lat.ferm_param(MASS, ph, phf)
    r, h, s, p = (lat.new_vec() for _ in range(4))
    start = v.clone()
    lmax = lat.ker_find_max_eigenvalue_openacc(u, pars, r, h, start)
    approx = shipped_order19(osb, lmax, MASS)
    n = approx.approx_order
    out = {"lattice": "x".join(map(str, job.gl)), "approx": "shipped x^(-1/4) order 19, rescaled to 1.05 x lambda_max", "lambda_max": lmax,
           "mass": MASS, "residue": RESIDUE, "shifts": n}
    sol, ps = lat.new_vec(n), lat.new_vec(n)
    src = v                                                       # halos valid by construction
    lat.multishift_invert(u, pars, approx, sol, src, RESIDUE, r, h, s, p, ps, 24)            # warm-up (graph capture, clocks)
    job.barrier(); t0 = time.perf_counter()
    st, cg = lat.multishift_invert(u, pars, approx, sol, src, RESIDUE, r, h, s, p, ps, 20000)
    job.barrier(); wall = job.maxr(time.perf_counter() - t0)
    it, act, loop_ms = lat.last_solve_stats()
    loop_ms = job.maxr(loop_ms)
    gbs = cgm_bytes(it, act, job.interior) / loop_ms / 1e6
    # true residual of the smallest and the largest shift, computed with the library's own operator
    res = []
    for i in (0, n - 1):
        lat.fermion_matrix_multiplication_shifted(u, s, sol[i], h, pars, float(approx.RA_b[i]))
        lat.combine_in1_minus_in2(src, s, h)
        res.append(float(np.sqrt(lat.l2norm2_global(h) / lat.l2norm2_global(src))))
    out["fp64"] = {"s_per_solve": wall, "iterations": cg, "status": st, "ms_per_iteration": loop_ms / max(it, 1),
                   "mean_active_shifts": act / max(it, 1), "hbm_GBps_per_gpu": gbs, "roofline_frac": gbs / peak,
                   "true_rel_residual_first_last_shift": res}
    key = "%s:fp64" % tag
    parity["cg_iters"][key] = cg
    if st != 1 or max(res) > 2 * RESIDUE:
        parity["failures"].append("%s: CG-M did not reach the residual (status %d, true residuals %r)" % (key, st, res))
    uf = None
    if want_fp32:
        uf = lat.new_conf(single=True); lat.convert_double_to_float_su3_soa(u, uf)
        solf, psf = lat.new_vec(n, single=True), lat.new_vec(n, single=True)
        rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5))
        srcf = src.to(torch.complex64)
        # inverter_wrappers.c:62-64 uses 8e-7 sqrt(sizeh) with the LOCAL sizeh (halos included), which changes with the number of
        # ranks; here the global half-volume, so that the FP32 solve is the same problem at every N
        resf = max(RESIDUE, 8e-7 * np.sqrt(job.interior * job.world))
        lat.multishift_invert(uf, pars, approx, solf, srcf, resf, rf, hf, sf, pf, psf, 24)
        job.barrier(); t0 = time.perf_counter()
        stf, cgf = lat.multishift_invert(uf, pars, approx, solf, srcf, resf, rf, hf, sf, pf, psf, 20000)
        job.barrier(); wallf = job.maxr(time.perf_counter() - t0)
        it, act, loop_ms = lat.last_solve_stats()
        loop_ms = job.maxr(loop_ms)
        gbs = cgm_bytes(it, act, job.interior, single=True) / loop_ms / 1e6
        out["fp32"] = {"s_per_solve": wallf, "iterations": cgf, "status": stf, "target_res": resf, "ms_per_iteration": loop_ms / max(it, 1),
                       "mean_active_shifts": act / max(it, 1), "hbm_GBps_per_gpu": gbs, "roofline_frac": gbs / peak}
        parity["cg_iters"]["%s:fp32" % tag] = cgf
        if want_accel:
            # singlePInvAccelMultiInv (inverter_wrappers.c:45-115): FP32 CG-M, then every shift refined to RESIDUE by the FP32-inner
            # mixed-precision CG (inverter_mixedp.c) starting from the FP32 solution
            ip = osb.InverterPackage()
            lat.setup_inverter_package_dp(ip, u, ps, n, r, h, s, p)
            lat.setup_inverter_package_sp(ip, uf, psf, n, rf, hf, sf, pf, of)
            lat.set_sp_globals(lat.new_vec(single=True), solf)
            lat.set_inverter_tricks(1, 1, 0.1, 10000)
            job.barrier(); t0 = time.perf_counter()
            tot = lat.inverter_multishift_wrapper(ip, pars, approx, sol, src, RESIDUE, 20000, osb.CONVERGENCE_NONCRITICAL)
            job.barrier(); walla = job.maxr(time.perf_counter() - t0)
            lat.set_inverter_tricks(0, 0, 0.1, 10000)
            res = []
            for i in (0, n - 1):
                lat.fermion_matrix_multiplication_shifted(u, s, sol[i], h, pars, float(approx.RA_b[i]))
                lat.combine_in1_minus_in2(src, s, h)
                res.append(float(np.sqrt(lat.l2norm2_global(h) / lat.l2norm2_global(src))))
            out["fp32_accelerated_fp64_refined"] = {"s_per_solve": walla, "wrapper_return": tot, "fp32_multishift_iterations": int(lat.last_solve_stats()[0]),
                                                    "refinement_iterations": lat.last_refinement_iterations(), "true_rel_residual_first_last_shift": res,
                                                    "note": "FP32 CG-M + per-shift FP32-inner mixed-precision CG to the FP64 residue (singlePInvAccelMultiInv + useMixedPrecision)"}
            parity.setdefault("cg_iters_not_compared", {})["%s:accel" % tag] = lat.last_refinement_iterations()     # its FP32 target follows the LOCAL sizeh (reference behaviour)
            if max(res) > 2 * RESIDUE:
                parity["failures"].append("%s:accel true residuals %r" % (tag, res))
    if job.rank == 0 and job.world == 1 and cpu_seconds > 0:
        per, kind, sample = cpu_cgm_per_site_iteration([approx.RA_b[i] for i in range(n)], cpu_seconds)
        tot_sites = job.interior * job.world
        out["cpu_baseline"] = {"value": per * tot_sites * cg, "unit": "s/solve", "cores": 1, "kind": kind, "sample": sample}
    return out


def operator_section(job, peak, steps, warm):
    """device-resident Doe+Deo and M^+M timings on job's lattice"""
    lat = job.lat
    u, ph = job.u, job.ph
    a, b, c = job.v, lat.new_vec(), lat.new_vec()
    ms = job.timeit(lambda: (lat.acc_Doe(u, b, a, ph), lat.acc_Deo(u, c, b, ph)), steps, warm)
    pars = lat.ferm_param(MASS, ph)
    tmp, out = lat.new_vec(), lat.new_vec()
    ms_mm = job.timeit(lambda: lat.fermion_matrix_multiplication(u, out, job.v, tmp, pars), max(10, steps // 2), 3)
    sites = 2 * job.interior * job.world
    return {"ms_per_step": ms, "gflops": FLOP_PER_SITE * sites / ms / 1e6, "hbm_GBps_per_gpu": BYTES_PER_SITE_FP64 * 2 * job.interior / ms / 1e6,
            "roofline_frac": BYTES_PER_SITE_FP64 * 2 * job.interior / ms / 1e6 / peak,
            "mdagm_ms": ms_mm, "mdagm_GBps_per_gpu": 1904.0 * job.interior / ms_mm / 1e6}


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
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference(args, emit)
    import torch
    import torch.distributed as dist
    import openstaple_b200 as osb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # everything (torch fills/copies, library kernels, timing events) on ONE non-default stream
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", local_rank)))
    sections = set(args.sections.split(","))
    peak, peak_src = peaks()
    gl = tuple(int(x) for x in args.lattice.split("x"))
    fixture = json.load(open(ITER_FIXTURE)) if os.path.exists(ITER_FIXTURE) else {}
    parity = {"tolerance": 1e-13, "max_rel_err": 0.0, "slices_compared": 0, "windows_per_rank": 0, "link_halo_exact": None,
              "cg_iters": {}, "cg_iters_expected": {}, "cg_iters_tolerance": 0.02, "failures": []}
    line = None

    def check_windows(job, tag):
        if args.no_parity:
            return
        b, o = job.lat.new_vec(), job.lat.new_vec()
        job.lat.acc_Doe(job.u, b, job.v, job.ph)
        job.lat.acc_Deo(job.u, o, b, job.ph)
        err, ncmp, nwin = parity_windows(job, b, o)
        err = job.maxr(err)
        parity["max_rel_err"] = max(parity["max_rel_err"], err)
        parity["slices_compared"] += ncmp; parity["windows_per_rank"] = nwin
        parity.setdefault("per_lattice", {})[tag] = err
        if job.link_halo_exact is not None:
            ok = job.maxr(0.0 if job.link_halo_exact else 1.0) == 0.0
            parity["link_halo_exact"] = ok if parity["link_halo_exact"] is None else (parity["link_halo_exact"] and ok)
            if not ok:
                parity["failures"].append("%s: communicate_su3_borders did not reproduce the generated link halos" % tag)
        if not err <= parity["tolerance"]:
            parity["failures"].append("%s: Doe / Deo.Doe differ from the oracle by %.3e" % (tag, err))

    # ======================================================================== headline: strong scaling on 64^3 x 128
    if "headline" in sections:
        job = Job(torch, dist, osb, gl, world, rank, local_rank)
        lat = job.lat
        u, v, ph = job.u, job.v, job.ph
        check_windows(job, args.lattice)
        a, b, c = v.clone(), lat.new_vec(), lat.new_vec()
        interior = job.interior

        def step():
            # the stock test re-applies the operators to the SAME input (deo_doe_test.c:236-269); feeding the result back would
            # grow it by up to lambda_max ~ 7 per step and overflow for large --steps
            lat.acc_Doe(u, b, a, ph)
            lat.acc_Deo(u, c, b, ph)

        warm = max(3, args.warmup)
        for _ in range(warm):
            step()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        job.barrier()
        l0 = lat.kernel_launches()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        job.barrier()
        launches = lat.kernel_launches() - l0
        ms_step = job.maxr(e0.elapsed_time(e1) / args.steps)
        clocks = sampler.stop() if rank == 0 else None
        sites_per_step = 2 * interior * world                      # Doe + Deo outputs, all ranks
        gflops = FLOP_PER_SITE * sites_per_step / (ms_step * 1e-3) / 1e9

        # ---- dominant kernel alone (roofline): Deo launches back to back on this stream (no exchange: acc_Deo_unsafe)
        del c
        nk = max(10, min(200, args.steps // 2))
        ms_kernel = job.timeit(lambda: lat.acc_Deo_unsafe(u, b, a, ph), nk, 2)
        achieved = BYTES_PER_SITE_FP64 * interior / (ms_kernel * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "dslash_kernel<double,0,EPI_NONE> (acc_Deo_unsafe on the local slab)", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "bytes_per_launch": BYTES_PER_SITE_FP64 * interior, "us_per_launch": ms_kernel * 1e3, "traffic": None,
                    "frac_of_nominal_8TBs": achieved / 8000.0,
                    "note": "peak = measured COPY bandwidth (half reads, half writes); this kernel is a 95 % read stream and exceeds it on "
                            "the faster boxes of the pool (frac up to 1.05); ncu: DRAM bytes = 0.9993 x the algorithmic bytes (profiles/dslash_traffic.json)"}
        tf = os.path.join(ROOT, "profiles", "dslash_traffic.json")
        if os.path.exists(tf):
            try:
                t = json.load(open(tf))
                if t.get("local_lattice") == "x".join(map(str, job.loc)):
                    roofline["traffic"] = t.get("dram_bytes_per_launch")
            except Exception:
                pass
        pars = lat.ferm_param(MASS, ph)
        tmp, out = lat.new_vec(), lat.new_vec()
        ms_mdagm = job.timeit(lambda: lat.fermion_matrix_multiplication(u, out, a, tmp, pars), nk, 3)
        del out

        # ---- end to end through the C ABI with host buffers
        h_in = lat.host_array((3, lat.sizeh), np.complex128)
        h_out = lat.host_array((3, lat.sizeh), np.complex128)
        h_in.np[...] = a.cpu().numpy()
        vec_bytes = 48 * lat.sizeh

        def e2e_plain():
            h_in.update_device()
            lat.acc_Doe(u, tmp, h_in, ph)
            lat.acc_Deo(u, h_out, tmp, ph)
            h_out.update_host()

        def e2e_streamed():
            lat.acc_Doe_Deo_streamed(u, h_out, h_in, tmp, ph, args.stream_chunk)

        def time_e2e(fn):
            for _ in range(3):
                fn()
            ne = max(5, min(20, args.steps // 5))
            job.barrier(); t0 = time.perf_counter(); e0.record()
            for _ in range(ne):
                fn()
            e1.record(); job.barrier()
            return job.maxr(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / ne)

        ms_plain = time_e2e(e2e_plain)
        r1lo, r1hi = lat.ranges[2], lat.ranges[3]
        plain_result = h_out.np[:, r1lo:r1hi].copy()
        ms_e2e = time_e2e(e2e_streamed)
        same = bool(np.array_equal(plain_result, h_out.np[:, r1lo:r1hi]))
        if job.maxr(0.0 if same else 1.0) != 0.0:
            parity["failures"].append("pipelined host round trip differs from the plain update-device/acc_Doe/acc_Deo/update-host sequence")
        del plain_result
        e2e = {"value": FLOP_PER_SITE * sites_per_step / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": vec_bytes, "d2h_bytes_per_step": vec_bytes, "ms_per_step": ms_e2e,
               "host_GBps_per_gpu_each_way": vec_bytes / ms_e2e / 1e6, "host_GBps_all_gpus_both_ways": 2 * vec_bytes * world / ms_e2e / 1e6,
               "unpipelined_ms_per_step": ms_plain,
               "unpipelined_value": FLOP_PER_SITE * sites_per_step / (ms_plain * 1e-3) / 1e9,
               "bit_identical_to_unpipelined": same,
               "note": "per rank and step: the local+halo box of the source goes up from and the result comes down to pinned host memory "
                       "through ONE C-ABI call (staple_acc_Doe_Deo_streamed: copies and Doe/Deo d3-chunk launches software-pipelined on three "
                       "streams, face slices first on D3 slabs); gauge field resident; bytes are per GPU"}
        h_in.free(); h_out.free()
        del tmp

        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            r = cpu_operator_throughput(args.cpu_seconds, 10 ** 6, 1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_value")}
        halo = "none"
        if world > 1:
            halo = "nvlink peer stores from the face blocks, the data words are their own arrival flags, staged halos inside the solvers" if getattr(lat, "p2p", False) else "nccl send/recv"
        line = {"metric": "deo_doe_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args.lattice, world), "halo": halo,
                "hbm_GBps_per_gpu": BYTES_PER_SITE_FP64 * 2 * interior / (ms_step * 1e-3) / 1e9,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "mdagm": {"ms": ms_mdagm, "algorithmic_GBps_per_gpu": 1904.0 * interior / (ms_mdagm * 1e-3) / 1e9,
                          "gflops": 1140.0 * interior * world / (ms_mdagm * 1e-3) / 1e9}}
        del a, b, u, v, ph
        # CG-M on the headline lattice (s/solve at 1/2/4/8 GPUs, the second half of BASELINE's metric)
        line["multishift"] = solver_section(job, peak, False, False, 0, args.lattice, fixture, parity)
        job.close(); del job
        torch.cuda.empty_cache()

    secondary = {}
    # ======================================================================== config 1: 32^4 on ONE GPU
    if "config1" in sections and world == 1:
        job = Job(torch, dist, osb, CONFIG1, 1, 0, local_rank)
        check_windows(job, "32x32x32x32")
        o = operator_section(job, peak, max(50, args.steps), 5)
        a, b = job.v.clone(), job.lat.new_vec()
        ms_kernel = job.timeit(lambda: job.lat.acc_Deo_unsafe(job.u, b, a, job.ph), 100, 5)
        o["deo_kernel"] = {"us_per_launch": ms_kernel * 1e3, "achieved_GBps": BYTES_PER_SITE_FP64 * job.interior / ms_kernel / 1e6,
                           "roofline_frac": BYTES_PER_SITE_FP64 * job.interior / ms_kernel / 1e6 / peak}
        # single-system solvers (the per-shift refinements of the accelerated wrapper, eo_inversion): restarted CG and FP32-inner
        # mixed-precision CG, iteration loop on the device vs read back by the host every iteration (staple_set_cg_device_loops)
        lat = job.lat
        phf = job.ph.to(torch.float32)
        pars = lat.ferm_param(MASS, job.ph, phf)
        uf = lat.new_conf(single=True); lat.convert_double_to_float_su3_soa(job.u, uf)
        ip = osb.InverterPackage()
        r, h, s, p = (lat.new_vec() for _ in range(4)); rf, hf, sf, pf, of = (lat.new_vec(single=True) for _ in range(5))
        lat.setup_inverter_package_dp(ip, job.u, lat.new_vec(1), 1, r, h, s, p)
        lat.setup_inverter_package_sp(ip, uf, lat.new_vec(1, single=True), 1, rf, hf, sf, pf, of)
        solv = {}
        for name, mixed in (("cg_fp64", 0), ("cg_mixed", 1)):
            for dev_loops in (1, 0):
                lat.L.staple_set_cg_device_loops(dev_loops)
                lat.set_inverter_tricks(0, mixed, 0.1, 10000)
                x = lat.new_vec()
                lat.inverter_wrapper(ip, pars, x, job.v, 1e-2, 20000, 1e-4, osb.CONVERGENCE_NONCRITICAL)      # warm-up
                x.zero_()
                job.barrier(); t0 = time.perf_counter(); l0 = lat.kernel_launches()
                its = lat.inverter_wrapper(ip, pars, x, job.v, RESIDUE, 20000, 1e-4, osb.CONVERGENCE_NONCRITICAL)
                job.barrier(); wall = time.perf_counter() - t0
                solv["%s_%s" % (name, "device_loop" if dev_loops else "host_loop")] = {
                    "s_per_solve": wall, "iterations": its, "ms_per_iteration": wall * 1e3 / max(its, 1), "launches": lat.kernel_launches() - l0}
        lat.L.staple_set_cg_device_loops(1); lat.set_inverter_tricks(0, 0, 0.1, 10000)
        o["single_system_solvers"] = solv
        secondary["config1_32x32x32x32"] = o
        job.close(); del job, a, b, uf, r, h, s, p, rf, hf, sf, pf, of
        torch.cuda.empty_cache()
    # ======================================================================== config 5: 64^3 x 16 strong scaling
    if "config5" in sections and CONFIG5[3] % (2 * world) == 0:
        job = Job(torch, dist, osb, CONFIG5, world, rank, local_rank)
        check_windows(job, "64x64x64x16")
        o = operator_section(job, peak, max(50, args.steps), 5)
        o["multishift"] = solver_section(job, peak, False, False, 0, "64x64x64x16", fixture, parity)
        secondary["config5_64x64x64x16"] = o
        job.close(); del job
        torch.cuda.empty_cache()
    # ======================================================================== config 3: 48^3 x 96 CG-M, FP64 / FP32 / accelerated
    if "config3" in sections and CONFIG3[3] % (2 * world) == 0:
        job = Job(torch, dist, osb, CONFIG3, world, rank, local_rank)
        check_windows(job, "48x48x48x96")
        secondary["config3_48x48x48x96"] = solver_section(job, peak, True, True, 0 if args.no_cpu_baseline else args.cpu_seconds,
                                                          "48x48x48x96", fixture, parity)
        job.close(); del job
        torch.cuda.empty_cache()

    # ---- CG-M iteration counts against the committed single-GPU counts
    for k, it in parity["cg_iters"].items():
        exp = fixture.get(k)
        parity["cg_iters_expected"][k] = exp
        if exp is not None and abs(it - exp) > parity["cg_iters_tolerance"] * exp:
            parity["failures"].append("%s: %d CG-M iterations against %d on one GPU" % (k, it, exp))
    parity["cg_iters_equal_to_single_gpu"] = all(parity["cg_iters_expected"][k] == it for k, it in parity["cg_iters"].items()) \
        if all(v is not None for v in parity["cg_iters_expected"].values()) else None
    parity["ok"] = not parity["failures"]
    if args.write_fixture and rank == 0 and world == 1:
        fixture.update(parity["cg_iters"])
        json.dump(fixture, open(ITER_FIXTURE, "w"), indent=1, sort_keys=True)
    if rank == 0:
        if line is None:
            line = {"metric": "deo_doe_gflops", "n_gpus": world, "note": "headline section skipped (--sections)"}
        line["secondary"] = secondary
        line["parity"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not parity["ok"]:
        sys.stderr.write("bench: PARITY FAILURE: %r\n" % (parity["failures"],))
        sys.exit(3)


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_host/*: the input files for the reference's OWN test programs (deo_doe_test,
inverter_multishift_test; built unmodified by oracle/build_ref_host.sh) and the results the pure-reference CPU build of
those programs writes.  tests/test_gpu_reference_host.py runs the SAME programs linked against libstaple_b200.so on the
B200 with the same input files and compares the files they write.  Run in the dev container only:

    python tests/golden/make_ref_host.py
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("STAPLE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "ref_host")
GEOM = (8, 8, 8, 8)


def make_input(n, deo_doe_iterations=3, ms_repetitions=1, benchmark=1, max_cg=40, save=1, eps_gen="3.0"):
    """build/input.example with: this geometry on one rank, identity direction map, the flavours and background field of
    tools/test (charges, chemical potential, E and B fields: general U(1) phases), no replicas section, results saved"""
    t = open(os.path.join(REF, "build", "input.example")).read()
    t = t[:t.index("#---- Hasenbusch Parallel Tempering params")]
    flav = open(os.path.join(REF, "tools", "test", "fermion_parameters.set")).read().strip() + "\n\n"
    t = re.sub(r"FlavourParameters\n.*?(?=BackgroundFieldParameters)", "", t, flags=re.S)
    t = t.replace("BackgroundFieldParameters", flav + "BackgroundFieldParameters", 1)

    def put(key, val):
        nonlocal t
        t, k = re.subn(r"^(%s\s+)\S+" % re.escape(key), lambda m: m.group(1) + str(val), t, count=1, flags=re.M)
        assert k == 1, key
    for key, val in (("ex", 5), ("ey", -5), ("ez", 1), ("bx", -5), ("by", 5), ("bz", 3), ("nx", n[0]), ("ny", n[1]), ("nz", n[2]),
                     ("nt", n[3]), ("xmap", 0), ("ymap", 1), ("zmap", 2), ("tmap", 3), ("NRanks", 1), ("NProcPerNode", 1),
                     ("residue_md", "1.0e-4"), ("residue_metro", "1.0e-8"), ("ExpMaxEigenvalue", "5.5"), ("EpsGen", eps_gen),
                     ("UseILDG", 0), ("VerbosityLv", 1), ("SaveDiagnostics", 0), ("DeoDoeIterations", deo_doe_iterations),
                     ("MultiShiftInverterRepetitions", ms_repetitions), ("BenchmarkMode", benchmark), ("SaveResults", save),
                     ("MaxCGIterations", max_cg), ("useMixedPrecision", 0), ("FakeShift", "1.0e-2")):
        put(key, val)
    # keep only what the parser reads: section names and `key value` pairs (the example's commentary stays in the reference)
    lines = [l.split("#")[0].rstrip() for l in t.splitlines()]
    return "\n".join(re.sub(r"\s+", " ", l) for l in lines if l.strip()) + "\n"


def read_vec3_ascii(path, single=False):
    """print_vec3_soa_wrapper's global ASCII format (io.c:553-614): one `re<TAB>im` line per colour and even site"""
    a = np.loadtxt(path, dtype=np.float64)
    return (a[:, 0] + 1j * a[:, 1]).reshape(-1, 3)


def run(prog, input_text, extra_files=()):
    exe = os.path.join(ROOT, "oracle", "_ref", "%s_ref_%dx%dx%dx%d" % ((prog,) + GEOM))
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "in.set"), "w").write(input_text)
        for name, text in extra_files:
            open(os.path.join(td, name), "w").write(text)
        r = subprocess.run([exe, "in.set"], cwd=td, capture_output=True, text=True, timeout=3600)
        files = {f: os.path.join(td, f) for f in os.listdir(td)}
        return r, {f: open(p, "rb").read() for f, p in files.items() if os.path.getsize(p) < 64 << 20}


def remez_text(r):
    """the .REMEZ layout the reference's reader parses (rationalapprox.c:96-106), from the numbers its own reader returned"""
    t = "\nApproximation to f(x) = (x)^(%d/%d)\n" % (r["num"], r["den"])
    t += "Order: %d\nLambda Min: %.16e\nLambda Max: %.16e\n" % (r["order"], r["lmin"], r["lmax"])
    t += "GMP Remez Precision: %d\nError: %.16e\nRA_a0 = %.16e\n" % (r["prec"], r["error"], r["a0"])
    for i, (x, y) in enumerate(zip(r["a"], r["b"])):
        t += "RA_a[%d] = %.16e, RA_b[%d] = %.16e\n" % (i, x, i, y)
    return t


def read_ratapproxes():
    """tools/test/ratapproxes/*.REMEZ through the reference's own reader -> {file name: numbers}"""
    import ctypes as C
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from make_golden import _RA
    from oracle.pyoracle import RefLib
    R = RefLib(4, 4, 4, 4)
    out = {}
    d = os.path.join(REF, "tools", "test", "ratapproxes")
    for f in sorted(os.listdir(d)):
        r = _RA()
        R.lib.ref_approx_read(C.byref(r), os.path.join(d, f).encode())
        out[f] = dict(num=r.exponent_num, den=r.exponent_den, order=r.approx_order, lmin=r.lambda_min, lmax=r.lambda_max,
                      prec=r.gmp_remez_precision, error=r.error, a0=r.RA_a0, a=list(r.RA_a[:r.approx_order]), b=list(r.RA_b[:r.approx_order]))
    return out


def parse_rhmc(stdout, gauge_obs_text):
    """what a trajectory of the reference's main leaves behind: the gauge_obs rows (iteration, accepted, plaquette, rectangle,
    Polyakov loop), the Metropolis energy differences and the CG-M iteration counts per trajectory"""
    rows = [[float(x) for x in l.split()] for l in gauge_obs_text.splitlines() if l.strip() and not l.startswith("#")]
    return {"gauge_obs": rows,
            "delta_action": [float(x) for x in re.findall(r"DELTA_ACTION = (-?[0-9.eE+-]+?)\. ", stdout)],
            "cgm_md": [int(x) for x in re.findall(r"CG-M iterations\[MD\]: (\d+)", stdout)],
            "cgm_fi": [int(x) for x in re.findall(r"CG-M iterations\[FI\]: (\d+)", stdout)],
            "cgm_li": [int(x) for x in re.findall(r"CG-M iterations\[LI\]: (\d+)", stdout)]}


def rhmc(n=(4, 4, 4, 4)):
    """The reference's production program (OpenAcc/main.c): three RHMC trajectories (one thermalisation + two with the
    Metropolis test) of 2+1 stout-smeared flavours from a near-cold start; zero background field so that the spectrum stays
    inside the shipped approximations on this tiny lattice."""
    import json
    subprocess.run([os.path.join(ROOT, "oracle", "build_ref_host.sh")] + [str(x) for x in n], check=True)
    t = make_input(n, benchmark=0, eps_gen="0.1")
    for k in ("ex", "ey", "ez", "bx", "by", "bz"):
        t = re.sub(r"^(%s )\S+" % k, r"\g<1>0", t, count=1, flags=re.M)
    t = re.sub(r"^(MuOverPiT )\S+", r"\g<1>0", t, flags=re.M)
    for k, v in (("Ntraj", 3), ("ThermNtraj", 1), ("NmdSteps", 8), ("GaugeSubSteps", 3), ("TrajLength", "0.5"), ("VerbosityLv", 1),
                 ("SaveConfInterval", 100), ("StoreConfInterval", 100), ("MeasEvery", 1000000), ("SaveAllAtEnd", 0), ("MeasCool", 0),
                 ("MeasStout", 0), ("CoolMeasSteps", 0), ("StoutMeasSteps", 0)):
        t, c = re.subn(r"^(%s )\S+" % k, lambda m: m.group(1) + str(v), t, count=1, flags=re.M)
        assert c == 1, k
    name = "rhmc_%dx%dx%dx%d" % n
    open(os.path.join(OUT, name + ".set"), "w").write(t)
    exe = os.path.join(ROOT, "oracle", "_ref", "main_ref_%dx%dx%dx%d" % n)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "in.set"), "w").write(t)
        for fname, r in read_ratapproxes().items():
            open(os.path.join(td, fname), "w").write(remez_text(r))
        r = subprocess.run([exe, "in.set"], cwd=td, capture_output=True, text=True, timeout=3600)
        assert r.returncode == 0, r.stdout[-3000:]
        obs = [f for f in os.listdir(td) if f.startswith("gauge_obs")]
        res = parse_rhmc(r.stdout, open(os.path.join(td, obs[0])).read())
    json.dump(res, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
    print("rhmc", res)


def rhmc_two_ranks(loc=(4, 4, 4, 4)):
    """the same RHMC input on two D3 slabs (global 4x4x4x8): the pure-reference NRANKS_D3 = 2 build of main as two processes
    under oracle/mpi_mini"""
    import json
    sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_mini"))
    from mpirun import launch
    subprocess.run([os.path.join(ROOT, "oracle", "build_ref_host.sh")] + [str(x) for x in loc] + ["2"], check=True, env=dict(os.environ, PROGS="main"))
    t = open(os.path.join(OUT, "rhmc_%dx%dx%dx%d.set" % loc)).read()
    for k, v in (("nt", 2 * loc[3]), ("NRanks", 2), ("NProcPerNode", 2)):
        t, c = re.subn(r"^(%s )\S+" % k, lambda m: m.group(1) + str(v), t, count=1, flags=re.M)
        assert c == 1, k
    name = "rhmc_%dx%dx%dx%d_r2" % loc
    open(os.path.join(OUT, name + ".set"), "w").write(t)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "in.set"), "w").write(t)
        for fname, r in read_ratapproxes().items():
            open(os.path.join(td, fname), "w").write(remez_text(r))
        rc = launch(2, [os.path.join(ROOT, "oracle", "_ref", "main_ref_%dx%dx%dx%d_r2" % loc), "in.set"], cwd=td)
        assert rc == 0, open(os.path.join(td, "stdout.0")).read()[-3000:]
        obs = [f for f in os.listdir(td) if f.startswith("gauge_obs")]
        res = parse_rhmc(open(os.path.join(td, "stdout.0")).read(), open(os.path.join(td, obs[0])).read())
    json.dump(res, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
    print("rhmc two ranks", res)


if __name__ == "__main__":
    if sys.argv[1:] == ["rhmc"]:
        rhmc(); rhmc_two_ranks(); sys.exit(0)
    subprocess.run([os.path.join(ROOT, "oracle", "build_ref_host.sh")] + [str(x) for x in GEOM], check=True)
    os.makedirs(OUT, exist_ok=True)
    text = make_input(GEOM)
    open(os.path.join(OUT, "deo_doe_%dx%dx%dx%d.set" % GEOM), "w").write(text)
    r, files = run("deo_doe_test", text)
    assert r.returncode == 0 and "Test completed" in r.stdout, r.stdout[-2000:]
    d = {}
    with tempfile.TemporaryDirectory() as td:
        for f in ("test_fermion", "test_fermion_result_doe2", "test_fermion_result_deo2", "test_fermion_result_fulldirac2",
                  "sp_test_fermion_result_doe2", "sp_test_fermion_result_deo2", "sp_test_fermion_result_fulldirac2"):
            open(os.path.join(td, f), "wb").write(files[f])
            d[f] = read_vec3_ascii(os.path.join(td, f))
    # inverter_multishift_test in benchmark mode: 15 equal shifts, MaxCGIterations iterations (its residue is 2e-144)
    approx = read_ratapproxes()
    import json
    json.dump(approx, open(os.path.join(OUT, "ratapproxes.json"), "w"), indent=0, sort_keys=True)
    extra = [(name, remez_text(r)) for name, r in approx.items()]
    r, files = run("inverter_multishift_test", text, extra)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    print(r.stdout[-1200:])
    shifts = sorted(f for f in files if re.match(r"fermion_shift_\d+\.dat$", f))
    with tempfile.TemporaryDirectory() as td:
        for f in shifts[:3]:
            open(os.path.join(td, f), "wb").write(files[f])
            d["ms_" + f.replace(".dat", "")] = read_vec3_ascii(os.path.join(td, f))
    d["ms_nshift_files"] = len(shifts)
    # the same program with BenchmarkMode 0: first flavour's approx_md, rescaled with the largest eigenvalue that
    # find_min_max_eigenvalue_soloopenacc measures (inverter_multishift_test.c:248-267), again MaxCGIterations iterations
    # (a smoother configuration: with EpsGen 3 the largest eigenvalue, 6.97, is outside what the shipped approximations cover)
    text0 = make_input(GEOM, benchmark=0, eps_gen="0.1")
    open(os.path.join(OUT, "inverter_mode0_%dx%dx%dx%d.set" % GEOM), "w").write(text0)
    r, files = run("inverter_multishift_test", text0, extra)
    assert r.returncode == 0 and "NOT ENTERING BENCHMARK MODE" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    m = re.search(r"Found eigenvalues of dirac operator: (\S+),\s+(\S+)", r.stdout)
    d["ms0_minmax"] = np.array([float(m.group(1)), float(m.group(2))])
    shifts0 = sorted((f for f in files if re.match(r"fermion_shift_\d+\.dat$", f)), key=lambda f: int(re.findall(r"\d+", f)[0]))
    with tempfile.TemporaryDirectory() as td:
        for f in (shifts0[0], shifts0[len(shifts0) // 2], shifts0[-1]):
            open(os.path.join(td, f), "wb").write(files[f])
            d["ms0_" + f.replace(".dat", "")] = read_vec3_ascii(os.path.join(td, f))
    d["ms0_nshift_files"] = len(shifts0)
    np.savez_compressed(os.path.join(OUT, "ref_host_results_%dx%dx%dx%d.npz" % GEOM), **d)
    print(sorted(files)); print({k: getattr(v, "shape", v) for k, v in d.items()})
    rhmc()
    rhmc_two_ranks()

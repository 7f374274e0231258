"""The drop-in claim, executed: the reference's OWN test programs -- src/tests_and_benchmarks/deo_doe_test.c and
inverter_multishift_test.c, unmodified, with the reference's parser, dSFMT generators, U(1) phase code and file writers --
linked against libstaple_b200.so in place of the object files of the subsystems it replaces (oracle/build_ref_host.sh; the
only foreign source is openstaple_b200/host/memory_wrapper_staple.c standing in for Include/memory_wrapper.c).  They run on the B200 with the input
file the pure-reference CPU build was run with in the dev container (tests/golden/make_ref_host.py), and the files they
write are compared with the files that build wrote:

  test_fermion                         the Gaussian source both programs generate (dSFMT, Seed 42): identical text
  test_fermion_result_{doe2,deo2,fulldirac2}   acc_Doe, acc_Deo, fermion_matrix_multiplication   FP64 relative 1e-13
  fermion_shift_N.dat                  multishift_invert, benchmark mode: 15 equal shifts, exactly MaxCGIterations
                                       iterations (its target residue is 2e-144)                  relative 1e-10

The binaries are built where /root/reference exists and travel to the GPU box under oracle/_ref (like the oracle .so
files); without them the tests skip."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, relerr

pytestmark = pytest.mark.gpu
GEOM = "8x8x8x8"
HOST_DIR = os.path.join(GOLDEN_DIR, "ref_host")


def _exe(prog):
    return os.path.join(ROOT, "oracle", "_ref", "%s_staple_%s" % (prog, GEOM))


def _read_vec3_ascii(path):
    a = np.loadtxt(path, dtype=np.float64)
    return (a[:, 0] + 1j * a[:, 1]).reshape(-1, 3)


def _remez_text(r):
    t = "\nApproximation to f(x) = (x)^(%d/%d)\n" % (r["num"], r["den"])
    t += "Order: %d\nLambda Min: %.16e\nLambda Max: %.16e\n" % (r["order"], r["lmin"], r["lmax"])
    t += "GMP Remez Precision: %d\nError: %.16e\nRA_a0 = %.16e\n" % (r["prec"], r["error"], r["a0"])
    for i, (x, y) in enumerate(zip(r["a"], r["b"])):
        t += "RA_a[%d] = %.16e, RA_b[%d] = %.16e\n" % (i, x, i, y)
    return t


def _run(prog, td, input_file=None):
    exe = _exe(prog)
    if not os.path.exists(exe):
        pytest.skip("no %s (built by oracle/build_ref_host.sh where the reference is present)" % os.path.basename(exe))
    open(os.path.join(td, "in.set"), "w").write(open(os.path.join(HOST_DIR, input_file or "deo_doe_%s.set" % GEOM)).read())
    for name, r in json.load(open(os.path.join(HOST_DIR, "ratapproxes.json"))).items():
        open(os.path.join(td, name), "w").write(_remez_text(r))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(x for x in (env.get("LD_LIBRARY_PATH", ""), "/usr/local/cuda/lib64") if x)
    r = subprocess.run([exe, "in.set"], cwd=td, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "hot path served by staple_b200" in r.stderr, r.stderr[-2000:]
    return r


@pytest.fixture(scope="module")
def golden_host():
    return dict(np.load(os.path.join(HOST_DIR, "ref_host_results_%s.npz" % GEOM)))


def test_reference_deo_doe_test_program(tmp_path, golden_host):
    td = str(tmp_path)
    r = _run("deo_doe_test", td)
    assert "Test completed" in r.stdout
    g = golden_host
    assert np.array_equal(_read_vec3_ascii(os.path.join(td, "test_fermion")), g["test_fermion"])      # same generated source
    for f in ("test_fermion_result_doe2", "test_fermion_result_deo2", "test_fermion_result_fulldirac2"):
        e = relerr(_read_vec3_ascii(os.path.join(td, f)), g[f])
        print("%s: %.1e" % (f, e))
        assert e < 1e-13, (f, e)
    # the program's own timing lines exist (blocking mode makes its gettimeofday timers meaningful)
    assert "Time for 1 application of Doe" in r.stdout
    # FP32 twins: the generated writers print 6 decimals (%f), so 2e-5 of the largest component is what the files can show
    for f in ("sp_test_fermion_result_doe2", "sp_test_fermion_result_deo2", "sp_test_fermion_result_fulldirac2"):
        e = relerr(_read_vec3_ascii(os.path.join(td, f)), g[f])
        print("%s: %.1e" % (f, e))
        assert e < 2e-5, (f, e)


def test_reference_inverter_multishift_test_program(tmp_path, golden_host):
    td = str(tmp_path)
    r = _run("inverter_multishift_test", td)
    assert "ENTERING BENCHMARK MODE" in r.stdout
    g = golden_host
    n = int(g["ms_nshift_files"])
    assert all(os.path.exists(os.path.join(td, "fermion_shift_%d.dat" % i)) for i in range(n))
    for k in [k for k in g if k.startswith("ms_fermion_shift_")]:
        e = relerr(_read_vec3_ascii(os.path.join(td, k[3:] + ".dat")), g[k])
        print("%s: %.1e" % (k, e))
        assert e < 1e-10, (k, e)

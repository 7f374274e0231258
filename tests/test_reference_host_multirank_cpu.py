"""The reference's own multi-rank programs on the CPU, without an MPI installation: oracle/mpi_mini (a minimal MPI over local
sockets, exactly the calls the reference makes) runs deo_doe_test built for NRANKS_D3 = 2 as two processes; the single-rank
build of the same program then reads the configuration and source the two ranks saved (the reference's own fixture-sharing
mechanism, deo_doe_test.c:191-219) and must write the same global result files -- D3 slabs, halo exchange
(communications.c:34-104), scatter/gather through rank 0 (communications.c:787-1100) all exercised by the reference's code.
This is also the CPU half of the multi-GPU drop-in test (tests/test_gpu_zz_reference_host_multirank.py)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_reference_host import HOST_DIR, _read_vec3_ascii, _remez_text

sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_mini"))
REF = os.path.join(ROOT, "oracle", "_ref")
FILES = ("test_fermion_result_doe2", "test_fermion_result_deo2", "test_fermion_result_fulldirac2")


def two_rank_input():
    """the 8^4 single-rank input with nt doubled and two ranks"""
    t = open(os.path.join(HOST_DIR, "deo_doe_8x8x8x8.set")).read()
    for k, v in (("nt", 16), ("NRanks", 2), ("NProcPerNode", 2)):
        t, n = re.subn(r"^(%s )\S+" % k, lambda m: m.group(1) + str(v), t, count=1, flags=re.M)
        assert n == 1
    return t


MS_FILES = ("fermion_shift_0.dat", "fermion_shift_7.dat", "fermion_shift_14.dat")


def run_two_ranks(kind, td, env=None, prog="deo_doe_test", files=FILES):
    import json
    from mpirun import launch
    exe = os.path.join(REF, "%s_%s_8x8x8x8_r2" % (prog, kind))
    if not os.path.exists(exe):
        pytest.skip("no " + os.path.basename(exe))
    open(os.path.join(td, "in.set"), "w").write(two_rank_input())
    for name, r in json.load(open(os.path.join(HOST_DIR, "ratapproxes.json"))).items():
        open(os.path.join(td, name), "w").write(_remez_text(r))
    rc = launch(2, [exe, "in.set"], cwd=td, env=env, timeout=900)
    assert rc == 0, (open(os.path.join(td, "stdout.0")).read()[-2000:], open(os.path.join(td, "stderr.0")).read()[-2000:],
                     open(os.path.join(td, "stderr.1")).read()[-2000:])
    return {f: _read_vec3_ascii(os.path.join(td, f)) for f in files}


def run_single_rank_on_saved_inputs(td, prog="deo_doe_test", files=FILES):
    exe = os.path.join(REF, "%s_ref_8x8x8x16" % prog)
    if not os.path.exists(exe):
        pytest.skip("no " + os.path.basename(exe))
    sd = os.path.join(td, "single"); os.makedirs(sd)
    for f in ["save_conf", "test_fermion"] + [f for f in os.listdir(td) if f.endswith(".REMEZ")]:
        os.link(os.path.join(td, f), os.path.join(sd, f))
    t = re.sub(r"^(NRanks|NProcPerNode) \S+", r"\g<1> 1", two_rank_input(), flags=re.M)
    open(os.path.join(sd, "in.set"), "w").write(t)
    r = subprocess.run([exe, "in.set"], cwd=sd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Fermion READ : OK" in r.stdout and "Read : OK" in r.stdout, r.stdout[-2000:]
    return {f: _read_vec3_ascii(os.path.join(sd, f)) for f in files}


def test_mini_mpi_unit(tmp_path):
    """ring Sendrecv of 8 MB per rank (everybody sends first), non-overtaking matching in posting order, reductions in rank order"""
    src = os.path.join(ROOT, "oracle", "mpi_mini", "selftest.c")
    exe = str(tmp_path / "selftest")
    subprocess.run(["gcc", "-O2", "-I" + os.path.dirname(src), src, os.path.join(os.path.dirname(src), "mpi_mini.c"), "-o", exe], check=True)
    from mpirun import launch
    for n in (1, 2, 4):
        assert launch(n, [exe], cwd=str(tmp_path), timeout=120) == 0, open(str(tmp_path / "stdout.0")).read()
        assert all("OK" in open(str(tmp_path / ("stdout.%d" % r))).read() for r in range(n))


def test_reference_two_rank_run_equals_its_single_rank_run(tmp_path):
    td = str(tmp_path)
    two = run_two_ranks("ref", td)
    one = run_single_rank_on_saved_inputs(td)
    for f in FILES:
        e = float(np.abs(two[f] - one[f]).max() / np.abs(one[f]).max())
        assert e < 1e-15, (f, e)


def test_reference_two_rank_multishift_equals_its_single_rank_run(tmp_path):
    """inverter_multishift_test, benchmark mode (15 equal shifts, MaxCGIterations iterations): two ranks with an MPI_Allreduce
    per scalar product against one rank on the same configuration and source.  The sums are accumulated in a different order
    (two partial sums instead of one), so agreement is to rounding amplified by 40 CG iterations, not to the bit."""
    td = str(tmp_path)
    two = run_two_ranks("ref", td, prog="inverter_multishift_test", files=MS_FILES)
    one = run_single_rank_on_saved_inputs(td, prog="inverter_multishift_test", files=MS_FILES)
    for f in MS_FILES:
        e = float(np.abs(two[f] - one[f]).max() / np.abs(one[f]).max())
        assert e < 1e-11, (f, e)


def test_multidev_glue_build_defines_the_rank_setup_symbols():
    """the `_staplemd_` programs (src/Mpi/multidev.c left out, openstaple_b200/host/multidev_staple.c in its place) define devinfo,
    pre_init_multidev1D, init_multidev1D and shutdown_multidev themselves and take the teardown and the readiness query from the
    library; the `_staple_` programs (the reference's multidev.c kept) need neither"""
    md = os.path.join(REF, "deo_doe_test_staplemd_8x8x8x8_r2")
    st = os.path.join(REF, "deo_doe_test_staple_8x8x8x8_r2")
    if not (os.path.exists(md) and os.path.exists(st)):
        pytest.skip("host programs not built")

    def symbols(exe):
        out = subprocess.run(["nm", exe], capture_output=True, text=True, check=True).stdout
        return {l.split()[-1]: l.split()[-2] for l in out.splitlines() if len(l.split()) >= 2}
    s = symbols(md)
    for name in ("pre_init_multidev1D", "init_multidev1D", "shutdown_multidev"):
        assert s.get(name) == "T", (name, s.get(name))
    assert s.get("devinfo") in ("B", "C", "D"), s.get("devinfo")
    assert s.get("staple_shutdown_multidev") == "U" and s.get("staple_init_multidev1D") == "U"
    assert s.get("staple_rank_layer_ready") == "U"
    assert "staple_shutdown_multidev" not in symbols(st)

"""CPU side of tests/test_gpu_reference_host.py: the reference's own test programs built by oracle/build_ref_host.sh.
  * the pure-reference build reproduces the committed results (the fixture is what that program writes);
  * the build linked against libstaple_b200.so starts, reaches the library through the allocation shim and -- with no GPU
    in this container -- stops there with the library's message: no CPU fallback hides behind the drop-in.
Skipped where the binaries are absent (they are built where /root/reference exists)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from test_gpu_reference_host import GEOM, HOST_DIR, _read_vec3_ascii, _remez_text  # noqa: F401  (helpers only; no GPU use at import)

pytestmark = []          # the helpers' module is marked gpu; these tests are not


def _prepare(td):
    import json
    open(os.path.join(td, "in.set"), "w").write(open(os.path.join(HOST_DIR, "deo_doe_%s.set" % GEOM)).read())
    for name, r in json.load(open(os.path.join(HOST_DIR, "ratapproxes.json"))).items():
        open(os.path.join(td, name), "w").write(_remez_text(r))


def _bin(prog, kind):
    p = os.path.join(ROOT, "oracle", "_ref", "%s_%s_%s" % (prog, kind, GEOM))
    if not os.path.exists(p):
        pytest.skip("no " + os.path.basename(p))
    return p


def test_pure_reference_program_reproduces_fixture(tmp_path):
    td = str(tmp_path); _prepare(td)
    r = subprocess.run([_bin("deo_doe_test", "ref"), "in.set"], cwd=td, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Test completed" in r.stdout
    g = dict(np.load(os.path.join(HOST_DIR, "ref_host_results_%s.npz" % GEOM)))
    for f in ("test_fermion", "test_fermion_result_doe2", "test_fermion_result_deo2", "test_fermion_result_fulldirac2"):
        assert np.array_equal(_read_vec3_ascii(os.path.join(td, f)), g[f]), f


def test_library_linked_program_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_reference_host.py runs the program for real")
    td = str(tmp_path); _prepare(td)
    exe = _bin("deo_doe_test", "staple")
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "libstaple_b200.so" in needed
    syms = dict(line.split()[::-1][:2] for line in subprocess.run(["nm", "-D", exe], capture_output=True, text=True).stdout.splitlines()
                if len(line.split()) >= 2)
    for name in ("acc_Deo", "acc_Doe", "acc_Deo_f", "fermion_matrix_multiplication", "inverter_multishift_wrapper",
                 "convert_double_to_float_su3_soa", "staple_init_geometry"):
        assert syms.get(name) == "U", (name, syms.get(name))        # undefined in the program: served by the shared library
    r = subprocess.run([exe, "in.set"], cwd=td, capture_output=True, text=True, timeout=900)
    assert r.returncode != 0 and "libstaple_b200: CUDA error" in r.stderr, r.stderr[-1500:]
    assert not os.path.exists(os.path.join(td, "test_fermion_result_doe2"))


def test_pure_reference_rhmc_main_reproduces_fixture(tmp_path):
    """the CPU half of tests/test_gpu_zz_reference_rhmc.py: the pure-reference build of OpenAcc/main.c reproduces the committed
    trajectory observables exactly (and the comparison code of the GPU test is exercised)"""
    import json
    import test_gpu_zz_reference_rhmc as T
    want = json.load(open(os.path.join(HOST_DIR, "rhmc_%s.json" % T.GEOM)))
    r, got = T.run_main("ref", str(tmp_path))
    assert got == want
    T.compare(got, want)
    assert all(row[1] == 1.0 for row in got["gauge_obs"]) and got["cgm_md"][0] > 100        # accepted trajectories, real solves


def test_pure_reference_two_rank_rhmc_main_reproduces_fixture(tmp_path):
    """... and on two D3 slabs as two processes under oracle/mpi_mini (the CPU half of the 2-GPU test)"""
    import json
    import test_gpu_zz_reference_rhmc as T
    want = json.load(open(os.path.join(HOST_DIR, "rhmc_%s_r2.json" % T.GEOM)))
    r, got = T.run_main("ref", str(tmp_path), ranks=2)
    assert got == want
    T.compare(got, want)

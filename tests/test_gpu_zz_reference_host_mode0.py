"""tests/test_gpu_reference_host.py continued: the reference's inverter_multishift_test program (library-linked) with
BenchmarkMode 0 -- the program measures the largest eigenvalue with find_min_max_eigenvalue_soloopenacc, rescales the first
flavour's approx_md with the reference's own rescale_rational_approximation and runs multishift_invert on those 8 distinct
shifts for MaxCGIterations iterations (inverter_multishift_test.c:248-267).  Compared with the files the pure-reference CPU
build wrote for the same input (tests/golden/make_ref_host.py).

Tolerance 1e-4: the eigenvalue comes out of a power iteration stopped by a 1e-5 relative criterion; if GPU and CPU rounding
ever stop it one iteration apart the rescaled shifts move by that much.  (Green on the B200 since the driver's round-1 run.)"""
import os
import re

import numpy as np
import pytest

from conftest import relerr
from test_gpu_reference_host import GEOM, HOST_DIR, _exe, _read_vec3_ascii, _run  # noqa: F401

pytestmark = pytest.mark.gpu


def test_reference_inverter_program_with_measured_spectrum(tmp_path):
    td = str(tmp_path)
    g = dict(np.load(os.path.join(HOST_DIR, "ref_host_results_%s.npz" % GEOM)))
    r = _run("inverter_multishift_test", td, input_file="inverter_mode0_%s.set" % GEOM)
    assert "NOT ENTERING BENCHMARK MODE" in r.stdout
    m = re.search(r"Found eigenvalues of dirac operator: (\S+),\s+(\S+)", r.stdout)
    lo, hi = float(m.group(1)), float(m.group(2))
    print("eigenvalues %.6e %.6e (reference build %.6e %.6e)" % (lo, hi, g["ms0_minmax"][0], g["ms0_minmax"][1]))
    assert lo == g["ms0_minmax"][0] and abs(hi / g["ms0_minmax"][1] - 1) < 3e-5
    n = int(g["ms0_nshift_files"])
    assert all(os.path.exists(os.path.join(td, "fermion_shift_%d.dat" % i)) for i in range(n))
    for k in [k for k in g if k.startswith("ms0_fermion_shift_")]:
        e = relerr(_read_vec3_ascii(os.path.join(td, k[4:] + ".dat")), g[k])
        print("%s: %.1e" % (k, e))
        assert e < 1e-4, (k, e)

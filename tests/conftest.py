import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_r1():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_4x4x4x4_r1.npz")))


@pytest.fixture(scope="session")
def golden_r2():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_4x4x4x4_r2.npz")))


@pytest.fixture(scope="session")
def golden_force():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_force_4x4x4x4_r1.npz")))


@pytest.fixture(scope="session")
def golden_stout():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_stout_4x4x4x4_r1.npz")))


@pytest.fixture(scope="session")
def golden_stoutforce():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_stoutforce_4x4x4x4_r1.npz")))


def relerr(a, b):
    """max |a-b| / max |b| -- the relative error used for every parity statement."""
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def golden_callers():
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_callers_4x4x4x4_r1.npz")))


def callers_flavours(gc, single=False):
    """the flavour list of tests/golden/make_golden.py:callers_single from the fixture's own arrays"""
    fl = []
    for i in range(2):
        fl.append(dict(mass=float(gc["fl%d_mass" % i]), ph=gc[("phf%d" if single else "ph%d") % i],
                       number_of_ps=int(gc["fl%d_number_of_ps" % i]), first_ps=int(gc["fl%d_first_ps" % i]),
                       ra_a=gc["fl%d_ra_a" % i], ra_b=gc["fl%d_ra_b" % i]))
    return fl

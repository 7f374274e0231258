"""The reference's PRODUCTION program -- OpenAcc/main.c, the whole RHMC: heat bath, molecular dynamics with the stouted
fermion force of 2+1 flavours, Metropolis test, gauge measurements -- built unmodified by oracle/build_ref_host.sh with the
replaced subsystems left out and libstaple_b200.so linked instead.  Its gauge sector (gauge force, link update, plaquettes)
is the reference's own CPU object code working on CUDA managed memory; every fermionic step (smearing, CG-M, force, Sigma'
-> Sigma, heat-bath and action inversions, reductions) runs on the B200 through the C ABI.  Three trajectories on 4^4 are
compared with what the pure-reference CPU build of the same program produced from the same input file
(tests/golden/ref_host/rhmc_4x4x4x4.json, written by tests/golden/make_ref_host.py): gauge_obs rows, Metropolis energy
differences, acceptances, CG-M iteration counts.

Tolerances: the two runs do the same arithmetic up to rounding, but MD solves stop at residue 1e-4 -- one iteration more or
less in a single solve moves the force by ~1e-5 -- so plaquette / rectangle 1e-5 relative, Delta H 1e-3 absolute, iteration
counts 2 %.

Status: the single-GPU test is green on the B200 since the driver's round-1 run (a plain test now); the pure-reference half is
exercised on the CPU (tests/test_reference_host_cpu.py)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT
from test_gpu_reference_host import HOST_DIR, _remez_text

pytestmark = pytest.mark.gpu
GEOM = "4x4x4x4"


def parse_rhmc(stdout, gauge_obs_text):
    import re
    rows = [[float(x) for x in l.split()] for l in gauge_obs_text.splitlines() if l.strip() and not l.startswith("#")]
    return {"gauge_obs": rows,
            "delta_action": [float(x) for x in re.findall(r"DELTA_ACTION = (-?[0-9.eE+-]+?)\. ", stdout)],
            "cgm_md": [int(x) for x in re.findall(r"CG-M iterations\[MD\]: (\d+)", stdout)],
            "cgm_fi": [int(x) for x in re.findall(r"CG-M iterations\[FI\]: (\d+)", stdout)],
            "cgm_li": [int(x) for x in re.findall(r"CG-M iterations\[LI\]: (\d+)", stdout)]}


def run_main(kind, td, ranks=1):
    """ranks = 2: the NRANKS_D3 = 2 build as two processes under oracle/mpi_mini (rank r on GPU r for the library-linked one)"""
    sfx = GEOM if ranks == 1 else "%s_r%d" % (GEOM, ranks)
    exe = os.path.join(ROOT, "oracle", "_ref", "main_%s_%s" % (kind, sfx))
    if not os.path.exists(exe):
        pytest.skip("no " + os.path.basename(exe))
    open(os.path.join(td, "in.set"), "w").write(open(os.path.join(HOST_DIR, "rhmc_%s.set" % sfx)).read())
    for name, r in json.load(open(os.path.join(HOST_DIR, "ratapproxes.json"))).items():
        open(os.path.join(td, name), "w").write(_remez_text(r))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(x for x in (env.get("LD_LIBRARY_PATH", ""), "/usr/local/cuda/lib64") if x)
    if ranks == 1:
        r = subprocess.run([exe, "in.set"], cwd=td, capture_output=True, text=True, timeout=1800, env=env)
        assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
        stdout, stderr = r.stdout, r.stderr
    else:
        import sys
        sys.path.insert(0, os.path.join(ROOT, "oracle", "mpi_mini"))
        from mpirun import launch
        rc = launch(ranks, [exe, "in.set"], cwd=td, env=env, timeout=1800)
        stdout, stderr = open(os.path.join(td, "stdout.0")).read(), open(os.path.join(td, "stderr.0")).read()
        assert rc == 0, (stdout[-3000:], stderr[-3000:], open(os.path.join(td, "stderr.1")).read()[-2000:])
    obs = [f for f in os.listdir(td) if f.startswith("gauge_obs")]

    class R:
        pass
    r = R(); r.stdout, r.stderr = stdout, stderr
    return r, parse_rhmc(stdout, open(os.path.join(td, obs[0])).read())


def compare(got, want):
    assert len(got["gauge_obs"]) == len(want["gauge_obs"]) == 3
    for a, b in zip(got["gauge_obs"], want["gauge_obs"]):
        assert a[0] == b[0] and a[1] == b[1], (a, b)                              # iteration, accepted
        assert abs(a[2] / b[2] - 1) < 1e-5 and abs(a[3] / b[3] - 1) < 1e-5, (a, b)   # plaquette, rectangle
        assert abs(a[4] - b[4]) < 1e-5 and abs(a[5] - b[5]) < 1e-5, (a, b)           # Polyakov loop
    assert len(got["delta_action"]) == len(want["delta_action"]) == 2
    assert all(abs(a - b) < 1e-3 for a, b in zip(got["delta_action"], want["delta_action"])), (got["delta_action"], want["delta_action"])
    for k in ("cgm_md", "cgm_fi", "cgm_li"):
        assert len(got[k]) == len(want[k])
        assert all(abs(a - b) <= max(2, 0.02 * b) for a, b in zip(got[k], want[k])), (k, got[k], want[k])


def test_reference_rhmc_main_with_the_library(tmp_path):
    want = json.load(open(os.path.join(HOST_DIR, "rhmc_%s.json" % GEOM)))
    r, got = run_main("staple", str(tmp_path))
    assert "hot path served by staple_b200" in r.stderr
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump({"got": got, "want": want}, open(os.path.join(out, "reference_rhmc_main_%s.json" % GEOM), "w"), indent=1)
    print("library-linked main:", got)
    compare(got, want)


def test_reference_rhmc_main_with_the_library_two_gpus(tmp_path):
    """the same program on two D3 slabs, one process per GPU under oracle/mpi_mini, against the pure-reference two-rank run"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    want = json.load(open(os.path.join(HOST_DIR, "rhmc_%s_r2.json" % GEOM)))
    r, got = run_main("staple", str(tmp_path), ranks=2)
    assert "hot path served by staple_b200" in r.stderr
    print("library-linked main, 2 ranks:", got)
    compare(got, want)

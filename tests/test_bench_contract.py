"""bench.py's JSON contract (task statement, section 4): the reference arm is run for real on the CPU (one bounded sample),
the product arm's line is checked on the latest committed B200 run (profiles/), and both must describe the same workload."""
import glob
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _latest_product_line():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_bench_n1.json")))
    if not files:
        pytest.skip("no round-2 B200 bench line committed yet")
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_product_line_carries_the_contract():
    d = _latest_product_line()
    assert BASE_KEYS | {"roofline", "clocks", "parity", "secondary"} <= set(d)
    assert d["metric"] == "deo_doe_gflops" and d["unit"] == "GFLOP/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["config"]["global_lattice"] == "64x64x64x128"
    assert d["vs_baseline"] is None and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.5 < r["frac"] <= 1.12      # the denominator is a measured COPY bandwidth (half reads, half writes); a 95 % read stream can exceed it
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port") and c["cores"] >= 1
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] == 48 * 16777216 and e["d2h_bytes_per_step"] == 48 * 16777216 and 0 < e["value"] < d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = d["parity"]
    assert p["ok"] is True and p["max_rel_err"] <= 1e-13 and p["slices_compared"] > 0
    assert {"config1_32x32x32x32", "config5_64x64x64x16", "config3_48x48x48x96"} <= set(d["secondary"])


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_64x64x64x4_r1.so")), reason="no reference build for the 64^3 x 4 sample slab")
def test_reference_arm_runs_on_the_cpu_and_matches_the_workload():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert BASE_KEYS | {"impl"} <= set(d) and d["impl"] == "reference"
    assert d["steps"] == 2 and d["warmup"] == 1              # honours --steps / --warmup (below the 90 s cap)
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.05 < d["value"] < 500                           # CPU cores: order 0.5-1 GFLOP/s each
    import bench
    assert d["config"] == bench.workload_config("64x64x64x128", 1)      # byte-identical `config` objects on both arms
    p = _latest_product_line()
    for k in ("metric", "unit", "higher_is_better", "dtype", "scaling", "config"):
        assert d[k] == p[k], k


def test_cpu_threads_do_not_disturb_one_another():
    """the reference arm runs the reference's Doe+Deo on several host threads at once (one source per thread, shared gauge field):
    same results as one after the other"""
    import threading
    import numpy as np
    from oracle.pyoracle import RefLib, gaussian_vec, random_su3_conf, have_ref
    if not have_ref(8, 8, 8, 8):
        pytest.skip("no reference build")
    R = RefLib(8, 8, 8, 8)
    u = random_su3_conf(R.sizeh, 1); ph = R.phases()
    vs = [gaussian_vec(R.sizeh, 10 + i) for i in range(4)]
    want = [R.dslash("acc_Deo", u, R.dslash("acc_Doe", u, v, ph), ph) for v in vs]
    got = [None] * 4

    def work(i):
        for _ in range(20):
            got[i] = R.dslash("acc_Deo", u, R.dslash("acc_Doe", u, vs[i], ph), ph)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_reference_arm_other_ranks_exit_without_work():
    """under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 and print nothing"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_workload_at_every_n():
    import bench
    for n in (1, 2, 4, 8):
        c = bench.workload_config("64x64x64x128", n)
        assert c["global_lattice"] == "64x64x64x128" and c["local_lattice"] == "64x64x64x%d" % (128 // n)
        assert json.dumps(c) == json.dumps(bench.workload_config("64x64x64x128", n))

#!/usr/bin/env python
"""Sweep the compile-time knobs of the Dslash kernel on the GPU box.
  here (CPU):   python scripts/tune_dslash.py build      -> build/variants/libstaple_<tag>.so
  on the box:   python scripts/tune_dslash.py run [lattice]  -> gpurun_out/tune_dslash.txt
"""
import itertools
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")
VARIANTS = [(b, m, l) for b, m, l in itertools.product((64, 128, 256), (1, 2), (0, 1, 3))
            if not (b == 256 and m == 2)]
VARIANTS = [(128, 1, 0), (128, 7, 0), (64, 14, 0), (64, 13, 0), (32, 28, 0), (128, 7, 2), (256, 3, 0), (192, 4, 0), (192, 5, 0),
            (96, 9, 0), (96, 10, 0), (160, 5, 0), (160, 6, 0), (224, 4, 0), (64, 12, 0), (64, 15, 0)]
if os.environ.get("STAPLE_TUNE_VARIANTS"):
    VARIANTS = [tuple(int(x) for x in v.split(",")) for v in os.environ["STAPLE_TUNE_VARIANTS"].split(";")]


def tag(v):
    return "b%d_m%d_l%d" % v


def build():
    from openstaple_b200.build import build as b
    os.makedirs(VDIR, exist_ok=True)
    for v in VARIANTS:
        flags = ["-DSTAPLE_DSLASH_BLOCK=%d" % v[0], "-DSTAPLE_DSLASH_MINBLOCKS=%d" % v[1], "-DSTAPLE_LINK_LOAD=%d" % v[2]]
        out = os.path.join(VDIR, "libstaple_%s.so" % tag(v))
        b(force=True, extra_flags=flags, out=out, tag="var_" + tag(v))
        r = subprocess.run("cuobjdump -res-usage %s | c++filt | grep -A1 'dslash_kernel<double, 0, 0>' | grep -o 'REG:[0-9]*\\|STACK:[0-9]*'" % out,
                           shell=True, capture_output=True, text=True)
        print("built", tag(v), " ".join(r.stdout.split()), flush=True)


SNIPPET = r"""
import sys, torch, numpy as np
sys.path.insert(0, %r)
import openstaple_b200 as osb
import bench
loc = tuple(int(x) for x in %r.split('x'))
torch.cuda.set_stream(torch.cuda.Stream())
lat = osb.Lattice(loc)
u, v = bench.make_fields(torch, lat, 1)
ph = lat.to_device(bench.staggered_phases(lat, 0))
a, b = v.clone(), lat.new_vec()
res = []
for name in ('acc_Doe_unsafe', 'acc_Deo_unsafe'):
    f = getattr(lat, name)
    for _ in range(5): f(u, b, a, ph)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(100): f(u, b, a, ph)
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) * 10)
n = lat.sizeh
print('RESULT doe %%.2f us deo %%.2f us  -> %%.0f / %%.0f GB/s' %% (res[0], res[1], 928 * n / res[0] / 1e3, 928 * n / res[1] / 1e3))
"""


def run(lattice):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    for v in VARIANTS:
        lib = os.path.join(VDIR, "libstaple_%s.so" % tag(v))
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, STAPLE_LIB=lib)
        r = subprocess.run([sys.executable, "-c", SNIPPET % (ROOT, lattice)], env=env, capture_output=True, text=True, cwd=ROOT)
        out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        line = "%-14s %s" % (tag(v), out[0] if out else "FAILED " + r.stderr[-300:])
        print(line, flush=True); lines.append(line)
    open(os.path.join(ROOT, "gpurun_out", "tune_dslash_%s.txt" % lattice), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "32x32x32x32")

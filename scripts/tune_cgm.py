#!/usr/bin/env python
"""Sweep the compile-time knobs of the CG-M shifted-vector pass (cgm_fused_kernel) on the GPU box.
  here (CPU):   python scripts/tune_cgm.py build        -> build/variants/libstaple_cgm_<tag>.so
  on the box:   python scripts/tune_cgm.py run [lattice] -> gpurun_out/tune_cgm_<lattice>.txt
Every variant runs the same fixed-length solve (all shifts kept active by an unreachable residue), so the time per
iteration is directly comparable: bytes/iteration = (2192 + 192 N) x sites.
"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")
# (block, minblocks, unroll, stream)
VARIANTS = [(256, 1, 2, 0), (256, 1, 2, 1), (256, 2, 2, 1), (128, 1, 2, 1), (128, 4, 2, 1), (512, 1, 2, 1),
            (256, 1, 1, 1), (256, 3, 1, 1), (256, 1, 3, 1), (256, 1, 4, 1), (128, 1, 4, 1), (128, 5, 1, 1)]
if os.environ.get("STAPLE_TUNE_VARIANTS"):
    VARIANTS = [tuple(int(x) for x in v.split(",")) for v in os.environ["STAPLE_TUNE_VARIANTS"].split(";")]


def tag(v):
    return "b%d_m%d_u%d_s%d" % v


def build():
    from openstaple_b200.build import build as b
    os.makedirs(VDIR, exist_ok=True)
    for v in VARIANTS:
        flags = ["-DSTAPLE_CGM_BLOCK=%d" % v[0], "-DSTAPLE_CGM_MINBLOCKS=%d" % v[1], "-DSTAPLE_CGM_UNROLL=%d" % v[2],
                 "-DSTAPLE_CGM_STREAM=%d" % v[3]]
        out = os.path.join(VDIR, "libstaple_cgm_%s.so" % tag(v))
        b(force=True, extra_flags=flags, out=out, tag="cgm_" + tag(v), only=["staple_solvers.cu"])
        r = subprocess.run("cuobjdump -res-usage %s | c++filt | grep -A1 'cgm_fused_kernel<double>' | grep -o 'REG:[0-9]*\\|STACK:[0-9]*'" % out,
                           shell=True, capture_output=True, text=True)
        print("built", tag(v), " ".join(r.stdout.split()), flush=True)


def run(lattice, iters):
    """one process: the fields are generated once, every variant is a separately loaded copy of the library"""
    import numpy as np
    import torch
    import openstaple_b200 as osb
    import openstaple_b200.lib as oslib
    import bench
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    loc = tuple(int(x) for x in lattice.split("x"))
    torch.cuda.set_stream(torch.cuda.Stream())
    real_stdout = os.dup(1); os.dup2(2, 1)      # the library prints the reference's warnings on stdout
    lines = []
    fields = None
    for v in VARIANTS:
        libp = os.path.join(VDIR, "libstaple_cgm_%s.so" % tag(v))
        if not os.path.exists(libp):
            continue
        os.environ["STAPLE_LIB"] = libp
        oslib._LIB = None
        lat = osb.Lattice(loc)
        if fields is None:
            u, src = bench.make_fields(torch, lat, 1)
            ph = lat.to_device(bench.staggered_phases(lat, 0))
            fields = (u, src, ph)
        u, src, ph = fields
        pars = lat.ferm_param(0.0507, ph)
        out = []
        for n in (19, 10, 4):
            shifts = np.geomspace(1e-4, 2.0, n)
            approx = osb.RationalApprox.make(1.0, np.ones(n), shifts)
            sol, ps = lat.new_vec(n), lat.new_vec(n)
            r, h, s, p = (lat.new_vec() for _ in range(4))
            lat.multishift_invert(u, pars, approx, sol, src, 1e-300, r, h, s, p, ps, 16)
            lat.multishift_invert(u, pars, approx, sol, src, 1e-300, r, h, s, p, ps, iters)
            it, act, ms = lat.last_solve_stats()
            gb = (2192.0 * it + 192.0 * act) * lat.sizeh / (ms * 1e-3) / 1e9
            out.append("N=%d %.4f ms/it %.0f GB/s" % (n, ms / it, gb))
            del sol, ps, r, h, s, p
        line = "%-18s %s" % (tag(v), " | ".join(out))
        os.write(real_stdout, (line + "\n").encode()); lines.append(line)
    open(os.path.join(ROOT, "gpurun_out", "tune_cgm_%s.txt" % lattice), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "32x32x32x32", int(sys.argv[3]) if len(sys.argv) > 3 else 200)

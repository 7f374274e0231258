#!/usr/bin/env python
"""timing-experiment variants of the library (never shipped): build/lib_<tag>.so"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from openstaple_b200.build import build
V = {"mr7": (["-DSTAPLE_DSLASH_MINBLOCKS_MR=7"], "staple_kernels.cu"), "faces6": (["-DSTAPLE_DSLASH_MINBLOCKS_FACES=6"], "staple_kernels.cu"),
     "sigma_index": (["-DSTAPLE_SIGMA_PER_LINK=0"], "staple_stout_force.cu"), "sigma_mb3": (["-DSTAPLE_SIGMA_MINBLOCKS=3"], "staple_stout_force.cu"),
     "sigma_mb4": (["-DSTAPLE_SIGMA_MINBLOCKS=4"], "staple_stout_force.cu")}
for tag in (sys.argv[1:] or V):
    out = os.path.join(ROOT, "build", "lib_%s.so" % tag)
    build(force=True, extra_flags=V[tag][0], out=out, tag="obj_" + tag, only=[V[tag][1]])
    print(out)

"""Build libstaple_b200.so (CUDA sm_100a) in-tree with nvcc.  No torch headers, no JIT cache:
the .so sits next to the sources so that it travels with the repo snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libstaple_b200.so")
SOURCES = ["staple_core.cu", "staple_kernels.cu", "staple_solvers.cu", "staple_force.cu", "staple_stout.cu", "staple_stout_force.cu", "staple_callers.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "staple_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None, tag="obj", only=None):
    """extra_flags/out/tag: tuning variants (scripts/tune_*.py) built next to the default library; `only` lists the
    sources the extra flags affect -- the other objects are taken from the default build (build/obj)."""
    lib = out or LIB
    if not force and not extra_flags and not _stale():
        return lib
    objdir = os.path.join(HERE, "..", "build", tag)
    os.makedirs(objdir, exist_ok=True)
    default_objdir = os.path.join(HERE, "..", "build", "obj")

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if only is not None and src not in only:
            dflt = os.path.join(default_objdir, src.replace(".cu", ".o"))
            if os.path.exists(dflt) and os.path.getmtime(dflt) >= os.path.getmtime(os.path.join(CSRC, src)):
                return dflt
        cmd = [NVCC] + FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    # -Bsymbolic-functions: calls between the library's own entry points (stout_wrapper -> stout_isotropic -> calc_loc_staples_...,
    # the force chain, the wrappers) bind inside the .so.  A host program keeps files such as plaquettes.c / su3_utilities.c that
    # define a few of the same names; without this flag its CPU definitions would interpose on the library's internal calls.
    # Data symbols (verbosity_lv, act_params, inverter_tricks, ... weak here) are NOT bound locally: the host's definitions win.
    cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic-functions",
                                                 "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

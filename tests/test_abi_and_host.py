"""CPU tests of the drop-in boundary and the host-side logic (no compute call needs a GPU):
  * libstaple_b200.so loads and exports every symbol include/staple_b200.h declares;
  * the struct layouts that cross the boundary equal the reference's (sizeof/offsetof taken from the
    reference's own headers by oracle/ref_shim.c:ref_abi, committed in tests/golden/ref_abi_approx.npz);
  * .REMEZ reader and rescale_rational_approximation mirror (rationalapprox.c:83-117, :145-194);
  * D3 slab sharding arithmetic (staple_geometry_plan) against the reference's ranges and halo offsets.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import openstaple_b200 as osb
from openstaple_b200.api import FermParam, InverterPackage, InvTricks, RationalApprox
from oracle.pyoracle import Restatement

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def abi():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_abi_approx.npz")))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "staple_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set()
    # plain prototypes
    for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", txt, flags=re.M):
        names.add(m.group(1))
    names -= {"defined", "STAPLE_DSLASH_DECL", "STAPLE_BLAS_DECL", "STAPLE_FORCE_DECL", "STAPLE_STOUT_DECL", "STAPLE_SF_DECL", "name", "if", "sizeof"}
    out = set(n for n in names if not n.isupper())
    # macro-generated families
    for m in re.finditer(r"^STAPLE_DSLASH_DECL\((\w+)\)", txt, flags=re.M):
        out |= {m.group(1), m.group(1) + "_f"}
    blas = re.search(r"#define STAPLE_BLAS_DECL\(V, S\)(.*?)\nSTAPLE_BLAS_DECL", txt, flags=re.S).group(1)
    for m in re.finditer(r"(\w+)##S\(", blas):
        out |= {m.group(1), m.group(1) + "_f"}
    force = re.search(r"#define STAPLE_FORCE_DECL\(S, SU3, VEC3, TAMAT\)(.*?)\nSTAPLE_FORCE_DECL", txt, flags=re.S).group(1)
    for m in re.finditer(r"(\w+)##S\(", force):
        out |= {m.group(1), m.group(1) + "_f"}
    stout = re.search(r"#define STAPLE_STOUT_DECL\(S, SU3, TAMAT\)(.*?)\nSTAPLE_STOUT_DECL", txt, flags=re.S).group(1)
    for m in re.finditer(r"(\w+)##S\(", stout):
        out |= {m.group(1), m.group(1) + "_f"}
    sf = re.search(r"#define STAPLE_SF_DECL\(S, SU3, TAMAT, THMAT\)(.*?)\nSTAPLE_SF_DECL", txt, flags=re.S).group(1)
    for m in re.finditer(r"(\w+)##S\(", sf):
        out |= {m.group(1), m.group(1) + "_f"}
    out -= {"name##_f", "name"}
    return sorted(out)


def test_library_exports_every_declared_symbol():
    L = osb.load_library()
    syms = declared_symbols()
    assert len(syms) > 126, syms
    for s_ in ("ker_openacc_compute_fermion_force", "ker_openacc_compute_fermion_force_f", "set_tamat_soa_to_zero",
               "multiply_conf_times_force_and_take_ta_nophase_f", "staple_acc_Doe_Deo_streamed", "stout_wrapper", "stout_isotropic_f",
               "calc_loc_staples_nnptrick_all_onlyferms", "exp_minus_QA_times_conf", "fermion_force_soloopenacc",
               "fermion_force_soloopenacc_f", "eo_inversion", "acc_Deo_wf", "acc_Doe_wf", "acc_Deo_wf_unsafe", "acc_Doe_wf_unsafe"):
        assert s_ in syms, s_
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    for g in ("verbosity_lv", "multishift_invert_iterations", "nMdInversionPerformed"):
        C.c_int.in_dll(L, g)
    for g in ("aux_th", "aux_ta", "aux_th_f", "aux_ta_f", "conf_acc_f", "auxbis_conf_acc", "glocal_staples", "gipdot"):
        assert C.c_void_p.in_dll(L, g).value is None          # the reference's parking-array globals, unset until the host sets them
    assert C.sizeof(osb.api.MdParam) == 64 and osb.api.MdParam.recycleInvsForce.offset == 52   # md_parameters.h:6-19
    InvTricks.in_dll(L, "inverter_tricks")
    assert b"sm_100a" in L.staple_version()
    # the reference's own entry-point names are all there (fermion_matrix.h, fermionic_utilities.h, inverter_*.h)
    for s in ("acc_Deo", "acc_Doe", "fermion_matrix_multiplication", "fermion_matrix_multiplication_shifted",
              "multishift_invert", "multishift_invert_f", "recombine_shifted_vec3_to_vec3", "ker_invert_openacc",
              "inverter_mixed_precision", "inverter_multishift_wrapper", "inverter_wrapper",
              "scal_prod_global", "real_scal_prod_global", "l2norm2_global", "communicate_fermion_borders",
              "communicate_su3_borders", "shutdown_multidev", "setup_inverter_package_dp", "setup_inverter_package_sp"):
        assert s in syms


def test_compute_without_init_fails_loudly():
    """No CPU fallback: a compute entry point called without an initialised CUDA context aborts."""
    import subprocess
    import sys
    code = ("import openstaple_b200 as o, numpy as np\n"
            "L = o.load_library(); a = np.zeros(96)\n"
            "L.l2norm2_global(a.ctypes.data)\nprint('SURVIVED')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout
    assert "staple_init_geometry" in r.stderr


def test_struct_layouts_match_reference(abi):
    a = [int(x) for x in abi["abi"]]; S = int(abi["abi_sizeh"])
    assert a[0] == 48 * S and a[1] == 144 * S and a[2] == 8 * S          # vec3_soa, su3_soa, double_soa
    assert a[3] == 24 * S and a[4] == 72 * S and a[5] == 4 * S           # _f twins
    assert a[22] == 16 * S and a[23] == 48 * S                           # c1 / r1 offsets (SoA strides used by the kernels)
    assert C.sizeof(FermParam) == a[6]
    assert (FermParam.ferm_mass.offset, FermParam.phases.offset, FermParam.phases_f.offset, FermParam.approx_md.offset) == tuple(a[7:11])
    assert C.sizeof(RationalApprox) == a[11]
    assert (RationalApprox.approx_order.offset, RationalApprox.RA_a0.offset, RationalApprox.RA_a.offset,
            RationalApprox.RA_b.offset) == tuple(a[12:16])
    assert C.sizeof(InverterPackage) == a[16]
    assert (InverterPackage.nshifts.offset, InverterPackage.loc_r.offset, InverterPackage.out_f.offset) == tuple(a[17:20])
    assert C.sizeof(InvTricks) == a[20] and InvTricks.mixedPrecisionDelta.offset == a[21]


def test_host_global_struct_layouts_match_reference(abi):
    """act_params and md_parameters are read by the library from the HOST program's own objects (weak/pre-emptible globals):
    the layouts include/staple_b200.h declares must be the reference's (action.h:6-18, md_parameters.h:6-19); tamat/thmat packing"""
    from openstaple_b200.api import ActionParam, MdParam
    a = [int(x) for x in abi["abi2"]]; S = int(abi["abi_sizeh"])
    assert C.sizeof(ActionParam) == a[0]
    assert (ActionParam.stout_steps.offset, ActionParam.stout_rho.offset, ActionParam.topo_action.offset,
            ActionParam.topo_file_path.offset, ActionParam.topo_stout_steps.offset, ActionParam.topo_rho.offset) == tuple(a[1:7])
    assert C.sizeof(MdParam) == a[7]
    assert (MdParam.residue_metro.offset, MdParam.singlePrecMD.offset, MdParam.max_cg_iterations.offset,
            MdParam.recycleInvsForce.offset) == tuple(a[8:12])
    assert a[12] == 64 * S and a[13] == 48 * S and a[14] == 64 * S and a[15] == 48 * S      # 3 complex + 2 real arrays per link
    # ... and the C declarations of the header itself, compiled
    import subprocess, tempfile
    src = r'''#include <stdio.h>
#include <stddef.h>
#include "staple_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(action_param), offsetof(action_param, stout_steps),
  offsetof(action_param, stout_rho), offsetof(action_param, topo_action), offsetof(action_param, topo_file_path),
  offsetof(action_param, topo_stout_steps), offsetof(action_param, topo_rho), sizeof(md_param), offsetof(md_param, residue_metro),
  offsetof(md_param, singlePrecMD), offsetof(md_param, max_cg_iterations), offsetof(md_param, recycleInvsForce)); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "a.c"), "w").write(src)
        subprocess.run(["gcc", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), os.path.join(td, "a.c"), "-o", os.path.join(td, "a")], check=True)
        out = subprocess.run([os.path.join(td, "a")], capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == a[:12]


def write_remez(path, g, tag):
    """the .REMEZ text layout the reference's reader expects (rationalapprox.c:96-106)."""
    with open(path, "w") as f:
        f.write("\nApproximation to f(x) = (x)^(%d/%d)\n" % (int(g[tag + "_num"]), int(g[tag + "_den"])))
        f.write("Order: %d\nLambda Min: %.16e\nLambda Max: %.16e\n" % (int(g[tag + "_order"]), float(g[tag + "_lmin"]), float(g[tag + "_lmax"])))
        f.write("GMP Remez Precision: %d\nError: %.16e\nRA_a0 = %.16e\n" % (int(g[tag + "_prec"]), float(g[tag + "_error"]), float(g[tag + "_a0"])))
        for i, (x, y) in enumerate(zip(g[tag + "_a"], g[tag + "_b"])):
            f.write("RA_a[%d] = %.16e, RA_b[%d] = %.16e\n" % (i, x, i, y))


@pytest.mark.parametrize("tag", ["m14", "p18", "m14o9"])
def test_remez_reader_and_rescale(abi, tag, tmp_path):
    g = abi
    p = str(tmp_path / "approx.REMEZ")
    write_remez(p, g, tag)
    r = RationalApprox.read(p)
    assert (r.exponent_num, r.exponent_den, r.approx_order) == (int(g[tag + "_num"]), int(g[tag + "_den"]), int(g[tag + "_order"]))
    assert r.lambda_min == float(g[tag + "_lmin"]) and r.RA_a0 == float(g[tag + "_a0"])
    assert np.array_equal(np.array(r.RA_a[:r.approx_order]), g[tag + "_a"])
    assert np.array_equal(np.array(r.RA_b[:r.approx_order]), g[tag + "_b"])
    out = r.rescaled(tuple(g["rescale_minmax"]))
    t = tag + "_rescaled"
    assert abs(out.RA_a0 / float(g[t + "_a0"]) - 1) < 1e-15
    assert np.allclose(np.array(out.RA_a[:r.approx_order]), g[t + "_a"], rtol=1e-15, atol=0)
    assert np.allclose(np.array(out.RA_b[:r.approx_order]), g[t + "_b"], rtol=1e-15, atol=0)
    assert abs(out.lambda_max / float(g[t + "_lmax"]) - 1) < 1e-15 and abs(out.lambda_min / float(g[t + "_lmin"]) - 1) < 1e-15


@pytest.mark.parametrize("loc_n,nr", [((8, 8, 8, 8), 1), ((8, 8, 8, 8), 2), ((4, 4, 4, 4), 2), ((64, 64, 64, 16), 8),
                                      ((64, 64, 64, 2), 8), ((48, 48, 48, 48), 2), ((32, 32, 32, 32), 1)])
def test_geometry_plan_matches_oracle(loc_n, nr):
    p = osb.geometry_plan(loc_n, nr)
    S = Restatement(*loc_n, nr=nr)
    assert p["nd"] == S.nd and p["sizeh"] == S.sizeh and p["vol3h"] == S.vol3h and p["d3_halo"] == S.d3_halo
    assert p["r0"] == (S.g.r0_lo, S.g.r0_hi) and p["r1"] == (S.g.r1_lo, S.g.r1_hi)
    if nr > 1:
        # first/last interior slice -> neighbour's inner halo slice (communications.c:51-96)
        v = p["vol3h"]; h = p["d3_halo"]
        assert p["send_L"] == h * v and p["recv_L"] == (h - 1) * v
        assert p["send_R"] == (h + loc_n[3] - 1) * v and p["recv_R"] == (h + loc_n[3]) * v and p["slab"] == v


def test_geometry_plan_worked_example_and_rejections(golden_r2):
    """8^3 x (2*8) -> nd3=12, sizeh=3072, R0=[512,2560), R1=[256,2816), halo offsets."""
    p = osb.geometry_plan((8, 8, 8, 8), 2)
    assert p == dict(nd=(8, 8, 8, 12), sizeh=3072, vol3h=256, r0=(512, 2560), r1=(256, 2816),
                     send_L=512, recv_R=2560, send_R=2304, recv_L=256, slab=256, d3_halo=2)
    g = golden_r2
    q = osb.geometry_plan(tuple(int(x) for x in g["loc_n"]), int(g["nranks"]))
    assert q["r0"] + q["r1"] == tuple(int(x) for x in g["ranges"]) and q["sizeh"] == int(g["sizeh"])
    for bad in (((7, 8, 8, 8), 1), ((8, 8, 8, 7), 2), ((8, 8, 8, 1), 1), ((8, 0, 8, 8), 1)):
        with pytest.raises(ValueError):
            osb.geometry_plan(*bad)
    with pytest.raises(ValueError):
        osb.geometry_plan((8, 8, 8, 8), 2, halo_width=1)     # Wilson multi-rank flips parities (io.c:595-597)


# ----------------------------------------------------------------------------- prototypes against the reference's headers
# headers whose EVERY function must be exported (the subsystems the library replaces, SURVEY 8b + the "next" rows built)
COMPLETE_HEADERS = ["OpenAcc/fermion_matrix.h", "OpenAcc/sp_fermion_matrix.h", "OpenAcc/fermionic_utilities.h",
                    "OpenAcc/sp_fermionic_utilities.h", "OpenAcc/inverter_multishift_full.h", "OpenAcc/sp_inverter_multishift_full.h",
                    "OpenAcc/inverter_full.h", "OpenAcc/sp_inverter_full.h", "OpenAcc/inverter_mixedp.h", "OpenAcc/inverter_wrappers.h",
                    "OpenAcc/inverter_package.h", "OpenAcc/float_double_conv.h", "OpenAcc/find_min_max.h",
                    "OpenAcc/fermion_force_utilities.h", "OpenAcc/sp_fermion_force_utilities.h", "OpenAcc/fermion_force.h",
                    "OpenAcc/sp_fermion_force.h", "OpenAcc/stouting.h", "OpenAcc/sp_stouting.h", "OpenAcc/field_times_fermion_matrix.h"]
# the only deliberate differences: C99 `double complex` returned as an ABI-identical {re, im} struct; MPI_Request arrays are
# opaque pointers here (the requests are CUDA events owned by the library)
TYPE_ALIAS = {"d_complex": "staple_dcomplex_ret", "MPI_Request*": "staple_request*"}


def test_prototypes_match_reference_headers():
    """names, argument order, argument types and return types of every entry point against the reference's own headers
    (tests/golden/ref_prototypes.json, written by tests/golden/make_golden.py:reference_prototypes)."""
    import json
    from prototypes import preprocess, prototypes
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_prototypes.json")))
    ours = prototypes(preprocess('#include "%s"\n' % os.path.join(ROOT, "include", "staple_b200.h")))
    alias = lambda t: TYPE_ALIAS.get(t, t)
    checked, bad, missing = 0, [], []
    for hdr, protos in ref.items():
        for name, (rt, params) in protos.items():
            if name not in ours:
                if hdr in COMPLETE_HEADERS:
                    missing.append((hdr, name))
                continue
            o_rt, o_params = ours[name]
            if alias(rt) != o_rt or [alias(p) for p in params] != o_params:
                bad.append((name, ours[name], [rt, params]))
            checked += 1
    assert not bad, bad
    assert not missing, missing
    assert checked >= 145, checked
    L = osb.load_library()
    assert not [n for hdr in COMPLETE_HEADERS for n in ref[hdr] if not hasattr(L, n)]


# ----------------------------------------------------------------------------- a C host program links and runs
REF_SCRATCH = os.path.join(os.environ.get("STAPLE_ORACLE_SCRATCH", os.path.join(os.environ.get("TMPDIR", "/tmp"), "staple_oracle_src")), "src")


def _gcc_host(tmp_path, extra, link=True):
    import subprocess
    libdir = os.path.join(ROOT, "openstaple_b200")
    exe = str(tmp_path / ("host" if link else "host.o"))
    cmd = ["gcc", "-std=gnu99", "-Wall", "-Werror=implicit-function-declaration", "-I" + os.path.join(ROOT, "include")] + extra + \
          [os.path.join(ROOT, "tests", "c_host", "host_link.c"), "-o", exe]
    cmd += ["-L" + libdir, "-lstaple_b200", "-Wl,-rpath," + libdir] if link else ["-c"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_c_host_program_links_and_runs(tmp_path):
    """INTEGRATION.md in practice: gcc compiles a C host against include/staple_b200.h, the link step resolves the entry points
    from libstaple_b200.so, and the program runs (host-side geometry arithmetic only -- no GPU)."""
    import subprocess
    osb.load_library()
    exe = _gcc_host(tmp_path, [])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("nd3 12 sizeh 3072 r0 512 2560 r1 256 2816 linked 49 "), r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_SCRATCH, "OpenAcc")), reason="reference headers (scratch copy with the generated sp_* twins) not present")
def test_header_coexists_with_reference_headers(tmp_path):
    """a translation unit may include the reference's own headers AND include/staple_b200.h: every declaration compatible
    (double complex return of scal_prod_global, MPI_Request* of the async exchanges, struct guards)"""
    flags = ["-DWITH_REFERENCE_HEADERS", "-fcommon", "-w", "-I" + os.path.join(ROOT, "oracle", "mpi_stub"), "-I" + REF_SCRATCH,
             "-DACTION_TYPE=TLSM", "-DNREPLICAS=1", "-DLOC_N0=8", "-DLOC_N1=8", "-DLOC_N2=8", "-DLOC_N3=8", "-DNRANKS_D3=2", "-DCOMMIT_HASH=t"]
    flags += ["-D%s%s=8" % (a, b) for a in ("DEODOE", "IMPSTAP", "STAP", "SIGMA") for b in ("TILE0", "TILE1", "TILE2", "GANG3")]
    _gcc_host(tmp_path, flags, link=False)


def test_library_internal_calls_cannot_be_interposed():
    """A host program keeps reference files (plaquettes.c, su3_utilities.c, ferm_meas.c) that define a few names the library also
    exports.  The library is linked -Bsymbolic-functions: calls between its own entry points carry no PLT relocation (they can
    not be captured by the host's CPU definitions -- no accidental CPU fallback), while the reference's data globals stay
    pre-emptible so that the host's own definitions are the ones the library reads."""
    import subprocess
    rel = subprocess.run(["readelf", "-rW", osb.library_path()], capture_output=True, text=True).stdout
    plt = [l.split()[4] for l in rel.splitlines() if "JUMP_SLO" in l and len(l.split()) > 4]
    own = set(declared_symbols())
    assert not [s for s in plt if s.split("@")[0] in own], [s for s in plt if s.split("@")[0] in own]
    glob = " ".join(l for l in rel.splitlines() if "GLOB_DAT" in l)
    for g in ("verbosity_lv", "inverter_tricks", "act_params", "md_parameters", "aux_th", "aux_ta", "conf_acc_f", "gl_stout_rho"):
        assert g in glob, g


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under openstaple_b200/ (Python or CUDA) may import, link or call it, and
    bench.py reaches it only in its cpu_baseline / --impl reference legs and in the in-run parity check of the GPU results"""
    import ast
    pkg = os.path.join(ROOT, "openstaple_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("pyoracle", "staggered_oracle", "oracle/", "oracle.", "_ref/", "libref_", "import oracle", "from oracle"):
                    assert needle not in txt, (os.path.join(dirpath, f), needle)
    needed = subprocess_out(["readelf", "-d", osb.library_path()])
    assert "oracle" not in needed and "libref" not in needed
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    users = set()
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("oracle"):
                users.add(fn.name)
    # the cpu_baseline legs (timing the reference on the host), the --impl reference arm, and the in-run parity CHECK of the
    # GPU results (the oracle as checker, never as the thing measured)
    assert users <= {"run_reference", "cpu_operator_throughput", "cpu_cgm_per_site_iteration", "parity_windows"}, users
    assert not [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]


def subprocess_out(cmd):
    import subprocess
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def _approx_from_numbers(r):
    a = RationalApprox.make(r["a0"], r["a"], r["b"], r["num"], r["den"])
    a.lambda_min, a.lambda_max, a.gmp_remez_precision, a.error = r["lmin"], r["lmax"], r["prec"], r["error"]
    return a


def test_rational_approx_host_functions():
    """filename / evaluate / renormalized / save of RationalApprox (rationalapprox.c:44-70, 120-143, 197-237) on the six
    approximations of tools/test (numbers parsed by the reference's own reader, tests/golden/ref_host/ratapproxes.json):
    against the reference build where it is present, and against x^(num/den) within the approximation's own error"""
    import json
    from oracle.pyoracle import ref_lib_path
    approxes = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_host", "ratapproxes.json")))
    lib = C.CDLL(ref_lib_path(4, 4, 4, 4)) if os.path.exists(ref_lib_path(4, 4, 4, 4)) else None
    if lib is not None:
        lib.rational_approx_evaluate.restype = C.c_double; lib.rational_approx_filename.restype = C.c_char_p
    for fname, r in approxes.items():
        a = _approx_from_numbers(r)
        p = a.exponent_num / a.exponent_den
        for x in (a.lambda_min * 1.01, 1e-3, 0.1, 1.0):
            if x >= a.lambda_min:
                assert abs(a.evaluate(x) / x ** p - 1) < 3 * a.error, (fname, x)
        n = a.renormalized()
        assert n.lambda_max == 1.0 and abs(n.evaluate(0.5) / 0.5 ** p - 1) < 3 * a.error
        assert a.filename().startswith("approx_%d_over_%d_mlogerr_" % (a.exponent_num, a.exponent_den)) and a.filename().endswith(fname[-17:])
        if lib is None:
            continue
        ref = RationalApprox.from_buffer_copy(a)
        assert lib.rational_approx_filename(C.c_double(a.error), C.c_int(a.exponent_num), C.c_int(a.exponent_den),
                                            C.c_double(a.lambda_min)).decode() == a.filename()
        for x in (1e-5, 0.003, 0.5, 1.0):
            assert lib.rational_approx_evaluate(C.byref(ref), C.c_double(x)) == a.evaluate(x)
        out = RationalApprox(); lib.renormalize_rational_approximation(C.byref(ref), C.byref(out))
        assert np.allclose(np.array(n.RA_a[:n.approx_order]), np.array(out.RA_a[:n.approx_order]), rtol=1e-15, atol=0)
        assert np.allclose(np.array(n.RA_b[:n.approx_order]), np.array(out.RA_b[:n.approx_order]), rtol=1e-15, atol=0)
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            p1, p2 = os.path.join(td, "a"), os.path.join(td, "b")
            a.save(p1); lib.rationalapprox_save(p2.encode(), C.byref(ref))
            assert open(p1, "rb").read() == open(p2, "rb").read()
            back = RationalApprox.read(p1)
            assert bytes(back) == bytes(a)


def test_ctypes_signatures_match_the_header():
    """every argtypes list openstaple_b200/lib.py declares has as many entries as the C prototype has parameters, floating-point
    parameters are bound as c_double / c_float (not as integers or pointers) and by-value structs are not bound as pointers"""
    from prototypes import preprocess, prototypes
    L = osb.load_library()
    ours = prototypes(preprocess('#include "%s"\n' % os.path.join(ROOT, "include", "staple_b200.h")))
    checked = 0
    for name, (rt, params) in ours.items():
        f = getattr(L, name, None)
        if f is None or f.argtypes is None:
            continue
        assert len(f.argtypes) == len(params), (name, len(f.argtypes), params)
        for a, p in zip(f.argtypes, params):
            if p == "double":
                assert a is C.c_double, (name, p, a)
            elif p == "float":
                assert a is C.c_float, (name, p, a)
            elif p in ("int", "constint"):
                assert a is C.c_int, (name, p, a)
            elif p == "inverter_package":
                assert issubclass(a, C.Structure), (name, p, a)
        if rt == "double":
            assert f.restype is C.c_double, name
        checked += 1
    assert checked > 100, checked


def test_lazily_bound_api_calls_match_the_header():
    """the Lattice methods that set argtypes at call time (by-value inverter_package, float res, ...) are driven against a
    recording stand-in for the library: the argtypes they set and the number of arguments they pass equal the C prototype"""
    from prototypes import preprocess, prototypes
    from openstaple_b200.api import Lattice, InverterPackage
    ours = prototypes(preprocess('#include "%s"\n' % os.path.join(ROOT, "include", "staple_b200.h")))
    calls = {}

    class Fn:
        def __init__(self, name):
            self.name, self.argtypes, self.restype = name, None, None

        def __call__(self, *a):
            calls[self.name] = (self.argtypes, len(a))
            return 0

    class Lib:
        def __getattr__(self, name):
            f = Fn(name); object.__setattr__(self, name, f); return f

    lat = Lattice.__new__(Lattice); lat.L = Lib(); lat._keep = []; lat.sizeh = 16
    ip, pars, approx = InverterPackage(), FermParam(), RationalApprox.make(1.0, [1.0], [0.1])
    arr = (FermParam * 1)()
    lat.eo_inversion(ip, pars, 1e-8, 10, 1, 2, 3, 4, 5, 6)
    lat.fermion_force_soloopenacc(1, 2, 3, 4, arr, 1, 5, 1e-6, 6, 7, ip, 100)
    lat.inverter_wrapper(ip, pars, 1, 2, 1e-8, 10, 0.0, 0)
    lat.inverter_multishift_wrapper(ip, pars, approx, 1, 2, 1e-8, 10, 0)
    lat.inverter_mixed_precision(ip, pars, 1, 2, 1e-8, 10, 0.0)
    lat.acc_Deo_wf(1, 2, 3, 4, 5, 6); lat.acc_Doe_wf_unsafe(1, 2, 3, 4, 5, 6)
    lat.ker_find_min_eigenvalue_openacc(1, pars, 2, 3, 4, 5.0)
    lat.setup_inverter_package_dp(ip, 1, 2, 3, 4, 5, 6, 7); lat.setup_inverter_package_sp(ip, 1, 2, 3, 4, 5, 6, 7, 8)
    assert len(calls) == 10, sorted(calls)
    for name, (argtypes, nargs) in calls.items():
        params = ours[name][1]
        assert nargs == len(params), (name, nargs, params)
        assert argtypes is not None and len(argtypes) == len(params), (name, argtypes, params)
        for a, p in zip(argtypes, params):
            if p == "double":
                assert a is C.c_double, (name, p)
            if p == "inverter_package":
                assert a is InverterPackage, (name, p)

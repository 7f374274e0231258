"""Generates tests/golden/*.npz by running the UNMODIFIED reference (gcc build of
/root/reference, oracle/build_ref.sh) on seeded inputs.  Run in the dev container only:

    python tests/golden/make_golden.py

The fixtures pin the CPU restatement (oracle/staggered_oracle.c) and the CUDA path on the GPU
box, where /root/reference does not exist.  Inputs are stored too, so nothing depends on the
numpy RNG stream.
"""
import ctypes as C
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.pyoracle import RefLib, gaussian_vec, ptr, random_su3_conf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EB = (5.0, -5.0, 1.0, -5.0, 5.0, 3.0)      # tools/test background field: ex ey ez bx by bz
MU, CHARGE = 1.0, 2.0                       # MuOverPiT 1, charge 2 (tools/test/fermion_parameters.set)
MASS = 0.0507
SHIFTS = np.array([2.0e-4, 3.0e-3, 4.0e-2, 0.5, 3.0])
RA_A = np.array([0.11, 0.23, 0.37, 0.41, 0.53]); RA_A0 = 0.7


def single_rank(n=(4, 4, 4, 4)):
    R = RefLib(*n)
    S = R.sizeh
    d = {}
    u = random_su3_conf(S, 11); v = gaussian_vec(S, 12); w = gaussian_vec(S, 13)
    uf = u.astype(np.complex64); vf = v.astype(np.complex64); wf = w.astype(np.complex64)
    d.update(u=u, v=v, w=w, mass=MASS, shifts=SHIFTS, ra_a=RA_A, ra_a0=RA_A0, eb=np.array(EB), mu=MU, charge=CHARGE,
             loc_n=np.array(n))
    for tag, args in (("p0", ((0,) * 6, 0.0, 0.0)), ("bf", (EB, MU, CHARGE))):
        ph = R.phases(*args); phf = R.phases_f(*args)
        d["ph_" + tag] = ph; d["phf_" + tag] = phf
        d["deo_" + tag] = R.dslash("acc_Deo", u, v, ph)
        d["doe_" + tag] = R.dslash("acc_Doe", u, v, ph)
        d["mdagm_" + tag] = R.mdagm(u, v, ph, MASS)
        d["mdagm_sh_" + tag] = R.mdagm(u, v, ph, MASS, shift=0.37)
        d["deo_f_" + tag] = R.dslash("acc_Deo", uf, vf, phf)
        d["doe_f_" + tag] = R.dslash("acc_Doe", uf, vf, phf)
        d["mdagm_f_" + tag] = R.mdagm(uf, vf, phf, MASS)
    ph = d["ph_bf"]; phf = d["phf_bf"]
    d["l2norm2"] = R.l2norm2(v); d["real_scal_prod"] = R.real_scal_prod(v, w)
    d["l2norm2_f"] = R.l2norm2(vf); d["real_scal_prod_f"] = R.real_scal_prod(vf, wf)
    out, cg, ok = R.multishift_invert(u, ph, MASS, (RA_A0, RA_A, SHIFTS), v, 1e-9, 5000)
    d["ms_out"] = out; d["ms_cg"] = cg; d["ms_ok"] = ok
    d["ms_recombined"] = R.recombine(out, v, (RA_A0, RA_A, SHIFTS))
    out, cg, ok = R.multishift_invert(uf, phf, MASS, (RA_A0, RA_A, SHIFTS), vf, 1e-4, 5000)
    d["ms_f_out"] = out; d["ms_f_cg"] = cg; d["ms_f_ok"] = ok
    sol, cg, ok = R.cg(u, ph, MASS, v, 1e-10, 5000, 0.01)
    d["cg_sol"] = sol; d["cg_cg"] = cg; d["cg_ok"] = ok
    R.set_inverter_tricks(0, 1, 0.1, 10000)
    sol, cg, ok = R.mixed_cg(u, uf, ph, phf, MASS, v, 1e-10, 5000, 0.01)
    d["mixed_sol"] = sol; d["mixed_cg"] = cg; d["mixed_ok"] = ok
    d["max_eig"] = R.max_eigenvalue(u, ph, MASS, w)
    np.savez_compressed(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n), **d)
    print("single rank", n, "ms_cg", d["ms_cg"], d["ms_f_cg"], "cg", d["cg_cg"], "mixed", d["mixed_cg"])


def multi_rank(loc=(4, 4, 4, 4), nr=2):
    gl = (loc[0], loc[1], loc[2], loc[3] * nr)
    G = RefLib(*gl)
    R = RefLib(*loc, nr)
    d = {}
    u = random_su3_conf(G.sizeh, 21); v = gaussian_vec(G.sizeh, 22)
    d.update(u=u, v=v, loc_n=np.array(loc), nranks=nr, eb=np.array(EB), mu=MU, charge=CHARGE, mass=MASS)
    phg = G.phases(EB, MU, CHARGE)
    d["doe_global"] = G.dslash("acc_Doe", u, v, phg)
    d["mdagm_global"] = G.mdagm(u, v, phg, MASS)
    d["l2norm2_global"] = G.l2norm2(v)
    out, cg, ok = G.multishift_invert(u, phg, MASS, (RA_A0, RA_A, SHIFTS), v, 1e-9, 5000)
    d["ms_out_global"] = out; d["ms_cg"] = cg
    ro = (C.c_long * 4)(); R.lib.ref_ranges(ro); d["ranges"] = np.array(list(ro))
    d["sizeh"] = R.sizeh; d["nd"] = np.array(R.nd)
    loc_out = []
    for r in range(nr):
        R.set_rank(r)
        d["gl_snum_r%d" % r] = np.array([R.lib.ref_lnh_to_gl_snum(0, 0, 0, d3, r) for d3 in range(R.nd[3])])
        lu = np.zeros((8, 3, 3, R.sizeh), np.complex128); R.lib.send_lnh_subconf_to_buffer(ptr(u), ptr(lu), r)
        lv = np.zeros((3, R.sizeh), np.complex128); R.lib.send_lnh_subfermion_to_buffer(ptr(v), ptr(lv), r)
        ph = R.phases(EB, MU, CHARGE)
        d["lnh_v_r%d" % r] = lv; d["lnh_ph_r%d" % r] = ph
        o = R.dslash("acc_Doe_unsafe", lu, lv, ph)
        d["doe_unsafe_r%d" % r] = o.copy()
        d["l2norm2_loc_r%d" % r] = R.l2norm2(lv)       # mailbox Allreduce returns the local contribution
        loc_out.append(o)
    R.lib.ref_mailbox_clear()
    for _ in range(2):
        for r in range(nr):
            R.set_rank(r); R.lib.communicate_fermion_borders(ptr(loc_out[r]))
    for r in range(nr):
        d["doe_exchanged_r%d" % r] = loc_out[r]
    np.savez_compressed(os.path.join(HERE, "ref_%dx%dx%dx%d_r%d.npz" % (loc + (nr,))), **d)
    print("multi rank", loc, nr, "ms_cg", cg)


def force_single(n=(4, 4, 4, 4)):
    """Fermion-force outer products (fermion_force_utilities.c) on the 4^4 fixture of single_rank(): inputs are that
    fixture's u, phases and multishift solutions (the vectors the MD force is really computed from)."""
    R = RefLib(*n)
    g = dict(np.load(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n)))
    u, ph, phf, sh = g["u"], g["ph_bf"], g["phf_bf"], g["ms_out"]
    d = {"ra_a": RA_A}
    rng = np.random.default_rng(77)
    aux0 = (rng.standard_normal((8, 3, 3, R.sizeh)) + 1j * rng.standard_normal((8, 3, 3, R.sizeh)))
    ta0 = rng.standard_normal((8, 8, R.sizeh))
    d["aux0"] = aux0; d["ta0"] = ta0
    aux = aux0.copy()
    s, h = R.compute_fermion_force(u, aux, sh, ph, RA_A)
    d["force_aux"] = aux.copy(); d["force_loc_s"] = s; d["force_loc_h"] = h
    one = aux0.copy(); R.direct_product(g["v"], g["w"], one, 0.37); d["direct_product"] = one
    pseudo = aux0[::-1].copy(); R.multiply_backfield_times_force(ph, aux, pseudo); d["backfield"] = pseudo.copy()
    R.accumulate_gl3(aux, pseudo); d["accumulated"] = pseudo.copy()
    ta = ta0.copy(); R.take_ta(u, pseudo, ta); d["ta"] = ta
    # FP32 twins (generated sp_fermion_force_utilities.c: pure float arithmetic)
    uf, shf = u.astype(np.complex64), sh.astype(np.complex64)
    auxf = aux0.astype(np.complex64)
    R.compute_fermion_force(uf, auxf, shf, phf, RA_A); d["force_aux_f"] = auxf.copy()
    pf = aux0[::-1].astype(np.complex64); R.multiply_backfield_times_force(phf, auxf, pf); d["backfield_f"] = pf.copy()
    taf = ta0.astype(np.float32); R.take_ta(uf, pf, taf); d["ta_f"] = taf
    np.savez_compressed(os.path.join(HERE, "ref_force_%dx%dx%dx%d_r1.npz" % n), **d)
    print("force", n, float(np.abs(d["force_aux"]).max()), float(np.abs(ta).max()))


def stout_single(n=(4, 4, 4, 4)):
    """Isotropic stout smearing (stouting.c) of the 4^4 fixture's links: one level with every intermediate the
    reference leaves behind, and the two-level stout_wrapper output; rho = 0.15 (tools/test input files) and a tiny
    rho that exercises the small-c1 branch of the Cayley-Hamilton exponential."""
    R = RefLib(*n)
    g = dict(np.load(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n)))
    u = g["u"]
    d = {"rho": 0.15, "rho_small": 1e-3}
    up, stap, aux, ta = R.stout_isotropic(u, 0.15)
    d.update(uprime=up, staples=stap, exp_aux=aux, tipdot=ta)
    d["uprime_small"] = R.stout_isotropic(u, 1e-3)[0]
    d["wrapper2"] = R.stout_wrapper(u, 0.15, 2)
    upf, stapf, auxf, taf = R.stout_isotropic(u.astype(np.complex64), 0.15)
    d.update(uprime_f=upf, tipdot_f=taf)
    np.savez_compressed(os.path.join(HERE, "ref_stout_%dx%dx%dx%d_r1.npz" % n), **d)
    print("stout", n, float(np.abs(ta).max()))


def stoutforce_single(n=(4, 4, 4, 4)):
    """Sigma' -> Sigma through one stout level (stouting.c:171-1305) on the 4^4 fixture: Q from stout_single(),
    a random gl(3) Sigma' as the force arriving from the level above."""
    R = RefLib(*n)
    g = dict(np.load(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n)))
    gs = dict(np.load(os.path.join(HERE, "ref_stout_%dx%dx%dx%d_r1.npz" % n)))
    u, ta, rho = g["u"], gs["tipdot"], float(gs["rho"])
    rng = np.random.default_rng(91)
    sp = rng.standard_normal((8, 3, 3, R.sizeh)) + 1j * rng.standard_normal((8, 3, 3, R.sizeh))
    d = {"sigma_prime": sp, "rho": rho}
    lam, tmp = R.compute_lambda(sp, u, ta)
    d["lambda"] = lam; d["lambda_tmp"] = tmp
    sg = sp.copy(); tmp2 = R.compute_sigma(lam, u, sg, ta, rho)
    d["sigma"] = sg; d["sigma_tmp"] = tmp2
    uf, taf, spf = u.astype(np.complex64), gs["tipdot_f"], sp.astype(np.complex64)
    lamf, _ = R.compute_lambda(spf, uf, taf); d["lambda_f"] = lamf
    sgf = spf.copy(); R.compute_sigma(lamf, uf, sgf, taf, rho); d["sigma_f"] = sgf
    np.savez_compressed(os.path.join(HERE, "ref_stoutforce_%dx%dx%dx%d_r1.npz" % n), **d)
    print("stoutforce", n, float(np.abs(lam).max()), float(np.abs(sg).max()))


def io_single(n=(4, 4, 4, 4)):
    """On-disk formats (io.c): the reference writes the 4^4 fixture's links as ASCII and as ILDG, and reads both
    files back; the file bytes and the read-back arrays are the fixture."""
    import tempfile
    R = RefLib(*n)
    g = dict(np.load(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n)))
    u = np.ascontiguousarray(g["u"])
    R.lib.ref_set_io(C.c_double(3.7), b"# embedded input file\nnx 4\n")
    d = {"beta": 3.7, "conf_id": 7, "input_file": "# embedded input file\nnx 4\n"}
    with tempfile.TemporaryDirectory() as td:
        pa, pi = os.path.join(td, "conf.ascii").encode(), os.path.join(td, "conf.ildg").encode()
        assert R.lib.print_su3_soa_ASCII(ptr(u), pa, C.c_int(7), None) == 0
        assert R.lib.print_su3_soa_ildg_binary(ptr(u), pi, C.c_int(7)) == 0
        d["ascii_bytes"] = np.frombuffer(open(pa, "rb").read(), dtype=np.uint8)
        d["ildg_bytes"] = np.frombuffer(open(pi, "rb").read(), dtype=np.uint8)
        back = np.zeros_like(u); cid = C.c_int(0)
        assert R.lib.read_su3_soa_ASCII(ptr(back), pa, C.byref(cid)) == 0
        d["ascii_read_back"] = back.copy(); d["ascii_conf_id"] = cid.value
        back = np.zeros_like(u); cid = C.c_int(0)
        assert R.lib.read_su3_soa_ildg_binary(ptr(back), pi, C.byref(cid)) == 0
        d["ildg_read_back"] = back.copy(); d["ildg_conf_id"] = cid.value
    np.savez_compressed(os.path.join(HERE, "ref_io_%dx%dx%dx%d_r1.npz" % n), **d)
    print("io", n, d["ascii_bytes"].size, d["ildg_bytes"].size, cid.value)


# flavours of the whole-force fixture: light (charge 2, two pseudofermions) and heavy (charge -1, one)
FORCE_FLAVOURS = [dict(mass=0.0507, charge=2.0, number_of_ps=2, first_ps=0, ra_a=[0.11, 0.23, 0.37], ra_b=[0.003, 0.04, 0.5]),
                  dict(mass=0.12, charge=-1.0, number_of_ps=1, first_ps=2, ra_a=[0.3, -0.2], ra_b=[0.01, 0.2])]


def callers_single(n=(4, 4, 4, 4)):
    """The two callers of the path on the 4^4 fixture: fermion_force_soloopenacc (fermion_force.c:166-357; two flavours,
    three pseudofermions, rho = 0.15, two / zero stout levels, FP64 and the _f twin), eo_inversion (Meas/ferm_meas.c:50-72)
    and the operator with a field (field_times_fermion_matrix.c)."""
    R = RefLib(*n)
    g = dict(np.load(os.path.join(HERE, "ref_%dx%dx%dx%d_r1.npz" % n)))
    u = g["u"]
    fin = np.stack([g["v"], g["w"], gaussian_vec(R.sizeh, 14)])
    d = {"ferm_in": fin, "rho": 0.15, "res": 1e-10, "res_f": 1e-5}
    fl = []; flf = []
    for i, f in enumerate(FORCE_FLAVOURS):
        ph = R.phases(EB, MU, f["charge"]); phf = R.phases_f(EB, MU, f["charge"])
        d["ph%d" % i] = ph; d["phf%d" % i] = phf
        for k in ("mass", "number_of_ps", "first_ps", "ra_a", "ra_b"):
            d["fl%d_%s" % (i, k)] = np.array(f[k])
        fl.append(dict(f, ph=ph)); flf.append(dict(f, ph=phf))
    for steps in (2, 0):
        ipdot, gl3, stout = R.fermion_force(u, fl, fin, 1e-10, 5000, 0.15, steps)
        d["ipdot_s%d" % steps] = ipdot; d["gl3_s%d" % steps] = gl3
    ipdot, gl3, _ = R.fermion_force(u.astype(np.complex64), flf, fin.astype(np.complex64), 1e-5, 5000, 0.15, 2)
    d["ipdot_s2_f"] = ipdot
    oe, oo = R.eo_inversion(u, d["ph0"], MASS, g["v"], g["w"], 1e-10, 5000)
    d["eo_out_e"] = oe; d["eo_out_o"] = oo
    rng = np.random.default_rng(15)
    fre, fim = rng.standard_normal((8, R.sizeh)), rng.standard_normal((8, R.sizeh))
    d["field_re"] = fre; d["field_im"] = fim
    d["deo_wf"] = R.dslash_wf("acc_Deo_wf", u, g["v"], d["ph0"], fre, fim)
    d["doe_wf"] = R.dslash_wf("acc_Doe_wf", u, g["v"], d["ph0"], fre, fim)
    np.savez_compressed(os.path.join(HERE, "ref_callers_%dx%dx%dx%d_r1.npz" % n), **d)
    print("callers", n, float(np.abs(d["ipdot_s2"]).max()), float(np.abs(d["ipdot_s0"]).max()), float(np.abs(oe).max()))


# headers of the reference that declare the hot path and its callers -- generated sp_* twins included
PROTO_HEADERS = ["OpenAcc/fermion_matrix.h", "OpenAcc/sp_fermion_matrix.h", "OpenAcc/fermionic_utilities.h", "OpenAcc/sp_fermionic_utilities.h",
                 "OpenAcc/inverter_multishift_full.h", "OpenAcc/sp_inverter_multishift_full.h", "OpenAcc/inverter_full.h",
                 "OpenAcc/sp_inverter_full.h", "OpenAcc/inverter_mixedp.h", "OpenAcc/inverter_wrappers.h", "OpenAcc/inverter_package.h",
                 "OpenAcc/float_double_conv.h", "OpenAcc/find_min_max.h", "OpenAcc/fermion_force_utilities.h",
                 "OpenAcc/sp_fermion_force_utilities.h", "OpenAcc/fermion_force.h", "OpenAcc/sp_fermion_force.h", "OpenAcc/stouting.h",
                 "OpenAcc/sp_stouting.h", "OpenAcc/plaquettes.h", "OpenAcc/sp_plaquettes.h", "OpenAcc/su3_utilities.h",
                 "OpenAcc/sp_su3_utilities.h", "OpenAcc/field_times_fermion_matrix.h", "Mpi/communications.h", "Mpi/sp_communications.h",
                 "Mpi/multidev.h", "Meas/ferm_meas.h"]


def reference_prototypes():
    """Every function prototype the reference's own headers declare for the path (preprocessed with the build's -D flags,
    MULTIDEVICE on), normalised by tests/prototypes.py -> tests/golden/ref_prototypes.json.  Pins the C ABI: names,
    argument order, argument and return types (tests/test_abi_and_host.py compares include/staple_b200.h against it)."""
    import json
    sys.path.insert(0, os.path.join(HERE, ".."))
    from prototypes import preprocess, prototypes
    from oracle.pyoracle import build_ref
    build_ref(4, 4, 4, 4, 2)                                  # makes sure the scratch copy with the generated sp_* headers exists
    scr = os.path.join(os.environ.get("STAPLE_ORACLE_SCRATCH", os.path.join(os.environ.get("TMPDIR", "/tmp"), "staple_oracle_src")), "src")
    flags = ["-fcommon", "-I" + os.path.join(HERE, "..", "..", "oracle", "mpi_stub"), "-I" + scr, "-DACTION_TYPE=TLSM", "-DNREPLICAS=1",
             "-DLOC_N0=4", "-DLOC_N1=4", "-DLOC_N2=4", "-DLOC_N3=4", "-DNRANKS_D3=2", "-DCOMMIT_HASH=oracle"]
    flags += ["-D%s%s=8" % (a, b) for a in ("DEODOE", "IMPSTAP", "STAP", "SIGMA") for b in ("TILE0", "TILE1", "TILE2", "GANG3")]
    per_header = {}
    for h in PROTO_HEADERS:
        names = set(re.findall(r"\b([A-Za-z_]\w*)\s*\(", open(os.path.join(scr, h)).read()))
        pr = prototypes(preprocess('#include "%s"\n' % h, flags))
        per_header[h] = {k: v for k, v in sorted(pr.items()) if k in names and not k.startswith(("MPI_", "__"))}
    json.dump(per_header, open(os.path.join(HERE, "ref_prototypes.json"), "w"), indent=0, sort_keys=True)
    print("prototypes", sum(len(v) for v in per_header.values()))


def border_fields(sizeh, rank):
    """deterministic local boxes for the border-exchange fixture: every element names its rank, array and site"""
    idx = np.arange(sizeh, dtype=np.float64)
    conf = np.zeros((8, 3, 3, sizeh), np.complex128)
    for k in range(8):
        for r in range(3):
            for c in range(3):
                conf[k, r, c] = (1000.0 * rank + 100 * k + 10 * r + c) + 1e-4 * idx + 1j * (idx + 0.5 * rank)
    ta = np.zeros((8, 8, sizeh), np.float64)
    for k in range(8):
        for j in range(8):
            ta[k, j] = 7000.0 * rank + 100 * k + j + 1e-4 * idx
    return conf, ta


def borders_multi(loc=(4, 4, 4, 4), nr=2):
    """the reference's own border exchanges (communications.c) between two ranks run in one process through the mailbox MPI of
    oracle/ref_shim.c: su3 (thickness 2 and 1), gl3, tamat, thmat (thickness 1)"""
    R = RefLib(*loc, nr)
    d = {"loc_n": np.array(loc), "nranks": nr, "sizeh": R.sizeh}

    def exchange(fn, arrays, *extra):
        R.lib.ref_mailbox_clear()
        for _ in range(2):
            for r in range(nr):
                R.set_rank(r); getattr(R.lib, fn)(ptr(arrays[r]), *extra)
        return arrays
    for name, fn, which, extra in (("su3_t2", "communicate_su3_borders", 0, (C.c_int(2),)), ("su3_t1", "communicate_su3_borders", 0, (C.c_int(1),)),
                                   ("gl3_t1", "communicate_gl3_borders", 0, (C.c_int(1),)), ("tamat_t1", "communicate_tamat_soa_borders", 1, (C.c_int(1),)),
                                   ("thmat_t1", "communicate_thmat_soa_borders", 1, (C.c_int(1),))):
        arrays = [border_fields(R.sizeh, r)[which].copy() for r in range(nr)]
        exchange(fn, arrays, *extra)
        for r in range(nr):
            d["%s_r%d" % (name, r)] = arrays[r]
    np.savez_compressed(os.path.join(HERE, "ref_borders_%dx%dx%dx%d_r%d.npz" % (loc + (nr,))), **d)
    print("borders", loc, nr, R.sizeh)


class _RA(C.Structure):      # RationalApprox/rationalapprox.h:15-26 (layout checked by ref_abi below)
    _fields_ = [("exponent_num", C.c_int), ("exponent_den", C.c_int), ("approx_order", C.c_int),
                ("lambda_min", C.c_double), ("lambda_max", C.c_double), ("gmp_remez_precision", C.c_int),
                ("error", C.c_double), ("RA_a0", C.c_double), ("RA_a", C.c_double * 25), ("RA_b", C.c_double * 25)]


def _ra_dict(tag, r):
    return {tag + "_num": r.exponent_num, tag + "_den": r.exponent_den, tag + "_order": r.approx_order,
            tag + "_lmin": r.lambda_min, tag + "_lmax": r.lambda_max, tag + "_prec": r.gmp_remez_precision,
            tag + "_error": r.error, tag + "_a0": r.RA_a0, tag + "_a": np.array(r.RA_a[:r.approx_order]),
            tag + "_b": np.array(r.RA_b[:r.approx_order])}


def abi_and_approx():
    """Struct layouts as the reference's headers define them (ref_abi), the shipped order-19 rational
    approximations as the reference's own reader parses them (rationalapprox.c:83-117) and
    rescale_rational_approximation (:145-194) applied to them."""
    R = RefLib(4, 4, 4, 4)
    o = (C.c_long * 24)(); R.lib.ref_abi(o)
    o2 = (C.c_long * 16)(); R.lib.ref_abi2(o2)
    d = {"abi": np.array(list(o)), "abi_sizeh": R.sizeh, "abi2": np.array(list(o2))}
    ref = os.environ.get("STAPLE_REFERENCE", "/root/reference")
    files = {"m14": "saved_approxs/approx_-1_over_4_order_19_mloglm_6.4.REMEZ",
             "p18": "saved_approxs/approx_1_over_8_order_19_mloglm_6.4.REMEZ",
             "m14o9": "saved_approxs/approx_-1_over_4_order_9_mloglm_6.4.REMEZ"}
    minmax = (C.c_double * 2)(0.0507 ** 2, 5.2)
    for tag, f in files.items():
        r = _RA(); assert C.sizeof(_RA) == o[11]
        R.lib.ref_approx_read(C.byref(r), os.path.join(ref, f).encode())
        d.update(_ra_dict(tag, r))
        out = _RA()
        R.lib.rescale_rational_approximation(C.byref(r), C.byref(out), minmax)
        d.update(_ra_dict(tag + "_rescaled", out))
    d["rescale_minmax"] = np.array(list(minmax))
    np.savez_compressed(os.path.join(HERE, "ref_abi_approx.npz"), **d)
    print("abi", list(o))


if __name__ == "__main__":
    which = sys.argv[1:] or ["single", "multi", "abi", "force", "stout", "stoutforce", "io", "callers", "prototypes", "borders"]
    if "single" in which:
        single_rank()
    if "multi" in which:
        multi_rank()
    if "abi" in which:
        abi_and_approx()
    if "force" in which:
        force_single()
    if "stout" in which:
        stout_single()
    if "stoutforce" in which:
        stoutforce_single()
    if "io" in which:
        io_single()
    if "callers" in which:
        callers_single()
    if "prototypes" in which:
        reference_prototypes()
    if "borders" in which:
        borders_multi()

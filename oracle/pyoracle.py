"""TEST INFRASTRUCTURE ONLY -- ctypes access to the two CPU oracles.

* ``RefLib``      : the UNMODIFIED reference sources compiled by gcc for one compile-time
                    geometry (oracle/_ref/libref_<geom>.so, see oracle/build_ref.sh).
* ``Restatement`` : our plain-C restatement (oracle/staggered_oracle.c), run-time geometry.

numpy layouts equal the reference ABI (struct_c_def.h:16-42):
  vec3_soa   -> complex128[3, sizeh]        su3_soa[8] -> complex128[8, 3, 3, sizeh]
  double_soa[8] -> float64[8, sizeh]        vec3_soa[N] -> complex128[N, 3, sizeh]
(complex64 / float32 for the ``_f`` twins).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("STAPLE_REFERENCE", "/root/reference")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- inputs
def random_su3_conf(sizeh, seed, dtype=np.complex128):
    """i.i.d. Haar-random SU(3) for every link: complex[8,3,3,sizeh] (row 2 = conj(r0 x r1))."""
    rng = np.random.default_rng(seed)
    n = 8 * sizeh
    z = rng.standard_normal((n, 3, 3)) + 1j * rng.standard_normal((n, 3, 3))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r, axis1=1, axis2=2)
    q = q * (d / np.abs(d))[:, None, :]
    det = np.linalg.det(q)
    q = q / (det ** (1.0 / 3.0))[:, None, None]
    q[:, 2, :] = np.conj(np.cross(q[:, 0, :], q[:, 1, :]))
    u = q.reshape(8, sizeh, 3, 3).transpose(0, 2, 3, 1)
    return np.ascontiguousarray(u).astype(dtype)


def gaussian_vec(sizeh, seed, n=None, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    shape = (3, sizeh) if n is None else (n, 3, sizeh)
    v = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)
    return np.ascontiguousarray(v).astype(dtype)


# ----------------------------------------------------------------------------- reference build
def ref_lib_path(n0, n1, n2, n3, nr=1):
    return os.path.join(HERE, "_ref", "libref_%dx%dx%dx%d_r%d.so" % (n0, n1, n2, n3, nr))


def build_ref(n0, n1, n2, n3, nr=1):
    """Compile the reference for one geometry if /root/reference is present; returns path or None."""
    path = ref_lib_path(n0, n1, n2, n3, nr)
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.run([os.path.join(HERE, "build_ref.sh")] + [str(x) for x in (n0, n1, n2, n3, nr)],
                       check=True, stdout=subprocess.DEVNULL)
    return path if os.path.exists(path) else None


def have_ref(n0, n1, n2, n3, nr=1):
    return os.path.exists(ref_lib_path(n0, n1, n2, n3, nr)) or os.path.isdir(os.path.join(REFERENCE, "src"))


class RefLib:
    """The reference's own functions for one (LOC_N0..3, NRANKS_D3) geometry."""

    def __init__(self, n0, n1, n2, n3, nr=1, rank=0):
        path = build_ref(n0, n1, n2, n3, nr)
        if path is None:
            raise FileNotFoundError("no reference build for this geometry: " + ref_lib_path(n0, n1, n2, n3, nr))
        self.lib = C.CDLL(path)
        o = (C.c_int * 12)()
        self.lib.ref_geometry(o)
        self.loc_n = tuple(o[0:4]); self.nranks = o[4]; self.nd = tuple(o[5:9])
        self.sizeh = o[9]; self.d3_halo = o[10]; self.gl_sizeh = o[11]
        self.vol3h = self.nd[0] * self.nd[1] * self.nd[2] // 2
        L = self.lib
        L.ref_ferm_param_new.restype = C.c_void_p
        L.ref_ferm_param_new.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        L.ref_approx_new.restype = C.c_void_p
        L.ref_approx_new.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.l2norm2_global.restype = C.c_double
        L.l2norm2_global_f.restype = C.c_double
        L.real_scal_prod_global.restype = C.c_double
        L.real_scal_prod_global_f.restype = C.c_double
        L.ker_find_max_eigenvalue_openacc.restype = C.c_double
        self.set_rank(rank)
        self.set_inverter_tricks(0, 0, 0.1, 10000)

    def set_rank(self, rank):
        if self.lib.ref_setup(C.c_int(rank)) != 0:
            raise RuntimeError("reference set_geom_glv failed")
        self.rank = rank

    def set_inverter_tricks(self, sp_accel, mixed, delta, restart):
        self.lib.ref_set_inverter_tricks(C.c_int(sp_accel), C.c_int(mixed), C.c_double(delta), C.c_int(restart))

    # --- inputs
    def phases(self, eb=(0, 0, 0, 0, 0, 0), im_chem_pot=0.0, charge=0.0):
        ph = np.zeros((8, self.sizeh), np.float64)
        self.lib.ref_phases(ptr(ph), *[C.c_double(x) for x in eb], C.c_double(im_chem_pot), C.c_double(charge))
        return ph

    def phases_f(self, eb=(0, 0, 0, 0, 0, 0), im_chem_pot=0.0, charge=0.0):
        ph = np.zeros((8, self.sizeh), np.float32)
        self.lib.ref_phases_f(ptr(ph), *[C.c_double(x) for x in eb], C.c_double(im_chem_pot), C.c_double(charge))
        return ph

    def ferm_param(self, mass, ph, ph_f=None):
        return C.c_void_p(self.lib.ref_ferm_param_new(C.c_double(mass), ptr(ph), ptr(ph_f)))

    def approx(self, a0, a, b):
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
        return C.c_void_p(self.lib.ref_approx_new(C.c_int(len(b)), C.c_double(a0), ptr(a), ptr(b)))

    def _sfx(self, a):
        return "_f" if a.dtype in (np.complex64, np.float32) else ""

    # --- operator
    def dslash(self, name, u, inp, ph, out=None):
        """name in acc_Deo, acc_Doe, acc_Deo_unsafe, acc_Doe_bulk, acc_Deo_d3p, ..."""
        out = np.zeros_like(inp) if out is None else out
        getattr(self.lib, name + self._sfx(inp))(ptr(u), ptr(out), ptr(inp), ptr(ph))
        return out

    def mdagm(self, u, inp, ph, mass, shift=None, ph_f=None):
        out = np.zeros_like(inp); tmp = np.zeros_like(inp)
        sfx = self._sfx(inp)
        pars = self.ferm_param(mass, ph if sfx == "" else None, ph if sfx else None)
        if shift is None:
            getattr(self.lib, "fermion_matrix_multiplication" + sfx)(ptr(u), ptr(out), ptr(inp), ptr(tmp), pars)
        else:
            getattr(self.lib, "fermion_matrix_multiplication_shifted" + sfx)(
                ptr(u), ptr(out), ptr(inp), ptr(tmp), pars, C.c_double(shift))
        return out

    # --- solvers
    def multishift_invert(self, u, ph, mass, approx_a0_a_b, inp, residuo, max_cg):
        a0, a, b = approx_a0_a_b
        n = len(b); sfx = self._sfx(inp)
        out = np.zeros((n,) + inp.shape, inp.dtype); ps = np.zeros_like(out)
        r, h, s, p = (np.zeros_like(inp) for _ in range(4))
        pars = self.ferm_param(mass, ph if sfx == "" else None, ph if sfx else None)
        cg = C.c_int(0)
        ok = getattr(self.lib, "multishift_invert" + sfx)(
            ptr(u), pars, self.approx(a0, a, b), ptr(out), ptr(inp), C.c_double(residuo),
            ptr(r), ptr(h), ptr(s), ptr(p), ptr(ps), C.c_int(max_cg), C.byref(cg))
        return out, cg.value, ok

    def recombine(self, shifted, inp, approx_a0_a_b):
        a0, a, b = approx_a0_a_b
        out = np.zeros_like(inp)
        getattr(self.lib, "recombine_shifted_vec3_to_vec3" + self._sfx(inp))(
            ptr(shifted), ptr(inp), ptr(out), self.approx(a0, a, b))
        return out

    def cg(self, u, ph, mass, inp, res, max_cg, shift, guess=None):
        sfx = self._sfx(inp)
        sol = np.zeros_like(inp) if guess is None else guess.copy()
        r, h, s, p = (np.zeros_like(inp) for _ in range(4))
        pars = self.ferm_param(mass, ph if sfx == "" else None, ph if sfx else None)
        cg = C.c_int(0)
        ok = getattr(self.lib, "ker_invert_openacc" + sfx)(
            ptr(u), pars, ptr(sol), ptr(inp), C.c_double(res), ptr(r), ptr(h), ptr(s), ptr(p),
            C.c_int(max_cg), C.c_double(shift), C.byref(cg))
        return sol, cg.value, ok

    def mixed_cg(self, u, u_f, ph, ph_f, mass, inp, res, max_cg, shift, guess=None):
        sol = np.zeros_like(inp) if guess is None else guess.copy()
        d = [np.zeros_like(inp) for _ in range(4)]
        f = [np.zeros(inp.shape, np.complex64) for _ in range(5)]
        st = np.zeros((1,) + inp.shape, inp.dtype); st_f = np.zeros((1,) + inp.shape, np.complex64)
        self.lib.ref_ip_dp(ptr(u), ptr(st), C.c_int(1), *[ptr(x) for x in d])
        self.lib.ref_ip_sp(ptr(u_f), ptr(st_f), C.c_int(1), *[ptr(x) for x in f])
        pars = self.ferm_param(mass, ph, ph_f)
        cg = C.c_int(0)
        ok = self.lib.ref_inverter_mixed_precision(pars, ptr(sol), ptr(inp), C.c_double(res), C.c_int(max_cg),
                                                   C.c_double(shift), C.byref(cg))
        return sol, cg.value, ok

    def max_eigenvalue(self, u, ph, mass, start):
        p = start.copy(); r = np.zeros_like(p); h = np.zeros_like(p)
        pars = self.ferm_param(mass, ph)
        return self.lib.ker_find_max_eigenvalue_openacc(ptr(u), pars, ptr(r), ptr(h), ptr(p))

    def min_eigenvalue(self, u, ph, mass, start, mx):
        p = start.copy(); r = np.zeros_like(p); h = np.zeros_like(p)
        self.lib.ker_find_min_eigenvalue_openacc.restype = C.c_double
        return self.lib.ker_find_min_eigenvalue_openacc(ptr(u), self.ferm_param(mass, ph), ptr(r), ptr(h), ptr(p), C.c_double(mx))

    # --- fermion-force outer products (fermion_force_utilities.c)
    def compute_fermion_force(self, u, aux, shiftmulti, ph, ra_a):
        """ker_openacc_compute_fermion_force: aux (gl3 field [8,3,3,sizeh]) accumulated in place; -> (loc_s, loc_h)."""
        sfx = self._sfx(u)
        ra_a = np.ascontiguousarray(ra_a, np.float64)
        pars = self.ferm_param(0.1, ph if sfx == "" else None, ph if sfx else None)
        self.lib.ref_ferm_param_set_md(pars, C.c_int(len(ra_a)), ptr(ra_a), ptr(np.zeros_like(ra_a)))
        s = np.zeros_like(shiftmulti[0]); h = np.zeros_like(s)
        getattr(self.lib, "ker_openacc_compute_fermion_force" + sfx)(ptr(u), ptr(aux), ptr(shiftmulti), ptr(s), ptr(h), pars)
        return s, h

    def direct_product(self, s, h, aux, a):
        ap = self.approx(1.0, [a], [0.0])
        getattr(self.lib, "direct_product_of_fermions_into_auxmat" + self._sfx(s))(ptr(s), ptr(h), ptr(aux), ap, C.c_int(0))

    def multiply_backfield_times_force(self, ph, aux, pseudo):
        sfx = self._sfx(aux)
        pars = self.ferm_param(0.1, ph if sfx == "" else None, ph if sfx else None)
        getattr(self.lib, "multiply_backfield_times_force" + sfx)(pars, ptr(aux), ptr(pseudo))

    def accumulate_gl3(self, aux, pseudo):
        getattr(self.lib, "accumulate_gl3soa_into_gl3soa" + self._sfx(aux))(ptr(aux), ptr(pseudo))

    def take_ta(self, u, aux, ta):
        getattr(self.lib, "multiply_conf_times_force_and_take_ta_nophase" + self._sfx(u))(ptr(u), ptr(aux), ptr(ta))

    # --- isotropic stout smearing (stouting.c, plaquettes.c:196-255, su3_utilities.c:210-237, cayley_hamilton.h)
    def _stout_globals(self, rho, steps, like):
        cd = like.dtype; rd = np.float32 if cd == np.complex64 else np.float64
        self._stout_keep = [np.zeros((8, 3, 3, self.sizeh), cd), np.zeros((8, 3, 3, self.sizeh), cd), np.zeros((8, 8, self.sizeh), rd)]
        a = [ptr(x) for x in self._stout_keep]
        if cd == np.complex64:
            self.lib.ref_set_stout(C.c_double(rho), C.c_int(steps), None, None, None, *a)
        else:
            self.lib.ref_set_stout(C.c_double(rho), C.c_int(steps), *a, None, None, None)

    def stout_isotropic(self, u, rho):
        """-> (uprime [rows 0,1 written], staples, exp_aux, tipdot) exactly as stout_isotropic leaves them"""
        self._stout_globals(rho, 1, u)
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        up = np.zeros_like(u); stap = np.zeros_like(u); aux = np.zeros_like(u); ta = np.zeros((8, 8, self.sizeh), rd)
        getattr(self.lib, "stout_isotropic" + self._sfx(u))(ptr(u), ptr(up), ptr(stap), ptr(aux), ptr(ta), C.c_int(0))
        return up, stap, aux, ta

    def stout_wrapper(self, u, rho, steps):
        """-> stout_conf_acc_arr [steps, 8, 3, 3, sizeh] (single rank: no border exchange involved)"""
        self._stout_globals(rho, steps, u)
        out = np.zeros((steps,) + u.shape, u.dtype)
        getattr(self.lib, "stout_wrapper" + self._sfx(u))(ptr(u), ptr(out), C.c_int(0))
        return out

    # --- stout force chain (stouting.c:171-1305)
    def compute_lambda(self, sp, u, ta):
        """-> (Lambda thmat [8,8,sizeh], TMP)"""
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        lam = np.zeros((8, 8, self.sizeh), rd); tmp = np.zeros_like(u)
        getattr(self.lib, "compute_lambda" + self._sfx(u))(ptr(lam), ptr(sp), ptr(u), ptr(ta), ptr(tmp))
        return lam, tmp

    def compute_sigma(self, lam, u, sg, ta, rho):
        """Sigma' (sg) -> Sigma in place; -> TMP"""
        self._stout_globals(rho, 1, u)
        tmp = np.zeros_like(u)
        getattr(self.lib, "compute_sigma" + self._sfx(u))(ptr(lam), ptr(u), ptr(sg), ptr(ta), ptr(tmp), C.c_int(0))
        return tmp

    # --- the callers of the path: whole fermion force (fermion_force.c:166-357), eo_inversion (Meas/ferm_meas.c:50-72)
    def _package(self, u, nshifts, single_too=False, u_f=None):
        """inverter_package with scratch vectors (kept alive on self); -> None (the shim holds the package)"""
        n = self.sizeh
        d = [np.zeros((3, n), np.complex128) for _ in range(4)]
        st = np.zeros((max(nshifts, 1), 3, n), np.complex128)
        self._ip_keep = [d, st]
        self.lib.ref_ip_dp(ptr(u), ptr(st), C.c_int(nshifts), *[ptr(x) for x in d])
        f = [np.zeros((3, n), np.complex64) for _ in range(5)]
        st_f = np.zeros((max(nshifts, 1), 3, n), np.complex64)
        self._ip_keep += [f, st_f, u_f]
        self.lib.ref_ip_sp(ptr(u_f), ptr(st_f), C.c_int(nshifts), *[ptr(x) for x in f])

    def _flavours(self, flavours, single):
        """flavours: list of dict(mass, ph, number_of_ps, first_ps, ra_a, ra_b) -> ferm_param[nflav]"""
        self.lib.ref_ferm_param_array_new.restype = C.c_void_p
        arr = C.c_void_p(self.lib.ref_ferm_param_array_new(C.c_int(len(flavours))))
        self._fl_keep = []
        for i, fl in enumerate(flavours):
            a = np.ascontiguousarray(fl["ra_a"], np.float64); b = np.ascontiguousarray(fl["ra_b"], np.float64)
            ph = np.ascontiguousarray(fl["ph"], np.float32 if single else np.float64)
            self._fl_keep += [a, b, ph]
            self.lib.ref_ferm_param_array_set(arr, C.c_int(i), C.c_double(fl["mass"]), None if single else ptr(ph),
                                              ptr(ph) if single else None, C.c_int(fl["number_of_ps"]),
                                              C.c_int(fl["first_ps"]), C.c_int(len(a)), ptr(a), ptr(b))
        return arr

    def fermion_force(self, u, flavours, ferm_in, res, max_cg, rho, steps):
        """fermion_force_soloopenacc[_f] (dtype of u selects): -> (ipdot tamat [8,8,sizeh], gl3_aux, stout levels)"""
        single = u.dtype == np.complex64
        cd = u.dtype; rd = np.float32 if single else np.float64
        n = self.sizeh
        self._stout_globals(rho, steps, u)
        th, ta = np.zeros((8, 8, n), rd), np.zeros((8, 8, n), rd)
        a = (None, None, ptr(th), ptr(ta), None) if single else (ptr(th), ptr(ta), None, None, None)
        self.lib.ref_set_force_globals(*a)
        nsh = max(len(fl["ra_b"]) for fl in flavours)
        if single:
            self._package(np.zeros_like(u, dtype=np.complex128), nsh, u_f=u)
        else:
            self._package(u, nsh)
        pars = self._flavours(flavours, single)
        stout = np.zeros((max(steps, 1),) + u.shape, cd)
        gl3, taux = np.zeros_like(u), np.zeros_like(u)
        ipdot = np.zeros((8, 8, n), rd)
        shiftmulti = np.zeros((nsh, 3, n), cd)
        ferm_in = np.ascontiguousarray(ferm_in, cd)
        fn = self.lib.ref_fermion_force_f if single else self.lib.ref_fermion_force
        fn(ptr(u), ptr(stout), ptr(gl3), ptr(ipdot), pars, C.c_int(len(flavours)), ptr(ferm_in), C.c_double(res),
           ptr(taux), ptr(shiftmulti), C.c_int(max_cg))
        return ipdot, gl3, stout

    def eo_inversion(self, u, ph, mass, in_e, in_o, res, max_cg):
        """-> (out_e, out_o): (D + m)(out_e, out_o) = (in_e, in_o)"""
        self._package(u, 1)
        pars = self.ferm_param(mass, ph)
        out_e, out_o, phi_e, phi_o = (np.zeros_like(in_e) for _ in range(4))
        self.lib.ref_eo_inversion(pars, C.c_double(res), C.c_int(max_cg), ptr(in_e), ptr(in_o), ptr(out_e), ptr(out_o),
                                  ptr(phi_e), ptr(phi_o))
        return out_e, out_o

    def dslash_wf(self, name, u, inp, ph, fre, fim):
        """name in acc_Deo_wf, acc_Doe_wf, acc_Deo_wf_unsafe, acc_Doe_wf_unsafe (field_times_fermion_matrix.c)"""
        out = np.zeros_like(inp)
        getattr(self.lib, name)(ptr(u), ptr(out), ptr(inp), ptr(ph), ptr(fre), ptr(fim))
        return out

    # --- reductions
    def l2norm2(self, a):
        return getattr(self.lib, "l2norm2_global" + self._sfx(a))(ptr(a))

    def real_scal_prod(self, a, b):
        return getattr(self.lib, "real_scal_prod_global" + self._sfx(a))(ptr(a), ptr(b))


# ----------------------------------------------------------------------------- restatement
class SoGeom(C.Structure):
    _fields_ = [("loc_n", C.c_int * 4), ("nranks_d3", C.c_int), ("halo_width", C.c_int),
                ("d3_halo", C.c_int), ("d3_fhalo", C.c_int), ("nd", C.c_int * 4),
                ("vol3h", C.c_long), ("sizeh", C.c_long), ("r0_lo", C.c_long), ("r0_hi", C.c_long),
                ("r1_lo", C.c_long), ("r1_hi", C.c_long), ("gl_n", C.c_int * 4)]


def build_restatement():
    path = os.path.join(HERE, "libstaggered_oracle.so")
    subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL)
    return path


class Restatement:
    """Plain-C restatement of the path (oracle/staggered_oracle.c) for one run-time geometry."""
    OPS = dict(in1xfactor_plus_in2=0, scale=1, add_factor_x_in2=2, in1xmass2_minus_in2_minus_in3=3,
               in1xmass_minus_in2=4, in1_minus_in2=5, assign=6, zero=7, fact1_minus_in2=8,
               in1_minus_in2_allxfact=9)

    def __init__(self, n0, n1, n2, n3, nr=1, halo_width=2):
        self.lib = C.CDLL(build_restatement())
        self.g = SoGeom()
        self.lib.so_geom_init(C.byref(self.g), n0, n1, n2, n3, nr, halo_width)
        self.sizeh = self.g.sizeh; self.vol3h = self.g.vol3h; self.nd = tuple(self.g.nd)
        self.loc_n = (n0, n1, n2, n3); self.nranks = nr; self.d3_halo = self.g.d3_halo
        L = self.lib
        for s in ("", "_f"):
            getattr(L, "so_l2norm2" + s).restype = C.c_double
            getattr(L, "so_real_scal_prod" + s).restype = C.c_double
        L.so_find_max_eigenvalue.restype = C.c_double
        L.so_snum.restype = C.c_long
        L.so_lnh_to_gl_snum.restype = C.c_long

    def _sfx(self, a):
        return "_f" if a.dtype in (np.complex64, np.float32) else ""

    def gp(self):
        return C.byref(self.g)

    def phases(self, rank=0, eb=(0, 0, 0, 0, 0, 0), im_chem_pot=0.0, charge=0.0, single=False):
        ph = np.zeros((8, self.sizeh), np.float32 if single else np.float64)
        e = (C.c_double * 6)(*eb)
        fn = self.lib.so_calc_u1_phases_f if single else self.lib.so_calc_u1_phases
        fn(self.gp(), C.c_int(rank), ptr(ph), e, C.c_double(im_chem_pot), C.c_double(charge))
        return ph

    def dslash(self, which, u, inp, ph, d3lo=None, d3hi=None, out=None):
        """which: 'deo' | 'doe'."""
        out = np.zeros_like(inp) if out is None else out
        d3lo = self.d3_halo if d3lo is None else d3lo
        d3hi = self.d3_halo + self.loc_n[3] if d3hi is None else d3hi
        getattr(self.lib, "so_" + which + self._sfx(inp))(self.gp(), ptr(u), ptr(out), ptr(inp), ptr(ph),
                                                         C.c_int(d3lo), C.c_int(d3hi))
        return out

    def mdagm(self, u, inp, ph, mass, shift=0.0):
        out = np.zeros_like(inp); tmp = np.zeros_like(inp)
        getattr(self.lib, "so_fermion_matrix_multiplication_shifted" + self._sfx(inp))(
            self.gp(), ptr(u), ptr(out), ptr(inp), ptr(tmp), ptr(ph), C.c_double(mass), C.c_double(shift))
        return out

    def axpy_like(self, op, out, a=None, b=None, c=None, f1=0.0):
        getattr(self.lib, "so_axpy_like" + self._sfx(out))(
            self.gp(), C.c_int(self.OPS[op]), ptr(out), ptr(a), ptr(b), ptr(c), C.c_double(f1), C.c_double(0))
        return out

    def l2norm2(self, a):
        return getattr(self.lib, "so_l2norm2" + self._sfx(a))(self.gp(), ptr(a))

    def real_scal_prod(self, a, b):
        return getattr(self.lib, "so_real_scal_prod" + self._sfx(a))(self.gp(), ptr(a), ptr(b))

    def multishift_invert(self, u, ph, mass, shifts, inp, residuo, max_cg):
        shifts = np.ascontiguousarray(shifts, np.float64); n = len(shifts)
        out = np.zeros((n,) + inp.shape, inp.dtype); ps = np.zeros_like(out)
        r, h, s, p = (np.zeros_like(inp) for _ in range(4))
        cg = C.c_int(0); rel = np.zeros(n)
        ok = getattr(self.lib, "so_multishift_invert" + self._sfx(inp))(
            self.gp(), ptr(u), ptr(ph), C.c_double(mass), C.c_int(n), ptr(shifts), ptr(out), ptr(inp),
            C.c_double(residuo), ptr(r), ptr(h), ptr(s), ptr(p), ptr(ps), C.c_int(max_cg), C.byref(cg), ptr(rel))
        return out, cg.value, ok, rel

    def recombine(self, shifted, inp, a0, a):
        a = np.ascontiguousarray(a, np.float64); out = np.zeros_like(inp)
        getattr(self.lib, "so_recombine" + self._sfx(inp))(self.gp(), ptr(shifted), ptr(inp), ptr(out),
                                                          C.c_int(len(a)), C.c_double(a0), ptr(a))
        return out

    def cg(self, u, ph, mass, inp, res, max_cg, shift, restarting_every=10000, guess=None):
        sol = np.zeros_like(inp) if guess is None else guess.copy()
        r, h, s, p = (np.zeros_like(inp) for _ in range(4))
        cg = C.c_int(0)
        ok = getattr(self.lib, "so_cg" + self._sfx(inp))(
            self.gp(), ptr(u), ptr(ph), C.c_double(mass), ptr(sol), ptr(inp), C.c_double(res), ptr(r), ptr(h),
            ptr(s), ptr(p), C.c_int(max_cg), C.c_double(shift), C.c_int(restarting_every), C.byref(cg))
        return sol, cg.value, ok

    def mixed_cg(self, u, u_f, ph, ph_f, mass, inp, res, max_cg, shift, mixed_delta=0.1, guess=None):
        sol = np.zeros_like(inp) if guess is None else guess.copy()
        d = [np.zeros_like(inp) for _ in range(3)]
        f = [np.zeros(inp.shape, np.complex64) for _ in range(5)]
        cg = C.c_int(0); mt = C.c_int(0)
        ok = self.lib.so_inverter_mixed_precision(
            self.gp(), ptr(u), ptr(u_f), ptr(ph), ptr(ph_f), C.c_double(mass), ptr(sol), ptr(inp),
            C.c_double(res), C.c_int(max_cg), C.c_double(shift), C.c_double(mixed_delta),
            *[ptr(x) for x in d], *[ptr(x) for x in f], C.byref(cg), C.byref(mt))
        return sol, cg.value, ok, mt.value

    def max_eigenvalue(self, u, ph, mass, start):
        p = start.copy(); r = np.zeros_like(p); h = np.zeros_like(p)
        return self.lib.so_find_max_eigenvalue(self.gp(), ptr(u), ptr(ph), C.c_double(mass), ptr(r), ptr(h), ptr(p), None)

    def min_eigenvalue(self, u, ph, mass, start, mx):
        """ker_find_min_eigenvalue_openacc (find_min_max.c:62-98): power iteration on max - M^+M from `start`"""
        p = start.copy(); m2 = mass * mass; delta = mx - m2
        norm = np.sqrt(self.l2norm2(p))
        while True:
            self.axpy_like("scale", p, f1=1.0 / norm)
            r = p.copy(); old = norm
            p = self.mdagm(u, r, ph, mass, delta - m2)
            norm = np.sqrt(self.l2norm2(p))
            if abs(old - norm) / norm <= 1.0e-5:
                return mx - norm

    # fermion-force outer products (tamat_soa[8] as reals [8, 8, sizeh], see staggered_oracle_impl.h)
    def compute_fermion_force(self, u, aux, shiftmulti, ph, ra_a):
        ra_a = np.ascontiguousarray(ra_a, np.float64)
        s = np.zeros_like(shiftmulti[0]); h = np.zeros_like(s)
        getattr(self.lib, "so_compute_fermion_force" + self._sfx(u))(
            self.gp(), ptr(u), ptr(aux), ptr(shiftmulti), ptr(s), ptr(h), ptr(ph), C.c_int(len(ra_a)), ptr(ra_a))
        return s, h

    def direct_product(self, s, h, aux, a):
        getattr(self.lib, "so_direct_product_of_fermions_into_auxmat" + self._sfx(s))(
            self.gp(), ptr(s), ptr(h), ptr(aux), C.c_double(a))

    def multiply_backfield_times_force(self, ph, aux, pseudo):
        getattr(self.lib, "so_multiply_backfield_times_force" + self._sfx(aux))(self.gp(), ptr(ph), ptr(aux), ptr(pseudo))

    def accumulate_gl3(self, aux, pseudo):
        getattr(self.lib, "so_accumulate_gl3soa_into_gl3soa" + self._sfx(aux))(self.gp(), ptr(aux), ptr(pseudo))

    def take_ta(self, u, aux, ta):
        getattr(self.lib, "so_multiply_conf_times_force_and_take_ta_nophase" + self._sfx(u))(
            self.gp(), ptr(u), ptr(aux), ptr(ta))

    # isotropic stout smearing
    def stout_isotropic(self, u, rho):
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        up = np.zeros_like(u); stap = np.zeros_like(u); aux = np.zeros_like(u); ta = np.zeros((8, 8, self.sizeh), rd)
        getattr(self.lib, "so_stout_isotropic" + self._sfx(u))(self.gp(), ptr(u), ptr(up), ptr(stap), ptr(aux), ptr(ta), C.c_double(rho))
        return up, stap, aux, ta

    def stout_wrapper(self, u, rho, steps):
        """stouting.c:27-72 for a single rank: level l smears level l-1 (level 0 smears u)"""
        out = np.zeros((steps,) + u.shape, u.dtype)
        src = u
        for l in range(steps):
            out[l] = self.stout_isotropic(src, rho)[0]
            src = out[l]
        return out

    # stout force chain
    def compute_lambda(self, sp, u, ta):
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        lam = np.zeros((8, 8, self.sizeh), rd); tmp = np.zeros_like(u)
        getattr(self.lib, "so_compute_lambda" + self._sfx(u))(self.gp(), ptr(lam), ptr(sp), ptr(u), ptr(ta), ptr(tmp))
        return lam, tmp

    def compute_sigma(self, lam, u, sg, ta, rho):
        tmp = np.zeros_like(u)
        getattr(self.lib, "so_compute_sigma" + self._sfx(u))(self.gp(), ptr(lam), ptr(u), ptr(sg), ptr(ta), ptr(tmp), C.c_double(rho))
        return tmp

    def sigma_prime_to_sigma(self, sigma, u, rho):
        """compute_sigma_from_sigma_prime_backinto_sigma_prime (fermion_force.c:52-163), single rank: staples of U,
        Q = rho TA(U staples), Lambda from Sigma', Sigma in place.  -> (Lambda, Q)"""
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        sfx = self._sfx(u)
        tmp = np.zeros_like(u); qa = np.zeros((8, 8, self.sizeh), rd)
        getattr(self.lib, "so_calc_loc_staples_onlyferms" + sfx)(self.gp(), ptr(u), ptr(tmp))
        getattr(self.lib, "so_rho_times_conf_times_staples_ta_part" + sfx)(self.gp(), ptr(u), ptr(tmp), ptr(qa), C.c_double(rho))
        lam = np.zeros((8, 8, self.sizeh), rd)
        getattr(self.lib, "so_compute_lambda" + sfx)(self.gp(), ptr(lam), ptr(sigma), ptr(u), ptr(qa), ptr(tmp))
        getattr(self.lib, "so_compute_sigma" + sfx)(self.gp(), ptr(lam), ptr(u), ptr(sigma), ptr(qa), ptr(tmp), C.c_double(rho))
        return lam, qa

    # --- the callers of the path (single rank)
    def fermion_force(self, u, flavours, ferm_in, res, max_cg, rho, steps):
        """fermion_force_soloopenacc (fermion_force.c:166-357) from the restated steps, in the reference's order.
        flavours: list of dict(mass, ph, number_of_ps, first_ps, ra_a, ra_b).  -> (ipdot, gl3_aux, stout levels, [cg])"""
        rd = np.float32 if u.dtype == np.complex64 else np.float64
        stout = self.stout_wrapper(u, rho, steps) if steps > 0 else np.zeros((1,) + u.shape, u.dtype)     # :197
        conf = stout[steps - 1] if steps > 0 else u                                                       # :198-201
        gl3 = np.zeros_like(u); ipdot = np.zeros((8, 8, self.sizeh), rd); cgs = []
        for fl in flavours:
            taux = np.zeros_like(u)                                                                       # :226
            ph = np.ascontiguousarray(fl["ph"], rd)
            for ips in range(fl["number_of_ps"]):
                out, cg, ok, _ = self.multishift_invert(conf, ph, fl["mass"], fl["ra_b"],
                                                        np.ascontiguousarray(ferm_in[fl["first_ps"] + ips], u.dtype), res, max_cg)
                cgs.append(cg)
                self.compute_fermion_force(conf, taux, out, ph, fl["ra_a"])                               # :249
            self.multiply_backfield_times_force(ph, taux, gl3)                                            # :257
        for lvl in range(steps, 1, -1):                                                                   # :275-292
            self.sigma_prime_to_sigma(gl3, stout[lvl - 2], rho)
        if steps > 0:
            self.sigma_prime_to_sigma(gl3, u, rho)                                                        # :294-300
        self.take_ta(u, gl3, ipdot)                                                                       # :306
        return ipdot, gl3, stout, cgs

    def eo_inversion(self, u, ph, mass, in_e, in_o, res, max_cg, restarting_every=10000):
        """Meas/ferm_meas.c:50-72 with inverter_wrapper -> ker_invert_openacc (useMixedPrecision = 0)"""
        phi_e = self.dslash("deo", u, in_o, ph)
        self.axpy_like("fact1_minus_in2", phi_e, in_e, f1=mass)            # phi_e = m in_e - phi_e
        out_e, cg, ok = self.cg(u, ph, mass, phi_e, res, max_cg, 0.0, restarting_every)
        phi_o = self.dslash("doe", u, out_e, ph)
        out_o = np.zeros_like(in_o)
        self.axpy_like("in1_minus_in2_allxfact", out_o, in_o, phi_o, f1=1.0 / mass)
        return out_e, out_o, cg

    def dslash_wf(self, which, u, inp, ph, fre, fim):
        """which: 'deo' | 'doe' (field_times_fermion_matrix.c:77-196)"""
        out = np.zeros_like(inp)
        getattr(self.lib, "so_" + which + "_wf")(self.gp(), ptr(u), ptr(out), ptr(inp), ptr(ph), ptr(fre), ptr(fim),
                                                 C.c_int(self.d3_halo), C.c_int(self.d3_halo + self.loc_n[3]))
        return out

    # multi-rank helpers (global <-> rank-local boxes)
    def scatter_vec(self, rank, gl):
        lnh = np.zeros((3, self.sizeh), np.complex128)
        self.lib.so_scatter_vec(self.gp(), C.c_int(rank), ptr(gl), ptr(lnh)); return lnh

    def gather_vec(self, rank, gl, lnh):
        self.lib.so_gather_vec(self.gp(), C.c_int(rank), ptr(gl), ptr(lnh))

    def scatter_conf(self, rank, gl):
        lnh = np.zeros((8, 3, 3, self.sizeh), np.complex128)
        self.lib.so_scatter_conf(self.gp(), C.c_int(rank), ptr(gl), ptr(lnh)); return lnh

    def exchange_halo(self, vecs, thickness=1):
        """vecs: list (one per rank) of complex128[ncomp, sizeh] arrays, exchanged in place."""
        ncomp = vecs[0].reshape(-1, self.sizeh).shape[0]
        arr = (C.c_void_p * len(vecs))(*[v.ctypes.data for v in vecs])
        self.lib.so_exchange_halo(self.gp(), arr, C.c_int(ncomp), C.c_int(thickness))

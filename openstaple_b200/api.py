"""Host-side mirror of the reference's operator / solver interface for the hot path.

Every method of :class:`Lattice` has the NAME and ARGUMENT ORDER of the reference C function it
forwards to (include/staple_b200.h cites file:line); array arguments are torch CUDA tensors in
the reference SoA layouts, :class:`HostArray` (pinned host memory with a device mirror, the
analogue of `posix_memalign_wrapper` + `#pragma acc enter data`), or raw integer addresses.

  vec3_soa      complex128[3, sizeh]          su3_soa[8]  complex128[8, 3, 3, sizeh]
  double_soa[8] float64[8, sizeh]             vec3_soa[N] complex128[N, 3, sizeh]
"""
import ctypes as C

import numpy as np

from .lib import load_library

MAX_APPROX_ORDER = 25
INVERTER_SUCCESS, INVERTER_FAILURE = 1, 0
CONVERGENCE_CRITICAL, CONVERGENCE_NONCRITICAL = 1, 0


class RationalApprox(C.Structure):      # RationalApprox/rationalapprox.h:15-26
    _fields_ = [("exponent_num", C.c_int), ("exponent_den", C.c_int), ("approx_order", C.c_int),
                ("lambda_min", C.c_double), ("lambda_max", C.c_double), ("gmp_remez_precision", C.c_int),
                ("error", C.c_double), ("RA_a0", C.c_double), ("RA_a", C.c_double * MAX_APPROX_ORDER),
                ("RA_b", C.c_double * MAX_APPROX_ORDER)]

    @classmethod
    def make(cls, a0, a, b, num=-1, den=4):
        r = cls()
        r.exponent_num, r.exponent_den, r.approx_order = num, den, len(b)
        r.lambda_min, r.lambda_max, r.RA_a0 = 0.0, 1.0, a0
        for i in range(len(b)):
            r.RA_a[i] = a[i]; r.RA_b[i] = b[i]
        return r

    @classmethod
    def read(cls, path):
        """.REMEZ text format, RationalApprox/rationalapprox.c:83-117."""
        import re
        txt = open(path).read()
        r = cls()
        m = re.search(r"\(x\)\^\((-?\d+)/(-?\d+)\)", txt); r.exponent_num, r.exponent_den = int(m.group(1)), int(m.group(2))
        r.approx_order = int(re.search(r"Order:\s*(\d+)", txt).group(1))
        r.lambda_min = float(re.search(r"Lambda Min:\s*(\S+)", txt).group(1))
        r.lambda_max = float(re.search(r"Lambda Max:\s*(\S+)", txt).group(1))
        r.gmp_remez_precision = int(re.search(r"GMP Remez Precision:\s*(\d+)", txt).group(1))
        r.error = float(re.search(r"Error:\s*(\S+)", txt).group(1))
        r.RA_a0 = float(re.search(r"RA_a0 = (\S+)", txt).group(1))
        for m in re.finditer(r"RA_a\[(\d+)\] = (\S+), RA_b\[\d+\] = (\S+)", txt):
            k = int(m.group(1)); r.RA_a[k] = float(m.group(2).rstrip(",")); r.RA_b[k] = float(m.group(3))
        return r

    def rescaled(self, minmax):
        """rescale_rational_approximation, RationalApprox/rationalapprox.c:145-194."""
        out = RationalApprox()
        power = self.exponent_num / self.exponent_den
        out.exponent_num, out.exponent_den, out.approx_order = self.exponent_num, self.exponent_den, self.approx_order
        out.gmp_remez_precision, out.error = self.gmp_remez_precision, self.error
        mx = minmax[1] * 1.05
        eps = mx ** power
        out.RA_a0 = self.RA_a0 * eps
        for k in range(self.approx_order):
            out.RA_a[k] = self.RA_a[k] * mx * eps
            out.RA_b[k] = self.RA_b[k] * mx
        out.lambda_min, out.lambda_max = self.lambda_min * mx, mx
        if out.lambda_min > minmax[0]:
            raise ValueError("mother rational approx does not cover the range")
        return out


    def renormalized(self):
        """renormalize_rational_approximation, rationalapprox.c:197-222: the same approximation with lambda_max = 1"""
        out = RationalApprox()
        power = self.exponent_num / self.exponent_den
        out.exponent_num, out.exponent_den, out.approx_order = self.exponent_num, self.exponent_den, self.approx_order
        out.gmp_remez_precision, out.error = self.gmp_remez_precision, self.error
        ratio = 1 / self.lambda_max
        eps = ratio ** power
        out.RA_a0 = self.RA_a0 * eps
        for k in range(self.approx_order):
            out.RA_a[k] = self.RA_a[k] * ratio * eps
            out.RA_b[k] = self.RA_b[k] * ratio
        out.lambda_min, out.lambda_max = self.lambda_min * ratio, 1.0
        return out

    def evaluate(self, x):
        """rational_approx_evaluate, rationalapprox.c:225-237: RA_a0 + sum_i RA_a[i] / (x + RA_b[i])"""
        res = self.RA_a0
        for k in range(self.approx_order):
            res += self.RA_a[k] / (x + self.RA_b[k])
        return res

    def filename(self):
        """rational_approx_filename, rationalapprox.c:44-70: the name the reference looks a mother approximation up by"""
        import math
        return "approx_%d_over_%d_mlogerr_%1.1f_mloglm_%1.1f.REMEZ" % (
            self.exponent_num, self.exponent_den, -math.log(self.error) / math.log(10.0), -math.log(self.lambda_min) / math.log(10.0))

    def save(self, path):
        """rationalapprox_save, rationalapprox.c:120-143: byte-identical .REMEZ text"""
        with open(path, "w") as f:
            f.write("\nApproximation to f(x) = (x)^(%i/%i)\n" % (self.exponent_num, self.exponent_den))
            f.write("Order: %i\n" % self.approx_order)
            f.write("Lambda Min: %18.16e\nLambda Max: %18.16e\n" % (self.lambda_min, self.lambda_max))
            f.write("GMP Remez Precision: %i\nError: %18.16e\n" % (self.gmp_remez_precision, self.error))
            f.write("RA_a0 = %18.16e\n" % self.RA_a0)
            for i in range(self.approx_order):
                f.write("RA_a[%d] = %18.16e, RA_b[%d] = %18.16e\n" % (i, self.RA_a[i], i, self.RA_b[i]))


class FermParam(C.Structure):           # Include/fermion_parameters.h:9-41
    _fields_ = [("ferm_mass", C.c_double), ("degeneracy", C.c_int), ("number_of_ps", C.c_int),
                ("name", C.c_char * 30), ("ferm_charge", C.c_double), ("ferm_im_chem_pot", C.c_double),
                ("index_of_the_first_ps", C.c_int), ("index_of_the_first_shift", C.c_int),
                ("phases", C.c_void_p), ("mag_re", C.c_void_p), ("mag_im", C.c_void_p),
                ("printed_bf_dbg_info", C.c_int), ("phases_f", C.c_void_p),
                ("approx_fi_mother", RationalApprox), ("approx_md_mother", RationalApprox),
                ("approx_li_mother", RationalApprox), ("approx_fi", RationalApprox),
                ("approx_md", RationalApprox), ("approx_li", RationalApprox)]


class InverterPackage(C.Structure):     # OpenAcc/inverter_package.h:12-29
    _fields_ = [("u", C.c_void_p), ("u_f", C.c_void_p), ("ferm_shift_temp", C.c_void_p),
                ("ferm_shift_temp_f", C.c_void_p), ("nshifts", C.c_int),
                ("loc_r", C.c_void_p), ("loc_h", C.c_void_p), ("loc_s", C.c_void_p), ("loc_p", C.c_void_p),
                ("loc_r_f", C.c_void_p), ("loc_h_f", C.c_void_p), ("loc_s_f", C.c_void_p), ("loc_p_f", C.c_void_p),
                ("out_f", C.c_void_p)]


class ActionParam(C.Structure):         # OpenAcc/action.h:6-18
    _fields_ = [("beta", C.c_double), ("stout_steps", C.c_int), ("stout_rho", C.c_double), ("topo_action", C.c_int),
                ("barrier", C.c_double), ("width", C.c_double), ("topo_file_path", C.c_char * 20),
                ("topo_stout_steps", C.c_int), ("topo_rho", C.c_double)]


class MdParam(C.Structure):             # OpenAcc/md_parameters.h:6-19
    _fields_ = [("no_md", C.c_int), ("gauge_scale", C.c_int), ("t", C.c_double), ("residue_metro", C.c_double),
                ("expected_max_eigenvalue", C.c_double), ("singlePrecMD", C.c_int), ("residue_md", C.c_double),
                ("max_cg_iterations", C.c_int), ("recycleInvsForce", C.c_int), ("extrapolateInvsForce", C.c_int)]


class InvTricks(C.Structure):           # Include/inverter_tricks.h:4-11
    _fields_ = [("singlePInvAccelMultiInv", C.c_int), ("useMixedPrecision", C.c_int),
                ("mixedPrecisionDelta", C.c_double), ("restartingEvery", C.c_int)]


def _addr(x):
    """device/host address of a tensor, HostArray, ctypes struct or int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, HostArray):
        return x.ptr
    if isinstance(x, C.Structure):
        return C.addressof(x)
    raise TypeError("cannot take the address of %r" % type(x))


def _sfx(x):
    import torch
    if hasattr(x, "dtype"):
        dt = x.dtype
        if dt in (torch.complex64, torch.float32, np.complex64, np.float32):
            return "_f"
    return ""


def tamat_fields(ta):
    """views of a tamat_soa[8] array [8, 8, sizeh] (numpy or torch): (c01, c02, c12) complex [8, sizeh], (ic00, ic11) real."""
    n = ta.shape[-1]
    if hasattr(ta, "numpy") and not isinstance(ta, np.ndarray):
        ta = ta.cpu().numpy()
    cdt = np.complex64 if ta.dtype == np.float32 else np.complex128
    c = [np.ascontiguousarray(ta[:, 2 * j:2 * j + 2, :]).reshape(8, 2 * n).view(cdt) for j in range(3)]
    return c[0], c[1], c[2], ta[:, 6, :], ta[:, 7, :]


def geometry_plan(loc_n, nranks_d3=1, halo_width=2):
    """Host-only sharding arithmetic of the D3 slab decomposition (staple_geometry_plan; needs no GPU):
    local+halo box, sizeh, reduction/update ranges and the fermion halo offsets of
    Mpi/communications.c:51-96, as a dict."""
    L = load_library()
    n = (C.c_int * 4)(*[int(x) for x in loc_n]); o = (C.c_long * 16)()
    if L.staple_geometry_plan(n, int(nranks_d3), int(halo_width), o) != 0:
        raise ValueError("unsupported geometry %r ranks %d halo %d" % (tuple(loc_n), nranks_d3, halo_width))
    return dict(nd=tuple(o[0:4]), sizeh=o[4], vol3h=o[5], r0=(o[6], o[7]), r1=(o[8], o[9]),
                send_L=o[10], recv_R=o[11], send_R=o[12], recv_L=o[13], slab=o[14], d3_halo=o[15])


class HostArray:
    """Pinned host array made `present` on the device: staple_posix_memalign (the replacement of
    Include/memory_wrapper.c:14-31 + `#pragma acc enter data create`).  ``.np`` is a numpy view."""

    def __init__(self, lattice, shape, dtype):
        self.L = lattice.L
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        if self.L.staple_posix_memalign(C.byref(p), 128, self.nbytes) != 0:
            raise MemoryError("staple_posix_memalign failed")
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.np = np.frombuffer(buf, dtype=self.dtype).reshape(shape)

    def update_device(self):            # #pragma acc update device(...)
        self.L.staple_acc_update_device(self.ptr, self.nbytes)

    def update_host(self):              # #pragma acc update host(...)
        self.L.staple_acc_update_host(self.ptr, self.nbytes)

    def free(self):
        if self.ptr:
            self.np = None
            self.L.staple_free(self.ptr); self.ptr = None


class Lattice:
    """One rank's view of the lattice: run-time stand-in for geom_defines.txt (LOC_N0..3, NRANKS_D3,
    HALO_WIDTH) plus the rank layer (Mpi/multidev.c).  Requires a CUDA device."""

    def __init__(self, loc_n, nranks_d3=1, halo_width=2, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("openstaple_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.L = load_library()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)          # make sure the primary context exists
        if self.L.staple_init_geometry(*[int(x) for x in loc_n], int(nranks_d3), int(halo_width), int(device)) != 0:
            raise ValueError("invalid geometry")
        self.loc_n = tuple(loc_n); self.nranks = nranks_d3; self.halo_width = halo_width
        nd = (C.c_int * 4)(); rg = (C.c_long * 4)()
        self.L.staple_geometry(nd, rg)
        self.nd = tuple(nd); self.ranges = tuple(rg)
        self.sizeh = self.L.staple_sizeh()
        self.d3_halo = halo_width if nranks_d3 > 1 else 0
        self.vol3h = self.nd[0] * self.nd[1] * self.nd[2] // 2
        self.rank = 0
        self.use_torch_stream()
        self.inverter_tricks = InvTricks.in_dll(self.L, "inverter_tricks")
        self._keep = []

    # ---- plumbing
    def use_torch_stream(self):
        """enqueue on torch's CURRENT stream (handle 0 = the legacy default stream), so library kernels are
        ordered with torch allocations, fills and copies of the same tensors."""
        self.L.staple_set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)

    def use_library_stream(self):
        self.L.staple_use_library_stream()

    def synchronize(self):
        self.L.staple_synchronize()

    def kernel_launches(self):
        return int(self.L.staple_kernel_launches())

    def init_multidev(self, dist, async_comm_fermion=1, p2p=0):
        """pre_init_multidev1D + init_multidev1D (Mpi/multidev.c:20-108) with torch.distributed as
        the out-of-band channel for the NCCL id."""
        rank, world = dist.get_rank(), dist.get_world_size()
        idt = self.torch.zeros(128, dtype=self.torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            self.L.staple_nccl_unique_id(buf)
            idt = self.torch.frombuffer(bytearray(buf.raw), dtype=self.torch.uint8).clone()
        if dist.get_backend() == "nccl":
            idt = idt.to(self.device); dist.broadcast(idt, 0); idt = idt.cpu()
        else:
            dist.broadcast(idt, 0)
        raw = bytes(idt.numpy().tobytes())
        self.L.staple_init_multidev1D(rank, world, C.c_char_p(raw), int(async_comm_fermion))
        self.rank = rank
        self.p2p = bool(self.L.staple_enable_p2p(int(p2p))) if world > 1 else False

    def init_loopback(self, p2p=1):
        """the D3-slab code path on one GPU: this rank is its own L and R neighbour (staple_init_loopback)"""
        if self.L.staple_init_loopback(int(p2p)) != 0:
            raise RuntimeError("staple_init_loopback failed")
        self.rank = 0
        self.p2p = p2p != 0

    def shutdown_multidev(self):
        self.L.shutdown_multidev()

    # ---- allocation helpers (torch owns device memory)
    def _cdt(self, single):
        return self.torch.complex64 if single else self.torch.complex128

    def new_vec(self, n=None, single=False):
        shape = (3, self.sizeh) if n is None else (n, 3, self.sizeh)
        return self.torch.zeros(shape, dtype=self._cdt(single), device=self.device)

    def new_conf(self, single=False):
        return self.torch.zeros((8, 3, 3, self.sizeh), dtype=self._cdt(single), device=self.device)

    def new_tamat(self, single=False):
        """tamat_soa[8] (struct_c_def.h:45-51) as reals: [8, 8, sizeh]; rows 0-5 = c01, c02, c12 (re/im interleaved
        per element), rows 6, 7 = ic00, ic11.  See :func:`tamat_fields`."""
        return self.torch.zeros((8, 8, self.sizeh), dtype=self.torch.float32 if single else self.torch.float64,
                                device=self.device)

    def to_device(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def load_configuration(self, path, fmt="ildg"):
        """Read a GLOBAL configuration file (ILDG/LIME or the reference's ASCII format, openstaple_b200/io.py), cut out
        this rank's local+halo box (halos included, Mpi/communications.c:1104-1148) and upload it in the su3_soa[8]
        layout -> (device tensor [8,3,3,sizeh], conf_id)."""
        from . import io as sio
        dims = (self.loc_n[0], self.loc_n[1], self.loc_n[2], self.loc_n[3] * self.nranks)
        if fmt == "ildg":
            conf, cid = sio.read_su3_soa_ildg_binary(path, dims)
        else:
            conf, cid, _ = sio.read_su3_soa_ASCII(path, dims)
        if self.nranks > 1:
            conf = sio.send_lnh_subconf_to_buffer(conf, self.rank, self.loc_n, self.nranks, self.halo_width)
        return self.to_device(conf), cid

    def calc_u1_phases(self, bf_pars=(0, 0, 0, 0, 0, 0), im_chem_pot=0.0, ferm_charge=0.0, single=False):
        """this rank's phase field (OpenAcc/backfield.c:20-187, host side: openstaple_b200/backfield.py) uploaded as
        double_soa[8] / float_soa[8] -> device tensor [8, sizeh]"""
        from .backfield import calc_u1_phases
        return self.to_device(calc_u1_phases(self.loc_n, bf_pars, im_chem_pot, ferm_charge, self.nranks, self.rank,
                                             self.halo_width, single))

    def host_array(self, shape, dtype):
        return HostArray(self, shape, dtype)

    def ferm_param(self, mass, phases=None, phases_f=None):
        p = FermParam()
        p.ferm_mass = mass; p.degeneracy = 1; p.number_of_ps = 1; p.name = b"flavour"
        p.phases = _addr(phases); p.phases_f = _addr(phases_f)
        self._keep.append((p, phases, phases_f))
        return p

    def set_inverter_tricks(self, singlePInvAccelMultiInv=0, useMixedPrecision=0, mixedPrecisionDelta=0.1,
                            restartingEvery=10000):
        t = self.inverter_tricks
        t.singlePInvAccelMultiInv, t.useMixedPrecision = singlePInvAccelMultiInv, useMixedPrecision
        t.mixedPrecisionDelta, t.restartingEvery = mixedPrecisionDelta, restartingEvery

    # ---- Dirac operator (OpenAcc/fermion_matrix.h:20-106)
    def _dslash(self, name, u, out, inp, backfield):
        getattr(self.L, name + _sfx(inp))(_addr(u), _addr(out), _addr(inp), _addr(backfield))

    def acc_Deo(self, u, out, inp, backfield): self._dslash("acc_Deo", u, out, inp, backfield)
    def acc_Doe(self, u, out, inp, backfield): self._dslash("acc_Doe", u, out, inp, backfield)
    def acc_Deo_unsafe(self, u, out, inp, backfield): self._dslash("acc_Deo_unsafe", u, out, inp, backfield)
    def acc_Doe_unsafe(self, u, out, inp, backfield): self._dslash("acc_Doe_unsafe", u, out, inp, backfield)
    def acc_Deo_bulk(self, u, out, inp, backfield): self._dslash("acc_Deo_bulk", u, out, inp, backfield)
    def acc_Doe_bulk(self, u, out, inp, backfield): self._dslash("acc_Doe_bulk", u, out, inp, backfield)
    def acc_Deo_d3p(self, u, out, inp, backfield): self._dslash("acc_Deo_d3p", u, out, inp, backfield)
    def acc_Doe_d3p(self, u, out, inp, backfield): self._dslash("acc_Doe_d3p", u, out, inp, backfield)
    def acc_Deo_d3m(self, u, out, inp, backfield): self._dslash("acc_Deo_d3m", u, out, inp, backfield)
    def acc_Doe_d3m(self, u, out, inp, backfield): self._dslash("acc_Doe_d3m", u, out, inp, backfield)

    # ---- operator "with a field" (OpenAcc/field_times_fermion_matrix.h:20-50), FP64 only
    def _dslash_wf(self, name, u, out, inp, phases, field_re, field_im):
        f = getattr(self.L, name); f.argtypes = [C.c_void_p] * 6; f.restype = None
        f(_addr(u), _addr(out), _addr(inp), _addr(phases), _addr(field_re), _addr(field_im))

    def acc_Deo_wf(self, u, out, inp, phases, field_re, field_im): self._dslash_wf("acc_Deo_wf", u, out, inp, phases, field_re, field_im)
    def acc_Doe_wf(self, u, out, inp, phases, field_re, field_im): self._dslash_wf("acc_Doe_wf", u, out, inp, phases, field_re, field_im)
    def acc_Deo_wf_unsafe(self, u, out, inp, phases, field_re, field_im): self._dslash_wf("acc_Deo_wf_unsafe", u, out, inp, phases, field_re, field_im)
    def acc_Doe_wf_unsafe(self, u, out, inp, phases, field_re, field_im): self._dslash_wf("acc_Doe_wf_unsafe", u, out, inp, phases, field_re, field_im)

    def acc_Deo_d3c(self, u, out, inp, backfield, off3, thick3):
        getattr(self.L, "acc_Deo_d3c" + _sfx(inp))(_addr(u), _addr(out), _addr(inp), _addr(backfield), off3, thick3)

    def acc_Doe_d3c(self, u, out, inp, backfield, off3, thick3):
        getattr(self.L, "acc_Doe_d3c" + _sfx(inp))(_addr(u), _addr(out), _addr(inp), _addr(backfield), off3, thick3)

    def acc_Doe_Deo_streamed(self, u, out, inp, tmp, backfield, chunk_slices=0):
        """deo_doe_test.c's `update device(in); acc_Doe; acc_Deo; update host(out)` round trip on HostArrays,
        pipelined over d3 chunks (staple_acc_Doe_Deo_streamed); returns with `out` valid on the host."""
        self.L.staple_acc_Doe_Deo_streamed(_addr(u), _addr(out), _addr(inp), _addr(tmp), _addr(backfield), int(chunk_slices))

    def fermion_matrix_multiplication(self, u, out, inp, temp1, pars, single=None):
        s = _sfx(inp) if single is None else ("_f" if single else "")
        getattr(self.L, "fermion_matrix_multiplication" + s)(_addr(u), _addr(out), _addr(inp), _addr(temp1), C.addressof(pars))

    def fermion_matrix_multiplication_shifted(self, u, out, inp, temp1, pars, shift, single=None):
        s = _sfx(inp) if single is None else ("_f" if single else "")
        getattr(self.L, "fermion_matrix_multiplication_shifted" + s)(
            _addr(u), _addr(out), _addr(inp), _addr(temp1), C.addressof(pars), float(shift))

    # ---- BLAS-1 (OpenAcc/fermionic_utilities.h:38-123)
    def scal_prod_global(self, a, b):
        r = getattr(self.L, "scal_prod_global" + _sfx(a))(_addr(a), _addr(b)); return complex(r.re, r.im)

    def real_scal_prod_global(self, a, b):
        return getattr(self.L, "real_scal_prod_global" + _sfx(a))(_addr(a), _addr(b))

    def l2norm2_global(self, a):
        return getattr(self.L, "l2norm2_global" + _sfx(a))(_addr(a))

    def combine_in1xfactor_plus_in2(self, in1, factor, in2, out):
        getattr(self.L, "combine_in1xfactor_plus_in2" + _sfx(out))(_addr(in1), float(factor), _addr(in2), _addr(out))

    def multiply_fermion_x_doublefactor(self, in1, factor):
        getattr(self.L, "multiply_fermion_x_doublefactor" + _sfx(in1))(_addr(in1), float(factor))

    def combine_add_factor_x_in2_to_in1(self, in1, in2, factor):
        getattr(self.L, "combine_add_factor_x_in2_to_in1" + _sfx(in1))(_addr(in1), _addr(in2), float(factor))

    def combine_in1xferm_mass2_minus_in2_minus_in3(self, in1, ferm_mass, in2, in3, out):
        getattr(self.L, "combine_in1xferm_mass2_minus_in2_minus_in3" + _sfx(out))(
            _addr(in1), float(ferm_mass), _addr(in2), _addr(in3), _addr(out))

    def combine_inside_loop(self, vect_out, vect_r, vect_s, vect_p, omega):
        getattr(self.L, "combine_inside_loop" + _sfx(vect_out))(_addr(vect_out), _addr(vect_r), _addr(vect_s), _addr(vect_p), float(omega))

    def combine_in1xferm_mass_minus_in2(self, in1, ferm_mass2, in2):
        getattr(self.L, "combine_in1xferm_mass_minus_in2" + _sfx(in2))(_addr(in1), float(ferm_mass2), _addr(in2))

    def combine_in1_minus_in2(self, in1, in2, out):
        getattr(self.L, "combine_in1_minus_in2" + _sfx(out))(_addr(in1), _addr(in2), _addr(out))

    def assign_in_to_out(self, in1, out):
        getattr(self.L, "assign_in_to_out" + _sfx(out))(_addr(in1), _addr(out))

    def set_vec3_soa_to_zero(self, fermion):
        getattr(self.L, "set_vec3_soa_to_zero" + _sfx(fermion))(_addr(fermion))

    def multiple_combine_in1_minus_in2x_factor_back_into_in1(self, out, inp, maxiter, flag, omegas):
        fl = (C.c_int * len(flag))(*flag); om = (C.c_double * len(omegas))(*omegas)
        getattr(self.L, "multiple_combine_in1_minus_in2x_factor_back_into_in1" + _sfx(out))(_addr(out), _addr(inp), maxiter, fl, om)

    def multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1(self, in1, maxiter, flag, gammas, in2, zeta_iii):
        fl = (C.c_int * len(flag))(*flag); g = (C.c_double * len(gammas))(*gammas); z = (C.c_double * len(zeta_iii))(*zeta_iii)
        getattr(self.L, "multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1" + _sfx(in1))(
            _addr(in1), maxiter, fl, g, _addr(in2), z)

    def combine_in1_x_fact1_minus_in2_back_into_in2(self, in1, fact1, in2):
        getattr(self.L, "combine_in1_x_fact1_minus_in2_back_into_in2" + _sfx(in2))(_addr(in1), float(fact1), _addr(in2))

    def combine_in1_minus_in2_allxfact(self, in1, in2, fact, out):
        getattr(self.L, "combine_in1_minus_in2_allxfact" + _sfx(out))(_addr(in1), _addr(in2), float(fact), _addr(out))

    def calc_new_trialsol_for_inversion_in_force(self, halfLen, inout, nPrecCalculations):
        getattr(self.L, "calc_new_trialsol_for_inversion_in_force" + _sfx(inout))(halfLen, _addr(inout), nPrecCalculations)

    # ---- fermion-force outer products (OpenAcc/fermion_force_utilities.h)
    def set_tamat_soa_to_zero(self, matrix): getattr(self.L, "set_tamat_soa_to_zero" + _sfx(matrix))(_addr(matrix))
    def set_su3_soa_to_zero(self, matrix): getattr(self.L, "set_su3_soa_to_zero" + _sfx(matrix))(_addr(matrix))

    def direct_product_of_fermions_into_auxmat(self, loc_s, loc_h, aux_u, approx, it):
        getattr(self.L, "direct_product_of_fermions_into_auxmat" + _sfx(loc_s))(
            _addr(loc_s), _addr(loc_h), _addr(aux_u), C.addressof(approx), int(it))

    def multiply_conf_times_force_and_take_ta_nophase(self, u, auxmat, ipdot):
        getattr(self.L, "multiply_conf_times_force_and_take_ta_nophase" + _sfx(u))(_addr(u), _addr(auxmat), _addr(ipdot))

    def multiply_backfield_times_force(self, tpars, auxmat, pseudo_ipdot):
        getattr(self.L, "multiply_backfield_times_force" + _sfx(auxmat))(C.addressof(tpars), _addr(auxmat), _addr(pseudo_ipdot))

    def accumulate_gl3soa_into_gl3soa(self, auxmat, pseudo_ipdot):
        getattr(self.L, "accumulate_gl3soa_into_gl3soa" + _sfx(auxmat))(_addr(auxmat), _addr(pseudo_ipdot))

    def ker_openacc_compute_fermion_force(self, u, aux_u, in_shiftmulti, loc_s, loc_h, tpars):
        """tpars.approx_md holds the shifts' residues RA_a (fermion_force_utilities.c:195)."""
        getattr(self.L, "ker_openacc_compute_fermion_force" + _sfx(u))(
            _addr(u), _addr(aux_u), _addr(in_shiftmulti), _addr(loc_s), _addr(loc_h), C.addressof(tpars))

    # ---- isotropic stout smearing (OpenAcc/stouting.h, plaquettes.h:18, su3_utilities.h:49)
    def set_stout(self, rho, steps, auxbis=None, staples=None, ipdot=None, single=False):
        """what main.c:276-277 and alloc_vars.c do for stout_wrapper: act_params.stout_rho/steps, gl_stout_rho and the
        parking arrays auxbis_conf_acc, glocal_staples, gipdot (the library's weak globals)."""
        ap = ActionParam.in_dll(self.L, "act_params")
        ap.stout_rho, ap.stout_steps, ap.topo_action = rho, steps, 0
        C.c_double.in_dll(self.L, "gl_stout_rho").value = rho
        C.c_double.in_dll(self.L, "gl_topo_rho").value = rho
        sfx = "_f" if single else ""
        for name, arr in (("auxbis_conf_acc", auxbis), ("glocal_staples", staples), ("gipdot", ipdot)):
            if arr is not None:
                C.c_void_p.in_dll(self.L, name + sfx).value = _addr(arr)
        self._keep.append((auxbis, staples, ipdot))

    def calc_loc_staples_nnptrick_all_onlyferms(self, u, loc_stap):
        getattr(self.L, "calc_loc_staples_nnptrick_all_onlyferms" + _sfx(u))(_addr(u), _addr(loc_stap))

    def RHO_times_conf_times_staples_ta_part(self, u, loc_stap, tipdot, istopo=0):
        getattr(self.L, "RHO_times_conf_times_staples_ta_part" + _sfx(u))(_addr(u), _addr(loc_stap), _addr(tipdot), int(istopo))

    def exp_minus_QA_times_conf(self, tu, QA, tu_out, exp_aux):
        getattr(self.L, "exp_minus_QA_times_conf" + _sfx(tu))(_addr(tu), _addr(QA), _addr(tu_out), _addr(exp_aux))

    def stout_isotropic(self, u, uprime, local_staples, auxiliary, tipdot, istopo=0):
        getattr(self.L, "stout_isotropic" + _sfx(u))(_addr(u), _addr(uprime), _addr(local_staples), _addr(auxiliary),
                                                     _addr(tipdot), int(istopo))

    def stout_wrapper(self, tconf_acc, tstout_conf_acc_arr, istopo=0):
        getattr(self.L, "stout_wrapper" + _sfx(tconf_acc))(_addr(tconf_acc), _addr(tstout_conf_acc_arr), int(istopo))

    # ---- stouted fermion force: Sigma' -> Sigma (OpenAcc/stouting.h, fermion_force.c:52-163)
    def compute_lambda(self, L, SP, U, QA, TMP):
        getattr(self.L, "compute_lambda" + _sfx(U))(_addr(L), _addr(SP), _addr(U), _addr(QA), _addr(TMP))

    def compute_sigma(self, L, U, S, QA, TMP, istopo=0):
        getattr(self.L, "compute_sigma" + _sfx(U))(_addr(L), _addr(U), _addr(S), _addr(QA), _addr(TMP), int(istopo))

    def compute_sigma_from_sigma_prime_backinto_sigma_prime(self, Sigma, Lambda, QA, U, TMP, istopo=0):
        getattr(self.L, "compute_sigma_from_sigma_prime_backinto_sigma_prime" + _sfx(U))(
            _addr(Sigma), _addr(Lambda), _addr(QA), _addr(U), _addr(TMP), int(istopo))

    # ---- conversions (OpenAcc/float_double_conv.c)
    def convert_double_to_float_vec3_soa(self, d, f): self.L.convert_double_to_float_vec3_soa(_addr(d), _addr(f))
    def convert_float_to_double_vec3_soa(self, f, d): self.L.convert_float_to_double_vec3_soa(_addr(f), _addr(d))

    def convert_double_to_float_su3_soa(self, d, f):
        """whole conf[8] in one call, like the reference (float_double_conv.c:122-150)."""
        self.L.convert_double_to_float_su3_soa(_addr(d), _addr(f))

    def convert_float_to_double_su3_soa(self, f, d):
        self.L.convert_float_to_double_su3_soa(_addr(f), _addr(d))

    def convert_double_to_float_tamat_soa(self, d, f):          # tamat_soa[8] / thmat_soa[8]: [8, 8, sizeh] reals
        self.L.convert_double_to_float_tamat_soa(_addr(d), _addr(f))

    def convert_float_to_double_tamat_soa(self, f, d):
        self.L.convert_float_to_double_tamat_soa(_addr(f), _addr(d))

    def convert_double_to_float_thmat_soa(self, d, f):
        self.L.convert_double_to_float_thmat_soa(_addr(d), _addr(f))

    def convert_float_to_double_thmat_soa(self, f, d):
        self.L.convert_float_to_double_thmat_soa(_addr(f), _addr(d))

    def convert_double_to_float_real_soa(self, d, f):
        for k in range(8):
            self.L.convert_double_to_float_real_soa(_addr(d) + k * self.sizeh * 8, _addr(f) + k * self.sizeh * 4)

    # ---- halo layer (Mpi/communications.h:12-30)
    def communicate_fermion_borders(self, v):
        getattr(self.L, "communicate_fermion_borders" + _sfx(v))(_addr(v))

    def communicate_su3_borders(self, u, thickness):
        getattr(self.L, "communicate_su3_borders" + _sfx(u))(_addr(u), thickness)

    # ---- solvers
    def multishift_invert(self, u, pars, approx, out, inp, residuo, loc_r, loc_h, loc_s, loc_p, shiftferm, max_cg):
        """-> (INVERTER_SUCCESS|FAILURE, cg) ; OpenAcc/inverter_multishift_full.c:23-252."""
        cg = C.c_int(0)
        st = getattr(self.L, "multishift_invert" + _sfx(inp))(
            _addr(u), C.addressof(pars), C.addressof(approx), _addr(out), _addr(inp), float(residuo), _addr(loc_r),
            _addr(loc_h), _addr(loc_s), _addr(loc_p), _addr(shiftferm), int(max_cg), C.byref(cg))
        return st, cg.value

    def recombine_shifted_vec3_to_vec3(self, in_shifted, inp, out, approx):
        getattr(self.L, "recombine_shifted_vec3_to_vec3" + _sfx(inp))(_addr(in_shifted), _addr(inp), _addr(out), C.addressof(approx))

    def ker_invert_openacc(self, u, pars, solution, inp, res, loc_r, loc_h, loc_s, loc_p, max_cg, shift):
        cg = C.c_int(0)
        st = getattr(self.L, "ker_invert_openacc" + _sfx(inp))(
            _addr(u), C.addressof(pars), _addr(solution), _addr(inp), float(res), _addr(loc_r), _addr(loc_h),
            _addr(loc_s), _addr(loc_p), int(max_cg), float(shift), C.byref(cg))
        return st, cg.value

    def setup_inverter_package_dp(self, ip, u, ferm_shift_temp, nshifts, loc_r, loc_h, loc_s, loc_p):
        self.L.setup_inverter_package_dp.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4
        self.L.setup_inverter_package_dp(C.addressof(ip), _addr(u), _addr(ferm_shift_temp), nshifts, _addr(loc_r),
                                         _addr(loc_h), _addr(loc_s), _addr(loc_p))
        ip._refs_dp = (u, ferm_shift_temp, loc_r, loc_h, loc_s, loc_p)     # the struct holds raw addresses only

    def setup_inverter_package_sp(self, ip, u_f, ferm_shift_temp_f, nshifts, loc_r_f, loc_h_f, loc_s_f, loc_p_f, out_f):
        self.L.setup_inverter_package_sp.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5
        self.L.setup_inverter_package_sp(C.addressof(ip), _addr(u_f), _addr(ferm_shift_temp_f), nshifts, _addr(loc_r_f),
                                         _addr(loc_h_f), _addr(loc_s_f), _addr(loc_p_f), _addr(out_f))
        ip._refs_sp = (u_f, ferm_shift_temp_f, loc_r_f, loc_h_f, loc_s_f, loc_p_f, out_f)

    def inverter_mixed_precision(self, ip, pars, solution, inp, res, max_cg, shift):
        f = self.L.inverter_mixed_precision
        f.argtypes = [InverterPackage, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int)]
        f.restype = C.c_int
        cg = C.c_int(0)
        st = f(ip, C.addressof(pars), _addr(solution), _addr(inp), float(res), int(max_cg), float(shift), C.byref(cg))
        return st, cg.value

    def inverter_multishift_wrapper(self, ip, pars, approx, out, inp, res, max_cg, convergence_importance):
        f = self.L.inverter_multishift_wrapper
        f.argtypes = [InverterPackage, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int]
        f.restype = C.c_int
        return f(ip, C.addressof(pars), C.addressof(approx), _addr(out), _addr(inp), float(res), int(max_cg),
                 int(convergence_importance))

    def inverter_wrapper(self, ip, pars, out, inp, res, max_cg, shift, convergence_importance):
        f = self.L.inverter_wrapper
        f.argtypes = [InverterPackage, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int]
        f.restype = C.c_int
        return f(ip, C.addressof(pars), _addr(out), _addr(inp), float(res), int(max_cg), float(shift),
                 int(convergence_importance))

    # ---- callers of the path: whole MD fermion force (OpenAcc/fermion_force.h:23-37), eo_inversion (Meas/ferm_meas.h)
    def ferm_param_array(self, flavours):
        """ferm_param[nflav] as fermion_force_soloopenacc walks it; flavours: list of dict(mass, phases, phases_f,
        number_of_ps, first_ps, ra_a, ra_b) with approx_md = (ra_a, ra_b)."""
        arr = (FermParam * len(flavours))()
        for p, fl in zip(arr, flavours):
            p.ferm_mass = fl["mass"]; p.degeneracy = 1; p.name = b"flavour"
            p.number_of_ps = fl.get("number_of_ps", 1); p.index_of_the_first_ps = fl.get("first_ps", 0)
            p.phases = _addr(fl.get("phases")); p.phases_f = _addr(fl.get("phases_f"))
            p.approx_md.approx_order = len(fl["ra_b"])
            for i, (a, b) in enumerate(zip(fl["ra_a"], fl["ra_b"])):
                p.approx_md.RA_a[i] = a; p.approx_md.RA_b[i] = b
        self._keep.append((arr, flavours))
        return arr

    def set_force_globals(self, aux_th=None, aux_ta=None, conf_acc_f=None, single=False, recycleInvsForce=0):
        """the parking arrays fermion_force_soloopenacc takes from alloc_vars (aux_th, aux_ta [+_f], conf_acc_f) and
        md_parameters.recycleInvsForce -- the library's weak globals."""
        sfx = "_f" if single else ""
        for name, arr in (("aux_th" + sfx, aux_th), ("aux_ta" + sfx, aux_ta), ("conf_acc_f", conf_acc_f)):
            if arr is not None:
                C.c_void_p.in_dll(self.L, name).value = _addr(arr)
        MdParam.in_dll(self.L, "md_parameters").recycleInvsForce = recycleInvsForce
        self._keep.append((aux_th, aux_ta, conf_acc_f))

    def fermion_force_soloopenacc(self, tconf_acc, tstout_conf_acc_arr, gl3_aux, tipdot_acc, tfermion_parameters, tNDiffFlavs,
                                  ferm_in_acc, res, taux_conf_acc, tferm_shiftmulti_acc, ipt, max_cg):
        single = _sfx(tconf_acc) == "_f"
        f = getattr(self.L, "fermion_force_soloopenacc" + ("_f" if single else ""))
        f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_float if single else C.c_double, C.c_void_p, C.c_void_p,
                                         InverterPackage, C.c_int]
        f.restype = None
        f(_addr(tconf_acc), _addr(tstout_conf_acc_arr), _addr(gl3_aux), _addr(tipdot_acc), C.addressof(tfermion_parameters),
          int(tNDiffFlavs), _addr(ferm_in_acc), float(res), _addr(taux_conf_acc), _addr(tferm_shiftmulti_acc), ipt, int(max_cg))

    def eo_inversion(self, ip, pars, res, max_cg, in_e, in_o, out_e, out_o, phi_e, phi_o):
        f = self.L.eo_inversion
        f.argtypes = [InverterPackage, C.c_void_p, C.c_double, C.c_int] + [C.c_void_p] * 6; f.restype = None
        f(ip, C.addressof(pars), float(res), int(max_cg), _addr(in_e), _addr(in_o), _addr(out_e), _addr(out_o),
          _addr(phi_e), _addr(phi_o))

    def last_refinement_iterations(self):
        return int(self.L.staple_last_refinement_iterations())

    def set_sp_globals(self, aux1_f, ferm_shiftmulti_acc_f):
        self._keep.append((aux1_f, ferm_shiftmulti_acc_f))
        self.L.staple_set_sp_globals(_addr(aux1_f), _addr(ferm_shiftmulti_acc_f))

    def ker_find_max_eigenvalue_openacc(self, u, pars, loc_r, loc_h, loc_p):
        return self.L.ker_find_max_eigenvalue_openacc(_addr(u), C.addressof(pars), _addr(loc_r), _addr(loc_h), _addr(loc_p))

    def ker_find_min_eigenvalue_openacc(self, u, pars, loc_r, loc_h, loc_p, mx):
        f = self.L.ker_find_min_eigenvalue_openacc
        f.argtypes = [C.c_void_p] * 5 + [C.c_double]; f.restype = C.c_double
        return f(_addr(u), C.addressof(pars), _addr(loc_r), _addr(loc_h), _addr(loc_p), float(mx))

    def last_solve_stats(self):
        it = C.c_int(0); act = C.c_longlong(0); ms = C.c_double(0)
        self.L.staple_last_solve_stats(C.byref(it), C.byref(act), C.byref(ms))
        return it.value, act.value, ms.value

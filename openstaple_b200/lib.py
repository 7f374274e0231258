"""ctypes loader for libstaple_b200.so.  There is no Python/CPU fallback: if the CUDA library is
missing, loading raises, and every compute entry point of the library aborts without a GPU."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    # STAPLE_LIB selects a tuning variant of the SAME CUDA library (scripts/tune_dslash.py); never a fallback
    return os.environ.get("STAPLE_LIB") or os.path.join(HERE, "libstaple_b200.so")


def load_library(build_if_missing=True):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path) and build_if_missing:
        from .build import build
        build()
    if not os.path.exists(path):
        raise RuntimeError("libstaple_b200.so is missing (run `python -m openstaple_b200.build`); "
                           "there is no CPU fallback for the hot path")
    _LIB = C.CDLL(path, mode=C.RTLD_LOCAL)   # never interpose on a host program (or the test oracle) by accident
    _declare(_LIB)
    return _LIB


class DComplex(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def _declare(L):
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    L.staple_init_geometry.argtypes = [i, i, i, i, i, i, i]; L.staple_init_geometry.restype = i
    L.staple_sizeh.restype = C.c_long
    L.staple_geometry.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_long)]
    L.staple_geometry_plan.argtypes = [C.POINTER(C.c_int), i, i, C.POINTER(C.c_long)]; L.staple_geometry_plan.restype = i
    L.staple_set_stream.argtypes = [vp]
    L.staple_get_stream.restype = vp
    L.staple_use_library_stream.argtypes = []
    L.staple_set_use_graphs.argtypes = [i]
    L.staple_set_cgm_fuse_tail.argtypes = [i]
    L.staple_set_cg_device_loops.argtypes = [i]
    L.staple_set_streamed_mode.argtypes = [i]
    L.staple_kernel_launches.restype = C.c_ulonglong
    L.staple_version.restype = C.c_char_p
    L.staple_posix_memalign.argtypes = [C.POINTER(vp), C.c_size_t, C.c_size_t]; L.staple_posix_memalign.restype = i
    L.staple_free.argtypes = [vp]
    for f in ("staple_acc_enter_data", "staple_acc_update_device", "staple_acc_update_host"):
        getattr(L, f).argtypes = [vp, C.c_size_t]
    L.staple_acc_exit_data.argtypes = [vp]
    L.staple_acc_deviceptr.argtypes = [vp]; L.staple_acc_deviceptr.restype = vp
    L.staple_acc_Doe_Deo_streamed.argtypes = [vp, vp, vp, vp, vp, i]; L.staple_acc_Doe_Deo_streamed.restype = None
    L.staple_nccl_unique_id.argtypes = [vp]; L.staple_nccl_unique_id.restype = i
    L.staple_init_multidev1D.argtypes = [i, i, vp, i]; L.staple_init_multidev1D.restype = i
    L.staple_myrank.restype = i
    L.staple_enable_p2p.argtypes = [i]; L.staple_enable_p2p.restype = i
    L.staple_init_loopback.argtypes = [i]; L.staple_init_loopback.restype = i
    L.staple_set_spin_timeout.argtypes = [d]
    for s in ("", "_f"):
        for f in ("acc_Deo", "acc_Doe", "acc_Deo_unsafe", "acc_Doe_unsafe", "acc_Deo_bulk", "acc_Doe_bulk",
                  "acc_Deo_d3p", "acc_Doe_d3p", "acc_Deo_d3m", "acc_Doe_d3m"):
            getattr(L, f + s).argtypes = [vp, vp, vp, vp]; getattr(L, f + s).restype = None
        for f in ("acc_Deo_d3c", "acc_Doe_d3c"):
            getattr(L, f + s).argtypes = [vp, vp, vp, vp, i, i]; getattr(L, f + s).restype = None
        getattr(L, "fermion_matrix_multiplication" + s).argtypes = [vp, vp, vp, vp, vp]
        getattr(L, "fermion_matrix_multiplication_shifted" + s).argtypes = [vp, vp, vp, vp, vp, C.c_float if s else d]
        getattr(L, "scal_prod_global" + s).argtypes = [vp, vp]; getattr(L, "scal_prod_global" + s).restype = DComplex
        getattr(L, "real_scal_prod_global" + s).argtypes = [vp, vp]; getattr(L, "real_scal_prod_global" + s).restype = d
        getattr(L, "l2norm2_global" + s).argtypes = [vp]; getattr(L, "l2norm2_global" + s).restype = d
        getattr(L, "combine_in1xfactor_plus_in2" + s).argtypes = [vp, d, vp, vp]
        getattr(L, "multiply_fermion_x_doublefactor" + s).argtypes = [vp, d]
        getattr(L, "combine_add_factor_x_in2_to_in1" + s).argtypes = [vp, vp, d]
        getattr(L, "combine_in1xferm_mass2_minus_in2_minus_in3" + s).argtypes = [vp, d, vp, vp, vp]
        getattr(L, "combine_inside_loop" + s).argtypes = [vp, vp, vp, vp, d]
        getattr(L, "combine_in1xferm_mass_minus_in2" + s).argtypes = [vp, d, vp]
        getattr(L, "combine_in1_minus_in2" + s).argtypes = [vp, vp, vp]
        getattr(L, "assign_in_to_out" + s).argtypes = [vp, vp]
        getattr(L, "set_vec3_soa_to_zero" + s).argtypes = [vp]
        getattr(L, "multiple_combine_in1_minus_in2x_factor_back_into_in1" + s).argtypes = [vp, vp, i, vp, vp]
        getattr(L, "multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1" + s).argtypes = [vp, i, vp, vp, vp, vp]
        getattr(L, "combine_in1_x_fact1_minus_in2_back_into_in2" + s).argtypes = [vp, d, vp]
        getattr(L, "combine_in1_minus_in2_allxfact" + s).argtypes = [vp, vp, d, vp]
        getattr(L, "calc_new_trialsol_for_inversion_in_force" + s).argtypes = [i, vp, i]
        getattr(L, "multishift_invert" + s).argtypes = [vp, vp, vp, vp, vp, d, vp, vp, vp, vp, vp, i, C.POINTER(i)]
        getattr(L, "multishift_invert" + s).restype = i
        getattr(L, "recombine_shifted_vec3_to_vec3" + s).argtypes = [vp, vp, vp, vp]
        getattr(L, "ker_invert_openacc" + s).argtypes = [vp, vp, vp, vp, d, vp, vp, vp, vp, i, d, C.POINTER(i)]
        getattr(L, "ker_invert_openacc" + s).restype = i
        getattr(L, "set_tamat_soa_to_zero" + s).argtypes = [vp]
        getattr(L, "set_su3_soa_to_zero" + s).argtypes = [vp]
        getattr(L, "direct_product_of_fermions_into_auxmat" + s).argtypes = [vp, vp, vp, vp, i]
        getattr(L, "multiply_conf_times_force_and_take_ta_nophase" + s).argtypes = [vp, vp, vp]
        getattr(L, "multiply_backfield_times_force" + s).argtypes = [vp, vp, vp]
        getattr(L, "accumulate_gl3soa_into_gl3soa" + s).argtypes = [vp, vp]
        getattr(L, "ker_openacc_compute_fermion_force" + s).argtypes = [vp, vp, vp, vp, vp, vp]
        getattr(L, "calc_loc_staples_nnptrick_all_onlyferms" + s).argtypes = [vp, vp]
        getattr(L, "RHO_times_conf_times_staples_ta_part" + s).argtypes = [vp, vp, vp, i]
        getattr(L, "exp_minus_QA_times_conf" + s).argtypes = [vp, vp, vp, vp]
        getattr(L, "stout_isotropic" + s).argtypes = [vp, vp, vp, vp, vp, i]
        getattr(L, "stout_wrapper" + s).argtypes = [vp, vp, i]
        getattr(L, "compute_lambda" + s).argtypes = [vp, vp, vp, vp, vp]
        getattr(L, "compute_sigma" + s).argtypes = [vp, vp, vp, vp, vp, i]
        getattr(L, "compute_sigma_from_sigma_prime_backinto_sigma_prime" + s).argtypes = [vp, vp, vp, vp, vp, i]
        for f in ("communicate_gl3_borders", "communicate_tamat_soa_borders", "communicate_thmat_soa_borders"):
            getattr(L, f + s).argtypes = [vp, i]
        getattr(L, "communicate_fermion_borders" + s).argtypes = [vp]
        getattr(L, "communicate_su3_borders" + s).argtypes = [vp, i]
    for f in ("convert_float_to_double_tamat_soa", "convert_double_to_float_tamat_soa", "convert_float_to_double_thmat_soa",
              "convert_double_to_float_thmat_soa", "convert_float_to_double_complex_soa", "convert_double_to_float_complex_soa",
              "convert_float_to_double_vec3", "convert_double_to_float_vec3"):
        getattr(L, f).argtypes = [vp, vp]; getattr(L, f).restype = None
    for f in ("convert_float_to_double_vec3_soa", "convert_double_to_float_vec3_soa", "convert_float_to_double_su3_soa",
              "convert_double_to_float_su3_soa", "convert_float_to_double_real_soa", "convert_double_to_float_real_soa",
              "combine_add_in2_into_in1_mixed_precision"):
        getattr(L, f).argtypes = [vp, vp]
    L.ker_find_max_eigenvalue_openacc.argtypes = [vp, vp, vp, vp, vp]; L.ker_find_max_eigenvalue_openacc.restype = d
    L.find_min_max_eigenvalue_soloopenacc.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(d)]
    L.staple_set_sp_globals.argtypes = [vp, vp]
    L.staple_last_solve_stats.argtypes = [C.POINTER(i), C.POINTER(C.c_longlong), C.POINTER(d)]
    L.communicate_fermion_borders_async.argtypes = [vp, vp, vp]
    L.communicate_su3_borders_async.argtypes = [vp, i, vp, vp]

/* A plain-C host program, as a maintainer of the reference would write it: includes the C ABI, links libstaple_b200.so with gcc
 * and uses the entry points under the reference's names.  Built and run by tests/test_abi_and_host.py WITHOUT a GPU: it only
 * calls the geometry planner (host arithmetic) and takes the address of the compute entry points so that the link step
 * has to resolve every one of them.  With -DWITH_REFERENCE_HEADERS the reference's own headers are included FIRST (the test
 * passes the reference's -D geometry macros): the two sets of declarations must be compatible. */
#ifdef WITH_REFERENCE_HEADERS
#define MULTIDEVICE
#include "OpenAcc/struct_c_def.h"
#include "OpenAcc/sp_struct_c_def.h"
#include "OpenAcc/fermion_matrix.h"
#include "OpenAcc/sp_fermion_matrix.h"
#include "OpenAcc/fermionic_utilities.h"
#include "OpenAcc/sp_fermionic_utilities.h"
#include "OpenAcc/inverter_multishift_full.h"
#include "OpenAcc/sp_inverter_multishift_full.h"
#include "OpenAcc/inverter_full.h"
#include "OpenAcc/sp_inverter_full.h"
#include "OpenAcc/inverter_mixedp.h"
#include "OpenAcc/inverter_wrappers.h"
#include "OpenAcc/inverter_package.h"
#include "OpenAcc/float_double_conv.h"
#include "OpenAcc/find_min_max.h"
#include "OpenAcc/fermion_force_utilities.h"
#include "OpenAcc/sp_fermion_force_utilities.h"
#include "OpenAcc/fermion_force.h"
#include "OpenAcc/sp_fermion_force.h"
#include "OpenAcc/stouting.h"
#include "OpenAcc/sp_stouting.h"
#include "OpenAcc/plaquettes.h"
#include "OpenAcc/su3_utilities.h"
#include "OpenAcc/field_times_fermion_matrix.h"
#include "OpenAcc/md_parameters.h"
#include "OpenAcc/action.h"
#include "Mpi/communications.h"
#include "Mpi/sp_communications.h"
#include "Mpi/multidev.h"
#include "Meas/ferm_meas.h"
#endif
#include <stdio.h>
#include "staple_b200.h"

typedef void (*fn)(void);
static fn table[] = {
	(fn) acc_Deo, (fn) acc_Doe, (fn) acc_Deo_f, (fn) acc_Doe_f, (fn) acc_Deo_unsafe, (fn) acc_Doe_bulk, (fn) acc_Deo_d3p, (fn) acc_Doe_d3m,
	(fn) acc_Deo_d3c, (fn) fermion_matrix_multiplication, (fn) fermion_matrix_multiplication_shifted, (fn) fermion_matrix_multiplication_shifted_f,
	(fn) scal_prod_global, (fn) real_scal_prod_global, (fn) l2norm2_global, (fn) l2norm2_global_f, (fn) combine_in1xfactor_plus_in2,
	(fn) multiple_combine_in1_minus_in2x_factor_back_into_in1, (fn) set_vec3_soa_to_zero, (fn) multishift_invert, (fn) multishift_invert_f,
	(fn) recombine_shifted_vec3_to_vec3, (fn) ker_invert_openacc, (fn) inverter_mixed_precision, (fn) inverter_multishift_wrapper,
	(fn) inverter_wrapper, (fn) setup_inverter_package_dp, (fn) setup_inverter_package_sp, (fn) convert_double_to_float_su3_soa,
	(fn) ker_find_max_eigenvalue_openacc, (fn) find_min_max_eigenvalue_soloopenacc, (fn) ker_openacc_compute_fermion_force,
	(fn) multiply_conf_times_force_and_take_ta_nophase, (fn) stout_wrapper, (fn) stout_isotropic, (fn) compute_lambda, (fn) compute_sigma,
	(fn) compute_sigma_from_sigma_prime_backinto_sigma_prime, (fn) fermion_force_soloopenacc, (fn) fermion_force_soloopenacc_f,
	(fn) eo_inversion, (fn) acc_Deo_wf, (fn) acc_Doe_wf, (fn) communicate_fermion_borders, (fn) communicate_su3_borders,
	(fn) communicate_fermion_borders_async, (fn) shutdown_multidev, (fn) staple_acc_update_device, (fn) staple_posix_memalign,
};

int main(void)
{
	/* the SURVEY's worked example: 8^3 x (2 ranks x 8), TLSM halo 2 -> nd3 = 12, sizeh = 3072 */
	const int loc_n[4] = { 8, 8, 8, 8 };
	long plan[16];
	if (staple_geometry_plan(loc_n, 2, 2, plan) != 0) return 1;
	int linked = 0;
	for (unsigned i = 0; i < sizeof(table) / sizeof(table[0]); i++) linked += table[i] != 0;
	printf("nd3 %ld sizeh %ld r0 %ld %ld r1 %ld %ld linked %d version %s\n", plan[3], plan[4], plan[6], plan[7], plan[8], plan[9], linked,
				 staple_version());
	inverter_package ip;               /* by-value struct of the reference API */
	ip.nshifts = 3;
	multishift_invert_iterations = 0;  /* exported global */
	return (int) sizeof(ip) == 0;
}

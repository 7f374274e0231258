// The two callers that drive the whole path from the reference's host program: the MD fermion force
// (OpenAcc/fermion_force.c:166-357) and the even/odd full-lattice inversion of the fermionic measurements
// (Meas/ferm_meas.c:50-72).  Host-side sequencing only: every step is an entry point of this library (stout smearing,
// CG-M, outer products, the Sigma' -> Sigma chain, TA projection), enqueued on the library stream; the only host
// synchronisations are the ones inside the solvers.
#include <sys/time.h>
#include "staple_internal.cuh"

using namespace staple;

static void need(const void *p, const char *fn, const char *what)
{
	if (p == nullptr) { fprintf(stderr, "%s: the global %s is not set\n", fn, what); exit(1); }
}

static double seconds_since(const timeval &t1)
{
	timeval t2; gettimeofday(&t2, nullptr);
	return (double) (t2.tv_sec - t1.tv_sec) + (double) (t2.tv_usec - t1.tv_usec) / 1.0e6;
}

// su3_soa[8] number `level` of the stout_conf_acc_arr container (fermion_force.c:200-201: &arr[8*(steps-1)])
template <typename SU3, typename C2>
static SU3 *stout_level(SU3 *arr, int level)
{
	return (SU3 *) ((char *) arr + (size_t) level * 8 * 9 * ctx().g.sizeh * sizeof(C2));
}

extern "C" {

// globals of the reference's host program (md_parameters.c, alloc_vars.c); weak: the host's own definitions win
__attribute__((weak)) md_param md_parameters = {};
__attribute__((weak)) int nMdInversionPerformed = 0;
__attribute__((weak)) thmat_soa *aux_th = nullptr;
__attribute__((weak)) tamat_soa *aux_ta = nullptr;
__attribute__((weak)) thmat_soa_f *aux_th_f = nullptr;
__attribute__((weak)) tamat_soa_f *aux_ta_f = nullptr;
__attribute__((weak)) su3_soa_f *conf_acc_f = nullptr;

// fermion_force.c:166-357
void fermion_force_soloopenacc(su3_soa *tconf_acc, su3_soa *tstout_conf_acc_arr, su3_soa *gl3_aux, tamat_soa *tipdot_acc,
															 ferm_param *tfermion_parameters, int tNDiffFlavs, const vec3_soa *ferm_in_acc, double res,
															 su3_soa *taux_conf_acc, vec3_soa *tferm_shiftmulti_acc, inverter_package ipt, const int max_cg)
{
	require_init("fermion_force_soloopenacc");
	if (verbosity_lv > 3) printf("DOUBLE PRECISION VERSION OF FERMION_FORCE_SOLOOPENACC\n");
	if (verbosity_lv > 2) printf("MPI%02d:\tCalculation of fermion force...\n", ctx().myrank);
	timeval t1; gettimeofday(&t1, nullptr);
	const size_t vbytes = sizeof(double2) * 3 * ctx().g.sizeh;

	su3_soa *conf_to_use;
	stout_wrapper(tconf_acc, tstout_conf_acc_arr, 0);                                                  // :197
	if (act_params.stout_steps > 0) conf_to_use = stout_level<su3_soa, double2>(tstout_conf_acc_arr, act_params.stout_steps - 1);
	else conf_to_use = tconf_acc;
	set_su3_soa_to_zero(gl3_aux);                                                                      // pseudo ipdot
	set_tamat_soa_to_zero(tipdot_acc);
	ipt.u = conf_to_use;
	if (1 == inverter_tricks.singlePInvAccelMultiInv || 1 == md_parameters.recycleInvsForce) {          // :212-219
		need(conf_acc_f, "fermion_force_soloopenacc", "conf_acc_f");
		if (0 == ctx().myrank && verbosity_lv > 2) printf("Converting gauge conf to single precision...\n");
		convert_double_to_float_su3_soa(conf_to_use, conf_acc_f);
		ipt.u_f = conf_acc_f;
	} else setup_inverter_package_sp(&ipt, 0, 0, 0, 0, 0, 0, 0, 0);                                    // by-value copy: this scope only

	for (int iflav = 0; iflav < tNDiffFlavs; iflav++) {
		set_su3_soa_to_zero(taux_conf_acc);
		const int ifps = tfermion_parameters[iflav].index_of_the_first_ps;
		for (int ips = 0; ips < tfermion_parameters[iflav].number_of_ps; ips++) {
			if (1 == md_parameters.recycleInvsForce && nMdInversionPerformed >= 2) {
				printf("ERROR, not implemented correctly! %s : %d", "fermion_force.c", 231); exit(1);       // :230-231
			}
			const vec3_soa *src = (const vec3_soa *) ((const char *) ferm_in_acc + (size_t) (ifps + ips) * vbytes);
			inverter_multishift_wrapper(ipt, &tfermion_parameters[iflav], &tfermion_parameters[iflav].approx_md, tferm_shiftmulti_acc,
																	src, res, max_cg, CONVERGENCE_NONCRITICAL);
			ker_openacc_compute_fermion_force(ipt.u, taux_conf_acc, tferm_shiftmulti_acc, ipt.loc_s, ipt.loc_h,
																				&tfermion_parameters[iflav]);
		}
		// staggered phases, back field and/or chemical potential
		multiply_backfield_times_force(&tfermion_parameters[iflav], taux_conf_acc, gl3_aux);
	}
	nMdInversionPerformed++;

	if (act_params.stout_steps > 0) need(aux_th, "fermion_force_soloopenacc", "aux_th"), need(aux_ta, "fermion_force_soloopenacc", "aux_ta");
	for (int lvl = act_params.stout_steps; lvl > 1; lvl--) {                                           // :275-292
		if (verbosity_lv > 1) printf("MPI%02d:\t\tSigma' to Sigma [lvl %d to lvl %d]\n", ctx().myrank, lvl, lvl - 1);
		conf_to_use = stout_level<su3_soa, double2>(tstout_conf_acc_arr, lvl - 2);
		compute_sigma_from_sigma_prime_backinto_sigma_prime(gl3_aux, aux_th, aux_ta, conf_to_use, taux_conf_acc, 0);
	}
	if (act_params.stout_steps > 0) {
		if (verbosity_lv > 1) printf("MPI%02d:\t\tSigma' to Sigma [lvl 1 to lvl 0]\n", ctx().myrank);
		compute_sigma_from_sigma_prime_backinto_sigma_prime(gl3_aux, aux_th, aux_ta, tconf_acc, taux_conf_acc, 0);
	}
	multiply_conf_times_force_and_take_ta_nophase(tconf_acc, gl3_aux, tipdot_acc);                     // :306
	if (verbosity_lv > 0) {
		// the reference times the (synchronous) OpenACC kernels with gettimeofday; here the work is enqueued, so the figure
		// is only meaningful after a synchronisation, done when somebody asked to see it
		STAPLE_CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
		printf("MPI%02d\t\tFULL FERMION FORCE COMPUTATION  PreKer->PostKer :%f sec  \n", ctx().myrank, seconds_since(t1));
		printf("MPI%02d:\t\tCompleted fermion force openacc\n", ctx().myrank);
	}
}

// generated sp_fermion_force.c:158-300
void fermion_force_soloopenacc_f(su3_soa_f *tconf_acc, su3_soa_f *tstout_conf_acc_arr, su3_soa_f *gl3_aux, tamat_soa_f *tipdot_acc,
																 ferm_param *tfermion_parameters, int tNDiffFlavs, const vec3_soa_f *ferm_in_acc, float res,
																 su3_soa_f *taux_conf_acc, vec3_soa_f *tferm_shiftmulti_acc, inverter_package ipt, const int max_cg)
{
	require_init("fermion_force_soloopenacc_f");
	if (verbosity_lv > 3) printf("SINGLE PRECISION VERSION OF FERMION_FORCE_SOLOOPENACC\n");
	if (verbosity_lv > 2) printf("MPI%02d:\tCalculation of fermion force...\n", ctx().myrank);
	timeval t1; gettimeofday(&t1, nullptr);
	const size_t vbytes = sizeof(float2) * 3 * ctx().g.sizeh;

	su3_soa_f *conf_to_use;
	stout_wrapper_f(tconf_acc, tstout_conf_acc_arr, 0);
	if (act_params.stout_steps > 0) conf_to_use = stout_level<su3_soa_f, float2>(tstout_conf_acc_arr, act_params.stout_steps - 1);
	else conf_to_use = tconf_acc;
	set_su3_soa_to_zero_f(gl3_aux);
	set_tamat_soa_to_zero_f(tipdot_acc);

	for (int iflav = 0; iflav < tNDiffFlavs; iflav++) {
		set_su3_soa_to_zero_f(taux_conf_acc);
		const int ifps = tfermion_parameters[iflav].index_of_the_first_ps;
		for (int ips = 0; ips < tfermion_parameters[iflav].number_of_ps; ips++) {
			int cg_return = 0;
			if (1 == md_parameters.recycleInvsForce && nMdInversionPerformed >= 2) {
				printf("ERROR, not implemented correctly! %s : %d", "sp_fermion_force.c", 208); exit(1);
			}
			const vec3_soa_f *src = (const vec3_soa_f *) ((const char *) ferm_in_acc + (size_t) (ifps + ips) * vbytes);
			const int converged = multishift_invert_f(conf_to_use, &tfermion_parameters[iflav], &tfermion_parameters[iflav].approx_md,
																								tferm_shiftmulti_acc, src, res, ipt.loc_r_f, ipt.loc_h_f, ipt.loc_s_f, ipt.loc_p_f,
																								ipt.ferm_shift_temp_f, max_cg, &cg_return);
			convergence_messages(CONVERGENCE_NONCRITICAL, converged);
			ker_openacc_compute_fermion_force_f(conf_to_use, taux_conf_acc, tferm_shiftmulti_acc, ipt.loc_s_f, ipt.loc_h_f,
																					&tfermion_parameters[iflav]);
		}
		multiply_backfield_times_force_f(&tfermion_parameters[iflav], taux_conf_acc, gl3_aux);
	}
	nMdInversionPerformed++;

	if (act_params.stout_steps > 0) need(aux_th_f, "fermion_force_soloopenacc_f", "aux_th_f"), need(aux_ta_f, "fermion_force_soloopenacc_f", "aux_ta_f");
	for (int lvl = act_params.stout_steps; lvl > 1; lvl--) {
		if (verbosity_lv > 1) printf("MPI%02d:\t\tSigma' to Sigma [lvl %d to lvl %d]\n", ctx().myrank, lvl, lvl - 1);
		conf_to_use = stout_level<su3_soa_f, float2>(tstout_conf_acc_arr, lvl - 2);
		compute_sigma_from_sigma_prime_backinto_sigma_prime_f(gl3_aux, aux_th_f, aux_ta_f, conf_to_use, taux_conf_acc, 0);
	}
	if (act_params.stout_steps > 0) {
		if (verbosity_lv > 1) printf("MPI%02d:\t\tSigma' to Sigma [lvl 1 to lvl 0]\n", ctx().myrank);
		compute_sigma_from_sigma_prime_backinto_sigma_prime_f(gl3_aux, aux_th_f, aux_ta_f, tconf_acc, taux_conf_acc, 0);
	}
	multiply_conf_times_force_and_take_ta_nophase_f(tconf_acc, gl3_aux, tipdot_acc);
	if (verbosity_lv > 0) {
		STAPLE_CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
		printf("MPI%02d\t\tFULL FERMION FORCE COMPUTATION  PreKer->PostKer :%f sec  \n", ctx().myrank, seconds_since(t1));
		printf("MPI%02d:\t\tCompleted fermion force openacc\n", ctx().myrank);
	}
}

// Meas/ferm_meas.c:50-72
void eo_inversion(inverter_package ip, ferm_param *tfermions_parameters, double res, int max_cg, vec3_soa *in_e, vec3_soa *in_o,
									vec3_soa *out_e, vec3_soa *out_o, vec3_soa *phi_e, vec3_soa *phi_o)
{
	require_init("eo_inversion");
	acc_Deo(ip.u, phi_e, in_o, tfermions_parameters->phases);
	combine_in1_x_fact1_minus_in2_back_into_in2(in_e, tfermions_parameters->ferm_mass, phi_e);
	set_vec3_soa_to_zero(out_e);
	inverter_wrapper(ip, tfermions_parameters, out_e, phi_e, res, max_cg, 0, CONVERGENCE_CRITICAL);
	acc_Doe(ip.u, phi_o, out_e, tfermions_parameters->phases);
	combine_in1_minus_in2_allxfact(in_o, phi_o, (double) 1 / tfermions_parameters->ferm_mass, out_o);
}

}   // extern "C"

/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Glue compiled TOGETHER WITH the unmodified OpenStaPLE hot-path sources (from
 * /root/reference, see oracle/build_ref.sh) into oracle/_ref/libref_<geom>.so.
 * It only (1) defines the few globals that the reference's main() programs
 * normally define, (2) offers small constructors for the reference's structs so
 * that the Python tests can call the reference functions (acc_Deo, acc_Doe,
 * fermion_matrix_multiplication, multishift_invert, ...) directly via ctypes,
 * and (3) provides a single-process "mailbox" MPI so that the reference's own
 * halo-exchange code (src/Mpi/communications.c:34-332) can be executed rank by
 * rank inside one process.  No arithmetic of the path is restated here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"
#include "OpenAcc/struct_c_def.h"
#include "OpenAcc/sp_struct_c_def.h"
#include "OpenAcc/geometry.h"
#include "Mpi/multidev.h"
#include "Include/fermion_parameters.h"
#include "Include/inverter_tricks.h"
#include "OpenAcc/backfield.h"
#include "OpenAcc/sp_backfield.h"
#include "Mpi/communications.h"
#include "OpenAcc/inverter_package.h"
#include "OpenAcc/inverter_wrappers.h"
#include "OpenAcc/inverter_mixedp.h"
#include "RationalApprox/rationalapprox.h"
#include "OpenAcc/action.h"
#include "OpenAcc/stouting.h"
#include "OpenAcc/alloc_settings.h"
#include "OpenAcc/io.h"
#include "Include/setting_file_parser.h"
#include "Include/debug.h"
#include "Include/montecarlo_parameters.h"
#include "OpenAcc/md_parameters.h"
#include "OpenAcc/fermion_force.h"
#include "OpenAcc/sp_fermion_force.h"
#include "OpenAcc/field_times_fermion_matrix.h"
#include "Meas/ferm_meas.h"

int verbosity_lv = 0;
vec3_soa_f *aux1_f = NULL;                 /* alloc_vars globals used by inverter_wrappers.c:60-71 */
vec3_soa_f *ferm_shiftmulti_acc_f = NULL;

/* globals that main.c / alloc_vars.c define in the reference's own programs and stouting.c reads (stouting.c:27-72) */
action_param act_params;
su3_soa *auxbis_conf_acc = NULL, *glocal_staples = NULL;
tamat_soa *gipdot = NULL;
su3_soa_f *auxbis_conf_acc_f = NULL, *glocal_staples_f = NULL;
tamat_soa_f *gipdot_f = NULL;
/* only called for verbosity_lv > 4 (stouting.c:42-46); su3_measurements.c is not part of this build */
void check_unitarity_device(const su3_soa *u, double *mx, double *avg) { *mx = 0; *avg = 0; }
void check_unitarity_device_f(const su3_soa_f *u, double *mx, double *avg) { *mx = 0; *avg = 0; }

/* stout parameters as main.c:276-277 sets them, and the parking arrays stout_wrapper takes from alloc_vars */
void ref_set_stout(double rho, int steps, void *auxbis, void *staples, void *ipdot, void *auxbis_f, void *staples_f,
									 void *ipdot_f)
{
	act_params.stout_rho = rho; act_params.stout_steps = steps; act_params.topo_action = 0;
	gl_stout_rho = rho; gl_topo_rho = rho;
	auxbis_conf_acc = (su3_soa *) auxbis; glocal_staples = (su3_soa *) staples; gipdot = (tamat_soa *) ipdot;
	auxbis_conf_acc_f = (su3_soa_f *) auxbis_f; glocal_staples_f = (su3_soa_f *) staples_f; gipdot_f = (tamat_soa_f *) ipdot_f;
}

/* globals read by io.c's writers (io.c:56-66, :498): no flavours, empty embedded input file */
alloc_settings alloc_info;
ferm_param *fermions_parameters = NULL;
char input_file_str[MAXLINES * MAXLINELENGTH];
void ref_set_io(double beta, const char *input_file)
{
	act_params.beta = beta; alloc_info.NDiffFlavs = 0;
	strncpy(input_file_str, input_file, sizeof(input_file_str) - 1);
}

void ref_geometry(int *o)
{
	o[0] = LOC_N0; o[1] = LOC_N1; o[2] = LOC_N2; o[3] = LOC_N3;
	o[4] = NRANKS_D3;
	o[5] = nd0; o[6] = nd1; o[7] = nd2; o[8] = nd3;
	o[9] = sizeh; o[10] = D3_HALO; o[11] = GL_SIZEH;
}


/* sizeof/offsetof of every struct that crosses the C ABI, as the reference's own headers define them */
#include <stddef.h>
void ref_abi(long *o)
{
	o[0] = sizeof(vec3_soa); o[1] = sizeof(su3_soa); o[2] = sizeof(double_soa);
	o[3] = sizeof(vec3_soa_f); o[4] = sizeof(su3_soa_f); o[5] = sizeof(float_soa);
	o[6] = sizeof(ferm_param); o[7] = offsetof(ferm_param, ferm_mass); o[8] = offsetof(ferm_param, phases);
	o[9] = offsetof(ferm_param, phases_f); o[10] = offsetof(ferm_param, approx_md);
	o[11] = sizeof(RationalApprox); o[12] = offsetof(RationalApprox, approx_order);
	o[13] = offsetof(RationalApprox, RA_a0); o[14] = offsetof(RationalApprox, RA_a); o[15] = offsetof(RationalApprox, RA_b);
	o[16] = sizeof(inverter_package); o[17] = offsetof(inverter_package, nshifts);
	o[18] = offsetof(inverter_package, loc_r); o[19] = offsetof(inverter_package, out_f);
	o[20] = sizeof(inv_tricks); o[21] = offsetof(inv_tricks, mixedPrecisionDelta);
	o[22] = offsetof(vec3_soa, c1); o[23] = offsetof(su3_soa, r1);
}

/* the globals' structs the library reads from the HOST program's own definitions (action.h:6-18, md_parameters.h:6-19) */
void ref_abi2(long *o)
{
	o[0] = sizeof(action_param); o[1] = offsetof(action_param, stout_steps); o[2] = offsetof(action_param, stout_rho);
	o[3] = offsetof(action_param, topo_action); o[4] = offsetof(action_param, topo_file_path);
	o[5] = offsetof(action_param, topo_stout_steps); o[6] = offsetof(action_param, topo_rho);
	o[7] = sizeof(md_param); o[8] = offsetof(md_param, residue_metro); o[9] = offsetof(md_param, singlePrecMD);
	o[10] = offsetof(md_param, max_cg_iterations); o[11] = offsetof(md_param, recycleInvsForce);
	o[12] = sizeof(tamat_soa); o[13] = offsetof(tamat_soa, ic00); o[14] = sizeof(thmat_soa); o[15] = offsetof(thmat_soa, rc00);
}

/* index helpers straight from the reference's header-only geometry (geometry_multidev.h:209-262) */
int ref_snum_acc(int d0, int d1, int d2, int d3) { return snum_acc(d0, d1, d2, d3); }
int ref_lnh_to_gl_snum(int d0, int d1, int d2, int d3, int rank)
{ return lnh_to_gl_snum(d0, d1, d2, d3, xyzt_rank(rank)); }
/* loop bounds exactly as spelled in fermionic_utilities.c:41 (reductions) and :188 (updates) */
void ref_ranges(long *o)
{
#if NRANKS_D3 > 1
	o[0] = (LNH_SIZEH-LOC_SIZEH)/2; o[1] = (LNH_SIZEH+LOC_SIZEH)/2;
#else
	o[0] = 0; o[1] = sizeh;
#endif
	o[2] = LNH_VOL3/2*(D3_HALO-D3_FERMION_HALO); o[3] = LNH_SIZEH-LNH_VOL3/2*(D3_HALO-D3_FERMION_HALO);
}

/* Select which rank of the D3 ring this process currently "is". */
int ref_setup(int rank)
{
	geom_par.gnx = GL_N0; geom_par.gny = GL_N1; geom_par.gnz = GL_N2; geom_par.gnt = GL_N3;
	geom_par.xmap = 0; geom_par.ymap = 1; geom_par.zmap = 2; geom_par.tmap = 3;
	devinfo.myrank = rank; devinfo.nranks = NRANKS_D3;
	devinfo.myrank_world = rank; devinfo.nranks_world = NRANKS_D3;
	devinfo.replica_idx = 0; devinfo.num_replicas = 1;
	int res = set_geom_glv(&geom_par);
	compute_nnp_and_nnm_openacc();      /* neighbour tables used by the fermion-force outer products */
#ifdef MULTIDEVICE
	devinfo.mpi_comm = 0;
	devinfo.async_comm_fermion = 0; devinfo.async_comm_gauge = 0;
	devinfo.proc_per_node = NRANKS_D3;
	devinfo.namelen = 0; devinfo.processor_name[0] = 0;
	init_multidev1D(&devinfo);
#endif
	return res;
}

void ref_set_async_fermion_comms(int on)
{
#ifdef MULTIDEVICE
	devinfo.async_comm_fermion = on;
#endif
}

void ref_set_verbosity(int v) { verbosity_lv = v; }

ferm_param *ref_ferm_param_new(double mass, double_soa *phases, float_soa *phases_f)
{
	ferm_param *p = (ferm_param *) calloc(1, sizeof(ferm_param));
	p->ferm_mass = mass; p->phases = phases; p->phases_f = phases_f;
	p->degeneracy = 1; p->number_of_ps = 1; strcpy(p->name, "oracle");
	return p;
}

/* approx_md of a ferm_param (read by ker_openacc_compute_fermion_force, fermion_force_utilities.c:195) */
void ref_ferm_param_set_md(ferm_param *p, int order, const double *a, const double *b)
{
	p->approx_md.approx_order = order;
	for (int i = 0; i < order; i++) { p->approx_md.RA_a[i] = a[i]; p->approx_md.RA_b[i] = b[i]; }
}

RationalApprox *ref_approx_new(int order, double a0, const double *a, const double *b)
{
	RationalApprox *r = (RationalApprox *) calloc(1, sizeof(RationalApprox));
	r->approx_order = order; r->RA_a0 = a0;
	for (int i = 0; i < order; i++) { r->RA_a[i] = a[i]; r->RA_b[i] = b[i]; }
	r->exponent_num = -1; r->exponent_den = 4; r->lambda_min = 0; r->lambda_max = 1;
	return r;
}

int ref_approx_read(RationalApprox *r, char *filename)
{
	return rationalapprox_read_custom_nomefile(r, filename);
}

void ref_approx_get(const RationalApprox *r, int *order, double *a0, double *a, double *b,
										int *num, int *den, double *lmin, double *lmax)
{
	*order = r->approx_order; *a0 = r->RA_a0; *num = r->exponent_num; *den = r->exponent_den;
	*lmin = r->lambda_min; *lmax = r->lambda_max;
	for (int i = 0; i < r->approx_order; i++) { a[i] = r->RA_a[i]; b[i] = r->RA_b[i]; }
}

void ref_set_inverter_tricks(int singlePInvAccelMultiInv, int useMixedPrecision,
														 double mixedPrecisionDelta, int restartingEvery)
{
	inverter_tricks.singlePInvAccelMultiInv = singlePInvAccelMultiInv;
	inverter_tricks.useMixedPrecision = useMixedPrecision;
	inverter_tricks.mixedPrecisionDelta = mixedPrecisionDelta;
	inverter_tricks.restartingEvery = restartingEvery;
}

void ref_set_sp_globals(vec3_soa_f *aux1, vec3_soa_f *shiftmulti)
{
	aux1_f = aux1; ferm_shiftmulti_acc_f = shiftmulti;
}

void ref_phases(double_soa *ph, double ex, double ey, double ez, double bx, double by, double bz,
								double im_chem_pot, double charge)
{
	bf_param b = { ex, ey, ez, bx, by, bz };
	calc_u1_phases(ph, b, im_chem_pot, charge);
}

void ref_phases_f(float_soa *ph, double ex, double ey, double ez, double bx, double by, double bz,
									double im_chem_pot, double charge)
{
	bf_param b = { ex, ey, ez, bx, by, bz };
	calc_u1_phases_f(ph, b, (float) im_chem_pot, (float) charge);
}

/* inverter_package travels by value in the reference API (inverter_package.h:12-29). */
static inverter_package the_ip;
void ref_ip_dp(su3_soa *u, vec3_soa *shift_temp, int nshift, vec3_soa *r, vec3_soa *h,
							 vec3_soa *s, vec3_soa *p)
{ setup_inverter_package_dp(&the_ip, u, shift_temp, nshift, r, h, s, p); }
void ref_ip_sp(su3_soa_f *u, vec3_soa_f *shift_temp, int nshift, vec3_soa_f *r, vec3_soa_f *h,
							 vec3_soa_f *s, vec3_soa_f *p, vec3_soa_f *out)
{ setup_inverter_package_sp(&the_ip, u, shift_temp, nshift, r, h, s, p, out); }
int ref_inverter_multishift_wrapper(ferm_param *pars, RationalApprox *approx, vec3_soa *out,
																		const vec3_soa *in, double res, int max_cg, int importance)
{ return inverter_multishift_wrapper(the_ip, pars, approx, out, in, res, max_cg, importance); }
int ref_inverter_wrapper(ferm_param *pars, vec3_soa *out, const vec3_soa *in, double res,
												 int max_cg, double shift, int importance)
{ return inverter_wrapper(the_ip, pars, out, in, res, max_cg, shift, importance); }
int ref_inverter_mixed_precision(ferm_param *pars, vec3_soa *solution, const vec3_soa *in,
																 double res, int max_cg, double shift, int *cg_return)
{ return inverter_mixed_precision(the_ip, pars, solution, in, res, max_cg, shift, cg_return); }

/* ---- the two callers of the path: fermion_force_soloopenacc (OpenAcc/fermion_force.c:166-357) and eo_inversion
 * (Meas/ferm_meas.c:50-72).  Globals that md_parameters.c / alloc_vars.c / debug.c / main.c define in the reference's
 * own programs; the dbg/diagnostic hooks are never reached with debug_settings zeroed (fermion_force.c:258,325). */
md_param md_parameters;
int nMdInversionPerformed = 0;
debug_settings_t debug_settings;
int md_dbg_print_count = 0, md_diag_count_fermion = 0, ipdot_f_reset = 0;
mc_params_t mc_params;
tamat_soa *aux_ta = NULL, *ipdot_f_old = NULL; thmat_soa *aux_th = NULL;
tamat_soa_f *aux_ta_f = NULL, *ipdot_f_old_f = NULL; thmat_soa_f *aux_th_f = NULL;
su3_soa_f *conf_acc_f = NULL;
su3_soa *gstout_conf_acc_arr = NULL;
vec3_soa *ferm_shiftmulti_acc = NULL, *kloc_r = NULL, *kloc_h = NULL, *kloc_s = NULL, *kloc_p = NULL;
vec3_soa_f *kloc_r_f = NULL, *kloc_h_f = NULL, *kloc_s_f = NULL, *kloc_p_f = NULL;
static void unreachable(const char *what) { printf("oracle shim: %s is not part of this build\n", what); exit(1); }
void dbg_print_su3_soa(su3_soa *const c, const char *n, int i) { unreachable("dbg_print_su3_soa"); }
void dbg_print_su3_soa_f(su3_soa_f *const c, const char *n, int i) { unreachable("dbg_print_su3_soa_f"); }
double calc_force_norm(const tamat_soa *t) { unreachable("calc_force_norm"); return 0; }
double calc_diff_force_norm(const tamat_soa *t, const tamat_soa *o) { unreachable("calc_diff_force_norm"); return 0; }
void copy_ipdot_into_old(const tamat_soa *t, tamat_soa *o) { unreachable("copy_ipdot_into_old"); }
float calc_force_norm_f(const tamat_soa_f *t) { unreachable("calc_force_norm_f"); return 0; }
float calc_diff_force_norm_f(const tamat_soa_f *t, const tamat_soa_f *o) { unreachable("calc_diff_force_norm_f"); return 0; }
void copy_ipdot_into_old_f(const tamat_soa_f *t, tamat_soa_f *o) { unreachable("copy_ipdot_into_old_f"); }
void generate_vec3_soa_gauss(vec3_soa *const v) { unreachable("generate_vec3_soa_gauss"); }
void generate_vec3_soa_z2noise(vec3_soa *const v) { unreachable("generate_vec3_soa_z2noise"); }

void ref_set_force_globals(void *th, void *ta, void *th_f, void *ta_f, void *conf_f)
{
	aux_th = (thmat_soa *) th; aux_ta = (tamat_soa *) ta; aux_th_f = (thmat_soa_f *) th_f; aux_ta_f = (tamat_soa_f *) ta_f;
	conf_acc_f = (su3_soa_f *) conf_f;
	memset(&md_parameters, 0, sizeof(md_parameters)); memset(&debug_settings, 0, sizeof(debug_settings));
	nMdInversionPerformed = 0;
}
int ref_md_inversions_performed(void) { return nMdInversionPerformed; }

/* an array of flavours as fermion_force_soloopenacc walks it (fermion_parameters.h:9-41) */
ferm_param *ref_ferm_param_array_new(int n) { return (ferm_param *) calloc(n, sizeof(ferm_param)); }
void ref_ferm_param_array_set(ferm_param *arr, int i, double mass, double_soa *phases, float_soa *phases_f, int number_of_ps,
															int index_of_the_first_ps, int order, const double *a, const double *b)
{
	ferm_param *p = &arr[i];
	p->ferm_mass = mass; p->phases = phases; p->phases_f = phases_f; p->degeneracy = 1; sprintf(p->name, "flav%d", i);
	p->number_of_ps = number_of_ps; p->index_of_the_first_ps = index_of_the_first_ps;
	p->approx_md.approx_order = order; p->approx_md.RA_a0 = 0;
	for (int k = 0; k < order; k++) { p->approx_md.RA_a[k] = a[k]; p->approx_md.RA_b[k] = b[k]; }
}
void ref_fermion_force(su3_soa *conf, su3_soa *stout_arr, su3_soa *gl3_aux, tamat_soa *ipdot, ferm_param *pars, int nflav,
											 const vec3_soa *ferm_in, double res, su3_soa *taux, vec3_soa *shiftmulti, int max_cg)
{ fermion_force_soloopenacc(conf, stout_arr, gl3_aux, ipdot, pars, nflav, ferm_in, res, taux, shiftmulti, the_ip, max_cg); }
void ref_fermion_force_f(su3_soa_f *conf, su3_soa_f *stout_arr, su3_soa_f *gl3_aux, tamat_soa_f *ipdot, ferm_param *pars, int nflav,
												 const vec3_soa_f *ferm_in, double res, su3_soa_f *taux, vec3_soa_f *shiftmulti, int max_cg)
{ fermion_force_soloopenacc_f(conf, stout_arr, gl3_aux, ipdot, pars, nflav, ferm_in, (float) res, taux, shiftmulti, the_ip, max_cg); }
void ref_eo_inversion(ferm_param *pars, double res, int max_cg, vec3_soa *in_e, vec3_soa *in_o, vec3_soa *out_e, vec3_soa *out_o,
											vec3_soa *phi_e, vec3_soa *phi_o)
{ eo_inversion(the_ip, pars, res, max_cg, in_e, in_o, out_e, out_o, phi_e, phi_o); }

#ifdef MULTIDEVICE
/* ---- single-process mailbox MPI ------------------------------------------------
 * Sends are appended to a FIFO keyed by (src,dst,tag) -- MPI's non-overtaking order: the reference re-uses tags 0-5 for every
 * component array of a field; receives take the OLDEST matching message if one is there, otherwise they are left pending
 * (nonblocking) or skipped (blocking).  Running the reference's exchange routine twice for every rank therefore completes
 * all transfers with the reference's own offsets, counts, tags and neighbour ranks (the second pass re-sends the same
 * interior slices, which the exchange never modifies).  ref_mailbox_clear() before every exchange drops what is left. */
#define MB_MAX 8192
typedef struct { int src, dst, tag, live; size_t bytes; void *data; } mb_msg;
static mb_msg mb[MB_MAX]; static int mb_n = 0;
typedef struct { int src, tag; size_t bytes; void *dst; int live; } mb_pend;
static mb_pend pend[MB_MAX]; static int pend_n = 0;
static long mb_missing = 0;

void ref_mailbox_clear(void)
{
	for (int i = 0; i < mb_n; i++) free(mb[i].data);
	mb_n = 0; pend_n = 0; mb_missing = 0;
}
long ref_mailbox_missing(void) { long m = mb_missing; mb_missing = 0; return m; }

static size_t dtsize(MPI_Datatype t) { return (size_t) t; }
static void mb_post(const void *buf, size_t bytes, int dst, int tag)
{
	if (mb_n == MB_MAX) { printf("oracle mailbox full\n"); exit(1); }
	mb[mb_n].src = devinfo.myrank; mb[mb_n].dst = dst; mb[mb_n].tag = tag; mb[mb_n].bytes = bytes; mb[mb_n].live = 1;
	mb[mb_n].data = malloc(bytes); memcpy(mb[mb_n].data, buf, bytes); mb_n++;
}
static int mb_fetch(void *buf, size_t bytes, int src, int tag)
{
	for (int i = 0; i < mb_n; i++)
		if (mb[i].live && mb[i].src == src && mb[i].dst == devinfo.myrank && mb[i].tag == tag) {
			if (mb[i].bytes != bytes) { printf("oracle mailbox size mismatch\n"); exit(1); }
			memcpy(buf, mb[i].data, bytes); mb[i].live = 0; return 1;
		}
	mb_missing++; return 0;
}
int MPI_Init(int *a, char ***b) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int e) { exit(e); }
int MPI_Barrier(MPI_Comm c) { return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { *r = devinfo.myrank; return 0; }
int MPI_Comm_size(MPI_Comm c, int *n) { *n = NRANKS_D3; return 0; }
int MPI_Comm_split(MPI_Comm c, int a, int b, MPI_Comm *o) { *o = 0; return 0; }
int MPI_Get_processor_name(char *n, int *l) { n[0] = 0; *l = 0; return 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int r, MPI_Comm c) { return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ memcpy(r, s, n * dtsize(t)); return 0; }   /* local contribution only; harness sums over ranks */
int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int dst, int stag, void *rb, int rn,
								 MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status *s)
{ mb_post(sb, sn * dtsize(st), dst, stag); mb_fetch(rb, rn * dtsize(rt), src, rtag); return 0; }
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ mb_post(b, n * dtsize(t), dst, tag); return 0; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ mb_fetch(b, n * dtsize(t), src, tag); return 0; }
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq)
{ mb_post(b, n * dtsize(t), dst, tag); *rq = -1; return 0; }
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq)
{
	if (pend_n == MB_MAX) pend_n = 0;
	pend[pend_n].src = src; pend[pend_n].tag = tag; pend[pend_n].bytes = n * dtsize(t);
	pend[pend_n].dst = b; pend[pend_n].live = 1; *rq = pend_n++; return 0;
}
int MPI_Wait(MPI_Request *rq, MPI_Status *s)
{
	if (*rq >= 0 && pend[*rq].live) {
		mb_fetch(pend[*rq].dst, pend[*rq].bytes, pend[*rq].src, pend[*rq].tag); pend[*rq].live = 0;
	}
	return 0;
}
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *s)
{ for (int i = 0; i < n; i++) MPI_Wait(&rq[i], s); return 0; }
#else
int MPI_Abort(MPI_Comm c, int e) { exit(e); }
#endif

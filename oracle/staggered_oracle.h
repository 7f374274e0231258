/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the OpenStaPLE fermion-solver hot
 * path (plain C99, single thread, run-time geometry).  Nothing in the product
 * (openstaple_b200/, libstaple_b200.so) includes, links or calls this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here against the reference's own gcc build (oracle/_ref, built from
 * /root/reference by oracle/build_ref.sh) whenever that is present, and against the
 * committed outputs of that build in tests/golden/ otherwise.
 *
 * Array conventions (identical to the reference ABI, struct_c_def.h:16-42):
 *   vector   v[c*sizeh + i]               c=0..2 colour, complex
 *   links    u[((k*3 + r)*3 + c)*sizeh+i] k=2*dir+parity, r=row (0,1 read; 2 ignored), c=col
 *   phases   ph[k*sizeh + i]              angle in radians
 *   arrays of vectors: element j at j*3*sizeh
 */
#ifndef STAGGERED_ORACLE_H_
#define STAGGERED_ORACLE_H_
#include <complex.h>

#define SO_MAX_APPROX_ORDER 25     /* rationalapprox.h:8 */

typedef struct so_geom_t {
	int loc_n[4];       /* LOC_N0..3 */
	int nranks_d3;      /* NRANKS_D3 */
	int halo_width;     /* HALO_WIDTH: 2 (TLSM) or 1 (Wilson) */
	int d3_halo;        /* D3_HALO: halo_width if nranks_d3>1 else 0 */
	int d3_fhalo;       /* D3_FERMION_HALO: 1 if nranks_d3>1 else 0 */
	int nd[4];          /* LNH_N0..3 */
	long vol3h;         /* nd0*nd1*nd2/2 */
	long sizeh;         /* nd0*nd1*nd2*nd3/2 */
	long r0_lo, r0_hi;  /* reduction range (fermionic_utilities.c:41,127) */
	long r1_lo, r1_hi;  /* update range    (fermionic_utilities.c:188) */
	int gl_n[4];        /* global lattice */
} so_geom;

void so_geom_init(so_geom *g, int n0, int n1, int n2, int n3, int nranks_d3, int halo_width);
long so_snum(const so_geom *g, int d0, int d1, int d2, int d3);
long so_lnh_to_gl_snum(const so_geom *g, int d0, int d1, int d2, int d3, int rank);

/* staggered + antiperiodic + U(1) phase angles, backfield.c:20-187 (identity xyzt map) */
void so_calc_u1_phases(const so_geom *g, int rank, double *ph, const double ebfield[6],
											 double im_chem_pot, double charge);
void so_calc_u1_phases_f(const so_geom *g, int rank, float *ph, const double ebfield[6],
												 double im_chem_pot, double charge);

/* global field -> rank-local box including halos, and back (communications.c:1104-1257) */
void so_scatter_vec(const so_geom *g, int rank, const double complex *gl, double complex *lnh);
void so_gather_vec(const so_geom *g, int rank, double complex *gl, const double complex *lnh);
void so_scatter_conf(const so_geom *g, int rank, const double complex *gl, double complex *lnh);
/* fermion halo exchange between the rank-local boxes of one vector (communications.c:34-104) */
void so_exchange_halo(const so_geom *g, double complex **ranks, int ncomp_arrays, int thickness);

#define SO_DECL(R, C, S) \
void so_deo##S(const so_geom *g, const C *u, C *out, const C *in, const R *ph, int d3lo, int d3hi); \
void so_doe##S(const so_geom *g, const C *u, C *out, const C *in, const R *ph, int d3lo, int d3hi); \
void so_fermion_matrix_multiplication_shifted##S(const so_geom *g, const C *u, C *out, const C *in, \
		C *tmp, const R *ph, double mass, double shift); \
double complex so_scal_prod##S(const so_geom *g, const C *a, const C *b); \
double so_real_scal_prod##S(const so_geom *g, const C *a, const C *b); \
double so_l2norm2##S(const so_geom *g, const C *a); \
void so_axpy_like##S(const so_geom *g, int op, C *out, const C *a, const C *b, const C *c, double f1, double f2); \
int so_multishift_invert##S(const so_geom *g, const C *u, const R *ph, double mass, int order, \
		const double *shifts, C *out, const C *in, double residuo, C *r, C *h, C *s, C *p, C *ps, \
		int max_cg, int *cg_return, double *true_rel_res2); \
void so_recombine##S(const so_geom *g, const C *in_shifted, const C *in, C *out, int order, \
		double a0, const double *a); \
int so_cg##S(const so_geom *g, const C *u, const R *ph, double mass, C *solution, const C *in, \
		double res, C *r, C *h, C *s, C *p, int max_cg, double shift, int restarting_every, int *cg_return); \
void so_direct_product_of_fermions_into_auxmat##S(const so_geom *g, const C *s, const C *h, C *aux, double a); \
void so_compute_fermion_force##S(const so_geom *g, const C *u, C *aux, const C *in_shiftmulti, C *s, C *h, \
		const R *ph, int order, const double *ra_a); \
void so_multiply_backfield_times_force##S(const so_geom *g, const R *ph, const C *aux, C *pseudo); \
void so_accumulate_gl3soa_into_gl3soa##S(const so_geom *g, const C *aux, C *pseudo); \
void so_multiply_conf_times_force_and_take_ta_nophase##S(const so_geom *g, const C *u, const C *aux, R *ta); \
void so_calc_loc_staples_onlyferms##S(const so_geom *g, const C *u, C *stap); \
void so_rho_times_conf_times_staples_ta_part##S(const so_geom *g, const C *u, const C *stap, R *ta, double rho); \
void so_exp_minus_QA_times_conf##S(const so_geom *g, const C *u, const R *ta, C *uout, C *expaux); \
void so_stout_isotropic##S(const so_geom *g, const C *u, C *uprime, C *stap, C *aux, R *ta, double rho); \
void so_compute_lambda##S(const so_geom *g, R *lam, const C *sp, const C *u, const R *ta, C *tmp); \
void so_compute_sigma##S(const so_geom *g, const R *lam, const C *u, C *sg, const R *ta, C *tmp, double rho);
SO_DECL(double, double complex, )
SO_DECL(float, float complex, _f)

/* operator "with a field" (field_times_fermion_matrix.c:77-196), FP64 only */
void so_deo_wf(const so_geom *g, const double complex *u, double complex *out, const double complex *in, const double *ph,
							 const double *fre, const double *fim, int d3lo, int d3hi);
void so_doe_wf(const so_geom *g, const double complex *u, double complex *out, const double complex *in, const double *ph,
							 const double *fre, const double *fim, int d3lo, int d3hi);

void so_convert_d2f(long n, const double complex *d, float complex *f);
void so_convert_f2d(long n, const float complex *f, double complex *d);

int so_inverter_mixed_precision(const so_geom *g, const double complex *u, const float complex *u_f,
		const double *ph, const float *ph_f, double mass, double complex *solution,
		const double complex *in, double res, int max_cg, double shift, double mixed_delta,
		double complex *r, double complex *h, double complex *s,
		float complex *r_f, float complex *h_f, float complex *s_f, float complex *p_f,
		float complex *out_f, int *cg_return, int *magic_touches);

double so_find_max_eigenvalue(const so_geom *g, const double complex *u, const double *ph, double mass,
		double complex *r, double complex *h, double complex *p, int *loops);

/* ops for so_axpy_like (all over the update range R1, fermionic_utilities.c:180-455) */
enum { SO_IN1XFACTOR_PLUS_IN2 = 0, /* out = a*f1 + b            */
       SO_SCALE,                   /* out = f1*out              */
       SO_ADD_FACTOR_X_IN2,        /* out += f1*a               */
       SO_IN1XMASS2_MINUS_IN2_MINUS_IN3, /* out = a*f1 - b - c  */
       SO_IN1XMASS_MINUS_IN2,      /* out = a*f1 - out          */
       SO_IN1_MINUS_IN2,           /* out = a - b               */
       SO_ASSIGN,                  /* out = a                   */
       SO_ZERO,                    /* out = 0 over [0,sizeh)    */
       SO_FACT1_MINUS_IN2,         /* out = f1*a - out          */
       SO_IN1_MINUS_IN2_ALLXFACT   /* out = f1*(a-b)            */ };
#endif

/* TEST INFRASTRUCTURE ONLY -- see staggered_oracle.h.  CPU restatement (plain C99)
 * of OpenStaPLE's staggered fermion-solver hot path; every function cites the
 * reference lines it follows.  Build: gcc -O3 -std=gnu99 -fPIC -shared (oracle/Makefile). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "staggered_oracle.h"

/* geometry.h:12-29, geometry_multidev.h:6-148 */
void so_geom_init(so_geom *g, int n0, int n1, int n2, int n3, int nranks_d3, int halo_width)
{
	g->loc_n[0] = n0; g->loc_n[1] = n1; g->loc_n[2] = n2; g->loc_n[3] = n3;
	g->nranks_d3 = nranks_d3; g->halo_width = halo_width;
	g->d3_halo = nranks_d3 > 1 ? halo_width : 0;
	g->d3_fhalo = nranks_d3 > 1 ? 1 : 0;
	g->nd[0] = n0; g->nd[1] = n1; g->nd[2] = n2; g->nd[3] = n3 + 2 * g->d3_halo;
	g->vol3h = (long) n0 * n1 * n2 / 2;
	g->sizeh = g->vol3h * g->nd[3];
	long loc_sizeh = g->vol3h * n3;
	if (nranks_d3 > 1) { g->r0_lo = (g->sizeh - loc_sizeh) / 2; g->r0_hi = (g->sizeh + loc_sizeh) / 2; }
	else { g->r0_lo = 0; g->r0_hi = g->sizeh; }
	g->r1_lo = g->vol3h * (g->d3_halo - g->d3_fhalo);
	g->r1_hi = g->sizeh - g->r1_lo;
	g->gl_n[0] = n0; g->gl_n[1] = n1; g->gl_n[2] = n2; g->gl_n[3] = n3 * nranks_d3;
}

/* geometry_multidev.h:219-229 */
long so_snum(const so_geom *g, int d0, int d1, int d2, int d3)
{
	return ((long) d0 + (long) g->nd[0] * (d1 + (long) g->nd[1] * (d2 + (long) g->nd[2] * d3))) / 2;
}

/* geometry_multidev.h:246-262 (only direction 3 is ever decomposed, :23-27) */
long so_lnh_to_gl_snum(const so_geom *g, int d0, int d1, int d2, int d3, int rank)
{
	int g3 = d3 + g->loc_n[3] * rank - g->d3_halo;
	g3 %= g->gl_n[3]; if (g3 < 0) g3 += g->gl_n[3];
	return ((long) d0 + (long) g->gl_n[0] * (d1 + (long) g->gl_n[1] * (d2 + (long) g->gl_n[2] * g3))) / 2;
}

/* backfield.c:20-187 with xmap..tmap = 0,1,2,3 */
#define SO_PHASES_BODY(RT, TWOPI, HALFC, ONEC) \
	const long n = g->sizeh; \
	const int tnx = g->gl_n[0], tny = g->gl_n[1], tnz = g->gl_n[2], tnt = g->gl_n[3]; \
	RT ex = eb[0], ey = eb[1], ez = eb[2], bx = eb[3], by = eb[4], bz = eb[5]; \
	RT chpotphase = (RT) im_chem_pot / tnt, q = (RT) charge; \
	for (int d3 = 0; d3 < g->nd[3]; d3++) for (int d2 = 0; d2 < g->nd[2]; d2++) \
	for (int d1 = 0; d1 < g->nd[1]; d1++) for (int d0 = 0; d0 < g->nd[0]; d0++) { \
		long idxh = so_snum(g, d0, d1, d2, d3); \
		int x = d0, y = d1, z = d2, t = d3; \
		if (g->nranks_d3 > 1) { \
			t += rank * g->loc_n[3] - g->d3_halo; \
			if (t > tnt - 1) t -= tnt; \
			if (t < 0) t += tnt; \
		} \
		int parity = (x + y + z + t) % 2; \
		RT arg; \
		arg = (z - tnz / 2 + 1) * by / (tnz * tnx); \
		if (x + 1 == tnx) { arg -= (y - tny / 2 + 1) * tnx * bz / (tnx * tny); arg -= (t - tnt / 2 + 1) * tnx * ex / (tnx * tnt); } \
		arg *= q; \
		ph[(0 + parity) * n + idxh] = arg; \
		arg = (x - tnx / 2 + 1) * bz / (tnx * tny); \
		if (y + 1 == tny) { arg -= (z - tnz / 2 + 1) * tny * bx / (tny * tnz); arg -= (t - tnt / 2 + 1) * tny * ey / (tny * tnt); } \
		arg *= q; if (x & 1) arg += HALFC; \
		ph[(2 + parity) * n + idxh] = arg; \
		arg = (y - tny / 2 + 1) * bx / (tny * tnz); \
		if (z + 1 == tnz) { arg -= (t - tnt / 2 + 1) * tnz * ez / (tnz * tnt); arg -= (x - tnx / 2 + 1) * tnz * by / (tnz * tnx); } \
		arg *= q; if ((x + y) & 1) arg += HALFC; \
		ph[(4 + parity) * n + idxh] = arg; \
		arg = (z - tnz / 2 + 1) * ez / (tnz * tnt); \
		arg += (y - tny / 2 + 1) * ey / (tny * tnt); \
		arg += (x - tnx / 2 + 1) * ex / (tnx * tnt); \
		arg *= q; if ((x + y + z) & 1) arg += HALFC; \
		arg += chpotphase * HALFC; \
		if (t + 1 == tnt) arg += HALFC; \
		ph[(6 + parity) * n + idxh] = arg; \
	} \
	for (long i = 0; i < 8 * n; i++) { \
		while (ph[i] > HALFC) ph[i] -= ONEC; \
		while (ph[i] < -HALFC) ph[i] += ONEC; \
	} \
	for (long i = 0; i < 8 * n; i++) ph[i] *= TWOPI;

void so_calc_u1_phases(const so_geom *g, int rank, double *ph, const double eb[6], double im_chem_pot,
											 double charge)
{
	SO_PHASES_BODY(double, 2 * 3.14159265358979323846, 0.5, 1.0)
}
void so_calc_u1_phases_f(const so_geom *g, int rank, float *ph, const double eb[6], double im_chem_pot,
												 double charge)
{
	SO_PHASES_BODY(float, 2 * 3.14159265358979323846f, 0.5f, 1.0f)
}

/* communications.c:1193-1257: local box (halos included) <- global field.  nd0..2 are never
 * decomposed, so a local d3 slice is a whole global d3 slice (even D3_HALO and even LOC_N3
 * keep local and global parity equal). */
void so_scatter_vec(const so_geom *g, int rank, const double complex *gl, double complex *lnh)
{
	const long glsizeh = g->vol3h * g->gl_n[3];
	for (int c = 0; c < 3; c++)
		for (int d3 = 0; d3 < g->nd[3]; d3++) {
			long gs = so_lnh_to_gl_snum(g, 0, 0, 0, d3, rank);
			memcpy(lnh + c * g->sizeh + d3 * g->vol3h, gl + c * glsizeh + gs, g->vol3h * sizeof(double complex));
		}
}
void so_gather_vec(const so_geom *g, int rank, double complex *gl, const double complex *lnh)
{
	const long glsizeh = g->vol3h * g->gl_n[3];
	for (int c = 0; c < 3; c++)
		for (int d3 = g->d3_halo; d3 < g->d3_halo + g->loc_n[3]; d3++) {
			long gs = so_lnh_to_gl_snum(g, 0, 0, 0, d3, rank);
			memcpy(gl + c * glsizeh + gs, lnh + c * g->sizeh + d3 * g->vol3h, g->vol3h * sizeof(double complex));
		}
}
/* communications.c:1104-1145: 8 link arrays x 9 entries */
void so_scatter_conf(const so_geom *g, int rank, const double complex *gl, double complex *lnh)
{
	const long glsizeh = g->vol3h * g->gl_n[3];
	for (int kc = 0; kc < 72; kc++)
		for (int d3 = 0; d3 < g->nd[3]; d3++) {
			long gs = so_lnh_to_gl_snum(g, 0, 0, 0, d3, rank);
			memcpy(lnh + kc * g->sizeh + d3 * g->vol3h, gl + kc * glsizeh + gs, g->vol3h * sizeof(double complex));
		}
}

/* communications.c:34-104: for every colour array send [off,+slab) to L which receives it at
 * [sizeh-off,+slab); send [sizeh-off-slab,+slab) to R which receives it at [off-slab,+slab). */
void so_exchange_halo(const so_geom *g, double complex **ranks, int ncomp_arrays, int thickness)
{
	const int nr = g->nranks_d3;
	const long slab = g->vol3h * thickness, off = g->vol3h * g->halo_width, n = g->sizeh;
	for (int r = 0; r < nr; r++) {
		int L = (r + nr - 1) % nr, Rr = (r + 1) % nr;
		for (int c = 0; c < ncomp_arrays; c++) {
			memcpy(ranks[L] + c * n + (n - off), ranks[r] + c * n + off, slab * sizeof(double complex));
			memcpy(ranks[Rr] + c * n + (off - slab), ranks[r] + c * n + (n - off - slab), slab * sizeof(double complex));
		}
	}
}

#define R double
#define C double complex
#define S(x) x
#define RCOS cos
#define RSIN sin
#define CONJ conj
#define HALF 0.5
#define RPOW pow
#define RACOS acos
#define RSQRT sqrt
#define RFABS fabs
static inline double so_cimag(double complex z) { return cimag(z); }
static inline float so_cimag_f(float complex z) { return cimagf(z); }
static inline double so_creal(double complex z) { return creal(z); }
static inline float so_creal_f(float complex z) { return crealf(z); }
#include "staggered_oracle_impl.h"
#undef R
#undef C
#undef S
#undef RCOS
#undef RSIN
#undef CONJ
#undef HALF
#undef RPOW
#undef RACOS
#undef RSQRT
#undef RFABS

#define R float
#define C float complex
#define S(x) x##_f
#define RCOS cosf
#define RSIN sinf
#define CONJ conjf
#define HALF 0.5f
#define RPOW powf
#define RACOS acosf
#define RSQRT sqrtf
#define RFABS fabsf
#include "staggered_oracle_impl.h"
#undef R
#undef C
#undef S
#undef RCOS
#undef RSIN
#undef CONJ
#undef HALF
#undef RPOW
#undef RACOS
#undef RSQRT
#undef RFABS

/* "wf" = with a field: the operator with every link's U(1) phase multiplied by a complex per-link field (the derivative of
 * the phase for the magnetic susceptibility), matvecmul.h:176-260 and field_times_fermion_matrix.c:77-196.  FP64 only
 * (the reference generates no _f twin of this file).  par = parity of the output sites. */
static void so_dslash_wf(const so_geom *g, int par, const double complex *u, double complex *out, const double complex *in,
												 const double *ph, const double *fre, const double *fim, int d3lo, int d3hi)
{
	const long n = g->sizeh;
	const int nd[4] = { g->nd[0], g->nd[1], g->nd[2], g->nd[3] };
	for (int d3 = d3lo; d3 < d3hi; d3++)
		for (int d2 = 0; d2 < nd[2]; d2++)
			for (int d1 = 0; d1 < nd[1]; d1++)
				for (int hd0 = 0; hd0 < nd[0] / 2; hd0++) {
					int c[4] = { 2 * hd0 + ((d1 + d2 + d3 + par) & 1), d1, d2, d3 };
					long idx = so_snum(g, c[0], c[1], c[2], c[3]);
					double complex acc[3] = { 0, 0, 0 };
					for (int fwd = 0; fwd < 2; fwd++)          /* backward hops subtracted first (:104-111), then forward (:117-124) */
						for (int mu = 0; mu < 4; mu++) {
							int cn[4] = { c[0], c[1], c[2], c[3] };
							if (fwd) cn[mu] = (c[mu] == nd[mu] - 1) ? 0 : c[mu] + 1; else cn[mu] = (c[mu] == 0) ? nd[mu] - 1 : c[mu] - 1;
							long in_idx = so_snum(g, cn[0], cn[1], cn[2], cn[3]);
							int k = fwd ? 2 * mu + par : 2 * mu + 1 - par;
							long im = fwd ? idx : in_idx;
							const double complex *um = u + (long) k * 9 * n;
							double arg = ph[(long) k * n + im];
							double complex phase = cos(arg) + I * sin(arg);
							phase *= (fre[(long) k * n + im] + I * fim[(long) k * n + im]);     /* not in U(1) anymore */
							if (!fwd) phase = conj(phase);
							double complex v0 = in[in_idx] * phase, v1 = in[n + in_idx] * phase, v2 = in[2 * n + in_idx] * phase;
							double complex m00 = um[im], m01 = um[n + im], m02 = um[2 * n + im];
							double complex m10 = um[3 * n + im], m11 = um[4 * n + im], m12 = um[5 * n + im];
							double complex m20 = conj(m01 * m12 - m02 * m11), m21 = conj(m02 * m10 - m00 * m12), m22 = conj(m00 * m11 - m01 * m10);
							if (fwd) {
								acc[0] += m00 * v0 + m01 * v1 + m02 * v2;
								acc[1] += m10 * v0 + m11 * v1 + m12 * v2;
								acc[2] += m20 * v0 + m21 * v1 + m22 * v2;
							} else {
								acc[0] -= conj(m00) * v0 + conj(m10) * v1 + conj(m20) * v2;
								acc[1] -= conj(m01) * v0 + conj(m11) * v1 + conj(m21) * v2;
								acc[2] -= conj(m02) * v0 + conj(m12) * v1 + conj(m22) * v2;
							}
						}
					out[idx] = acc[0] * 0.5; out[n + idx] = acc[1] * 0.5; out[2 * n + idx] = acc[2] * 0.5;
				}
}
void so_deo_wf(const so_geom *g, const double complex *u, double complex *out, const double complex *in, const double *ph,
							 const double *fre, const double *fim, int d3lo, int d3hi)
{ so_dslash_wf(g, 0, u, out, in, ph, fre, fim, d3lo, d3hi); }
void so_doe_wf(const so_geom *g, const double complex *u, double complex *out, const double complex *in, const double *ph,
							 const double *fre, const double *fim, int d3lo, int d3hi)
{ so_dslash_wf(g, 1, u, out, in, ph, fre, fim, d3lo, d3hi); }

/* float_double_conv.c:9-33 */
void so_convert_d2f(long n, const double complex *d, float complex *f)
{ for (long i = 0; i < n; i++) f[i] = (float) creal(d[i]) + (float) cimag(d[i]) * I; }
void so_convert_f2d(long n, const float complex *f, double complex *d)
{ for (long i = 0; i < n; i++) d[i] = (double) crealf(f[i]) + (double) cimagf(f[i]) * I; }

/* inverter_mixedp.c:24-181 ("magic touch" reliable updates; SAFETY_MARGIN 0.9) */
int so_inverter_mixed_precision(const so_geom *g, const double complex *u, const float complex *u_f,
		const double *ph, const float *ph_f, double mass, double complex *solution,
		const double complex *in, double res, int max_cg, double shift, double mixed_delta,
		double complex *r, double complex *h, double complex *s,
		float complex *r_f, float complex *h_f, float complex *s_f, float complex *p_f,
		float complex *out_f, int *cg_return, int *magic_touches)
{
	const long vs = 3 * g->sizeh, n = g->sizeh;
	int cg = 0, touches = 0;
	double delta, alpha, lambda, omega, gammag, last_max = 0;
	double source_norm = so_l2norm2(g, in);
	so_fermion_matrix_multiplication_shifted(g, u, s, solution, h, ph, mass, shift);
	so_axpy_like(g, SO_IN1_MINUS_IN2, r, in, s, 0, 0, 0);
	so_convert_d2f(vs, r, r_f);
	so_axpy_like_f(g, SO_ASSIGN, p_f, r_f, 0, 0, 0, 0);
	delta = so_l2norm2_f(g, r_f);
	so_axpy_like_f(g, SO_ZERO, out_f, 0, 0, 0, 0, 0);
	do {
		cg++;
		so_fermion_matrix_multiplication_shifted_f(g, u_f, s_f, p_f, h_f, ph_f, mass, shift);
		alpha = so_real_scal_prod_f(g, p_f, s_f);
		omega = delta / alpha;
		so_axpy_like_f(g, SO_IN1XFACTOR_PLUS_IN2, out_f, p_f, out_f, 0, omega, 0);
		if (last_max < delta) last_max = delta;
		if (delta < mixed_delta * last_max) {
			for (int c = 0; c < 3; c++)          /* combine_add_in2_into_in1_mixed_precision (:24-34) */
				for (long t = g->r1_lo; t < g->r1_hi; t++) solution[c * n + t] += (double complex) out_f[c * n + t];
			so_fermion_matrix_multiplication_shifted(g, u, s, solution, h, ph, mass, shift);
			so_axpy_like(g, SO_IN1_MINUS_IN2, r, in, s, 0, 0, 0);
			so_convert_d2f(vs, r, r_f);
			so_axpy_like_f(g, SO_ZERO, out_f, 0, 0, 0, 0, 0);
			last_max = 0; touches++;
		} else so_axpy_like_f(g, SO_IN1XFACTOR_PLUS_IN2, r_f, s_f, r_f, 0, -omega, 0);
		lambda = so_l2norm2_f(g, r_f);
		gammag = lambda / delta;
		delta = lambda;
		so_axpy_like_f(g, SO_IN1XFACTOR_PLUS_IN2, p_f, p_f, r_f, 0, gammag, 0);
	} while (sqrt(lambda / source_norm) > res * 0.9 && cg < max_cg);
	for (int c = 0; c < 3; c++)
		for (long t = g->r1_lo; t < g->r1_hi; t++) solution[c * n + t] += (double complex) out_f[c * n + t];
	so_fermion_matrix_multiplication_shifted(g, u, s, solution, h, ph, mass, shift);
	so_axpy_like(g, SO_IN1_MINUS_IN2, h, in, s, 0, 0, 0);
	double giusto = so_l2norm2(g, h) / source_norm;
	*cg_return = cg;
	if (magic_touches) *magic_touches = touches;
	return sqrt(giusto) <= res ? 1 : 0;
}

/* find_min_max.c:21-60: power iteration for lambda_max(M^+M) */
double so_find_max_eigenvalue(const so_geom *g, const double complex *u, const double *ph, double mass,
		double complex *r, double complex *h, double complex *p, int *loops)
{
	int loop_count = 0;
	double norm = sqrt(so_l2norm2(g, p)), old_norm;
	do {
		so_axpy_like(g, SO_SCALE, p, 0, 0, 0, 1.0 / norm, 0);
		so_axpy_like(g, SO_ASSIGN, r, p, 0, 0, 0, 0);
		old_norm = norm;
		so_fermion_matrix_multiplication_shifted(g, u, p, r, h, ph, mass, 0.0);
		norm = sqrt(so_l2norm2(g, p));
		old_norm = fabs(old_norm - norm) / norm;
		loop_count++;
	} while (old_norm > 1.0e-5);
	if (loops) *loops = loop_count;
	return norm;
}

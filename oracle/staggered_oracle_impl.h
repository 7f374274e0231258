/* TEST INFRASTRUCTURE ONLY -- precision-generic body of the CPU restatement.
 * Included twice by staggered_oracle.c with
 *   R = real type, C = complex type, S(x) = name suffixing, RCOS/RSIN, HALF.
 * FP32 rule (double_to_single_transformer.py:108-145): the operator (matvecmul.h,
 * fermion_matrix.c) is pure float; BLAS-1 and the solvers keep double scalars,
 * accumulators and factors and only the stored vectors are float.
 */

/* U(x) e^{i th} v   -- matvecmul.h:88-126 */
static inline void S(so_mat_vec_arg)(const so_geom *g, const C *um, long im, const C *v, long iv,
																		 const R *phk, C o[3])
{
	const long n = g->sizeh;
	R th = phk[im];
	C phase = RCOS(th) + I * RSIN(th);
	C v0 = v[iv] * phase, v1 = v[n + iv] * phase, v2 = v[2 * n + iv] * phase;
	C m00 = um[im], m01 = um[n + im], m02 = um[2 * n + im];
	C m10 = um[3 * n + im], m11 = um[4 * n + im], m12 = um[5 * n + im];
	C m20 = CONJ(m01 * m12 - m02 * m11);
	C m21 = CONJ(m02 * m10 - m00 * m12);
	C m22 = CONJ(m00 * m11 - m01 * m10);
	o[0] = m00 * v0 + m01 * v1 + m02 * v2;
	o[1] = m10 * v0 + m11 * v1 + m12 * v2;
	o[2] = m20 * v0 + m21 * v1 + m22 * v2;
}

/* U(x)^dagger e^{-i th} v   -- matvecmul.h:129-172 */
static inline void S(so_conjmat_vec_arg)(const so_geom *g, const C *um, long im, const C *v, long iv,
																				 const R *phk, C o[3])
{
	const long n = g->sizeh;
	R th = phk[im];
	C phase = RCOS(th) + I * RSIN(th);
	C v0 = v[iv] * CONJ(phase), v1 = v[n + iv] * CONJ(phase), v2 = v[2 * n + iv] * CONJ(phase);
	C m00 = um[im], m01 = um[n + im], m02 = um[2 * n + im];
	C m10 = um[3 * n + im], m11 = um[4 * n + im], m12 = um[5 * n + im];
	C c20 = m01 * m12 - m02 * m11;     /* = conj(third row) used directly as third column of U^+ */
	C c21 = m02 * m10 - m00 * m12;
	C c22 = m00 * m11 - m01 * m10;
	o[0] = CONJ(m00) * v0 + CONJ(m10) * v1 + c20 * v2;
	o[1] = CONJ(m01) * v0 + CONJ(m11) * v1 + c21 * v2;
	o[2] = CONJ(m02) * v0 + CONJ(m12) * v1 + c22 * v2;
}

/* fermion_matrix.c:47-157 (unsafe), :271-718 (bulk/d3p/d3m/d3c via the d3 range).
 * par = parity of the OUTPUT sites: 0 -> Deo, 1 -> Doe. */
static void S(so_dslash)(const so_geom *g, int par, const C *u, C *out, const C *in, const R *ph,
												 int d3lo, int d3hi)
{
	const long n = g->sizeh;
	const int nd0 = g->nd[0], nd1 = g->nd[1], nd2 = g->nd[2], nd3 = g->nd[3];
	for (int d3 = d3lo; d3 < d3hi; d3++)
		for (int d2 = 0; d2 < nd2; d2++)
			for (int d1 = 0; d1 < nd1; d1++)
				for (int hd0 = 0; hd0 < nd0 / 2; hd0++) {
					int d0 = 2 * hd0 + ((d1 + d2 + d3 + par) & 1);
					int c[4] = { d0, d1, d2, d3 };
					const int nd[4] = { nd0, nd1, nd2, nd3 };
					long idx = so_snum(g, d0, d1, d2, d3);
					C acc[3] = { 0, 0, 0 }, t[3];
					for (int mu = 0; mu < 4; mu++) {      /* backward hops, subtracted first (:74-77) */
						int cm[4] = { c[0], c[1], c[2], c[3] };
						cm[mu] = (c[mu] == 0) ? nd[mu] - 1 : c[mu] - 1;
						long im = so_snum(g, cm[0], cm[1], cm[2], cm[3]);
						int k = 2 * mu + 1 - par;
						S(so_conjmat_vec_arg)(g, u + (long) k * 9 * n, im, in, im, ph + (long) k * n, t);
						acc[0] -= t[0]; acc[1] -= t[1]; acc[2] -= t[2];
					}
					for (int mu = 0; mu < 4; mu++) {      /* forward hops (:86-89) */
						int cp[4] = { c[0], c[1], c[2], c[3] };
						cp[mu] = (c[mu] == nd[mu] - 1) ? 0 : c[mu] + 1;
						long ip = so_snum(g, cp[0], cp[1], cp[2], cp[3]);
						int k = 2 * mu + par;
						S(so_mat_vec_arg)(g, u + (long) k * 9 * n, idx, in, ip, ph + (long) k * n, t);
						acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2];
					}
					out[idx] = acc[0] * HALF; out[n + idx] = acc[1] * HALF; out[2 * n + idx] = acc[2] * HALF;
				}
}

void S(so_deo)(const so_geom *g, const C *u, C *out, const C *in, const R *ph, int d3lo, int d3hi)
{ S(so_dslash)(g, 0, u, out, in, ph, d3lo, d3hi); }
void S(so_doe)(const so_geom *g, const C *u, C *out, const C *in, const R *ph, int d3lo, int d3hi)
{ S(so_dslash)(g, 1, u, out, in, ph, d3lo, d3hi); }

/* fermionic_utilities.c:180-313,379-420 : element-wise updates, double factors */
void S(so_axpy_like)(const so_geom *g, int op, C *out, const C *a, const C *b, const C *c, double f1,
										 double f2)
{
	const long n = g->sizeh;
	long lo = g->r1_lo, hi = g->r1_hi;
	if (op == SO_ZERO) { lo = 0; hi = n; }
	(void) f2;
	for (int col = 0; col < 3; col++)
		for (long i = lo; i < hi; i++) {
			long j = col * n + i;
			switch (op) {
			case SO_IN1XFACTOR_PLUS_IN2: out[j] = (a[j] * f1) + b[j]; break;
			case SO_SCALE: out[j] = f1 * out[j]; break;
			case SO_ADD_FACTOR_X_IN2: out[j] += f1 * a[j]; break;
			case SO_IN1XMASS2_MINUS_IN2_MINUS_IN3: out[j] = (a[j] * f1) - b[j] - c[j]; break;
			case SO_IN1XMASS_MINUS_IN2: out[j] = (a[j] * f1) - out[j]; break;
			case SO_IN1_MINUS_IN2: out[j] = a[j] - b[j]; break;
			case SO_ASSIGN: out[j] = a[j]; break;
			case SO_ZERO: out[j] = 0; break;
			case SO_FACT1_MINUS_IN2: out[j] = f1 * a[j] - out[j]; break;
			case SO_IN1_MINUS_IN2_ALLXFACT: out[j] = f1 * (a[j] - b[j]); break;
			}
		}
}

/* fermion_matrix.c:723-746 */
void S(so_fermion_matrix_multiplication_shifted)(const so_geom *g, const C *u, C *out, const C *in,
		C *tmp, const R *ph, double mass, double shift)
{
	int lo = g->d3_halo, hi = g->d3_halo + g->loc_n[3];
	S(so_doe)(g, u, tmp, in, ph, lo, hi);
	S(so_deo)(g, u, out, tmp, ph, lo, hi);
	/* the generated FP32 twin takes `float shift` (sp_fermion_matrix.c:735-746): its callers' double shifts are rounded */
	S(so_axpy_like)(g, SO_IN1XMASS_MINUS_IN2, out, in, 0, 0, mass * mass + (R) shift, 0);
}

/* reductions, fermionic_utilities.c:32-175 (+ fermionic_utilities.h:15-35): double accumulators */
double complex S(so_scal_prod)(const so_geom *g, const C *a, const C *b)
{
	const long n = g->sizeh;
	double re = 0, im = 0;
	for (long t = g->r0_lo; t < g->r0_hi; t++) {
		double complex s = conj(a[t]) * b[t];       /* conj() promotes float complex to double complex */
		s += conj(a[n + t]) * b[n + t];
		s += conj(a[2 * n + t]) * b[2 * n + t];
		re += creal(s); im += cimag(s);
	}
	return re + im * I;
}
double S(so_real_scal_prod)(const so_geom *g, const C *a, const C *b)
{
	const long n = g->sizeh;
	double res = 0;
	for (long t = g->r0_lo; t < g->r0_hi; t++) {
		double s = 0;
		for (int c = 0; c < 3; c++) {
			double complex x = a[c * n + t], y = b[c * n + t];
			if (c == 0) s = creal(x) * creal(y) + cimag(x) * cimag(y);
			else s += creal(x) * creal(y) + cimag(x) * cimag(y);
		}
		res += s;
	}
	return res;
}
double S(so_l2norm2)(const so_geom *g, const C *a)
{
	const long n = g->sizeh;
	double res = 0;
	for (long t = g->r0_lo; t < g->r0_hi; t++) {
		double s = 0;
		for (int c = 0; c < 3; c++) {
			double complex x = a[c * n + t];
			if (c == 0) s = creal(x) * creal(x) + cimag(x) * cimag(x);
			else s += creal(x) * creal(x) + cimag(x) * cimag(x);
		}
		res += s;
	}
	return res;
}

/* CG-M, inverter_multishift_full.c:23-252.  ps = shiftferm[order], out[order].
 * true_rel_res2[i] (optional) = |in - (M^+M + b_i) x_i|^2 / |in|^2 of the post-loop check (:211-229). */
int S(so_multishift_invert)(const so_geom *g, const C *u, const R *ph, double mass, int order,
		const double *shifts, C *out, const C *in, double residuo, C *r, C *h, C *s, C *p, C *ps,
		int max_cg, int *cg_return, double *true_rel_res2)
{
	const long vs = 3 * g->sizeh;
	double zeta_i[SO_MAX_APPROX_ORDER], zeta_ii[SO_MAX_APPROX_ORDER], zeta_iii[SO_MAX_APPROX_ORDER];
	double omegas[SO_MAX_APPROX_ORDER], gammas[SO_MAX_APPROX_ORDER];
	int flag[SO_MAX_APPROX_ORDER];
	double alpha, delta, lambda, omega, omega_save, gammag, fact;
	int cg = 0, maxiter = 0;

	for (int i = 0; i < order; i++) { flag[i] = 1; S(so_axpy_like)(g, SO_ZERO, out + i * vs, 0, 0, 0, 0, 0); }
	S(so_axpy_like)(g, SO_ASSIGN, r, in, 0, 0, 0, 0);
	S(so_axpy_like)(g, SO_ASSIGN, p, r, 0, 0, 0, 0);
	delta = S(so_l2norm2)(g, r);
	double source_norm = S(so_l2norm2)(g, in);
	omega = 1.0;
	for (int i = 0; i < order; i++) {
		S(so_axpy_like)(g, SO_ASSIGN, ps + i * vs, in, 0, 0, 0, 0);
		zeta_i[i] = 1.0; zeta_ii[i] = 1.0; gammas[i] = 0.0;
	}
	gammag = 0.0;
	for (int i = 0; i < order; i++) if (flag[i] == 1) maxiter = i + 1;

	do {
		cg++;
		S(so_fermion_matrix_multiplication_shifted)(g, u, s, p, h, ph, mass, 0.0);
		alpha = S(so_real_scal_prod)(g, p, s);
		omega_save = omega;
		omega = -delta / alpha;
		for (int i = 0; i < maxiter; i++)
			if (flag[i] == 1) {
				zeta_iii[i] = (zeta_i[i] * zeta_ii[i] * omega_save) /
					(omega * gammag * (zeta_i[i] - zeta_ii[i]) + zeta_i[i] * omega_save * (1.0 - shifts[i] * omega));
				omegas[i] = omega * zeta_iii[i] / zeta_ii[i];
			}
		for (int i = 0; i < maxiter; i++)                 /* out_i -= omega_i ps_i */
			if (flag[i] == 1) S(so_axpy_like)(g, SO_ADD_FACTOR_X_IN2, out + i * vs, ps + i * vs, 0, 0, -omegas[i], 0);
		S(so_axpy_like)(g, SO_ADD_FACTOR_X_IN2, r, s, 0, 0, omega, 0);
		lambda = S(so_l2norm2)(g, r);
		gammag = lambda / delta;
		S(so_axpy_like)(g, SO_IN1XFACTOR_PLUS_IN2, p, p, r, 0, gammag, 0);
		for (int i = 0; i < order; i++)
			if (flag[i] == 1) gammas[i] = gammag * zeta_iii[i] * omegas[i] / (zeta_ii[i] * omega);
		for (int i = 0; i < maxiter; i++)                 /* ps_i = gamma_i ps_i + zeta_i^+ r */
			if (flag[i] == 1) {
				C *q = ps + i * vs;
				for (int col = 0; col < 3; col++)
					for (long t = g->r1_lo; t < g->r1_hi; t++) {
						long j = col * g->sizeh + t;
						q[j] = gammas[i] * q[j] + zeta_iii[i] * r[j];
					}
			}
		maxiter = 0;
		for (int i = 0; i < order; i++)
			if (flag[i] == 1) {
				fact = sqrt(delta * zeta_ii[i] * zeta_ii[i] / source_norm);
				if (fact < residuo * 0.95) flag[i] = 0;
				else maxiter = i + 1;
				zeta_i[i] = zeta_ii[i];
				zeta_ii[i] = zeta_iii[i];
			}
		delta = lambda;
	} while (maxiter > 0 && cg < max_cg);

	int check = 1;
	for (int i = 0; i < order; i++) {
		S(so_axpy_like)(g, SO_ASSIGN, p, out + i * vs, 0, 0, 0, 0);
		S(so_fermion_matrix_multiplication_shifted)(g, u, s, p, h, ph, mass, shifts[i]);
		S(so_axpy_like)(g, SO_IN1_MINUS_IN2, h, in, s, 0, 0, 0);
		double rel = S(so_l2norm2)(g, h) / source_norm;
		if (true_rel_res2) true_rel_res2[i] = rel;
		check *= (rel <= 1) ? 1 : 0;
	}
	*cg_return = cg;
	return check == 1 ? 1 : 0;
}

/* inverter_multishift_full.c:254-282 (over all sizeh) */
void S(so_recombine)(const so_geom *g, const C *in_shifted, const C *in, C *out, int order, double a0,
										 const double *a)
{
	const long vs = 3 * g->sizeh;
	for (long j = 0; j < vs; j++) {
		out[j] = in[j] * a0;
		for (int i = 0; i < order; i++) out[j] += a[i] * in_shifted[i * vs + j];
	}
}

/* restarted CG on (M^+M + shift), inverter_full.c:19-132 */
int S(so_cg)(const so_geom *g, const C *u, const R *ph, double mass, C *solution, const C *in,
		double res, C *r, C *h, C *s, C *p, int max_cg, double shift, int restarting_every, int *cg_return)
{
	int cg = 0;
	double delta, alpha, lambda = 0, omega, gammag;
	double source_norm = S(so_l2norm2)(g, in);
	do {
		S(so_fermion_matrix_multiplication_shifted)(g, u, s, solution, h, ph, mass, shift);
		S(so_axpy_like)(g, SO_IN1_MINUS_IN2, r, in, s, 0, 0, 0);
		S(so_axpy_like)(g, SO_ASSIGN, p, r, 0, 0, 0, 0);
		delta = S(so_l2norm2)(g, r);
		int cg_restarted = 0;
		do {
			cg++; cg_restarted++;
			S(so_fermion_matrix_multiplication_shifted)(g, u, s, p, h, ph, mass, shift);
			alpha = S(so_real_scal_prod)(g, p, s);
			omega = delta / alpha;
			S(so_axpy_like)(g, SO_IN1XFACTOR_PLUS_IN2, solution, p, solution, 0, omega, 0);
			S(so_axpy_like)(g, SO_IN1XFACTOR_PLUS_IN2, r, s, r, 0, -omega, 0);
			lambda = S(so_l2norm2)(g, r);
			gammag = lambda / delta;
			delta = lambda;
			S(so_axpy_like)(g, SO_IN1XFACTOR_PLUS_IN2, p, p, r, 0, gammag, 0);
		} while (sqrt(lambda / source_norm) > res * 0.95 && cg_restarted < restarting_every);
	} while (sqrt(lambda / source_norm) > res && cg < max_cg);
	S(so_fermion_matrix_multiplication_shifted)(g, u, s, solution, h, ph, mass, shift);
	S(so_axpy_like)(g, SO_IN1_MINUS_IN2, h, in, s, 0, 0, 0);
	double current_res = S(so_l2norm2)(g, h) / source_norm;
	*cg_return = cg;
	return sqrt(current_res) <= res ? 1 : 0;
}

/* ------------------------------------------------------------------ fermion force outer products
 * gl(3) fields (aux_u, pseudo_ipdot) use the su3_soa[8] layout with ALL three rows meaningful:
 *   aux[((k*3 + r)*3 + c)*sizeh + i].   tamat_soa[8] (struct_c_def.h:45-51) as reals:
 *   ta[k*8*sizeh + ...]: c01 complex @0, c02 @2*sizeh, c12 @4*sizeh, ic00 real @6*sizeh, ic11 @7*sizeh. */

/* fermion_force_utilities.h:16-42  aux(idxh) += fer_l(idl) (x) [factor * conj(fer_r(idr))] */
static inline void S(so_directprod)(const so_geom *g, C *auxk, long idxh, const C *fl, long idl, const C *fr,
																		long idr, R factor)
{
	const long n = g->sizeh;
	C r[3], l[3];
	for (int c = 0; c < 3; c++) { r[c] = factor * CONJ(fr[c * n + idr]); l[c] = fl[c * n + idl]; }
	for (int a = 0; a < 3; a++)
		for (int c = 0; c < 3; c++) auxk[(a * 3 + c) * n + idxh] += l[a] * r[c];
}

static long S(so_nnp)(const so_geom *g, int d0, int d1, int d2, int d3, int mu)   /* geometry.c:49-61 */
{
	int c[4] = { d0, d1, d2, d3 };
	c[mu] = (c[mu] == g->nd[mu] - 1) ? 0 : c[mu] + 1;
	return so_snum(g, c[0], c[1], c[2], c[3]);
}

/* fermion_force_utilities.c:31-95: even sites  aux[2mu]  (x) += a * h(x+mu) (x) conj(s(x))
 *                                  odd sites   aux[2mu+1](x) += -a * s(x+mu) (x) conj(h(x)) */
void S(so_direct_product_of_fermions_into_auxmat)(const so_geom *g, const C *s, const C *h, C *aux, double a)
{
	const long n = g->sizeh;
	for (int par = 0; par < 2; par++)
		for (int d3 = g->d3_halo; d3 < g->nd[3] - g->d3_halo; d3++)
			for (int d2 = 0; d2 < g->nd[2]; d2++)
				for (int d1 = 0; d1 < g->nd[1]; d1++)
					for (int hd0 = 0; hd0 < g->nd[0] / 2; hd0++) {
						int d0 = 2 * hd0 + ((d1 + d2 + d3 + par) & 1);
						long idxh = so_snum(g, d0, d1, d2, d3);
						for (int mu = 0; mu < 4; mu++) {
							long ip = S(so_nnp)(g, d0, d1, d2, d3, mu);
							C *auxk = aux + (long) (2 * mu + par) * 9 * n;
							if (par == 0) S(so_directprod)(g, auxk, idxh, h, ip, s, idxh, (R) a);
							else S(so_directprod)(g, auxk, idxh, s, ip, h, idxh, (R) -a);
						}
					}
}

/* fermion_force_utilities.c:183-201 (acc_Doe without exchange: single-rank geometries only) */
void S(so_compute_fermion_force)(const so_geom *g, const C *u, C *aux, const C *in_shiftmulti, C *s, C *h,
																 const R *ph, int order, const double *ra_a)
{
	const long vs = 3 * g->sizeh;
	for (int iter = 0; iter < order; iter++) {
		S(so_axpy_like)(g, SO_ASSIGN, s, in_shiftmulti + iter * vs, 0, 0, 0, 0);
		S(so_doe)(g, u, h, s, ph, g->d3_halo, g->d3_halo + g->loc_n[3]);
		S(so_direct_product_of_fermions_into_auxmat)(g, s, h, aux, ra_a[iter]);
	}
}

/* fermion_force_utilities.c:123-153 + .h:185-203: pseudo_ipdot += e^{i theta} aux over the local interior */
void S(so_multiply_backfield_times_force)(const so_geom *g, const R *ph, const C *aux, C *pseudo)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	for (int k = 0; k < 8; k++)
		for (long i = lo; i < hi; i++) {
			R arg = ph[k * n + i];
			C phase = RCOS(arg) + I * RSIN(arg);
			for (int e = 0; e < 9; e++) pseudo[(k * 9 + e) * n + i] += aux[(k * 9 + e) * n + i] * phase;
		}
}

/* fermion_force_utilities.c:155-180 */
void S(so_accumulate_gl3soa_into_gl3soa)(const so_geom *g, const C *aux, C *pseudo)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	for (int k = 0; k < 8; k++)
		for (long i = lo; i < hi; i++)
			for (int e = 0; e < 9; e++) pseudo[(k * 9 + e) * n + i] += aux[(k * 9 + e) * n + i];
}

/* fermion_force_utilities.c:97-121 + .h:108-152: ipdot -= TA(U * aux), U's third row rebuilt */
void S(so_multiply_conf_times_force_and_take_ta_nophase)(const so_geom *g, const C *u, const C *aux, R *ta)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	const R one_by_three = (R) 0.33333333333333333333333;   /* common_defines.h:65-66 */
	for (int k = 0; k < 8; k++) {
		const C *uk = u + (long) k * 9 * n, *ak = aux + (long) k * 9 * n;
		R *tk = ta + (long) k * 8 * n;
		C *c01 = (C *) tk, *c02 = (C *) (tk + 2 * n), *c12 = (C *) (tk + 4 * n);
		R *ic00 = tk + 6 * n, *ic11 = tk + 7 * n;
		for (long i = lo; i < hi; i++) {
			C m[3][3], x[3][3], p[3][3];
			for (int c = 0; c < 3; c++) { m[0][c] = uk[c * n + i]; m[1][c] = uk[(3 + c) * n + i]; }
			m[2][0] = CONJ(m[0][1] * m[1][2] - m[0][2] * m[1][1]);
			m[2][1] = CONJ(m[0][2] * m[1][0] - m[0][0] * m[1][2]);
			m[2][2] = CONJ(m[0][0] * m[1][1] - m[0][1] * m[1][0]);
			for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) x[r][c] = ak[(r * 3 + c) * n + i];
			for (int r = 0; r < 3; r++)
				for (int c = 0; c < 3; c++) p[r][c] = m[r][0] * x[0][c] + m[r][1] * x[1][c] + m[r][2] * x[2][c];
			c01[i] -= HALF * (p[0][1] - CONJ(p[1][0]));
			c02[i] -= HALF * (p[0][2] - CONJ(p[2][0]));
			c12[i] -= HALF * (p[1][2] - CONJ(p[2][1]));
			R tr = S(so_cimag)(p[0][0]) + S(so_cimag)(p[1][1]) + S(so_cimag)(p[2][2]);
			ic00[i] -= S(so_cimag)(p[0][0]) - one_by_three * tr;
			ic11[i] -= S(so_cimag)(p[1][1]) - one_by_three * tr;
		}
	}
}

/* ------------------------------------------------------------------ isotropic stout smearing
 * Produces the smeared links the Dirac operator reads.  C_ZERO = 5/3 (ACTION_TYPE TLSM, common_defines.h:73). */
#define SO_C_ZERO ((R) 5.0 * (R) 0.33333333333333333333333)

static inline void S(so_load_link)(const C *uk, long n, long i, C m[3][3])   /* rows 0,1 + conj(r0 x r1) */
{
	for (int c = 0; c < 3; c++) { m[0][c] = uk[c * n + i]; m[1][c] = uk[(3 + c) * n + i]; }
	m[2][0] = CONJ((m[0][1] * m[1][2]) - (m[0][2] * m[1][1]));
	m[2][1] = CONJ((m[0][2] * m[1][0]) - (m[0][0] * m[1][2]));
	m[2][2] = CONJ((m[0][0] * m[1][1]) - (m[0][1] * m[1][0]));
}
static inline void S(so_dagger)(C m[3][3])
{
	for (int r = 0; r < 3; r++) m[r][r] = CONJ(m[r][r]);
	for (int r = 0; r < 3; r++)
		for (int c = r + 1; c < 3; c++) { C t = CONJ(m[r][c]); m[r][c] = CONJ(m[c][r]); m[c][r] = t; }
}
/* first two rows of a*b, third row rebuilt (su3_utilities.h:660-750, :878-975) */
static inline void S(so_mul2)(C a[3][3], C b[3][3], C o[3][3])
{
	for (int r = 0; r < 2; r++)
		for (int c = 0; c < 3; c++) o[r][c] = a[r][0] * b[0][c] + a[r][1] * b[1][c] + a[r][2] * b[2][c];
	o[2][0] = CONJ((o[0][1] * o[1][2]) - (o[0][2] * o[1][1]));
	o[2][1] = CONJ((o[0][2] * o[1][0]) - (o[0][0] * o[1][2]));
	o[2][2] = CONJ((o[0][0] * o[1][1]) - (o[0][1] * o[1][0]));
}
static long S(so_shift)(const so_geom *g, const int c0[4], int mu, int smu, int nu, int snu)   /* geometry.c:14-61 wrap */
{
	int c[4] = { c0[0], c0[1], c0[2], c0[3] };
	if (smu) c[mu] = (c[mu] + smu + g->nd[mu]) % g->nd[mu];
	if (snu) c[nu] = (c[nu] + snu + g->nd[nu]) % g->nd[nu];
	return so_snum(g, c[0], c[1], c[2], c[3]);
}

/* plaquettes.c:196-255: loc_stap[2mu+p](x) += C_ZERO * sum_{nu != mu} [ U_nu(x+mu) U_mu(x+nu)^+ U_nu(x)^+
 *                                                                    + U_nu(x+mu-nu)^+ U_mu(x-nu)^+ U_nu(x-nu) ] */
void S(so_calc_loc_staples_onlyferms)(const so_geom *g, const C *u, C *stap)
{
	const long n = g->sizeh;
	static const int perp[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
	for (int d3 = g->d3_halo; d3 < g->nd[3] - g->d3_halo; d3++)
		for (int d2 = 0; d2 < g->nd[2]; d2++)
			for (int d1 = 0; d1 < g->nd[1]; d1++)
				for (int d0 = 0; d0 < g->nd[0]; d0++) {
					const int x[4] = { d0, d1, d2, d3 };
					const long idxh = so_snum(g, d0, d1, d2, d3);
					const int p = (d0 + d1 + d2 + d3) % 2;
					for (int mu = 0; mu < 4; mu++) {
						C *sk = stap + (long) (2 * mu + p) * 9 * n;
						for (int it = 0; it < 3; it++) {
							const int nu = perp[mu][it];
							const long ipmu = S(so_shift)(g, x, mu, 1, nu, 0), ipnu = S(so_shift)(g, x, mu, 0, nu, 1);
							const long imnu = S(so_shift)(g, x, mu, 0, nu, -1), ipmumnu = S(so_shift)(g, x, mu, 1, nu, -1);
							C a[3][3], b[3][3], c[3][3], ab[3][3], abc[3][3];
							/* right: U_nu(x+mu) [2nu+!p] * U_mu(x+nu)^+ [2mu+!p] * U_nu(x)^+ [2nu+p] */
							S(so_load_link)(u + (long) (2 * nu + !p) * 9 * n, n, ipmu, a);
							S(so_load_link)(u + (long) (2 * mu + !p) * 9 * n, n, ipnu, b); S(so_dagger)(b);
							S(so_load_link)(u + (long) (2 * nu + p) * 9 * n, n, idxh, c); S(so_dagger)(c);
							S(so_mul2)(a, b, ab); S(so_mul2)(ab, c, abc);
							for (int e = 0; e < 9; e++) sk[e * n + idxh] += SO_C_ZERO * abc[e / 3][e % 3];
							/* left: U_nu(x+mu-nu)^+ [2nu+p] * U_mu(x-nu)^+ [2mu+!p] * U_nu(x-nu) [2nu+!p] */
							S(so_load_link)(u + (long) (2 * nu + p) * 9 * n, n, ipmumnu, a); S(so_dagger)(a);
							S(so_load_link)(u + (long) (2 * mu + !p) * 9 * n, n, imnu, b); S(so_dagger)(b);
							S(so_load_link)(u + (long) (2 * nu + !p) * 9 * n, n, imnu, c);
							S(so_mul2)(a, b, ab); S(so_mul2)(ab, c, abc);
							for (int e = 0; e < 9; e++) sk[e * n + idxh] += SO_C_ZERO * abc[e / 3][e % 3];
						}
					}
				}
}

/* su3_utilities.c:210-237 + su3_utilities.h:1097-1150: tipdot = (rho/C_ZERO) TA(U * staples), assigned */
void S(so_rho_times_conf_times_staples_ta_part)(const so_geom *g, const C *u, const C *stap, R *ta, double rho)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	const R one_by_three = (R) 0.33333333333333333333333, tmp = (R) rho / SO_C_ZERO;
	for (int k = 0; k < 8; k++) {
		R *tk = ta + (long) k * 8 * n;
		C *c01 = (C *) tk, *c02 = (C *) (tk + 2 * n), *c12 = (C *) (tk + 4 * n);
		R *ic00 = tk + 6 * n, *ic11 = tk + 7 * n;
		for (long i = lo; i < hi; i++) {
			C m[3][3], x[3][3], p[3][3];
			S(so_load_link)(u + (long) k * 9 * n, n, i, m);
			for (int e = 0; e < 9; e++) x[e / 3][e % 3] = stap[((long) k * 9 + e) * n + i];
			for (int r = 0; r < 3; r++)
				for (int c = 0; c < 3; c++) p[r][c] = m[r][0] * x[0][c] + m[r][1] * x[1][c] + m[r][2] * x[2][c];
			c01[i] = tmp * (HALF * (p[0][1] - CONJ(p[1][0])));
			c02[i] = tmp * (HALF * (p[0][2] - CONJ(p[2][0])));
			c12[i] = tmp * (HALF * (p[1][2] - CONJ(p[2][1])));
			R tr = S(so_cimag)(p[0][0]) + S(so_cimag)(p[1][1]) + S(so_cimag)(p[2][2]);
			ic00[i] = tmp * (S(so_cimag)(p[0][0]) - one_by_three * tr);
			ic11[i] = tmp * (S(so_cimag)(p[1][1]) - one_by_three * tr);
		}
	}
}

/* cayley_hamilton.h:24-180: rows 0,1 of exp(-QA) = exp(iQ), Q = i QA hermitian traceless (Morningstar-Peardon) */
static void S(so_ch_exp)(C q01, C q02, C q12, R i00, R i11, C e[2][3])
{
	const R i22 = -i00 - i11;
	R c0 = -S(so_creal)(i00 * i11 * i22 + 2 * S(so_cimag)(q01 * q12 * CONJ(q02)) - i00 * q12 * CONJ(q12)
											- i11 * q02 * CONJ(q02) - q01 * CONJ(q01) * i22);
	const R c1 = HALF * (2 * S(so_creal)(i00 * i00 + i11 * i11 + i00 * i11 + q01 * CONJ(q01) + q02 * CONJ(q02) + q12 * CONJ(q12)));
	const R c0max = 2 * RPOW(c1 / 3, (R) 1.5);
	C f0, f1, f2;
	if (c1 < (R) 4.0e-3) {
		f0 = (1 - c0 * c0 / 720) + ((R) 1.0 * I) * (-c0 * (1 - c1 * (1 - c1 / 42) / 20) / 6);
		f1 = (c0 * (1 - c1 * (1 - 3 * c1 / 112) / 15) / 24) + ((R) 1.0 * I) * (1 - c1 * (1 - c1 * (1 - c1 / 42) / 20) / 6 - c0 * c0 / 5040);
		f2 = (HALF * (-1 + c1 * (1 - c1 * (1 - c1 / 56) / 30) / 12 + c0 * c0 / 20160)) + ((R) 1.0 * I) * (HALF * (c0 * (1 - c1 * (1 - c1 / 48) / 21) / 60));
	} else {
		int sign = 1;
		if (c0 < 0) { sign = -1; c0 = -c0; }
		const R eps = (c0max - c0) / c0max;
		R theta;
		if (eps < 0) theta = 0;
		else if (eps < 1e-3) theta = RSQRT(2 * eps) * (1 + ((R) 1.0 / 12 + ((R) 3.0 / 160 + ((R) 5.0 / 896 + ((R) 35.0 / 18432 + (R) 63.0 / 90112 * eps) * eps) * eps) * eps) * eps);
		else theta = RACOS(c0 / c0max);
		const R u = RSQRT(c1 / 3) * RCOS(theta / 3), w = RSQRT(c1) * RSIN(theta / 3);
		const R u2 = u * u, w2 = w * w, u2mw2 = u2 - w2, w2p3u2 = w2 + 3 * u2, w2m3u2 = w2 - 3 * u2;
		const R cu = RCOS(u), c2u = RCOS(2 * u), su = RSIN(u), s2u = RSIN(2 * u), cw = RCOS(w);
		R xi0w;
		if (RFABS(w) < (R) 0.05) { R t0 = w * w, t1 = 1 - t0 / 42, t2 = (R) 1.0 - t0 / 20 * t1; xi0w = 1 - t0 / 6 * t2; }
		else xi0w = RSIN(w) / w;
		const R denom = 1 / (9 * u * u - w * w);
		f0 = (u2mw2 * c2u + cu * 8 * u2 * cw + 2 * su * u * w2p3u2 * xi0w) + ((R) 1.0 * I) * (u2mw2 * s2u + -su * 8 * u2 * cw + cu * 2 * u * w2p3u2 * xi0w);
		f0 *= denom;
		f1 = (2 * u * c2u + -cu * 2 * u * cw + -su * w2m3u2 * xi0w) + ((R) 1.0 * I) * (2 * u * s2u + su * 2 * u * RCOS(w) + -cu * w2m3u2 * xi0w);
		f1 *= denom;
		f2 = (c2u + -cu * cw + -3 * su * u * xi0w) + ((R) 1.0 * I) * (s2u + su * cw + -cu * 3 * u * xi0w);
		f2 *= denom;
		if (sign == -1) { f0 = CONJ(f0); f1 = -CONJ(f1); f2 = CONJ(f2); }
	}
	e[0][0] = f0 - f1 * i00 + f2 * (i00 * i00 + q01 * CONJ(q01) + q02 * CONJ(q02));
	e[0][1] = (f1 * I) * q01 + f2 * (q02 * CONJ(q12) + ((R) -1.0 * I) * q01 * (i00 + i11));
	e[0][2] = (f1 * I) * q02 + f2 * (-q01 * q12 + ((R) 1.0 * I) * q02 * i11);
	e[1][0] = (-f1 * I) * CONJ(q01) + f2 * (q12 * CONJ(q02) + ((R) 1.0 * I) * CONJ(q01) * (i00 + i11));
	e[1][1] = f0 - f1 * i11 + f2 * (i11 * i11 + q01 * CONJ(q01) + q12 * CONJ(q12));
	e[1][2] = (f1 * I) * q12 + f2 * (((R) 1.0 * I) * i00 * q12 + q02 * CONJ(q01));
}

/* stouting.c:107-167: exp_aux = rows 0,1 of exp(-QA); u_out rows 0,1 = exp_aux * U (row 2 of both untouched) */
void S(so_exp_minus_QA_times_conf)(const so_geom *g, const C *u, const R *ta, C *uout, C *expaux)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	for (int k = 0; k < 8; k++) {
		const R *tk = ta + (long) k * 8 * n;
		const C *c01 = (const C *) tk, *c02 = (const C *) (tk + 2 * n), *c12 = (const C *) (tk + 4 * n);
		for (long i = lo; i < hi; i++) {
			C e[2][3], m[3][3];
			S(so_ch_exp)(c01[i], c02[i], c12[i], tk[6 * n + i], tk[7 * n + i], e);
			S(so_load_link)(u + (long) k * 9 * n, n, i, m);
			for (int r = 0; r < 2; r++)
				for (int c = 0; c < 3; c++) {
					expaux[((long) k * 9 + r * 3 + c) * n + i] = e[r][c];
					uout[((long) k * 9 + r * 3 + c) * n + i] = e[r][0] * m[0][c] + e[r][1] * m[1][c] + e[r][2] * m[2][c];
				}
		}
	}
}

/* stouting.c:74-100 */
void S(so_stout_isotropic)(const so_geom *g, const C *u, C *uprime, C *stap, C *aux, R *ta, double rho)
{
	memset(stap, 0, sizeof(C) * 72 * g->sizeh);
	S(so_calc_loc_staples_onlyferms)(g, u, stap);
	S(so_rho_times_conf_times_staples_ta_part)(g, u, stap, ta, rho);
	S(so_exp_minus_QA_times_conf)(g, u, ta, uprime, aux);
}
#undef SO_C_ZERO

/* ------------------------------------------------------------------ stout force chain Sigma' -> Sigma (stouting.c:171-1305)
 * thmat_soa[8] (struct_c_def.h:52-58) as reals, same packing as tamat: c01 @0, c02 @2*sizeh, c12 @4*sizeh (complex),
 * rc00 @6*sizeh, rc11 @7*sizeh.  The reference spells every 3x3 product out on the packed components; here the
 * same algebra is done on full matrices: Q = i*QA (hermitian, traceless), Lambda hermitian traceless.
 * Pinned against the reference build to rounding (1e-14), not bit for bit -- the operation order differs. */
static void S(so_q_from_qa)(const R *tk, long n, long i, C q[3][3])   /* Q = i * QA */
{
	const C c01 = ((const C *) tk)[i], c02 = ((const C *) (tk + 2 * n))[i], c12 = ((const C *) (tk + 4 * n))[i];
	const R i00 = tk[6 * n + i], i11 = tk[7 * n + i];
	q[0][0] = -i00; q[1][1] = -i11; q[2][2] = i00 + i11;
	q[0][1] = I * c01; q[1][0] = -I * CONJ(c01);
	q[0][2] = I * c02; q[2][0] = -I * CONJ(c02);
	q[1][2] = I * c12; q[2][1] = -I * CONJ(c12);
}
static void S(so_herm_from_thmat)(const R *tk, long n, long i, C l[3][3])
{
	const C c01 = ((const C *) tk)[i], c02 = ((const C *) (tk + 2 * n))[i], c12 = ((const C *) (tk + 4 * n))[i];
	const R r00 = tk[6 * n + i], r11 = tk[7 * n + i];
	l[0][0] = r00; l[1][1] = r11; l[2][2] = -r00 - r11;
	l[0][1] = c01; l[1][0] = CONJ(c01); l[0][2] = c02; l[2][0] = CONJ(c02); l[1][2] = c12; l[2][1] = CONJ(c12);
}
static void S(so_mm)(C a[3][3], C b[3][3], C o[3][3])
{
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) o[r][c] = a[r][0] * b[0][c] + a[r][1] * b[1][c] + a[r][2] * b[2][c];
}
/* Cayley-Hamilton coefficients f_j and their derivatives b_1j, b_2j (stouting.c:179-327, hep-lat/0311018 eqs. 57-69) */
static void S(so_ch_coeffs)(C q[3][3], C f[3], C b1[3], C b2[3])
{
	C q2[3][3]; S(so_mm)(q, q, q2);
	R c0 = S(so_creal)(q[0][0] * (q[1][1] * q[2][2] - q[1][2] * q[2][1]) - q[0][1] * (q[1][0] * q[2][2] - q[1][2] * q[2][0])
										 + q[0][2] * (q[1][0] * q[2][1] - q[1][1] * q[2][0]));          /* det Q */
	const R c1 = HALF * S(so_creal)(q2[0][0] + q2[1][1] + q2[2][2]);                     /* Tr Q^2 / 2 */
	const R c0max = 2 * RPOW(c1 / 3, (R) 1.5);
	if (c1 < (R) 4e-3) {
		f[0] = (1 - c0 * c0 / 720) + I * (-c0 * (1 - c1 * (1 - c1 / 42) / 20) / 6);
		f[1] = (c0 * (1 - c1 * (1 - 3 * c1 / 112) / 15) / 24) + I * (1 - c1 * (1 - c1 * (1 - c1 / 42) / 20) / 6 - c0 * c0 / 5040);
		f[2] = (HALF * (-1 + c1 * (1 - c1 * (1 - c1 / 56) / 30) / 12 + c0 * c0 / 20160)) + I * (HALF * (c0 * (1 - c1 * (1 - c1 / 48) / 21) / 60));
		b1[0] = 0 + I * (c0 / 120 * (1 - c1 / 21));
		b1[1] = (-c0 / 360 * (1 - 3 * c1 / 56)) + I * ((R) -1.0 / 6 * (1 - c1 / 10 * ((R) 1.0 - c1 / 28)));
		b1[2] = (HALF * ((R) 1.0 / 12 * (1 - 2 * c1 / 30 * (1 - 3 * c1 / 112)))) + I * (HALF * (-c0 / 1260 * (1 - c1 / 24)));
		b2[0] = (-c0 / 360) + I * ((R) -1.0 / 6 * (1 - c1 / 20 * (1 - c1 / 42)));
		b2[1] = ((R) 1.0 / 24 * (1 - c1 / 15 * (1 - 3 * c1 / 112))) + I * (-c0 / 2520);
		b2[2] = (HALF * c0 / 10080) + I * (HALF * ((R) 1.0 / 60 * (1 - c1 / 21 * (1 - c1 / 48))));
		return;
	}
	int sign = 1;
	if (c0 < 0) { sign = -1; c0 = -c0; }
	const R eps = (c0max - c0) / c0max;
	R theta;
	if (eps < 0) theta = 0;
	else if (eps < 1e-3) theta = RSQRT(2 * eps) * (1 + ((R) 1.0 / 12 + ((R) 3.0 / 160 + ((R) 5.0 / 896 + ((R) 35.0 / 18432 + (R) 63.0 / 90112 * eps) * eps) * eps) * eps) * eps);
	else theta = RACOS(c0 / c0max);
	const R u = RSQRT(c1 / 3) * RCOS(theta / 3), w = RSQRT(c1) * RSIN(theta / 3);
	const R u2 = u * u, w2 = w * w, u2mw2 = u2 - w2, w2p3u2 = w2 + 3 * u2, w2m3u2 = w2 - 3 * u2;
	const R cu = RCOS(u), c2u = RCOS(2 * u), su = RSIN(u), s2u = RSIN(2 * u), cw = RCOS(w);
	R xi0w, xi1w;
	if (RFABS(w) < (R) 0.05) { R t0 = w * w, t1 = 1 - t0 / 42, t2 = (R) 1.0 - t0 / 20 * t1; xi0w = 1 - t0 / 6 * t2; }
	else xi0w = RSIN(w) / w;
	if (RFABS(w) < (R) 0.05) xi1w = -(1 - w2 * (1 - w2 * (1 - w2 / 54) / 28) / 10) / 3;
	else xi1w = cw / w2 - RSIN(w) / (w2 * w);
	const R denom = 1 / (9 * u * u - w * w);
	f[0] = ((u2mw2 * c2u + cu * 8 * u2 * cw + 2 * su * u * w2p3u2 * xi0w) + I * (u2mw2 * s2u + -su * 8 * u2 * cw + cu * 2 * u * w2p3u2 * xi0w)) * denom;
	f[1] = ((2 * u * c2u + -cu * 2 * u * cw + -su * w2m3u2 * xi0w) + I * (2 * u * s2u + su * 2 * u * cw + -cu * w2m3u2 * xi0w)) * denom;
	f[2] = ((c2u + -cu * cw + -3 * su * u * xi0w) + I * (s2u + su * cw + -cu * 3 * u * xi0w)) * denom;
	C r1[3], r2[3];
	r1[0] = (2 * c2u * u + s2u * (-2 * u2 + 2 * w2) + 2 * cu * u * (8 * cw + 3 * u2 * xi0w + w2 * xi0w) + su * (-8 * cw * u2 + 18 * u2 * xi0w + 2 * w2 * xi0w))
		+ I * (-8 * cw * (2 * su * u + cu * u2) + 2 * (s2u * u + c2u * u2 - c2u * w2) + 2 * (9 * cu * u2 - 3 * su * u * u2 + cu * w2 - su * u * w2) * xi0w);
	r1[1] = (2 * c2u - 4 * s2u * u + su * (2 * cw * u + 6 * u * xi0w) + cu * (-2 * cw + 3 * u2 * xi0w - w2 * xi0w))
		+ I * (2 * s2u + 4 * c2u * u + 2 * cw * (su + cu * u) + (6 * cu * u - 3 * su * u2 + su * w2) * xi0w);
	r1[2] = (-2 * s2u + cw * su - 3 * (su + cu * u) * xi0w) + I * (2 * c2u + cu * cw + (-3 * cu + 3 * su * u) * xi0w);
	r2[0] = (-2 * c2u + 2 * cw * su * u + 2 * su * u * xi0w - 8 * cu * u2 * xi0w + 6 * su * u * u2 * xi1w)
		+ I * (2 * (-s2u + 4 * su * u2 * xi0w + cu * u * (cw + xi0w + 3 * u2 * xi1w)));
	r2[1] = (2 * cu * u * xi0w + su * (-cw - xi0w + 3 * u2 * xi1w)) + I * (-2 * su * u * xi0w - cu * (cw + xi0w - 3 * u2 * xi1w));
	r2[2] = (cu * xi0w - 3 * su * u * xi1w) + I * (-(su * xi0w) - 3 * cu * u * xi1w);
	for (int j = 0; j < 3; j++) {
		b1[j] = HALF * denom * denom * (2 * u * r1[j] + (3 * u * u - w * w) * r2[j] - 2 * (15 * u * u + w * w) * f[j]);   /* (57) */
		b2[j] = HALF * denom * denom * (r1[j] - 3 * u * r2[j] - 24 * u * f[j]);                                            /* (58) */
	}
	if (sign == -1) {
		b1[0] = CONJ(b1[0]); b1[1] = -CONJ(b1[1]); b1[2] = CONJ(b1[2]);
		b2[0] = -CONJ(b2[0]); b2[1] = CONJ(b2[1]); b2[2] = -CONJ(b2[2]);
		f[0] = CONJ(f[0]); f[1] = -CONJ(f[1]); f[2] = CONJ(f[2]);
	}
}

/* stouting.c:171-548: Lambda = traceless hermitian part of
 * Gamma = Tr(B1 U Sigma') Q + Tr(B2 U Sigma') Q^2 + f1 U Sigma' + f2 (Q U Sigma' + U Sigma' Q);  TMP is left = U Sigma' */
void S(so_compute_lambda)(const so_geom *g, R *lam, const C *sp, const C *u, const R *ta, C *tmp)
{
	const long n = g->sizeh, lo = (long) g->d3_halo * g->vol3h, hi = (long) (g->nd[3] - g->d3_halo) * g->vol3h;
	const R one_by_three = (R) 0.33333333333333333333333;
	for (int k = 0; k < 8; k++)
		for (long i = lo; i < hi; i++) {
			C q[3][3], q2[3][3], f[3], b1[3], b2[3], m[3][3], s[3][3], us[3][3], B[3][3], t[3][3], gm[3][3], qus[3][3], usq[3][3];
			S(so_q_from_qa)(ta + (long) k * 8 * n, n, i, q);
			S(so_ch_coeffs)(q, f, b1, b2);
			S(so_mm)(q, q, q2);
			S(so_load_link)(u + (long) k * 9 * n, n, i, m);
			for (int e = 0; e < 9; e++) s[e / 3][e % 3] = sp[((long) k * 9 + e) * n + i];
			S(so_mm)(m, s, us);
			C tr[2];
			for (int w = 0; w < 2; w++) {
				const C *b = w ? b2 : b1;
				for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B[r][c] = (r == c ? b[0] : 0) + b[1] * q[r][c] + b[2] * q2[r][c];
				S(so_mm)(B, us, t);
				tr[w] = t[0][0] + t[1][1] + t[2][2];
			}
			S(so_mm)(q, us, qus); S(so_mm)(us, q, usq);
			for (int r = 0; r < 3; r++)
				for (int c = 0; c < 3; c++) {
					gm[r][c] = tr[0] * q[r][c] + tr[1] * q2[r][c] + f[1] * us[r][c] + f[2] * (qus[r][c] + usq[r][c]);
					tmp[((long) k * 9 + r * 3 + c) * n + i] = us[r][c];
				}
			R *lk = lam + (long) k * 8 * n;
			lk[6 * n + i] = (2 * S(so_creal)(gm[0][0]) - S(so_creal)(gm[1][1]) - S(so_creal)(gm[2][2])) * one_by_three;
			lk[7 * n + i] = (2 * S(so_creal)(gm[1][1]) - S(so_creal)(gm[0][0]) - S(so_creal)(gm[2][2])) * one_by_three;
			((C *) lk)[i] = (gm[0][1] + CONJ(gm[1][0])) * HALF;
			((C *) (lk + 2 * n))[i] = (gm[0][2] + CONJ(gm[2][0])) * HALF;
			((C *) (lk + 4 * n))[i] = (gm[1][2] + CONJ(gm[2][1])) * HALF;
		}
}

/* stouting.c:550-1305: Sigma = Sigma' exp(iQ) + i rho sum_{nu != mu} [ right and left staples with one Lambda inserted ]
 *   right, A = U_nu(x+mu), B = U_mu(x+nu)^+, C = U_nu(x)^+ :  ABC (L_mu(x) - L_nu(x)) + L_nu(x+mu) ABC - A B L_mu(x+nu) C
 *   left,  A = U_nu(x+mu-nu)^+, B = U_mu(x-nu)^+, C = U_nu(x-nu) :
 *                           A B (L_nu(x-nu) - L_mu(x-nu)) C + A B C L_mu(x) - A L_nu(x+mu-nu) B C
 * TMP is left = exp(iQ). */
void S(so_compute_sigma)(const so_geom *g, const R *lam, const C *u, C *sg, const R *ta, C *tmp, double rho)
{
	const long n = g->sizeh;
	static const int perp[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
	const C irho = I * (R) rho;
	for (int d3 = g->d3_halo; d3 < g->nd[3] - g->d3_halo; d3++)
		for (int d2 = 0; d2 < g->nd[2]; d2++)
			for (int d1 = 0; d1 < g->nd[1]; d1++)
				for (int d0 = 0; d0 < g->nd[0]; d0++) {
					const int x[4] = { d0, d1, d2, d3 };
					const long idxh = so_snum(g, d0, d1, d2, d3);
					const int p = (d0 + d1 + d2 + d3) % 2;
					for (int mu = 0; mu < 4; mu++) {
						const int k = 2 * mu + p;
						C q[3][3], q2[3][3], f[3], b1[3], b2[3], e[3][3], s[3][3], res[3][3];
						S(so_q_from_qa)(ta + (long) k * 8 * n, n, idxh, q);
						S(so_ch_coeffs)(q, f, b1, b2);
						S(so_mm)(q, q, q2);
						for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) e[r][c] = (r == c ? f[0] : 0) + f[1] * q[r][c] + f[2] * q2[r][c];
						e[2][0] = CONJ(e[0][1] * e[1][2] - e[0][2] * e[1][1]);      /* third row rebuilt (:642-644) */
						e[2][1] = CONJ(e[0][2] * e[1][0] - e[0][0] * e[1][2]);
						e[2][2] = CONJ(e[0][0] * e[1][1] - e[0][1] * e[1][0]);
						for (int t = 0; t < 9; t++) { s[t / 3][t % 3] = sg[((long) k * 9 + t) * n + idxh]; tmp[((long) k * 9 + t) * n + idxh] = e[t / 3][t % 3]; }
						S(so_mm)(s, e, res);
						C lmu[3][3]; S(so_herm_from_thmat)(lam + (long) k * 8 * n, n, idxh, lmu);
						for (int it = 0; it < 3; it++) {
							const int nu = perp[mu][it];
							const long ipmu = S(so_shift)(g, x, mu, 1, nu, 0), ipnu = S(so_shift)(g, x, mu, 0, nu, 1);
							const long imnu = S(so_shift)(g, x, mu, 0, nu, -1), ipmumnu = S(so_shift)(g, x, mu, 1, nu, -1);
							C a[3][3], b[3][3], c[3][3], ab[3][3], abc[3][3], l1[3][3], l2[3][3], t1[3][3], t2[3][3];
							/* right */
							S(so_load_link)(u + (long) (2 * nu + !p) * 9 * n, n, ipmu, a);
							S(so_load_link)(u + (long) (2 * mu + !p) * 9 * n, n, ipnu, b); S(so_dagger)(b);
							S(so_load_link)(u + (long) (2 * nu + p) * 9 * n, n, idxh, c); S(so_dagger)(c);
							S(so_mm)(a, b, ab); S(so_mm)(ab, c, abc);
							S(so_herm_from_thmat)(lam + (long) (2 * nu + p) * 8 * n, n, idxh, l1);             /* E = L_nu(x) */
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) l2[r][cc] = lmu[r][cc] - l1[r][cc];
							S(so_mm)(abc, l2, t1);
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] += irho * t1[r][cc];
							S(so_herm_from_thmat)(lam + (long) (2 * nu + !p) * 8 * n, n, ipmu, l1);            /* F = L_nu(x+mu) */
							S(so_mm)(l1, abc, t1);
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] += irho * t1[r][cc];
							S(so_herm_from_thmat)(lam + (long) (2 * mu + !p) * 8 * n, n, ipnu, l1);            /* G = L_mu(x+nu) */
							S(so_mm)(ab, l1, t1); S(so_mm)(t1, c, t2);
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] -= irho * t2[r][cc];
							/* left */
							S(so_load_link)(u + (long) (2 * nu + p) * 9 * n, n, ipmumnu, a); S(so_dagger)(a);
							S(so_load_link)(u + (long) (2 * mu + !p) * 9 * n, n, imnu, b); S(so_dagger)(b);
							S(so_load_link)(u + (long) (2 * nu + !p) * 9 * n, n, imnu, c);
							S(so_mm)(a, b, ab); S(so_mm)(ab, c, abc);
							S(so_herm_from_thmat)(lam + (long) (2 * nu + !p) * 8 * n, n, imnu, l1);            /* G = L_nu(x-nu) */
							S(so_herm_from_thmat)(lam + (long) (2 * mu + !p) * 8 * n, n, imnu, l2);            /* E = L_mu(x-nu) */
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) l1[r][cc] -= l2[r][cc];
							S(so_mm)(ab, l1, t1); S(so_mm)(t1, c, t2);
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] += irho * t2[r][cc];
							S(so_mm)(abc, lmu, t1);                                                              /* D = L_mu(x) */
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] += irho * t1[r][cc];
							S(so_herm_from_thmat)(lam + (long) (2 * nu + p) * 8 * n, n, ipmumnu, l1);          /* F = L_nu(x+mu-nu) */
							S(so_mm)(a, l1, t1); S(so_mm)(t1, b, t2); S(so_mm)(t2, c, t1);
							for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) res[r][cc] -= irho * t1[r][cc];
						}
						for (int t = 0; t < 9; t++) sg[((long) k * 9 + t) * n + idxh] = res[t / 3][t % 3];
					}
				}
}

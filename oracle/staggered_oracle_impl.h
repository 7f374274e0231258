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
	S(so_axpy_like)(g, SO_IN1XMASS_MINUS_IN2, out, in, 0, 0, mass * mass + shift, 0);
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

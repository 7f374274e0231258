// Isotropic stout smearing, the producer of the links the Dirac operator reads:
//   OpenAcc/stouting.c:27-167            stout_wrapper, stout_isotropic, exp_minus_QA_times_conf
//   OpenAcc/plaquettes.c:196-255         calc_loc_staples_nnptrick_all_onlyferms  (+ su3_utilities.h:660-975)
//   OpenAcc/su3_utilities.c:210-237      RHO_times_conf_times_staples_ta_part     (+ su3_utilities.h:1097-1150)
//   OpenAcc/cayley_hamilton.h:24-180     CH_exponential_antihermitian_soa_nissalike (Morningstar-Peardon)
// One thread per half-lattice index, looping over its eight links.  The six staples of a link are accumulated in registers and,
// inside stout_isotropic, projected straight to Q = (rho/C_ZERO) TA(U S): the reference's 8 x 6 read-modify-write
// passes over the staple field become one store.  Neighbouring threads share 18 of the 19 links a thread reads, so
// the kernel runs out of L1/L2 and is bound by FP64 arithmetic (~2.7 kflop per link), not by HBM.
#include "staple_internal.cuh"
#include <cmath>

// globals the reference's stout_wrapper reads (action.h:6-23, alloc_vars.h:23,54-55); weak: the host's own win
extern "C" {
__attribute__((weak)) action_param act_params = {};   // OpenAcc/action.h:6-18, layout in staple_b200.h
__attribute__((weak)) double gl_stout_rho = 0.0, gl_topo_rho = 0.0;
__attribute__((weak)) su3_soa *auxbis_conf_acc = nullptr, *glocal_staples = nullptr;
__attribute__((weak)) tamat_soa *gipdot = nullptr;
__attribute__((weak)) su3_soa_f *auxbis_conf_acc_f = nullptr, *glocal_staples_f = nullptr;
__attribute__((weak)) tamat_soa_f *gipdot_f = nullptr;
}

namespace staple {

constexpr int kStoutBlock = 128;

template <typename T> __device__ __forceinline__ cplx_t<T> mks(T x, T y);
template <> __device__ __forceinline__ double2 mks<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mks<float>(float x, float y) { return make_float2(x, y); }

struct StoutGeom { int nd0, nd1, nd2, nd3; unsigned int vol3h; long sizeh; unsigned int lo, cnt; int tlsm; };
static StoutGeom stout_geom()
{
	const Geom &g = ctx().g;
	StoutGeom s;
	s.nd0 = g.nd0; s.nd1 = g.nd1; s.nd2 = g.nd2; s.nd3 = g.nd3; s.vol3h = (unsigned int) g.vol3h; s.sizeh = g.sizeh;
	s.lo = (unsigned int) (g.d3_halo * g.vol3h); s.cnt = (unsigned int) (g.loc_n3 * g.vol3h);
	// the gauge action is a compile-time choice of the reference that also fixes HALO_WIDTH (geometry_multidev.h:6-12: 2 for
	// ACTION_TYPE TLSM, 1 for WILSON) and C_ZERO (common_defines.h:70-85: 5/3 or 1): the run-time halo_width carries it here
	s.tlsm = g.halo_width == 2 ? 1 : 0;
	return s;
}

template <typename C> __device__ __forceinline__ C cmul_(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> __device__ __forceinline__ C conj_(C a) { a.y = -a.y; return a; }

// rows 0,1 from memory, third row = conj(r0 x r1)   (su3_utilities.h, everywhere)
template <typename T, bool THIRD = true>
__device__ __forceinline__ void load_link(const cplx_t<T> *uk, long n, unsigned int i, cplx_t<T> m[3][3])
{
	using C = cplx_t<T>;
#pragma unroll
	for (int c = 0; c < 3; c++) { m[0][c] = __ldg(uk + c * n + i); m[1][c] = __ldg(uk + (3 + c) * n + i); }
	if (!THIRD) return;            // a plain left factor only contributes its first two rows
	C a, b;
	a = cmul_(m[0][1], m[1][2]); b = cmul_(m[0][2], m[1][1]); m[2][0] = mks<T>(a.x - b.x, -(a.y - b.y));
	a = cmul_(m[0][2], m[1][0]); b = cmul_(m[0][0], m[1][2]); m[2][1] = mks<T>(a.x - b.x, -(a.y - b.y));
	a = cmul_(m[0][0], m[1][1]); b = cmul_(m[0][1], m[1][0]); m[2][2] = mks<T>(a.x - b.x, -(a.y - b.y));
}
// first two rows of op(a) * op(b), op = dagger where the flag says so; third row rebuilt
template <typename T, bool DA, bool DB, bool THIRD = true>
__device__ __forceinline__ void mul2(const cplx_t<T> a[3][3], const cplx_t<T> b[3][3], cplx_t<T> o[3][3])
{
	using C = cplx_t<T>;
#pragma unroll
	for (int r = 0; r < 2; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = mks<T>(0, 0);
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const C x = DA ? conj_(a[j][r]) : a[r][j];
				const C y = DB ? conj_(b[c][j]) : b[j][c];
				acc.x += x.x * y.x - x.y * y.y; acc.y += x.x * y.y + x.y * y.x;
			}
			o[r][c] = acc;
		}
	if (!THIRD) return;            // an intermediate product that is only used as a plain left factor
	C p, q;
	p = cmul_(o[0][1], o[1][2]); q = cmul_(o[0][2], o[1][1]); o[2][0] = mks<T>(p.x - q.x, -(p.y - q.y));
	p = cmul_(o[0][2], o[1][0]); q = cmul_(o[0][0], o[1][2]); o[2][1] = mks<T>(p.x - q.x, -(p.y - q.y));
	p = cmul_(o[0][0], o[1][1]); q = cmul_(o[0][1], o[1][0]); o[2][2] = mks<T>(p.x - q.x, -(p.y - q.y));
}

__device__ __forceinline__ unsigned int snum_dev(const StoutGeom &g, int d0, int d1, int d2, int d3)
{
	return (unsigned int) (d0 + g.nd0 * (d1 + g.nd1 * (d2 + g.nd2 * d3))) >> 1;
}
__device__ __forceinline__ int wrap(int x, int n) { return x < 0 ? x + n : (x >= n ? x - n : x); }

// Q = tmp * TA(P) into the five tamat arrays (su3_utilities.h:1144-1150), P a full 3x3 product
template <typename T, bool ASSIGN>
__device__ __forceinline__ void store_ta(T *tk, long n, unsigned int i, const cplx_t<T> p[3][3], T tmp)
{
	using C = cplx_t<T>;
	C *c01 = (C *) tk + i, *c02 = (C *) (tk + 2 * n) + i, *c12 = (C *) (tk + 4 * n) + i;
	T *ic00 = tk + 6 * n + i, *ic11 = tk + 7 * n + i;
	const T half = (T) 0.5, third = (T) 0.33333333333333333333333;
	const T tr = p[0][0].y + p[1][1].y + p[2][2].y;
	const C q01 = mks<T>(tmp * (half * (p[0][1].x - p[1][0].x)), tmp * (half * (p[0][1].y + p[1][0].y)));
	const C q02 = mks<T>(tmp * (half * (p[0][2].x - p[2][0].x)), tmp * (half * (p[0][2].y + p[2][0].y)));
	const C q12 = mks<T>(tmp * (half * (p[1][2].x - p[2][1].x)), tmp * (half * (p[1][2].y + p[2][1].y)));
	const T i00 = tmp * (p[0][0].y - third * tr), i11 = tmp * (p[1][1].y - third * tr);
	if (ASSIGN) { *c01 = q01; *c02 = q02; *c12 = q12; *ic00 = i00; *ic11 = i11; }
}

// MODE 0: calc_loc_staples_nnptrick_all_onlyferms -- loc_stap += C_ZERO * staples               (plaquettes.c:196-255)
// MODE 1: the first two thirds of stout_isotropic  -- loc_stap  = C_ZERO * staples (the reference zeroes it first),
//                                                     tipdot    = (rho/C_ZERO) TA(U * loc_stap)   (stouting.c:83-90)
template <typename T, int MODE>
__global__ void __launch_bounds__(kStoutBlock, 3) stout_staples_kernel(const cplx_t<T> *u, cplx_t<T> *stap, T *ta, T rho, StoutGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kStoutBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int idx = g.lo + t;
	const long n = g.sizeh;
	const int nd0h = g.nd0 >> 1;
	const int hd0 = idx % nd0h;
	unsigned int q = idx / nd0h;
	const int d1 = q % g.nd1; q /= g.nd1;
	const int d2 = q % g.nd2;
	const int d3 = q / g.nd2;
	const int nd[4] = { g.nd0, g.nd1, g.nd2, g.nd3 };
	const T c_zero = g.tlsm ? (T) 5.0 * (T) 0.33333333333333333333333 : (T) 1.0;   // C_ZERO (common_defines.h:73, :82)
	// all eight links of this half-lattice index (both parities, four directions) in ONE thread: the 8 x 18 link
	// reads around the two sites then hit L1/L2 while they are hot.  With the link index on grid.y instead, each
	// of the eight sweeps re-streamed most of the configuration from HBM (ncu: 3.3 GB read for 0.4 GB of links).
#pragma unroll 1
	for (int k = 0; k < 8; k++) {
		const int mu = k >> 1, p = k & 1;
		const int x[4] = { 2 * hd0 + ((d1 + d2 + d3 + p) & 1), d1, d2, d3 };
		C s[3][3];
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) s[r][c] = MODE == 0 ? stap[((long) k * 9 + r * 3 + c) * n + idx] : mks<T>(0, 0);
		auto site = [&](int dmu, int nu, int dnu) {
			int y[4] = { x[0], x[1], x[2], x[3] };
			y[mu] = wrap(y[mu] + dmu, nd[mu]);
			y[nu] = wrap(y[nu] + dnu, nd[nu]);
			return snum_dev(g, y[0], y[1], y[2], y[3]);
		};
#pragma unroll 1
		for (int it = 0; it < 3; it++) {
			const int nu = it + (it >= mu ? 1 : 0);          // perp_dirs[mu][it]
			C a[3][3], b[3][3], ab[3][3];
			// right: U_nu(x+mu) U_mu(x+nu)^+ U_nu(x)^+
			load_link<T, false>(u + (long) (2 * nu + !p) * 9 * n, n, site(1, nu, 0), a);
			load_link<T>(u + (long) (2 * mu + !p) * 9 * n, n, site(0, nu, 1), b);
			mul2<T, false, true, false>(a, b, ab);
			load_link<T>(u + (long) (2 * nu + p) * 9 * n, n, idx, b);
			mul2<T, false, true>(ab, b, a);
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int c = 0; c < 3; c++) { s[r][c].x += c_zero * a[r][c].x; s[r][c].y += c_zero * a[r][c].y; }
			// left: U_nu(x+mu-nu)^+ U_mu(x-nu)^+ U_nu(x-nu)
			const unsigned int imnu = site(0, nu, -1);
			load_link<T>(u + (long) (2 * nu + p) * 9 * n, n, site(1, nu, -1), a);
			load_link<T>(u + (long) (2 * mu + !p) * 9 * n, n, imnu, b);
			mul2<T, true, true, false>(a, b, ab);
			load_link<T>(u + (long) (2 * nu + !p) * 9 * n, n, imnu, b);
			mul2<T, false, false>(ab, b, a);
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int c = 0; c < 3; c++) { s[r][c].x += c_zero * a[r][c].x; s[r][c].y += c_zero * a[r][c].y; }
		}
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) stap[((long) k * 9 + r * 3 + c) * n + idx] = s[r][c];
		if (MODE == 1) {
			C m[3][3], pr[3][3];
			load_link<T>(u + (long) k * 9 * n, n, idx, m);
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int c = 0; c < 3; c++) {
					C acc = cmul_(m[r][0], s[0][c]);
					const C b1 = cmul_(m[r][1], s[1][c]), b2 = cmul_(m[r][2], s[2][c]);
					acc.x += b1.x; acc.y += b1.y; acc.x += b2.x; acc.y += b2.y;
					pr[r][c] = acc;
				}
			store_ta<T, true>(ta + (long) k * 8 * n, n, idx, pr, rho / c_zero);
		}
	}
}

// RHO_times_conf_times_staples_ta_part (su3_utilities.c:210-237) on a staple field already in memory
template <typename T>
__global__ void __launch_bounds__(kStoutBlock) stout_rho_ta_kernel(const cplx_t<T> *u, const cplx_t<T> *stap, T *ta, T rho, StoutGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kStoutBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int idx = g.lo + t;
	const long n = g.sizeh;
	const int k = blockIdx.y;
	C m[3][3], pr[3][3], s[3][3];
	load_link<T>(u + (long) k * 9 * n, n, idx, m);
#pragma unroll
	for (int e = 0; e < 9; e++) s[e / 3][e % 3] = stap[((long) k * 9 + e) * n + idx];
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = cmul_(m[r][0], s[0][c]);
			const C b1 = cmul_(m[r][1], s[1][c]), b2 = cmul_(m[r][2], s[2][c]);
			acc.x += b1.x; acc.y += b1.y; acc.x += b2.x; acc.y += b2.y;
			pr[r][c] = acc;
		}
	const T c_zero = g.tlsm ? (T) 5.0 * (T) 0.33333333333333333333333 : (T) 1.0;
	store_ta<T, true>(ta + (long) k * 8 * n, n, idx, pr, rho / c_zero);
}

template <typename T> struct MathOf;
template <> struct MathOf<double> {
	static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
	static __device__ __forceinline__ double acos_(double x) { return acos(x); }
	static __device__ __forceinline__ double pow15(double x) { return pow(x, 1.5); }
	static __device__ __forceinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
};
template <> struct MathOf<float> {
	static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
	static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
	static __device__ __forceinline__ float pow15(float x) { return powf(x, 1.5f); }
	static __device__ __forceinline__ void sincos_(float x, float *s, float *c) { sincosf(x, s, c); }
};

// exp_minus_QA_times_conf (stouting.c:107-167) with CH_exponential_antihermitian_soa_nissalike (cayley_hamilton.h:52-180):
// exp_aux rows 0,1 = exp(-QA) = exp(iQ), tu_out rows 0,1 = exp_aux * U.  The coefficient formulas follow the reference
// line by line (MILC/NISSA small-c1 series below 4e-3, the eps < 1e-3 series for theta, xi0 series for |w| < 0.05).
template <typename T>
__global__ void __launch_bounds__(kStoutBlock) stout_exp_kernel(const cplx_t<T> *u, const T *ta, cplx_t<T> *uout, cplx_t<T> *expaux, StoutGeom g)
{
	using C = cplx_t<T>;
	using M = MathOf<T>;
	const unsigned int t = blockIdx.x * kStoutBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int i = g.lo + t;
	const long n = g.sizeh;
	const int k = blockIdx.y;
	const T *tk = ta + (long) k * 8 * n;
	const C q01 = *((const C *) tk + i), q02 = *((const C *) (tk + 2 * n) + i), q12 = *((const C *) (tk + 4 * n) + i);
	const T i00 = tk[6 * n + i], i11 = tk[7 * n + i], i22 = -i00 - i11;
	const T n01 = q01.x * q01.x + q01.y * q01.y, n02 = q02.x * q02.x + q02.y * q02.y, n12 = q12.x * q12.x + q12.y * q12.y;
	// det(Q) (:24-37) and Tr(Q^2)/2 (:39-50)
	const C t3 = cmul_(cmul_(q01, q12), conj_(q02));
	T c0 = -(i00 * i11 * i22 + 2 * t3.y - i00 * n12 - i11 * n02 - n01 * i22);
	const T c1 = (T) 0.5 * (2 * (i00 * i00 + i11 * i11 + i00 * i11 + n01 + n02 + n12));
	const T c0max = 2 * M::pow15(c1 / 3);
	C f0, f1, f2;
	if (c1 < (T) 4.0e-3) {
		f0 = mks<T>(1 - c0 * c0 / 720, -c0 * (1 - c1 * (1 - c1 / 42) / 20) / 6);
		f1 = mks<T>(c0 * (1 - c1 * (1 - 3 * c1 / 112) / 15) / 24, 1 - c1 * (1 - c1 * (1 - c1 / 42) / 20) / 6 - c0 * c0 / 5040);
		f2 = mks<T>((T) 0.5 * (-1 + c1 * (1 - c1 * (1 - c1 / 56) / 30) / 12 + c0 * c0 / 20160), (T) 0.5 * (c0 * (1 - c1 * (1 - c1 / 48) / 21) / 60));
	} else {
		int sign = 1;
		if (c0 < 0) { sign = -1; c0 = -c0; }
		const T eps = (c0max - c0) / c0max;
		T theta;
		if (eps < 0) theta = 0;
		else if (eps < (T) 1e-3)
			theta = M::sqrt_(2 * eps) * (1 + ((T) 1.0 / 12 + ((T) 3.0 / 160 + ((T) 5.0 / 896 + ((T) 35.0 / 18432 + (T) 63.0 / 90112 * eps) * eps) * eps) * eps) * eps);
		else theta = M::acos_(c0 / c0max);
		T st3, ct3; M::sincos_(theta / 3, &st3, &ct3);
		const T uu = M::sqrt_(c1 / 3) * ct3, w = M::sqrt_(c1) * st3;
		const T u2 = uu * uu, w2 = w * w, u2mw2 = u2 - w2, w2p3u2 = w2 + 3 * u2, w2m3u2 = w2 - 3 * u2;
		T su, cu, s2u, c2u, sw, cw;
		M::sincos_(uu, &su, &cu); M::sincos_(2 * uu, &s2u, &c2u); M::sincos_(w, &sw, &cw);
		T xi0w;
		if (fabs((double) w) < 0.05) { const T t0 = w * w, t1 = 1 - t0 / 42, t2 = (T) 1.0 - t0 / 20 * t1; xi0w = 1 - t0 / 6 * t2; }
		else xi0w = sw / w;
		const T denom = 1 / (9 * uu * uu - w * w);
		f0 = mks<T>((u2mw2 * c2u + cu * 8 * u2 * cw + 2 * su * uu * w2p3u2 * xi0w) * denom,
								(u2mw2 * s2u + -su * 8 * u2 * cw + cu * 2 * uu * w2p3u2 * xi0w) * denom);
		f1 = mks<T>((2 * uu * c2u + -cu * 2 * uu * cw + -su * w2m3u2 * xi0w) * denom,
								(2 * uu * s2u + su * 2 * uu * cw + -cu * w2m3u2 * xi0w) * denom);
		f2 = mks<T>((c2u + -cu * cw + -3 * su * uu * xi0w) * denom, (s2u + su * cw + -cu * 3 * uu * xi0w) * denom);
		if (sign == -1) { f0 = conj_(f0); f1 = mks<T>(-f1.x, f1.y); f2 = conj_(f2); }
	}
	// eq. 19: exp(iQ) = f0 + f1 Q + f2 Q^2 written out on the tamat components (:139-172)
	auto add = [](C a, C b) { return mks<T>(a.x + b.x, a.y + b.y); };
	auto scl = [](C a, T s) { return mks<T>(a.x * s, a.y * s); };
	auto muli = [](C a) { return mks<T>(-a.y, a.x); };           // i * a
	const C if1 = muli(f1);
	C e[2][3];
	e[0][0] = add(mks<T>(f0.x - f1.x * i00, f0.y - f1.y * i00), scl(f2, i00 * i00 + n01 + n02));
	e[0][1] = add(cmul_(if1, q01), cmul_(f2, add(cmul_(q02, conj_(q12)), scl(muli(q01), -(i00 + i11)))));
	e[0][2] = add(cmul_(if1, q02), cmul_(f2, add(scl(cmul_(q01, q12), (T) -1), scl(muli(q02), i11))));
	e[1][0] = add(scl(cmul_(if1, conj_(q01)), (T) -1), cmul_(f2, add(cmul_(q12, conj_(q02)), scl(muli(conj_(q01)), i00 + i11))));
	e[1][1] = add(mks<T>(f0.x - f1.x * i11, f0.y - f1.y * i11), scl(f2, i11 * i11 + n01 + n12));
	e[1][2] = add(cmul_(if1, q12), cmul_(f2, add(scl(muli(q12), i00), cmul_(q02, conj_(q01)))));
	C m[3][3];
	load_link<T>(u + (long) k * 9 * n, n, i, m);
#pragma unroll
	for (int r = 0; r < 2; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = cmul_(e[r][0], m[0][c]);
			const C b1 = cmul_(e[r][1], m[1][c]), b2 = cmul_(e[r][2], m[2][c]);
			acc.x += b1.x; acc.y += b1.y; acc.x += b2.x; acc.y += b2.y;
			expaux[((long) k * 9 + r * 3 + c) * n + i] = e[r][c];
			uout[((long) k * 9 + r * 3 + c) * n + i] = acc;
		}
}

template <typename T>
static void stout_isotropic_t(const cplx_t<T> *u, cplx_t<T> *uprime, cplx_t<T> *stap, cplx_t<T> *aux, T *ta, int istopo)
{
	const StoutGeom g = stout_geom();
	const dim3 grid((g.cnt + kStoutBlock - 1) / kStoutBlock, 8);
	const T rho = (T) (istopo ? gl_topo_rho : gl_stout_rho);
	// set_su3_soa_to_zero(local_staples) (stouting.c:85): the halo slices of the parking field are zero afterwards
	STAPLE_CUDA_CHECK(cudaMemsetAsync(stap, 0, sizeof(cplx_t<T>) * 72 * g.sizeh, ctx().stream));
	stout_staples_kernel<T, 1><<<grid.x, kStoutBlock, 0, ctx().stream>>>(u, stap, ta, rho, g);
	stout_exp_kernel<T><<<grid, kStoutBlock, 0, ctx().stream>>>(u, ta, uprime, aux, g);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch(2);
}

}   // namespace staple

using namespace staple;

#define DD(p) ((double2 *) dev(p, #p))
#define DF(p) ((float2 *) dev(p, #p))
#define CDD(p) ((const double2 *) dev(p, #p))
#define CDF(p) ((const float2 *) dev(p, #p))

extern "C" {

void communicate_su3_borders(su3_soa *lnh_conf, int thickness);
void communicate_su3_borders_f(su3_soa_f *lnh_conf, int thickness);

#define STAPLE_STOUT_DEF(S, T, D, CD, SU3, TAMAT, AUXBIS, GSTAP, GIPDOT)                                                    \
	void calc_loc_staples_nnptrick_all_onlyferms##S(const SU3 *u, SU3 *loc_stap)                                              \
	{                                                                                                                         \
		require_init("calc_loc_staples_nnptrick_all_onlyferms");                                                                \
		const StoutGeom g = stout_geom();                                                                                       \
		stout_staples_kernel<T, 0><<<(g.cnt + kStoutBlock - 1) / kStoutBlock, kStoutBlock, 0, ctx().stream>>>(                  \
			CD(u), D(loc_stap), nullptr, (T) 0, g);                                                                               \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                  \
	}                                                                                                                         \
	void RHO_times_conf_times_staples_ta_part##S(const SU3 *u, const SU3 *loc_stap, TAMAT *tipdot, int istopo)                \
	{                                                                                                                         \
		require_init("RHO_times_conf_times_staples_ta_part");                                                                   \
		const StoutGeom g = stout_geom();                                                                                       \
		stout_rho_ta_kernel<T><<<dim3((g.cnt + kStoutBlock - 1) / kStoutBlock, 8), kStoutBlock, 0, ctx().stream>>>(             \
			CD(u), CD(loc_stap), (T *) dev(tipdot, "tipdot"), (T) (istopo ? gl_topo_rho : gl_stout_rho), g);                      \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                  \
	}                                                                                                                         \
	void exp_minus_QA_times_conf##S(const SU3 *tu, const TAMAT *QA, SU3 *tu_out, SU3 *exp_aux)                                \
	{                                                                                                                         \
		require_init("exp_minus_QA_times_conf");                                                                                \
		const StoutGeom g = stout_geom();                                                                                       \
		stout_exp_kernel<T><<<dim3((g.cnt + kStoutBlock - 1) / kStoutBlock, 8), kStoutBlock, 0, ctx().stream>>>(                \
			CD(tu), (const T *) dev(QA, "QA"), D(tu_out), D(exp_aux), g);                                                         \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                  \
	}                                                                                                                         \
	void stout_isotropic##S(const SU3 *u, SU3 *uprime, SU3 *local_staples, SU3 *auxiliary, TAMAT *tipdot, const int istopo)   \
	{                                                                                                                         \
		require_init("stout_isotropic");                                                                                        \
		if (verbosity_lv > 1 && 0 == ctx().myrank) printf("Isotropic stouting...\n");                                           \
		stout_isotropic_t<T>(CD(u), D(uprime), D(local_staples), D(auxiliary), (T *) dev(tipdot, "tipdot"), istopo);            \
		if (verbosity_lv > 1 && 0 == ctx().myrank) printf("Isotropic stouting done\n");                                         \
	}                                                                                                                         \
	/* stouting.c:27-72: level 0 smears tconf_acc, level l smears level l-1; link halos (thickness GAUGE_HALO = 2) */        \
	/* are exchanged after every level.  Parking arrays and parameters: the reference's globals (weak here). */              \
	void stout_wrapper##S(const SU3 *tconf_acc, SU3 *tstout_conf_acc_arr, const int istopo)                                   \
	{                                                                                                                         \
		require_init("stout_wrapper");                                                                                          \
		const int stoutsteps = (istopo & act_params.topo_action) ? act_params.topo_stout_steps : act_params.stout_steps;        \
		if (verbosity_lv > 1 && 0 == ctx().myrank) printf(":Stouting gauge conf %d times.\n", stoutsteps);                      \
		if (stoutsteps > 0 && (!AUXBIS || !GSTAP || !GIPDOT)) {                                                                 \
			fprintf(stderr, "stout_wrapper: " #AUXBIS ", " #GSTAP ", " #GIPDOT " (alloc_vars globals) are not set\n"); exit(1);  \
		}                                                                                                                       \
		const size_t conf_bytes = sizeof(T) * 2 * 72 * ctx().g.sizeh;                                                           \
		for (int level = 0; level < stoutsteps; level++) {                                                                      \
			const SU3 *src = level == 0 ? tconf_acc : (const SU3 *) ((const char *) tstout_conf_acc_arr + (size_t) (level - 1) * conf_bytes); \
			SU3 *dst = (SU3 *) ((char *) tstout_conf_acc_arr + (size_t) level * conf_bytes);                                      \
			stout_isotropic##S(src, dst, AUXBIS, GSTAP, GIPDOT, istopo);                                                          \
			if (ctx().nranks > 1) communicate_su3_borders##S(dst, 2);                                                             \
		}                                                                                                                       \
	}
STAPLE_STOUT_DEF(, double, DD, CDD, su3_soa, tamat_soa, auxbis_conf_acc, glocal_staples, gipdot)
STAPLE_STOUT_DEF(_f, float, DF, CDF, su3_soa_f, tamat_soa_f, auxbis_conf_acc_f, glocal_staples_f, gipdot_f)

}   // extern "C"

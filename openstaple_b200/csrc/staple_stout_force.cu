// The stouted fermion force: Sigma' -> Sigma through one smearing level (the force counterpart of stout_isotropic,
// SURVEY 8f row N4) and the chain that drives it.
//   OpenAcc/stouting.c:171-548     compute_loc_Lambda / compute_lambda   (hep-lat/0311018 eqs. 57-73)
//   OpenAcc/stouting.c:550-1305    compute_sigma_local_PEZZO1, the six RIGHT_/LEFT_ staple helpers, compute_sigma
//   OpenAcc/fermion_force.c:52-163 compute_sigma_from_sigma_prime_backinto_sigma_prime
// The reference spells every 3x3 product out on the packed tamat / thmat components; here the same algebra runs on
// full 3x3 matrices in registers: Q = i*QA (hermitian traceless), Lambda hermitian traceless, links with the third
// row rebuilt.  compute_sigma: one thread per half-lattice index looping over its eight links (as the staple
// kernel does, so the ~150 link and Lambda reads around the two sites stay in L1/L2); FP64-bound.
#include "staple_internal.cuh"
#include <cmath>

namespace staple {

constexpr int kSfBlock = 128;

template <typename T> __device__ __forceinline__ cplx_t<T> mkq(T x, T y);
template <> __device__ __forceinline__ double2 mkq<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mkq<float>(float x, float y) { return make_float2(x, y); }

struct SfGeom { int nd0, nd1, nd2, nd3; long sizeh; unsigned int lo, cnt; };
static SfGeom sf_geom()
{
	const Geom &g = ctx().g;
	SfGeom s;
	s.nd0 = g.nd0; s.nd1 = g.nd1; s.nd2 = g.nd2; s.nd3 = g.nd3; s.sizeh = g.sizeh;
	s.lo = (unsigned int) (g.d3_halo * g.vol3h); s.cnt = (unsigned int) (g.loc_n3 * g.vol3h);
	return s;
}

template <typename C> __device__ __forceinline__ C cm(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> __device__ __forceinline__ C cj(C a) { a.y = -a.y; return a; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }

template <typename T, bool DA = false, bool DB = false>
__device__ __forceinline__ void mm(const cplx_t<T> a[3][3], const cplx_t<T> b[3][3], cplx_t<T> o[3][3])   // op(a) * op(b)
{
	using C = cplx_t<T>;
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = mkq<T>(0, 0);
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const C x = DA ? cj(a[j][r]) : a[r][j];
				const C y = DB ? cj(b[c][j]) : b[j][c];
				acc.x += x.x * y.x - x.y * y.y; acc.y += x.x * y.y + x.y * y.x;
			}
			o[r][c] = acc;
		}
}
template <typename T>
__device__ __forceinline__ void load_su3(const cplx_t<T> *uk, long n, unsigned int i, cplx_t<T> m[3][3])
{
	using C = cplx_t<T>;
#pragma unroll
	for (int c = 0; c < 3; c++) { m[0][c] = __ldg(uk + c * n + i); m[1][c] = __ldg(uk + (3 + c) * n + i); }
	m[2][0] = cj(csub(cm(m[0][1], m[1][2]), cm(m[0][2], m[1][1])));
	m[2][1] = cj(csub(cm(m[0][2], m[1][0]), cm(m[0][0], m[1][2])));
	m[2][2] = cj(csub(cm(m[0][0], m[1][1]), cm(m[0][1], m[1][0])));
}
// Q = i*QA from the packed tamat components
template <typename T>
__device__ __forceinline__ void load_q(const T *tk, long n, unsigned int i, cplx_t<T> q[3][3])
{
	using C = cplx_t<T>;
	const C c01 = __ldg((const C *) tk + i), c02 = __ldg((const C *) (tk + 2 * n) + i), c12 = __ldg((const C *) (tk + 4 * n) + i);
	const T i00 = __ldg(tk + 6 * n + i), i11 = __ldg(tk + 7 * n + i);
	q[0][0] = mkq<T>(-i00, 0); q[1][1] = mkq<T>(-i11, 0); q[2][2] = mkq<T>(i00 + i11, 0);
	q[0][1] = mkq<T>(-c01.y, c01.x); q[1][0] = mkq<T>(-c01.y, -c01.x);      // i c, -i conj(c)
	q[0][2] = mkq<T>(-c02.y, c02.x); q[2][0] = mkq<T>(-c02.y, -c02.x);
	q[1][2] = mkq<T>(-c12.y, c12.x); q[2][1] = mkq<T>(-c12.y, -c12.x);
}
// hermitian traceless Lambda from the packed thmat components
template <typename T>
__device__ __forceinline__ void load_herm(const T *tk, long n, unsigned int i, cplx_t<T> l[3][3])
{
	using C = cplx_t<T>;
	const C c01 = __ldg((const C *) tk + i), c02 = __ldg((const C *) (tk + 2 * n) + i), c12 = __ldg((const C *) (tk + 4 * n) + i);
	const T r00 = __ldg(tk + 6 * n + i), r11 = __ldg(tk + 7 * n + i);
	l[0][0] = mkq<T>(r00, 0); l[1][1] = mkq<T>(r11, 0); l[2][2] = mkq<T>(-r00 - r11, 0);
	l[0][1] = c01; l[1][0] = cj(c01); l[0][2] = c02; l[2][0] = cj(c02); l[1][2] = c12; l[2][1] = cj(c12);
}

template <typename T> struct Mth;
template <> struct Mth<double> {
	static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
	static __device__ __forceinline__ double acos_(double x) { return acos(x); }
	static __device__ __forceinline__ double pow15(double x) { return pow(x, 1.5); }
	static __device__ __forceinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
};
template <> struct Mth<float> {
	static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
	static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
	static __device__ __forceinline__ float pow15(float x) { return powf(x, 1.5f); }
	static __device__ __forceinline__ void sincos_(float x, float *s, float *c) { sincosf(x, s, c); }
};

// Cayley-Hamilton coefficients f_j of exp(iQ) and, if WITH_B, their derivatives b_1j, b_2j (stouting.c:179-327;
// the small-c1 series, the eps < 1e-3 series for theta and the xi0/xi1 series for |w| < 0.05 as in the reference)
template <typename T, bool WITH_B>
__device__ __noinline__ void ch_coeffs(const cplx_t<T> q[3][3], const cplx_t<T> q2[3][3], cplx_t<T> f[3], cplx_t<T> b1[3], cplx_t<T> b2[3])
{
	using C = cplx_t<T>;
	using M = Mth<T>;
	const C d0 = csub(cm(q[1][1], q[2][2]), cm(q[1][2], q[2][1]));
	const C d1 = csub(cm(q[1][0], q[2][2]), cm(q[1][2], q[2][0]));
	const C d2 = csub(cm(q[1][0], q[2][1]), cm(q[1][1], q[2][0]));
	T c0 = cm(q[0][0], d0).x - cm(q[0][1], d1).x + cm(q[0][2], d2).x;          // det Q (real)
	const T c1 = (T) 0.5 * (q2[0][0].x + q2[1][1].x + q2[2][2].x);                // Tr Q^2 / 2
	const T c0max = 2 * M::pow15(c1 / 3);
	if (c1 < (T) 4e-3) {
		f[0] = mkq<T>(1 - c0 * c0 / 720, -c0 * (1 - c1 * (1 - c1 / 42) / 20) / 6);
		f[1] = mkq<T>(c0 * (1 - c1 * (1 - 3 * c1 / 112) / 15) / 24, 1 - c1 * (1 - c1 * (1 - c1 / 42) / 20) / 6 - c0 * c0 / 5040);
		f[2] = mkq<T>((T) 0.5 * (-1 + c1 * (1 - c1 * (1 - c1 / 56) / 30) / 12 + c0 * c0 / 20160), (T) 0.5 * (c0 * (1 - c1 * (1 - c1 / 48) / 21) / 60));
		if (WITH_B) {
			b1[0] = mkq<T>(0, c0 / 120 * (1 - c1 / 21));
			b1[1] = mkq<T>(-c0 / 360 * (1 - 3 * c1 / 56), (T) -1.0 / 6 * (1 - c1 / 10 * ((T) 1.0 - c1 / 28)));
			b1[2] = mkq<T>((T) 0.5 * ((T) 1.0 / 12 * (1 - 2 * c1 / 30 * (1 - 3 * c1 / 112))), (T) 0.5 * (-c0 / 1260 * (1 - c1 / 24)));
			b2[0] = mkq<T>(-c0 / 360, (T) -1.0 / 6 * (1 - c1 / 20 * (1 - c1 / 42)));
			b2[1] = mkq<T>((T) 1.0 / 24 * (1 - c1 / 15 * (1 - 3 * c1 / 112)), -c0 / 2520);
			b2[2] = mkq<T>((T) 0.5 * c0 / 10080, (T) 0.5 * ((T) 1.0 / 60 * (1 - c1 / 21 * (1 - c1 / 48))));
		}
		return;
	}
	int sign = 1;
	if (c0 < 0) { sign = -1; c0 = -c0; }
	const T eps = (c0max - c0) / c0max;
	T theta;
	if (eps < 0) theta = 0;
	else if (eps < (T) 1e-3)
		theta = M::sqrt_(2 * eps) * (1 + ((T) 1.0 / 12 + ((T) 3.0 / 160 + ((T) 5.0 / 896 + ((T) 35.0 / 18432 + (T) 63.0 / 90112 * eps) * eps) * eps) * eps) * eps);
	else theta = M::acos_(c0 / c0max);
	T st3, ct3; M::sincos_(theta / 3, &st3, &ct3);
	const T u = M::sqrt_(c1 / 3) * ct3, w = M::sqrt_(c1) * st3;
	const T u2 = u * u, w2 = w * w, u2mw2 = u2 - w2, w2p3u2 = w2 + 3 * u2, w2m3u2 = w2 - 3 * u2;
	T su, cu, s2u, c2u, sw, cw;
	M::sincos_(u, &su, &cu); M::sincos_(2 * u, &s2u, &c2u); M::sincos_(w, &sw, &cw);
	const bool smallw = fabs((double) w) < 0.05;
	T xi0w, xi1w = 0;
	if (smallw) { const T t0 = w * w, t1 = 1 - t0 / 42, t2 = (T) 1.0 - t0 / 20 * t1; xi0w = 1 - t0 / 6 * t2; }
	else xi0w = sw / w;
	if (WITH_B) xi1w = smallw ? -(1 - w2 * (1 - w2 * (1 - w2 / 54) / 28) / 10) / 3 : cw / w2 - sw / (w2 * w);
	const T denom = 1 / (9 * u * u - w * w);
	f[0] = mkq<T>((u2mw2 * c2u + cu * 8 * u2 * cw + 2 * su * u * w2p3u2 * xi0w) * denom, (u2mw2 * s2u + -su * 8 * u2 * cw + cu * 2 * u * w2p3u2 * xi0w) * denom);
	f[1] = mkq<T>((2 * u * c2u + -cu * 2 * u * cw + -su * w2m3u2 * xi0w) * denom, (2 * u * s2u + su * 2 * u * cw + -cu * w2m3u2 * xi0w) * denom);
	f[2] = mkq<T>((c2u + -cu * cw + -3 * su * u * xi0w) * denom, (s2u + su * cw + -cu * 3 * u * xi0w) * denom);
	if (WITH_B) {
		C r1[3], r2[3];
		r1[0] = mkq<T>(2 * c2u * u + s2u * (-2 * u2 + 2 * w2) + 2 * cu * u * (8 * cw + 3 * u2 * xi0w + w2 * xi0w) + su * (-8 * cw * u2 + 18 * u2 * xi0w + 2 * w2 * xi0w),
									 -8 * cw * (2 * su * u + cu * u2) + 2 * (s2u * u + c2u * u2 - c2u * w2) + 2 * (9 * cu * u2 - 3 * su * u * u2 + cu * w2 - su * u * w2) * xi0w);
		r1[1] = mkq<T>(2 * c2u - 4 * s2u * u + su * (2 * cw * u + 6 * u * xi0w) + cu * (-2 * cw + 3 * u2 * xi0w - w2 * xi0w),
									 2 * s2u + 4 * c2u * u + 2 * cw * (su + cu * u) + (6 * cu * u - 3 * su * u2 + su * w2) * xi0w);
		r1[2] = mkq<T>(-2 * s2u + cw * su - 3 * (su + cu * u) * xi0w, 2 * c2u + cu * cw + (-3 * cu + 3 * su * u) * xi0w);
		r2[0] = mkq<T>(-2 * c2u + 2 * cw * su * u + 2 * su * u * xi0w - 8 * cu * u2 * xi0w + 6 * su * u * u2 * xi1w,
									 2 * (-s2u + 4 * su * u2 * xi0w + cu * u * (cw + xi0w + 3 * u2 * xi1w)));
		r2[1] = mkq<T>(2 * cu * u * xi0w + su * (-cw - xi0w + 3 * u2 * xi1w), -2 * su * u * xi0w - cu * (cw + xi0w - 3 * u2 * xi1w));
		r2[2] = mkq<T>(cu * xi0w - 3 * su * u * xi1w, -(su * xi0w) - 3 * cu * u * xi1w);
		const T hd2 = (T) 0.5 * denom * denom, k1 = 3 * u * u - w * w, k2 = 2 * (15 * u * u + w * w);
#pragma unroll
		for (int j = 0; j < 3; j++) {
			b1[j] = mkq<T>(hd2 * (2 * u * r1[j].x + k1 * r2[j].x - k2 * f[j].x), hd2 * (2 * u * r1[j].y + k1 * r2[j].y - k2 * f[j].y));   // (57)
			b2[j] = mkq<T>(hd2 * (r1[j].x - 3 * u * r2[j].x - 24 * u * f[j].x), hd2 * (r1[j].y - 3 * u * r2[j].y - 24 * u * f[j].y));     // (58)
		}
	}
	if (sign == -1) {
		if (WITH_B) {
			b1[0] = cj(b1[0]); b1[1] = mkq<T>(-b1[1].x, b1[1].y); b1[2] = cj(b1[2]);
			b2[0] = mkq<T>(-b2[0].x, b2[0].y); b2[1] = cj(b2[1]); b2[2] = mkq<T>(-b2[2].x, b2[2].y);
		}
		f[0] = cj(f[0]); f[1] = mkq<T>(-f[1].x, f[1].y); f[2] = cj(f[2]);
	}
}

// compute_lambda (stouting.c:516-548 over compute_loc_Lambda :171-514): one thread per link
//   Gamma = Tr(B1 U S') Q + Tr(B2 U S') Q^2 + f1 U S' + f2 (Q U S' + U S' Q),  B_i = b_i0 + b_i1 Q + b_i2 Q^2
//   Lambda = (Gamma + Gamma^+)/2 - Tr(...)/6 ;  TMP is left = U S' as in the reference
template <typename T>
__global__ void __launch_bounds__(kSfBlock) stout_lambda_kernel(T *lam, const cplx_t<T> *sp, const cplx_t<T> *u, const T *ta, cplx_t<T> *tmp, SfGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kSfBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int i = g.lo + t;
	const long n = g.sizeh;
	const int k = blockIdx.y;
	C q[3][3], q2[3][3], f[3], b1[3], b2[3], m[3][3], s[3][3], us[3][3];
	load_q<T>(ta + (long) k * 8 * n, n, i, q);
	mm<T>(q, q, q2);
	ch_coeffs<T, true>(q, q2, f, b1, b2);
	load_su3<T>(u + (long) k * 9 * n, n, i, m);
#pragma unroll
	for (int e = 0; e < 9; e++) s[e / 3][e % 3] = sp[((long) k * 9 + e) * n + i];
	mm<T>(m, s, us);
	// tr_w = Tr(B_w U S') = sum_rc B_w[r][c] us[c][r]
	C tr1 = mkq<T>(0, 0), tr2 = mkq<T>(0, 0);
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C B1 = cadd(cm(b1[1], q[r][c]), cm(b1[2], q2[r][c])), B2 = cadd(cm(b2[1], q[r][c]), cm(b2[2], q2[r][c]));
			if (r == c) { B1 = cadd(B1, b1[0]); B2 = cadd(B2, b2[0]); }
			tr1 = cadd(tr1, cm(B1, us[c][r])); tr2 = cadd(tr2, cm(B2, us[c][r]));
		}
	C qus[3][3], usq[3][3], gm[3][3];
	mm<T>(q, us, qus); mm<T>(us, q, usq);
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			gm[r][c] = cadd(cadd(cm(tr1, q[r][c]), cm(tr2, q2[r][c])), cadd(cm(f[1], us[r][c]), cm(f[2], cadd(qus[r][c], usq[r][c]))));
			tmp[((long) k * 9 + r * 3 + c) * n + i] = us[r][c];
		}
	T *lk = lam + (long) k * 8 * n;
	const T third = (T) 0.33333333333333333333333, half = (T) 0.5;
	lk[6 * n + i] = (2 * gm[0][0].x - gm[1][1].x - gm[2][2].x) * third;
	lk[7 * n + i] = (2 * gm[1][1].x - gm[0][0].x - gm[2][2].x) * third;
	((C *) lk)[i] = mkq<T>((gm[0][1].x + gm[1][0].x) * half, (gm[0][1].y - gm[1][0].y) * half);
	((C *) (lk + 2 * n))[i] = mkq<T>((gm[0][2].x + gm[2][0].x) * half, (gm[0][2].y - gm[2][0].y) * half);
	((C *) (lk + 4 * n))[i] = mkq<T>((gm[1][2].x + gm[2][1].x) * half, (gm[1][2].y - gm[2][1].y) * half);
}

__device__ __forceinline__ unsigned int sf_snum(const SfGeom &g, int d0, int d1, int d2, int d3)
{
	return (unsigned int) (d0 + g.nd0 * (d1 + g.nd1 * (d2 + g.nd2 * d3))) >> 1;
}
__device__ __forceinline__ int sf_wrap(int x, int n) { return x < 0 ? x + n : (x >= n ? x - n : x); }

template <typename T>
__device__ __forceinline__ void acc_irho(cplx_t<T> res[3][3], const cplx_t<T> t[3][3], T rho)   // res += i*rho*t
{
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) { res[r][c].x -= rho * t[r][c].y; res[r][c].y += rho * t[r][c].x; }
}

// res += i*rho * op(a) * op(b), accumulated entry by entry (no temporary matrix)
template <typename T, bool DA = false, bool DB = false>
__device__ __forceinline__ void mm_acc_irho(cplx_t<T> res[3][3], const cplx_t<T> a[3][3], const cplx_t<T> b[3][3], T rho)
{
	using C = cplx_t<T>;
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = mkq<T>(0, 0);
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const C x = DA ? cj(a[j][r]) : a[r][j];
				const C y = DB ? cj(b[c][j]) : b[j][c];
				acc.x += x.x * y.x - x.y * y.y; acc.y += x.x * y.y + x.y * y.x;
			}
			res[r][c].x -= rho * acc.y; res[r][c].y += rho * acc.x;
		}
}
// o -= op(a) * op(b)
template <typename T, bool DA = false, bool DB = false>
__device__ __forceinline__ void mm_sub(cplx_t<T> o[3][3], const cplx_t<T> a[3][3], const cplx_t<T> b[3][3])
{
	using C = cplx_t<T>;
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = o[r][c];
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const C x = DA ? cj(a[j][r]) : a[r][j];
				const C y = DB ? cj(b[c][j]) : b[j][c];
				acc.x -= x.x * y.x - x.y * y.y; acc.y -= x.x * y.y + x.y * y.x;
			}
			o[r][c] = acc;
		}
}
// o += op(a) * op(b)
template <typename T, bool DA = false, bool DB = false>
__device__ __forceinline__ void mm_add(cplx_t<T> o[3][3], const cplx_t<T> a[3][3], const cplx_t<T> b[3][3])
{
	using C = cplx_t<T>;
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = o[r][c];
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const C x = DA ? cj(a[j][r]) : a[r][j];
				const C y = DB ? cj(b[c][j]) : b[j][c];
				acc.x += x.x * y.x - x.y * y.y; acc.y += x.x * y.y + x.y * y.x;
			}
			o[r][c] = acc;
		}
}

// compute_sigma (stouting.c:1175-1305): Sigma = Sigma' exp(iQ) + i rho sum_{nu != mu} [ staples with one Lambda inserted ]
//   right, A = U_nu(x+mu), B = U_mu(x+nu)^+, C = U_nu(x)^+ :  ABC (L_mu(x) - L_nu(x)) + L_nu(x+mu) ABC - A B L_mu(x+nu) C
//   left,  A = U_nu(x+mu-nu)^+, B = U_mu(x-nu)^+, C = U_nu(x-nu) :
//                           A B (L_nu(x-nu) - L_mu(x-nu)) C + A B C L_mu(x) - A L_nu(x+mu-nu) B C
// TMP is left = exp(iQ) (third row rebuilt) as in the reference.
// tuning knobs (scripts/bench_sigma.py): one thread per LINK (grid.y = 8) instead of per index looping over its eight links;
// minimum CTAs per SM (register bound)
#ifndef STAPLE_SIGMA_PER_LINK
#define STAPLE_SIGMA_PER_LINK 1
#endif
#ifndef STAPLE_SIGMA_MINBLOCKS
#define STAPLE_SIGMA_MINBLOCKS 1
#endif
template <typename T>
__global__ void __launch_bounds__(kSfBlock, STAPLE_SIGMA_MINBLOCKS) stout_sigma_kernel(const T *lam, const cplx_t<T> *u, cplx_t<T> *sg, const T *ta, cplx_t<T> *tmp, T rho, SfGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kSfBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int idx = g.lo + t;
	const long n = g.sizeh;
	const int nd0h = g.nd0 >> 1;
	const int hd0 = idx % nd0h;
	unsigned int qq = idx / nd0h;
	const int d1 = qq % g.nd1; qq /= g.nd1;
	const int d2 = qq % g.nd2;
	const int d3 = qq / g.nd2;
	const int nd[4] = { g.nd0, g.nd1, g.nd2, g.nd3 };
#if STAPLE_SIGMA_PER_LINK
	const int k0 = blockIdx.y, k1 = blockIdx.y + 1;
#else
	const int k0 = 0, k1 = 8;
#endif
#pragma unroll 1
	for (int k = k0; k < k1; k++) {
		const int mu = k >> 1, p = k & 1;
		const int x[4] = { 2 * hd0 + ((d1 + d2 + d3 + p) & 1), d1, d2, d3 };
		auto site = [&](int dmu, int nu, int dnu) {
			int y[4] = { x[0], x[1], x[2], x[3] };
			y[mu] = sf_wrap(y[mu] + dmu, nd[mu]);
			y[nu] = sf_wrap(y[nu] + dnu, nd[nu]);
			return sf_snum(g, y[0], y[1], y[2], y[3]);
		};
		C res[3][3];
		{	// PIECE1 = Sigma' exp(iQ)   (:550-668)
			C q[3][3], q2[3][3], f[3], e[3][3], s[3][3];
			load_q<T>(ta + (long) k * 8 * n, n, idx, q);
			mm<T>(q, q, q2);
			ch_coeffs<T, false>(q, q2, f, nullptr, nullptr);
#pragma unroll
			for (int r = 0; r < 2; r++)
#pragma unroll
				for (int c = 0; c < 3; c++) {
					e[r][c] = cadd(cm(f[1], q[r][c]), cm(f[2], q2[r][c]));
					if (r == c) e[r][c] = cadd(e[r][c], f[0]);
				}
			e[2][0] = cj(csub(cm(e[0][1], e[1][2]), cm(e[0][2], e[1][1])));
			e[2][1] = cj(csub(cm(e[0][2], e[1][0]), cm(e[0][0], e[1][2])));
			e[2][2] = cj(csub(cm(e[0][0], e[1][1]), cm(e[0][1], e[1][0])));
#pragma unroll
			for (int w = 0; w < 9; w++) { s[w / 3][w % 3] = sg[((long) k * 9 + w) * n + idx]; tmp[((long) k * 9 + w) * n + idx] = e[w / 3][w % 3]; }
			mm<T>(s, e, res);
		}
		C lmu[3][3];
		load_herm<T>(lam + (long) k * 8 * n, n, idx, lmu);
#pragma unroll 1
		for (int it = 0; it < 3; it++) {
			const int nu = it + (it >= mu ? 1 : 0);
			const unsigned int ipmu = site(1, nu, 0), ipnu = site(0, nu, 1), imnu = site(0, nu, -1), ipmumnu = site(1, nu, -1);
			C a[3][3], b[3][3], c[3][3], pm[3][3], l1[3][3], xm[3][3];
			// ---- right: P = A B^+ ;  i rho [ P (C^+ (D - E) - G C^+) + F (P C^+) ]      (12 products per plane instead of 15)
			load_su3<T>(u + (long) (2 * nu + !p) * 9 * n, n, ipmu, a);       // A = U_nu(x+mu)
			load_su3<T>(u + (long) (2 * mu + !p) * 9 * n, n, ipnu, b);       // B = U_mu(x+nu)
			mm<T, false, true>(a, b, pm);                                     // P = A B^+
			load_su3<T>(u + (long) (2 * nu + p) * 9 * n, n, idx, c);         // C = U_nu(x)
			load_herm<T>(lam + (long) (2 * nu + p) * 8 * n, n, idx, l1);     // E = L_nu(x)
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int cc = 0; cc < 3; cc++) l1[r][cc] = csub(lmu[r][cc], l1[r][cc]);
			mm<T, true, false>(c, l1, xm);                                    // X = C^+ (D - E)
			load_herm<T>(lam + (long) (2 * mu + !p) * 8 * n, n, ipnu, l1);   // G = L_mu(x+nu)
			mm_sub<T, false, true>(xm, l1, c);                                // X -= G C^+
			mm_acc_irho<T>(res, pm, xm, rho);                                 // + i rho P X
			mm<T, false, true>(pm, c, xm);                                    // T = P C^+
			load_herm<T>(lam + (long) (2 * nu + !p) * 8 * n, n, ipmu, l1);   // F = L_nu(x+mu)
			mm_acc_irho<T>(res, l1, xm, rho);                                 // + i rho F T
			// ---- left: i rho A^+ [ B^+ ((G - E) C + C D) - F (B^+ C) ]
			load_su3<T>(u + (long) (2 * nu + !p) * 9 * n, n, imnu, c);       // C = U_nu(x-nu)
			load_herm<T>(lam + (long) (2 * nu + !p) * 8 * n, n, imnu, l1);   // G = L_nu(x-nu)
			load_herm<T>(lam + (long) (2 * mu + !p) * 8 * n, n, imnu, a);    // E = L_mu(x-nu)
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int cc = 0; cc < 3; cc++) l1[r][cc] = csub(l1[r][cc], a[r][cc]);
			mm<T>(l1, c, xm);                                                 // Y = (G - E) C
			mm_add<T>(xm, c, lmu);                                            // Y += C D
			load_su3<T>(u + (long) (2 * mu + !p) * 9 * n, n, imnu, b);       // B = U_mu(x-nu)
			mm<T, true, false>(b, xm, pm);                                    // Z = B^+ Y
			mm<T, true, false>(b, c, xm);                                     // W = B^+ C
			load_herm<T>(lam + (long) (2 * nu + p) * 8 * n, n, ipmumnu, l1); // F = L_nu(x+mu-nu)
			mm_sub<T>(pm, l1, xm);                                            // Z -= F W
			load_su3<T>(u + (long) (2 * nu + p) * 9 * n, n, ipmumnu, a);     // A = U_nu(x+mu-nu)
			mm_acc_irho<T, true, false>(res, a, pm, rho);                     // + i rho A^+ Z
		}
#pragma unroll
		for (int w = 0; w < 9; w++) sg[((long) k * 9 + w) * n + idx] = res[w / 3][w % 3];
	}
}

}   // namespace staple

using namespace staple;

#define DD(p) ((double2 *) dev(p, #p))
#define DF(p) ((float2 *) dev(p, #p))
#define CDD(p) ((const double2 *) dev(p, #p))
#define CDF(p) ((const float2 *) dev(p, #p))

extern "C" {

extern double gl_stout_rho, gl_topo_rho;

#define STAPLE_SF_DEF(S, T, C2, D, CD, SU3, TAMAT, THMAT)                                                                    \
	void compute_lambda##S(THMAT *L, const SU3 *SP, const SU3 *U, const TAMAT *QA, SU3 *TMP)                                   \
	{                                                                                                                          \
		require_init("compute_lambda");                                                                                          \
		const SfGeom g = sf_geom();                                                                                              \
		stout_lambda_kernel<T><<<dim3((g.cnt + kSfBlock - 1) / kSfBlock, 8), kSfBlock, 0, ctx().stream>>>(                       \
			(T *) dev(L, "L"), CD(SP), CD(U), (const T *) dev(QA, "QA"), D(TMP), g);                                               \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                   \
	}                                                                                                                          \
	void compute_sigma##S(const THMAT *L, const SU3 *U, SU3 *Sg, const TAMAT *QA, SU3 *TMP, const int istopo)                  \
	{                                                                                                                          \
		require_init("compute_sigma");                                                                                           \
		const SfGeom g = sf_geom();                                                                                              \
		stout_sigma_kernel<T><<<dim3((g.cnt + kSfBlock - 1) / kSfBlock, STAPLE_SIGMA_PER_LINK ? 8 : 1), kSfBlock, 0, ctx().stream>>>( \
			(const T *) dev(L, "L"), CD(U), D(Sg), (const T *) dev(QA, "QA"), D(TMP), (T) (istopo ? gl_topo_rho : gl_stout_rho), g); \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                   \
	}                                                                                                                          \
	/* communications.c gl3 / tamat / thmat borders, thickness 1: contiguous d3 slabs of every component array */            \
	void communicate_gl3_borders##S(SU3 *lnh_conf, int thickness)                                                              \
	{                                                                                                                          \
		require_init("communicate_gl3_borders");                                                                                 \
		exchange_slices(dev(lnh_conf, "lnh_conf"), sizeof(C2), ctx().g.sizeh, 72, thickness, ctx().stream);                      \
	}                                                                                                                          \
	static void packed5_borders##S(void *base, int thickness)                                                                  \
	{	/* per link: three complex arrays then two real arrays = 8*sizeh reals */                                              \
		const long n = ctx().g.sizeh;                                                                                            \
		for (int k = 0; k < 8; k++) {                                                                                            \
			char *p = (char *) base + (size_t) k * 8 * n * sizeof(T);                                                              \
			exchange_slices(p, sizeof(C2), n, 3, thickness, ctx().stream);                                                         \
			exchange_slices(p + (size_t) 6 * n * sizeof(T), sizeof(T), n, 2, thickness, ctx().stream);                             \
		}                                                                                                                        \
	}                                                                                                                          \
	void communicate_tamat_soa_borders##S(TAMAT *lnh_ipdot, int thickness)                                                     \
	{ require_init("communicate_tamat_soa_borders"); packed5_borders##S(dev(lnh_ipdot, "lnh_ipdot"), thickness); }            \
	void communicate_thmat_soa_borders##S(THMAT *lnh_ipdot, int thickness)                                                     \
	{ require_init("communicate_thmat_soa_borders"); packed5_borders##S(dev(lnh_ipdot, "lnh_ipdot"), thickness); }            \
	/* fermion_force.c:52-163: staples of U, Q = rho TA(U staples), Lambda from Sigma', then Sigma; with NRANKS_D3 > 1  */    \
	/* the reference exchanges the borders of the staples, Q, Lambda and Sigma (thickness 1) at the same points.        */    \
	void compute_sigma_from_sigma_prime_backinto_sigma_prime##S(SU3 *Sigma, THMAT *Lambda, TAMAT *QA, const SU3 *U, SU3 *TMP, \
																															 const int istopo)                                           \
	{                                                                                                                          \
		require_init("compute_sigma_from_sigma_prime_backinto_sigma_prime");                                                     \
		if (verbosity_lv > 2) printf("MPI%02d:\t\tSIGMA_PRIME --> SIGMA\n", ctx().myrank);                                       \
		set_su3_soa_to_zero##S(TMP);                                                                                             \
		calc_loc_staples_nnptrick_all_onlyferms##S(U, TMP);                                                                      \
		if (ctx().nranks > 1) communicate_gl3_borders##S(TMP, 1);                                                                \
		RHO_times_conf_times_staples_ta_part##S(U, TMP, QA, istopo);                                                             \
		if (ctx().nranks > 1) communicate_tamat_soa_borders##S(QA, 1);                                                           \
		compute_lambda##S(Lambda, Sigma, U, QA, TMP);                                                                            \
		if (ctx().nranks > 1) communicate_thmat_soa_borders##S(Lambda, 1);                                                       \
		compute_sigma##S(Lambda, U, Sigma, QA, TMP, istopo);                                                                     \
		if (ctx().nranks > 1) communicate_gl3_borders##S(Sigma, 1);                                                              \
	}
STAPLE_SF_DEF(, double, double2, DD, CDD, su3_soa, tamat_soa, thmat_soa)
STAPLE_SF_DEF(_f, float, float2, DF, CDF, su3_soa_f, tamat_soa_f, thmat_soa_f)

}   // extern "C"

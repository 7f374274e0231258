// Fermion-force outer products, the step right after every MD multishift solve:
//   OpenAcc/fermion_force_utilities.c:17-201 + fermion_force_utilities.h:16-216
//     set_tamat_soa_to_zero, direct_product_of_fermions_into_auxmat, multiply_conf_times_force_and_take_ta_nophase,
//     multiply_backfield_times_force, accumulate_gl3soa_into_gl3soa, ker_openacc_compute_fermion_force
//   OpenAcc/su3_utilities.c set_su3_soa_to_zero (the companion initialiser, fermion_force.c:211)
// gl(3) fields (aux_u, pseudo_ipdot) use the su3_soa[8] layout with all three rows meaningful; tamat_soa[8] is
// {c01, c02, c12 complex[sizeh]; ic00, ic11 real[sizeh]} (struct_c_def.h:45-51).
// All kernels are HBM-bound streaming kernels: one thread per half-lattice index of the local interior,
// every SoA stream contiguous across the warp.
#include "staple_internal.cuh"

namespace staple {

constexpr int kForceBlock = 128;

template <typename T> __device__ __forceinline__ cplx_t<T> mkf(T x, T y);
template <> __device__ __forceinline__ double2 mkf<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mkf<float>(float x, float y) { return make_float2(x, y); }

struct ForceGeom { unsigned int nd0h, nd1, nd2, nd3, vol3h; long sizeh; unsigned int lo, cnt; };

static ForceGeom force_geom()
{
	const Geom &g = ctx().g;
	ForceGeom f;
	f.nd0h = g.nd0h; f.nd1 = g.nd1; f.nd2 = g.nd2; f.nd3 = g.nd3; f.vol3h = (unsigned int) g.vol3h; f.sizeh = g.sizeh;
	f.lo = (unsigned int) (g.d3_halo * g.vol3h); f.cnt = (unsigned int) (g.loc_n3 * g.vol3h);   // d3 in [D3_HALO, nd3-D3_HALO)
	return f;
}

// aux(idx) += l (x) r   (fermion_force_utilities.h:16-42, r = factor*conj(fer_r) prepared by the caller)
template <typename T>
__device__ __forceinline__ void outer_acc(cplx_t<T> m[9], const cplx_t<T> l[3], const cplx_t<T> r[3])
{
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			m[a * 3 + c].x += l[a].x * r[c].x - l[a].y * r[c].y;
			m[a * 3 + c].y += l[a].x * r[c].y + l[a].y * r[c].x;
		}
}

struct ForcePair { double a[2]; };

// direct_product_of_fermions_into_auxmat (fermion_force_utilities.c:31-95) for B shifts at once: the even site
// with this idxh owns aux[2mu](idxh) += a h(x+mu) (x) conj(s(x)), the odd one aux[2mu+1](idxh) += -a s(x+mu) (x) conj(h(x)).
// Every link matrix is read and written exactly once per launch whatever B is, in the reference's accumulation order.
template <typename T, int B>
__global__ void __launch_bounds__(kForceBlock) force_outer_kernel(cplx_t<T> *aux, const cplx_t<T> *s0, const cplx_t<T> *h0,
																																	 const cplx_t<T> *s1, const cplx_t<T> *h1, ForcePair f, ForceGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kForceBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const unsigned int idx = g.lo + t;
	const long n = g.sizeh;
	const unsigned int hd0 = idx % g.nd0h;
	unsigned int q = idx / g.nd0h;
	const unsigned int d1 = q % g.nd1; q /= g.nd1;
	const unsigned int d2 = q % g.nd2;
	const unsigned int d3 = q / g.nd2;
	const unsigned int s1s = g.nd0h, s2s = g.nd0h * g.nd1, s3s = g.vol3h;
	const unsigned int ip1 = d1 == g.nd1 - 1 ? idx - s1s * (g.nd1 - 1) : idx + s1s;
	const unsigned int ip2 = d2 == g.nd2 - 1 ? idx - s2s * (g.nd2 - 1) : idx + s2s;
	const unsigned int ip3 = d3 == g.nd3 - 1 ? idx - s3s * (g.nd3 - 1) : idx + s3s;
	const C *sv[2] = { s0, s1 }, *hv[2] = { h0, h1 };
#pragma unroll
	for (int par = 0; par < 2; par++) {
		// d0 = 2*hd0 + rp: the +0 neighbour of a site in an rp = 1 row has the next idxh, in an rp = 0 row the same one
		const unsigned int rp = (d1 + d2 + d3 + par) & 1u;
		const unsigned int ip0 = rp ? (hd0 == g.nd0h - 1 ? idx - (g.nd0h - 1) : idx + 1) : idx;
		const unsigned int ipm[4] = { ip0, ip1, ip2, ip3 };
		C r[B][3];
#pragma unroll
		for (int b = 0; b < B; b++) {
			const C *fr = par == 0 ? sv[b] : hv[b];
			const T fac = (T) (par == 0 ? f.a[b] : -f.a[b]);
#pragma unroll
			for (int c = 0; c < 3; c++) { const C v = __ldg(fr + c * n + idx); r[b][c] = mkf<T>(fac * v.x, fac * -v.y); }
		}
#pragma unroll
		for (int mu = 0; mu < 4; mu++) {
			C *ak = aux + (long) (2 * mu + par) * 9 * n + idx;
			C m[9];
#pragma unroll
			for (int e = 0; e < 9; e++) m[e] = ak[e * n];
#pragma unroll
			for (int b = 0; b < B; b++) {
				const C *fl = par == 0 ? hv[b] : sv[b];
				C l[3];
#pragma unroll
				for (int c = 0; c < 3; c++) l[c] = __ldg(fl + c * n + ipm[mu]);
				outer_acc<T>(m, l, r[b]);
			}
#pragma unroll
			for (int e = 0; e < 9; e++) ak[e * n] = m[e];
		}
	}
}

// multiply_conf_times_force_and_take_ta_nophase (fermion_force_utilities.c:97-121, .h:108-152):
// ipdot[k] -= TA(U[k] * aux[k]) over the local interior, third row of U rebuilt as conj(r0 x r1).
template <typename T>
__global__ void __launch_bounds__(kForceBlock) force_ta_kernel(const cplx_t<T> *u, const cplx_t<T> *aux, T *ta, ForceGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kForceBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const long n = g.sizeh, i = g.lo + t;
	const int k = blockIdx.y;
	const C *uk = u + (long) k * 9 * n + i, *ak = aux + (long) k * 9 * n + i;
	C m[3][3], x[3][3], p[3][3];
#pragma unroll
	for (int c = 0; c < 3; c++) { m[0][c] = __ldcs(uk + c * n); m[1][c] = __ldcs(uk + (3 + c) * n); }
#pragma unroll
	for (int e = 0; e < 9; e++) x[e / 3][e % 3] = __ldcs(ak + e * n);
	auto cm = [](C a, C b) { return mkf<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
	auto cj = [](C a) { return mkf<T>(a.x, -a.y); };
	auto sub = [](C a, C b) { return mkf<T>(a.x - b.x, a.y - b.y); };
	m[2][0] = cj(sub(cm(m[0][1], m[1][2]), cm(m[0][2], m[1][1])));
	m[2][1] = cj(sub(cm(m[0][2], m[1][0]), cm(m[0][0], m[1][2])));
	m[2][2] = cj(sub(cm(m[0][0], m[1][1]), cm(m[0][1], m[1][0])));
#pragma unroll
	for (int r = 0; r < 3; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) {
			C acc = cm(m[r][0], x[0][c]);
			const C b1 = cm(m[r][1], x[1][c]), b2 = cm(m[r][2], x[2][c]);
			acc.x += b1.x; acc.y += b1.y; acc.x += b2.x; acc.y += b2.y;
			p[r][c] = acc;
		}
	T *tk = ta + (long) k * 8 * n;
	C *c01 = (C *) tk + i, *c02 = (C *) (tk + 2 * n) + i, *c12 = (C *) (tk + 4 * n) + i;
	T *ic00 = tk + 6 * n + i, *ic11 = tk + 7 * n + i;
	const T half = (T) 0.5, third = (T) 0.33333333333333333333333;   // common_defines.h:65-66
	C v = *c01; v.x -= half * (p[0][1].x - p[1][0].x); v.y -= half * (p[0][1].y + p[1][0].y); *c01 = v;
	v = *c02; v.x -= half * (p[0][2].x - p[2][0].x); v.y -= half * (p[0][2].y + p[2][0].y); *c02 = v;
	v = *c12; v.x -= half * (p[1][2].x - p[2][1].x); v.y -= half * (p[1][2].y + p[2][1].y); *c12 = v;
	const T tr = p[0][0].y + p[1][1].y + p[2][2].y;
	*ic00 -= p[0][0].y - third * tr;
	*ic11 -= p[1][1].y - third * tr;
}

// multiply_backfield_times_force (fermion_force_utilities.c:123-153): pseudo += e^{i theta} aux   (PHASE = true)
// accumulate_gl3soa_into_gl3soa (:155-180):                          pseudo += aux               (PHASE = false)
template <typename T, bool PHASE>
__global__ void __launch_bounds__(kForceBlock) force_accum_kernel(const T *ph, const cplx_t<T> *aux, cplx_t<T> *pseudo, ForceGeom g)
{
	using C = cplx_t<T>;
	const unsigned int t = blockIdx.x * kForceBlock + threadIdx.x;
	if (t >= g.cnt) return;
	const long n = g.sizeh, i = g.lo + t;
	const int k = blockIdx.y;
	T s = 0, c = 1;
	if (PHASE) {
		const T arg = ph[(long) k * n + i];
		if (sizeof(T) == 8) { double ds, dc; sincos((double) arg, &ds, &dc); s = (T) ds; c = (T) dc; }
		else { float fs, fc; sincosf((float) arg, &fs, &fc); s = (T) fs; c = (T) fc; }
	}
	const C *ak = aux + (long) k * 9 * n + i;
	C *pk = pseudo + (long) k * 9 * n + i;
#pragma unroll
	for (int e = 0; e < 9; e++) {
		const C a = __ldcs(ak + e * n);
		C o = pk[e * n];
		if (PHASE) { o.x += a.x * c - a.y * s; o.y += a.x * s + a.y * c; }
		else { o.x += a.x; o.y += a.y; }
		pk[e * n] = o;
	}
}

template <typename T>
static void force_outer(cplx_t<T> *aux, const cplx_t<T> *s0, const cplx_t<T> *h0, double a0, const cplx_t<T> *s1,
												const cplx_t<T> *h1, double a1)
{
	const ForceGeom g = force_geom();
	const unsigned int grid = (g.cnt + kForceBlock - 1) / kForceBlock;
	ForcePair f; f.a[0] = a0; f.a[1] = a1;
	if (s1 == nullptr) force_outer_kernel<T, 1><<<grid, kForceBlock, 0, ctx().stream>>>(aux, s0, h0, s0, h0, f, g);
	else force_outer_kernel<T, 2><<<grid, kForceBlock, 0, ctx().stream>>>(aux, s0, h0, s1, h1, f, g);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}

// ker_openacc_compute_fermion_force (fermion_force_utilities.c:183-201).  The reference does, per shift,
// copy -> acc_Doe -> outer products (aux_u read and written once per shift).  Here the copy is dropped (the
// operator reads in_shiftmulti[iter] directly) and shifts go in PAIRS: Doe of the first into loc_s, Doe of the
// second into loc_h, one outer-product launch for both -- aux_u moves half as often, the additions keep the
// reference's order.  On return loc_s = in_shiftmulti[last] and loc_h = Doe(loc_s), as after the reference's loop.
template <typename T>
static void compute_fermion_force(const cplx_t<T> *u, cplx_t<T> *aux, const cplx_t<T> *shiftmulti, cplx_t<T> *loc_s,
																	cplx_t<T> *loc_h, const T *ph, const RationalApprox *approx)
{
	const long vs = 3 * ctx().g.sizeh;
	const int order = approx->approx_order;
	int iter = 0;
	for (; iter + 2 <= order; iter += 2) {
		const cplx_t<T> *sa = shiftmulti + (long) iter * vs, *sb = sa + vs;
		apply_dslash<T>(1, EPI_NONE, u, loc_s, sa, ph, nullptr, 0.0, -1, nullptr);
		apply_dslash<T>(1, EPI_NONE, u, loc_h, sb, ph, nullptr, 0.0, -1, nullptr);
		force_outer<T>(aux, sa, loc_s, approx->RA_a[iter], sb, loc_h, approx->RA_a[iter + 1]);
	}
	if (iter < order) {
		const cplx_t<T> *sa = shiftmulti + (long) iter * vs;
		apply_dslash<T>(1, EPI_NONE, u, loc_h, sa, ph, nullptr, 0.0, -1, nullptr);
		force_outer<T>(aux, sa, loc_h, approx->RA_a[iter], nullptr, nullptr, 0.0);
	}
	if (order > 0) blas<T>(OP_ASSIGN, loc_s, shiftmulti + (long) (order - 1) * vs, nullptr, nullptr, 0.0);
}

}   // namespace staple

using namespace staple;

#define DD(p) ((double2 *) dev(p, #p))
#define DF(p) ((float2 *) dev(p, #p))
#define CDD(p) ((const double2 *) dev(p, #p))
#define CDF(p) ((const float2 *) dev(p, #p))

extern "C" {

#define STAPLE_FORCE_DEF(S, T, C2, D, CD, SU3, VEC3, TAMAT, PHASES)                                                       \
	void set_tamat_soa_to_zero##S(TAMAT *matrix)                                                                            \
	{                                                                                                                       \
		require_init("set_tamat_soa_to_zero");                                                                                \
		STAPLE_CUDA_CHECK(cudaMemsetAsync(dev(matrix, "matrix"), 0, sizeof(T) * 8 * 8 * ctx().g.sizeh, ctx().stream)); blocking_point(); \
	}                                                                                                                       \
	void set_su3_soa_to_zero##S(SU3 *matrix)                                                                                \
	{                                                                                                                       \
		require_init("set_su3_soa_to_zero");                                                                                  \
		STAPLE_CUDA_CHECK(cudaMemsetAsync(dev(matrix, "matrix"), 0, sizeof(C2) * 8 * 9 * ctx().g.sizeh, ctx().stream)); blocking_point(); \
	}                                                                                                                       \
	void direct_product_of_fermions_into_auxmat##S(const VEC3 *loc_s, const VEC3 *loc_h, SU3 *aux_u,                        \
																								 const RationalApprox *approx, int iter)                                  \
	{                                                                                                                       \
		require_init("direct_product_of_fermions_into_auxmat");                                                               \
		force_outer<T>(D(aux_u), CD(loc_s), CD(loc_h), approx->RA_a[iter], nullptr, nullptr, 0.0);                            \
	}                                                                                                                       \
	void multiply_conf_times_force_and_take_ta_nophase##S(const SU3 *u, const SU3 *auxmat, TAMAT *ipdot)                    \
	{                                                                                                                       \
		require_init("multiply_conf_times_force_and_take_ta_nophase");                                                        \
		const ForceGeom g = force_geom();                                                                                     \
		force_ta_kernel<T><<<dim3((g.cnt + kForceBlock - 1) / kForceBlock, 8), kForceBlock, 0, ctx().stream>>>(               \
			CD(u), CD(auxmat), (T *) dev(ipdot, "ipdot"), g);                                                                   \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                \
	}                                                                                                                       \
	void multiply_backfield_times_force##S(ferm_param *tpars, const SU3 *auxmat, SU3 *pseudo_ipdot)                         \
	{                                                                                                                       \
		require_init("multiply_backfield_times_force");                                                                       \
		const ForceGeom g = force_geom();                                                                                     \
		force_accum_kernel<T, true><<<dim3((g.cnt + kForceBlock - 1) / kForceBlock, 8), kForceBlock, 0, ctx().stream>>>(      \
			(const T *) dev(tpars->PHASES, "tpars->" #PHASES), CD(auxmat), D(pseudo_ipdot), g);                                 \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                \
	}                                                                                                                       \
	void accumulate_gl3soa_into_gl3soa##S(const SU3 *auxmat, SU3 *pseudo_ipdot)                                             \
	{                                                                                                                       \
		require_init("accumulate_gl3soa_into_gl3soa");                                                                        \
		const ForceGeom g = force_geom();                                                                                     \
		force_accum_kernel<T, false><<<dim3((g.cnt + kForceBlock - 1) / kForceBlock, 8), kForceBlock, 0, ctx().stream>>>(     \
			nullptr, CD(auxmat), D(pseudo_ipdot), g);                                                                           \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                                \
	}                                                                                                                       \
	void ker_openacc_compute_fermion_force##S(const SU3 *u, SU3 *aux_u, const VEC3 *in_shiftmulti, VEC3 *loc_s,             \
																						VEC3 *loc_h, ferm_param *tpars)                                               \
	{                                                                                                                       \
		require_init("ker_openacc_compute_fermion_force");                                                                    \
		compute_fermion_force<T>(CD(u), D(aux_u), CD(in_shiftmulti), D(loc_s), D(loc_h),                                      \
														 (const T *) dev(tpars->PHASES, "tpars->" #PHASES), &tpars->approx_md);                       \
	}
STAPLE_FORCE_DEF(, double, double2, DD, CDD, su3_soa, vec3_soa, tamat_soa, phases)
STAPLE_FORCE_DEF(_f, float, float2, DF, CDF, su3_soa_f, vec3_soa_f, tamat_soa_f, phases_f)

}   // extern "C"

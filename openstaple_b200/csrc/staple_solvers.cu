// Solvers of the staggered hot path:
//   multishift_invert[_f]        <- OpenAcc/inverter_multishift_full.c:23-252   (CG-M)
//   ker_invert_openacc[_f]       <- OpenAcc/inverter_full.c:19-132              (restarted CG)
//   inverter_mixed_precision     <- OpenAcc/inverter_mixedp.c:41-181
//   wrappers / package           <- OpenAcc/inverter_wrappers.c:24-159, inverter_package.c:18-72
//   power iteration              <- OpenAcc/find_min_max.c:21-117
//
// CG-M keeps ALL scalar recurrences (alpha, omega, zeta_i, gamma_i, convergence flags) in device
// memory: the reduction kernels leave their sums in device slots, one-warp kernels advance the
// recurrences, and the vector kernels read the coefficients from that control block.  The host only
// enqueues iterations and polls a `done` flag every few iterations, so there is no host round trip
// inside an iteration (the reference has >= 4 per iteration).  Kernels launched after convergence see
// `done` and return immediately; the iteration count is the device-side counter, so it is exact.
#include "staple_internal.cuh"
#include <cmath>
#include <cstring>

namespace staple {

constexpr int kBlasBlock = 256;
template <typename T> __device__ __forceinline__ cplx_t<T> mkc(double x, double y);
template <> __device__ __forceinline__ double2 mkc<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mkc<float>(double x, double y) { return make_float2((float) x, (float) y); }

// setup (:65-104): delta=(r,r), source_norm=(in,in) arrive in result slots
__global__ void cgm_init_kernel(CgmCtl *c, const double *delta_slot, const double *srcnorm_slot)
{
	if (threadIdx.x != 0) return;
	c->delta = *delta_slot; c->source_norm = *srcnorm_slot;
	c->omega = 1.0; c->gammag = 0.0; c->alpha = 0.0; c->lambda = 0.0; c->omega_save = 1.0;
	for (int i = 0; i < c->order; i++) {
		c->flag[i] = 1; c->zeta_i[i] = 1.0; c->zeta_ii[i] = 1.0; c->zeta_iii[i] = 1.0;
		c->gammas[i] = 0.0; c->omegas[i] = 0.0; c->pgam[i] = 1.0; c->pzeta[i] = 0.0;
	}
	c->maxiter = c->order; c->cg = 0; c->done = 0; c->pending = 0; c->active_sum = 0;
}

// standalone one-warp forms of the recurrences (NCCL all-reduce path, where the sum over ranks is a library call
// between the reduction kernel and its consumer)
__global__ void cgm_after_alpha_kernel(CgmCtl *c, double *alpha_slot, RedView red)
{
	if (c->done) return;
	cgm_after_alpha_warp(c, alpha_slot, red);
}
__global__ void cgm_after_lambda_kernel(CgmCtl *c, double *lambda_slot, RedView red)
{
	if (c->done) return;
	cgm_after_lambda_warp(c, lambda_slot, red);
}

__device__ __forceinline__ void block_sum1(double &v, double *sm)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__syncthreads();
	if (lane == 0) sm[warp] = v;
	__syncthreads();
	if (warp == 0) {
		double x = lane < nwarp ? sm[lane] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		v = x;
	}
}

// The ONE pass over the shifted vectors of an iteration (192 B per site and active shift in FP64):
//   ps_i  <- pgam_i ps_i + pzeta_i r     the update the reference does at the end of the PREVIOUS iteration
//                                        (:152-157, multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1)
//   out_i <- out_i - omega_i ps_i        (:139,     multiple_combine_in1_minus_in2x_factor_back_into_in1)
//   r     <- r + omega s ; lambda = |r|^2 over the reduction range   (:140-142)
// Element for element the same arithmetic in the same order as the reference (ps_i is rounded to T before
// it enters out_i, as it is when it goes through memory there); shifts that have converged are skipped:
// their ps_i can no longer reach an output.
// tuning knobs of the shifted-vector pass (scripts/tune_cgm.py sweeps them on the GPU)
#ifndef STAPLE_CGM_BLOCK
#define STAPLE_CGM_BLOCK 256
#endif
#ifndef STAPLE_CGM_MINBLOCKS
#define STAPLE_CGM_MINBLOCKS 1
#endif
#ifndef STAPLE_CGM_UNROLL
#define STAPLE_CGM_UNROLL 2      // shifts per trip: 6*UNROLL independent 16-byte loads in flight per thread
#endif
#ifndef STAPLE_CGM_STREAM
#define STAPLE_CGM_STREAM 0      // 1: out_i / ps_i with evict-first loads and stores (they are touched once per iteration)
#endif
constexpr int kCgmBlock = STAPLE_CGM_BLOCK;

template <typename V> __device__ __forceinline__ V cgm_ld(const V *p)
{
#if STAPLE_CGM_STREAM
	return __ldcs(p);
#else
	return *p;
#endif
}
template <typename V> __device__ __forceinline__ void cgm_st(V *p, V v)
{
#if STAPLE_CGM_STREAM
	__stcs(p, v);
#else
	*p = v;
#endif
}

// The vector updates of the solvers are "real scalar x complex + complex", i.e. component-wise on the re/im parts, so a kernel
// may treat an array as elements E of any number of components: double2 / float2 = one site per thread, float4 = TWO consecutive
// FP32 sites per thread (16-byte loads and stores: the FP32 pass otherwise has half the bytes in flight per thread and was
// measured at 0.82 of the roofline where the FP64 pass reaches 0.95+).
template <typename E> struct Elem;
template <> struct Elem<double2> { static constexpr int N = 2; using S = double; };
template <> struct Elem<float2> { static constexpr int N = 2; using S = float; };
template <> struct Elem<float4> { static constexpr int N = 4; using S = float; };
template <typename E> __device__ __forceinline__ typename Elem<E>::S &comp(E &v, int k) { return reinterpret_cast<typename Elem<E>::S *>(&v)[k]; }
template <typename E> __device__ __forceinline__ typename Elem<E>::S comp(const E &v, int k) { return reinterpret_cast<const typename Elem<E>::S *>(&v)[k]; }

// U shifts of one element: all loads first, then the arithmetic of inverter_multishift_full.c:139,152-157
template <typename E, int U>
__device__ __forceinline__ void cgm_shift_group(E *out, E *ps, const E rv[3], long i, long n,
																								const int *s_ix, const double *s_om, const double *s_g, const double *s_z,
																								int k, int pending)
{
	using S = typename Elem<E>::S;
	E q[U][3], o[U][3];
#pragma unroll
	for (int u2 = 0; u2 < U; u2++) {
		const long base = (long) s_ix[k + u2] * 3 * n + i;
#pragma unroll
		for (int col = 0; col < 3; col++) { q[u2][col] = cgm_ld(ps + base + col * n); o[u2][col] = cgm_ld(out + base + col * n); }
	}
#pragma unroll
	for (int u2 = 0; u2 < U; u2++) {
		const long base = (long) s_ix[k + u2] * 3 * n + i;
		const double f = s_om[k + u2], g = s_g[k + u2], z = s_z[k + u2];
#pragma unroll
		for (int col = 0; col < 3; col++) {
			E qq = q[u2][col], oo = o[u2][col];
			if (pending) {
#pragma unroll
				for (int e = 0; e < Elem<E>::N; e++) comp(qq, e) = (S) (g * comp(qq, e) + z * comp(rv[col], e));
				cgm_st(ps + base + col * n, qq);
			}
#pragma unroll
			for (int e = 0; e < Elem<E>::N; e++) comp(oo, e) = (S) (comp(oo, e) - f * comp(qq, e));
			cgm_st(out + base + col * n, oo);
		}
	}
}

// all index arguments (lo, cnt, n, r0_lo, r0_hi) in units of E
template <typename E>
__global__ void __launch_bounds__(kCgmBlock, Elem<E>::N == 4 ? 2 : STAPLE_CGM_MINBLOCKS) cgm_fused_kernel(CgmCtl *c, E *out, E *ps, E *r,
																																const E *s, long lo, long cnt, long n, long r0_lo,
																																long r0_hi, double *partials, unsigned int *ticket,
																																double *result, int fuse_tail, RedView red)
{
	if (c->done) return;
	__shared__ double sm[32];
	__shared__ bool last;
	__shared__ double s_om[MAX_APPROX_ORDER], s_g[MAX_APPROX_ORDER], s_z[MAX_APPROX_ORDER];
	__shared__ int s_ix[MAX_APPROX_ORDER], s_nact;
	const int maxiter = c->maxiter;
	if (threadIdx.x == 0) {       // compact list of the still-active shifts: no divergence, unrollable loop
		int k = 0;
		for (int ia = 0; ia < maxiter; ia++)
			if (c->flag[ia] == 1) { s_ix[k] = ia; s_om[k] = c->omegas[ia]; s_g[k] = c->pgam[ia]; s_z[k] = c->pzeta[ia]; k++; }
		s_nact = k;
	}
	const double omega = c->omega;
	const int pending = c->pending;
	__syncthreads();
	const int nact = s_nact;
	const long t = (long) blockIdx.x * kCgmBlock + threadIdx.x;
	double nrm = 0.0;
	if (t < cnt) {
		using S = typename Elem<E>::S;
		const long i = lo + t;
		E rv[3];
#pragma unroll
		for (int col = 0; col < 3; col++) {
			const long j = col * n + i;
			rv[col] = r[j];
			const E sv = s[j];
			E rn;
#pragma unroll
			for (int e = 0; e < Elem<E>::N; e++) comp(rn, e) = (S) (comp(rv[col], e) + omega * comp(sv, e));
			r[j] = rn;
			if (i >= r0_lo && i < r0_hi) {
#pragma unroll
				for (int e = 0; e < Elem<E>::N; e++) nrm += (double) comp(rn, e) * comp(rn, e);
			}
		}
		int k = 0;
		for (; k + STAPLE_CGM_UNROLL <= nact; k += STAPLE_CGM_UNROLL)
			cgm_shift_group<E, STAPLE_CGM_UNROLL>(out, ps, rv, i, n, s_ix, s_om, s_g, s_z, k, pending);
		for (; k < nact; k++) cgm_shift_group<E, 1>(out, ps, rv, i, n, s_ix, s_om, s_g, s_z, k, pending);
	}
	block_sum1(nrm, sm);
	if (threadIdx.x == 0) {
		partials[blockIdx.x] = nrm;
		__threadfence();
		last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		double acc = 0.0;
		for (unsigned int k0 = threadIdx.x; k0 < gridDim.x; k0 += 8 * blockDim.x) {
			double x[8];
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const unsigned int k = k0 + j * blockDim.x;
				x[j] = k < gridDim.x ? __ldcg(partials + k) : 0.0;
			}
#pragma unroll
			for (int j = 0; j < 8; j++) acc += x[j];
		}
		block_sum1(acc, sm);
		if (threadIdx.x == 0) { result[0] = acc; *ticket = 0u; }
		// every other block has finished (and read the control block at its start): lambda is final, advance the
		// recurrences here instead of in a one-warp kernel of their own
		if (fuse_tail) {
			__syncthreads();
			if (threadIdx.x < 32) cgm_after_lambda_warp(c, result, red);
		}
	}
}

// p = r + gammag p  (:147-148, combine_in1xfactor_plus_in2(loc_p, gammag, loc_r, loc_p)) with gammag on the device.
// D3 slabs with staged halos (pv.on): the new p of the two interior face slices is ALSO stored into the neighbours' staging
// slots -- the exchange that the next iteration's Doe consumes; the solver's vector updates then run over the interior only
// (the reference keeps halos consistent by updating them redundantly, fermionic_utilities.c:188: at LOC_N3 = 2 that doubles
// the traffic of every update).
template <typename E> __device__ __forceinline__ E stageable_e(E v);
template <> __device__ __forceinline__ double2 stageable_e<double2>(double2 v) { return stageable(v); }
template <> __device__ __forceinline__ float2 stageable_e<float2>(float2 v) { return stageable(v); }
template <> __device__ __forceinline__ float4 stageable_e<float4>(float4 v)
{
	const float2 a = stageable(make_float2(v.x, v.y)), b = stageable(make_float2(v.z, v.w));
	return make_float4(a.x, a.y, b.x, b.y);
}
// index arguments (lo, cnt, n and the PushView's top_lo / bot_lo / vol3h) in units of E
template <typename E>
__global__ void __launch_bounds__(kBlasBlock) cgm_pupdate_kernel(const CgmCtl *c, E *p, const E *r, long lo,
																																	long cnt, long n, PushView pv)
{
	using S = typename Elem<E>::S;
	if (c->done) return;
	const double gammag = c->gammag;
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	const long i = lo + t;
	E *peer = nullptr;
	long pi = 0;
	if (pv.on) {
		const bool top = i >= pv.top_lo && i < pv.top_lo + pv.vol3h, bot = i >= pv.bot_lo && i < pv.bot_lo + pv.vol3h;
		if (top || bot) {
			peer = (E *) ((top ? pv.peer_top : pv.peer_bot) + (pv.seq & 1ull) * pv.parity_bytes);
			pi = i - (top ? pv.top_lo : pv.bot_lo);
		}
	}
#pragma unroll
	for (int col = 0; col < 3; col++) {
		const long j = col * n + i;
		const E pv_ = p[j], rv = r[j];
		E pn;
#pragma unroll
		for (int e = 0; e < Elem<E>::N; e++) comp(pn, e) = (S) (comp(pv_, e) * gammag + comp(rv, e));
		p[j] = pn;
		if (peer != nullptr) peer[col * pv.vol3h + pi] = stageable_e<E>(pn);
	}
}
// ps_i = in for all i (the order assign_in_to_out calls of :89-95 in one pass)
template <typename T>
__global__ void __launch_bounds__(kBlasBlock) broadcast_kernel(cplx_t<T> *dst, const cplx_t<T> *src, int order, long lo,
																															 long cnt, long n)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	for (int col = 0; col < 3; col++) {
		const cplx_t<T> v = src[col * n + lo + t];
		for (int ia = 0; ia < order; ia++) dst[(long) ia * 3 * n + col * n + lo + t] = v;
	}
}

static CgmCtl *g_d_ctl = nullptr;
static CgmCtl *g_h_ctl = nullptr;   // pinned, 2 snapshots
static cudaEvent_t g_ev_snap[2] = { nullptr, nullptr }, g_ev_t0 = nullptr, g_ev_t1 = nullptr;

static void ensure_ctl()
{
	if (g_d_ctl) return;
	STAPLE_CUDA_CHECK(cudaMalloc(&g_d_ctl, sizeof(CgmCtl)));
	STAPLE_CUDA_CHECK(cudaHostAlloc(&g_h_ctl, 2 * sizeof(CgmCtl), cudaHostAllocDefault));
	for (int i = 0; i < 2; i++) STAPLE_CUDA_CHECK(cudaEventCreateWithFlags(&g_ev_snap[i], cudaEventDisableTiming));
	STAPLE_CUDA_CHECK(cudaEventCreate(&g_ev_t0));
	STAPLE_CUDA_CHECK(cudaEventCreate(&g_ev_t1));
}

void set_spin_timeout_solvers(unsigned long long ns)
{
	STAPLE_CUDA_CHECK(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof(ns)));
}

void release_cg_state();
void release_solver_state()        // staple_shutdown
{
	release_cg_state();
	if (!g_d_ctl) return;
	cudaFree(g_d_ctl); cudaFreeHost(g_h_ctl);
	for (int i = 0; i < 2; i++) cudaEventDestroy(g_ev_snap[i]);
	cudaEventDestroy(g_ev_t0); cudaEventDestroy(g_ev_t1);
	g_d_ctl = nullptr; g_h_ctl = nullptr; g_ev_snap[0] = g_ev_snap[1] = g_ev_t0 = g_ev_t1 = nullptr;
}

template <typename T> struct PhasesOf;
template <> struct PhasesOf<double> { static const double *get(ferm_param *p) { return (const double *) dev(p->phases, "pars->phases"); } };
template <> struct PhasesOf<float> { static const float *get(ferm_param *p) { return (const float *) dev(p->phases_f, "pars->phases_f"); } };

enum { SLOT_TMP = 0, SLOT_DELTA = 1, SLOT_SRC = 2, SLOT_ALPHA = 3, SLOT_LAMBDA = 4 };

// the reference's generated FP32 operator takes `float shift` (sp_fermion_matrix.c:735-746): every double shift its FP32
// callers pass (sp_inverter_full.c:58,78,114, sp_inverter_multishift_full.c:217, inverter_mixedp.c:101) is rounded to float
template <typename T> static inline double shift_as_seen(double shift) { return shift; }
template <> inline double shift_as_seen<float>(double shift) { return (double) (float) shift; }

template <typename T>
static int multishift_impl(const cplx_t<T> *u, ferm_param *pars, RationalApprox *approx, cplx_t<T> *out,
													 const cplx_t<T> *in, double residuo, cplx_t<T> *loc_r, cplx_t<T> *loc_h, cplx_t<T> *loc_s,
													 cplx_t<T> *loc_p, cplx_t<T> *shiftferm, const int max_cg, int *cg_return)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	ensure_ctl();
	const int order = approx->approx_order;
	if (order > MAX_APPROX_ORDER || order < 1) { fprintf(stderr, "multishift_invert: bad approx_order %d\n", order); exit(1); }
	if (verbosity_lv > 3) printf("%s PRECISION VERSION OF MULTISHIFT INVERTER\n", sizeof(T) == 8 ? "DOUBLE" : "SINGLE");
	const T *ph = PhasesOf<T>::get(pars);
	const double m2 = pars->ferm_mass * pars->ferm_mass;
	const long n = g.sizeh, lo = g.r1_lo, cnt = g.r1_hi - g.r1_lo;
	const unsigned int grid = (unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock);
	cudaStream_t st = c.stream;

	// trial solution out = 0 (all sizeh, :67-70); r = p = in; delta = (r,r); source_norm = (in,in)
	STAPLE_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(cplx_t<T>) * 3 * n * order, st));
	blas<T>(OP_ASSIGN, loc_r, in, nullptr, nullptr, 0.0);
	blas<T>(OP_ASSIGN, loc_p, loc_r, nullptr, nullptr, 0.0);
	reduce_local<T>(RED_L2NORM2, loc_r, nullptr, SLOT_DELTA);
	allreduce_results(SLOT_DELTA, 1, st);
	reduce_local<T>(RED_L2NORM2, in, nullptr, SLOT_SRC);
	allreduce_results(SLOT_SRC, 1, st);
	broadcast_kernel<T><<<grid, kBlasBlock, 0, st>>>(shiftferm, in, order, lo, cnt, n);
	count_launch();
	CgmCtl *h = g_h_ctl;
	memset(h, 0, sizeof(CgmCtl));
	h->order = order; h->max_cg = max_cg; h->residuo = residuo;
	for (int i = 0; i < order; i++) h->shifts[i] = approx->RA_b[i];
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(g_d_ctl, h, sizeof(CgmCtl), cudaMemcpyHostToDevice, st));
	cgm_init_kernel<<<1, 32, 0, st>>>(g_d_ctl, result(SLOT_DELTA), result(SLOT_SRC));
	count_launch();
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));   // h is reused for snapshots below

	if (verbosity_lv > 0 && 0 == c.myrank) printf("STARTING CG-M:\n");
	STAPLE_CUDA_CHECK(cudaEventRecord(g_ev_t0, st));
	const int batch = 8;
	int issued = 0, snap = 0, pending[2] = { 0, 0 };
	bool finished = false;
	// multi-GPU with the peer mailboxes: the all-reduce of alpha / lambda is the prologue of the one-warp kernel
	// that advances the recurrences (K12-14 + C4 of the reference fused), no NCCL call inside the iteration
	const bool fuse_red = c.nranks > 1 && c.p2p.on && c.p2p.d_redq != nullptr;
	const RedView red = fuse_red ? make_redview() : single_rank_redview();
	// single rank, or sums over ranks through the peer mailboxes: the recurrences ride in the tail of the kernel
	// whose grid reduction feeds them (Deo -> alpha, shifted pass -> lambda): 4 launches per iteration.  With NCCL
	// all-reduces the sum is a library call between producer and consumer: 6 launches + 2 collectives.
	const bool fuse_tail = (c.nranks == 1 || fuse_red) && c.cgm_fuse_tail;
	// D3 slabs over peer memory: per iteration two exchanges, both consumed from the staging area by the d3 hops of the next
	// operator's face blocks (no unpack, no wait inside the producing launch): p (pushed by the kernel that updates it) and
	// h = Doe p.  s = M^+M p is not exchanged at all.
	const bool interior = halo_lazy_ok();
	// vector updates: the reference's range R1 (interior + one halo slice each side) -- or, with staged halos, the interior only
	const long ulo = interior ? g.r0_lo : lo, ucnt = interior ? g.r0_hi - g.r0_lo : cnt;
	const unsigned int ugrid = (unsigned int) ((ucnt + kBlasBlock - 1) / kBlasBlock), ugrid_f = (unsigned int) ((ucnt + kCgmBlock - 1) / kCgmBlock);
	if (interior) p2p_push_faces(loc_p, sizeof(cplx_t<T>), st);      // the first Doe consumes the halos of p staged, like every other
	// FP32: two sites per thread when every range boundary is even (vol3h even: all of them are multiples of vol3h)
	const bool pack2 = sizeof(T) == 4 && g.vol3h % 2 == 0 && ulo % 2 == 0 && ucnt % 2 == 0 && n % 2 == 0 && g.r0_lo % 2 == 0 && g.r0_hi % 2 == 0;

	auto enqueue_batch = [&]() {
		if (fuse_tail) { c.cgm_hook = g_d_ctl; c.cgm_hook_red = red; }
		for (int b = 0; b < batch; b++) {
			// s = (M^+M) p, alpha = Re(p,s) fused in the Deo epilogue (:113-118)
			apply_mdagm<T>(u, loc_s, loc_p, loc_h, ph, m2, SLOT_ALPHA, &g_d_ctl->done, interior);
			if (!fuse_tail) {
				if (!fuse_red) allreduce_results(SLOT_ALPHA, 1, st);
				cgm_after_alpha_kernel<<<1, 32, 0, st>>>(g_d_ctl, result(SLOT_ALPHA), red);
				count_launch();
			}
			if (pack2)       // FP32: two consecutive sites per thread (float4), every index in units of two sites
				cgm_fused_kernel<float4><<<(unsigned int) ((ucnt / 2 + kCgmBlock - 1) / kCgmBlock), kCgmBlock, 0, st>>>(
					g_d_ctl, (float4 *) out, (float4 *) shiftferm, (float4 *) loc_r, (const float4 *) loc_s, ulo / 2, ucnt / 2, n / 2, g.r0_lo / 2,
					g.r0_hi / 2, partials(SLOT_LAMBDA), ticket(SLOT_LAMBDA), result(SLOT_LAMBDA), fuse_tail ? 1 : 0, red);
			else
				cgm_fused_kernel<cplx_t<T>><<<ugrid_f, kCgmBlock, 0, st>>>(g_d_ctl, out, shiftferm, loc_r, loc_s, ulo, ucnt, n, g.r0_lo,
																																	g.r0_hi, partials(SLOT_LAMBDA), ticket(SLOT_LAMBDA),
																																	result(SLOT_LAMBDA), fuse_tail ? 1 : 0, red);
			if (!fuse_tail) {
				if (!fuse_red) allreduce_results(SLOT_LAMBDA, 1, st);
				cgm_after_lambda_kernel<<<1, 32, 0, st>>>(g_d_ctl, result(SLOT_LAMBDA), red);
				count_launch();
			}
			PushView pv = make_pushview(interior);         // the exchange this p update produces (its number fixes the staging parity)
			if (pack2) {
				pv.top_lo /= 2; pv.bot_lo /= 2; pv.vol3h /= 2;
				cgm_pupdate_kernel<float4><<<(unsigned int) ((ucnt / 2 + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, st>>>(
					g_d_ctl, (float4 *) loc_p, (const float4 *) loc_r, ulo / 2, ucnt / 2, n / 2, pv);
			} else
				cgm_pupdate_kernel<cplx_t<T>><<<ugrid, kBlasBlock, 0, st>>>(g_d_ctl, loc_p, loc_r, ulo, ucnt, n, pv);
			if (interior) c.p2p.h_seq += 1;
			count_launch(2);
		}
		c.cgm_hook = nullptr;
	};
	// Single GPU: a batch is a fixed sequence of launches on one stream whose every data dependence (flags,
	// coefficients, `done`) lives in device memory, so it is captured ONCE into a CUDA graph and replayed --
	// the iteration is launch-gap bound on small lattices.  Multi-GPU batches qualify too when halos and
	// all-reduces go through the peer mailboxes (their sequence numbers are device-resident, and no NCCL call
	// is left inside the iteration).  The legacy default stream cannot be captured: direct launches there.
	cudaGraphExec_t gexec = nullptr;
	const unsigned long long launches_before = c.launches;
	unsigned long long launches_per_batch = 0;
	const bool capturable = c.nranks == 1 || (c.p2p.on && c.p2p_single_launch && c.p2p.d_redq != nullptr);
	if (capturable && st != nullptr && c.use_graphs) {
		cudaGraph_t graph = nullptr;
		const unsigned long long seq_before = c.p2p.h_seq;
		if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
			enqueue_batch();
			cudaError_t e = cudaStreamEndCapture(st, &graph);
			launches_per_batch = c.launches - launches_before;
			c.launches = launches_before;
			// a graph freezes the staging parities of its launches: replayable only if it holds an even number of exchanges
			const bool even = ((c.p2p.h_seq - seq_before) & 1ull) == 0;
			if (e == cudaSuccess && graph != nullptr && even && cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
			if (graph) cudaGraphDestroy(graph);
		}
		c.p2p.h_seq = seq_before;      // nothing was executed
		cudaGetLastError();
	}
	while (!finished) {
		if (gexec) { STAPLE_CUDA_CHECK(cudaGraphLaunch(gexec, st)); c.launches += launches_per_batch; }
		else enqueue_batch();
		issued += batch;
		STAPLE_CUDA_CHECK(cudaGetLastError());
		STAPLE_CUDA_CHECK(cudaMemcpyAsync(&g_h_ctl[snap], g_d_ctl, sizeof(CgmCtl), cudaMemcpyDeviceToHost, st));
		STAPLE_CUDA_CHECK(cudaEventRecord(g_ev_snap[snap], st));
		pending[snap] = 1;
		const int other = snap ^ 1;
		// look at the previous batch's snapshot while this one runs
		if (pending[other]) {
			STAPLE_CUDA_CHECK(cudaEventSynchronize(g_ev_snap[other]));
			pending[other] = 0;
			if (g_h_ctl[other].done) finished = true;
		}
		if (!finished && issued >= max_cg) {
			STAPLE_CUDA_CHECK(cudaEventSynchronize(g_ev_snap[snap]));
			pending[snap] = 0;
			finished = true;   // max_cg reached: the device sets done itself at cg == max_cg
		}
		snap = other;
	}
	STAPLE_CUDA_CHECK(cudaEventRecord(g_ev_t1, st));
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(&g_h_ctl[0], g_d_ctl, sizeof(CgmCtl), cudaMemcpyDeviceToHost, st));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));
	if (gexec) cudaGraphExecDestroy(gexec);
	const int cg = g_h_ctl[0].cg;
	const double source_norm = g_h_ctl[0].source_norm;
	float ms = 0;
	cudaEventElapsedTime(&ms, g_ev_t0, g_ev_t1);
	c.last_iterations = cg; c.last_active = g_h_ctl[0].active_sum; c.last_loop_ms = ms;
	multishift_invert_iterations += cg;

	if (cg == max_cg && 0 == c.myrank) printf("WARNING: maximum number of iterations reached in invert\n");
	if (verbosity_lv > 0 && 0 == c.myrank)
		printf("Terminated multishift_invert ( target res = %1.1e,source_norm = %1.1e )\tCG count %d\n", residuo,
					 source_norm, cg);

	// interior-only updates: the solutions leave with valid halos all the same (the reference's do, through its updates over R1)
	if (interior)
		for (int i = 0; i < order; i++) p2p_exchange_fermion(out + (long) i * 3 * n, sizeof(cplx_t<T>), st);
	// post-loop verification of every shifted system (:211-229)
	int check = 1;
	if (verbosity_lv > 2 && 0 == c.myrank) printf("Relative Res:");
	for (int i = 0; i < order; i++) {
		blas<T>(OP_ASSIGN, loc_p, out + (long) i * 3 * n, nullptr, nullptr, 0.0);
		apply_mdagm<T>(u, loc_s, loc_p, loc_h, ph, m2 + shift_as_seen<T>(approx->RA_b[i]), -1, nullptr);
		blas<T>(OP_IN1_MINUS_IN2, loc_h, in, loc_s, nullptr, 0.0);
		const double giustoono = reduce_global<T>(RED_L2NORM2, loc_h, nullptr).re / source_norm;
		check *= (giustoono <= 1) ? 1 : 0;
		if (verbosity_lv > 2 && 0 == c.myrank && residuo != 0) printf("\t%1.1e", sqrt(giustoono) / residuo);
	}
	if (verbosity_lv > 2 && 0 == c.myrank) printf("\n");
	if (verbosity_lv > 0 && 0 == c.myrank)
		printf("Inverter Multishift timings:\nTiming Loops    : %f / %d (%f per iteration)\n", ms * 1e-3, cg, ms * 1e-3 / (cg > 0 ? cg : 1));
	*cg_return = cg;
	return check == 1 ? INVERTER_SUCCESS : INVERTER_FAILURE;
}

// ------------------------------------------------------------------ single-system CG, device resident
// solution += omega p ; r -= omega s ; lambda = |r|^2 over the reduction range   (inverter_full.c:86-92; inverter_mixedp.c:110,131-133
// with x = the FP32 accumulator `out`).  In a "touch" iteration of the mixed-precision solver r is not updated here -- it is
// recomputed in double precision by mp_refresh_kernel, which then also owns the lambda reduction.
// (E: double2 / float2 = one site per thread, float4 = two consecutive FP32 sites; index arguments in units of E)
template <typename E>
__global__ void __launch_bounds__(kBlasBlock) cg_update_kernel(CgCtl *c, E *x, E *r, const E *p, const E *s,
																															 long lo, long cnt, long n, long r0_lo, long r0_hi, double *partials,
																															 unsigned int *ticket, double *result, RedView red)
{
	using S = typename Elem<E>::S;
	if (c->done) return;
	__shared__ double sm[32];
	__shared__ bool last;
	const double omega = c->omega;
	const bool touch = c->touch != 0;
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	double nrm = 0.0;
	if (t < cnt) {
		const long i = lo + t;
		E pv[3], xv[3], sv[3], rv[3];
#pragma unroll
		for (int col = 0; col < 3; col++) { pv[col] = p[col * n + i]; xv[col] = x[col * n + i]; }
		if (!touch) {
#pragma unroll
			for (int col = 0; col < 3; col++) { sv[col] = s[col * n + i]; rv[col] = r[col * n + i]; }
		}
#pragma unroll
		for (int col = 0; col < 3; col++) {
			const long j = col * n + i;
			E xn;
#pragma unroll
			for (int e = 0; e < Elem<E>::N; e++) comp(xn, e) = (S) (comp(pv[col], e) * omega + comp(xv[col], e));
			x[j] = xn;
			if (!touch) {
				E rn;
#pragma unroll
				for (int e = 0; e < Elem<E>::N; e++) comp(rn, e) = (S) (comp(sv[col], e) * (-omega) + comp(rv[col], e));
				r[j] = rn;
				if (i >= r0_lo && i < r0_hi) {
#pragma unroll
					for (int e = 0; e < Elem<E>::N; e++) nrm += (double) comp(rn, e) * comp(rn, e);
				}
			}
		}
	}
	if (touch) return;
	block_sum1(nrm, sm);
	if (threadIdx.x == 0) {
		partials[blockIdx.x] = nrm;
		__threadfence();
		last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		double acc = 0.0;
		for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) acc += __ldcg(partials + k);
		block_sum1(acc, sm);
		if (threadIdx.x == 0) { result[0] = acc; *ticket = 0u; }
		__syncthreads();
		if (threadIdx.x < 32) cg_after_lambda_warp(c, result, red);
	}
}

// p = r + gammag p  (inverter_full.c:96, inverter_mixedp.c:139)
template <typename E>
__global__ void __launch_bounds__(kBlasBlock) cg_pupdate_kernel(const CgCtl *c, E *p, const E *r, long lo, long cnt, long n)
{
	using S = typename Elem<E>::S;
	if (c->done) return;
	const double gammag = c->gammag;
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	E pv[3], rv[3];
#pragma unroll
	for (int col = 0; col < 3; col++) { pv[col] = p[col * n + lo + t]; rv[col] = r[col * n + lo + t]; }
#pragma unroll
	for (int col = 0; col < 3; col++) {
		E pn;
#pragma unroll
		for (int e = 0; e < Elem<E>::N; e++) comp(pn, e) = (S) (comp(pv[col], e) * gammag + comp(rv[col], e));
		p[col * n + lo + t] = pn;
	}
}

// mixed precision, last kernel of every iteration (one thread): this iteration is over -- if the NEXT one has to refresh the
// residual in double precision ("magic touch", decided in this iteration's lambda tail), the batch stops here and the host
// enqueues that iteration
__global__ void cg_promote_kernel(CgCtl *c)
{
	if (c->done) return;
	c->touch = 0; c->no_touch = 1;
	if (c->touch_next) { c->done = 1; c->paused = 1; }
}
// first kernel of a magic-touch iteration (one thread)
__global__ void cg_resume_kernel(CgCtl *c) { c->done = 0; c->paused = 0; c->touch = 1; c->no_touch = 0; c->touch_next = 0; }

// "magic touch", first half (inverter_mixedp.c:115-116, :126): solution += out over the update range, out = 0 everywhere
__global__ void __launch_bounds__(kBlasBlock) mp_accumulate_kernel(const CgCtl *c, double2 *sol, float2 *out, long r1_lo, long r1_hi, long n)
{
	if (c->no_touch) return;
	const long i = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (i >= n) return;
#pragma unroll
	for (int col = 0; col < 3; col++) {
		const long j = col * n + i;
		if (i >= r1_lo && i < r1_hi) { const float2 o = out[j]; const double2 x = sol[j]; sol[j] = make_double2(x.x + (double) o.x, x.y + (double) o.y); }
		out[j] = make_float2(0.f, 0.f);
	}
}
// second half (:118-125, :133): d_r = in - d_s ; r_f = (float) d_r ; lambda = |r_f|^2
__global__ void __launch_bounds__(kBlasBlock) mp_refresh_kernel(CgCtl *c, const double2 *in, const double2 *d_s, double2 *d_r, float2 *r_f,
																																long lo, long cnt, long n, long r0_lo, long r0_hi, double *partials,
																																unsigned int *ticket, double *result, RedView red)
{
	if (c->no_touch) return;
	__shared__ double sm[32];
	__shared__ bool last;
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	double nrm = 0.0;
	if (t < cnt) {
		const long i = lo + t;
#pragma unroll
		for (int col = 0; col < 3; col++) {
			const long j = col * n + i;
			const double2 a = in[j], b = d_s[j];
			const double2 d = make_double2(a.x - b.x, a.y - b.y);
			d_r[j] = d;
			const float2 f = make_float2((float) d.x, (float) d.y);
			r_f[j] = f;
			if (i >= r0_lo && i < r0_hi) nrm += (double) f.x * f.x + (double) f.y * f.y;
		}
	}
	block_sum1(nrm, sm);
	if (threadIdx.x == 0) {
		partials[blockIdx.x] = nrm;
		__threadfence();
		last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		double acc = 0.0;
		for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) acc += __ldcg(partials + k);
		block_sum1(acc, sm);
		if (threadIdx.x == 0) { result[0] = acc; *ticket = 0u; }
		__syncthreads();
		if (threadIdx.x < 32) cg_after_lambda_warp(c, result, red);
	}
}

// launches of the two update kernels: FP32 with two sites per thread where every range boundary is even
template <typename T>
static void launch_cg_update(CgCtl *ctl, cplx_t<T> *x, cplx_t<T> *r, const cplx_t<T> *p, const cplx_t<T> *s, const RedView &red, cudaStream_t st)
{
	const Geom &g = ctx().g;
	const long n = g.sizeh, lo = g.r1_lo, cnt = g.r1_hi - g.r1_lo;
	const bool pack2 = sizeof(T) == 4 && n % 2 == 0 && lo % 2 == 0 && cnt % 2 == 0 && g.r0_lo % 2 == 0 && g.r0_hi % 2 == 0;
	if (pack2)
		cg_update_kernel<float4><<<(unsigned int) ((cnt / 2 + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, st>>>(
			ctl, (float4 *) x, (float4 *) r, (const float4 *) p, (const float4 *) s, lo / 2, cnt / 2, n / 2, g.r0_lo / 2, g.r0_hi / 2,
			partials(SLOT_LAMBDA), ticket(SLOT_LAMBDA), result(SLOT_LAMBDA), red);
	else
		cg_update_kernel<cplx_t<T>><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, st>>>(
			ctl, x, r, p, s, lo, cnt, n, g.r0_lo, g.r0_hi, partials(SLOT_LAMBDA), ticket(SLOT_LAMBDA), result(SLOT_LAMBDA), red);
	count_launch();
}
template <typename T>
static void launch_cg_pupdate(const CgCtl *ctl, cplx_t<T> *p, const cplx_t<T> *r, cudaStream_t st)
{
	const Geom &g = ctx().g;
	const long n = g.sizeh, lo = g.r1_lo, cnt = g.r1_hi - g.r1_lo;
	const bool pack2 = sizeof(T) == 4 && n % 2 == 0 && lo % 2 == 0 && cnt % 2 == 0;
	if (pack2)
		cg_pupdate_kernel<float4><<<(unsigned int) ((cnt / 2 + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, st>>>(ctl, (float4 *) p, (const float4 *) r, lo / 2, cnt / 2, n / 2);
	else
		cg_pupdate_kernel<cplx_t<T>><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, st>>>(ctl, p, r, lo, cnt, n);
	count_launch();
}

static CgCtl *g_d_cg = nullptr, *g_h_cg = nullptr;     // device block; pinned: [0],[1] snapshots, [2] upload staging
static cudaEvent_t g_ev_cgsnap[2] = { nullptr, nullptr };
static void ensure_cg_ctl()
{
	if (g_d_cg) return;
	STAPLE_CUDA_CHECK(cudaMalloc(&g_d_cg, sizeof(CgCtl)));
	STAPLE_CUDA_CHECK(cudaHostAlloc(&g_h_cg, 3 * sizeof(CgCtl), cudaHostAllocDefault));
	for (int i = 0; i < 2; i++) STAPLE_CUDA_CHECK(cudaEventCreateWithFlags(&g_ev_cgsnap[i], cudaEventDisableTiming));
}
void release_cg_state()
{
	if (!g_d_cg) return;
	cudaFree(g_d_cg); cudaFreeHost(g_h_cg);
	for (int i = 0; i < 2; i++) cudaEventDestroy(g_ev_cgsnap[i]);
	g_d_cg = nullptr; g_h_cg = nullptr; g_ev_cgsnap[0] = g_ev_cgsnap[1] = nullptr;
}

// device-resident iteration loops need the sums over ranks inside the kernels' tails: one rank, or the peer mailboxes
static bool cg_device_resident()
{
	const Ctx &c = ctx();
	return c.cg_device_loops && (c.nranks == 1 || c.loopback || (c.p2p.on && c.p2p.d_redq != nullptr && c.p2p_single_launch));
}

// Runs batches of iterations (captured once into a CUDA graph where the stream allows it) until the device sets `done`.
// Small lattices: batches of 8, the host looks at a pinned snapshot of the control block of the PREVIOUS batch while the next
// one runs (no host synchronisation inside the loop; up to two batches of no-op launches after convergence cost microseconds).
// Large lattices (an iteration takes a millisecond, and every no-op launch still dispatches ~10^5 empty blocks): batches of 2
// and the snapshot of the batch itself.
struct CgBatchRunner {
	cudaGraphExec_t gexec = nullptr;
	unsigned long long launches_per_batch = 0;
	int batch = 8, parity = 0;
	bool lag = true;
	template <typename F> void prepare(F enqueue_iteration)
	{
		Ctx &c = ctx();
		cudaStream_t st = c.stream;
		if (c.g.sizeh >= (1l << 21)) { batch = 2; lag = false; }
		if (st == nullptr || !c.use_graphs) return;
		const unsigned long long before = c.launches, seq_before = c.p2p.h_seq;
		cudaGraph_t graph = nullptr;
		if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
			for (int b = 0; b < batch; b++) enqueue_iteration();
			cudaError_t e = cudaStreamEndCapture(st, &graph);
			launches_per_batch = c.launches - before;
			c.launches = before;
			// a graph freezes the staging parities of its launches: replayable only if it holds an even number of exchanges, and
			// only from the parity it was captured at (run() checks)
			const bool even = ((c.p2p.h_seq - seq_before) & 1ull) == 0;
			if (e == cudaSuccess && graph != nullptr && even && cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
			if (graph) cudaGraphDestroy(graph);
		}
		c.p2p.h_seq = seq_before;      // nothing was executed
		parity = (int) (seq_before & 1ull);
		cudaGetLastError();
	}
	template <typename F> CgCtl run(F enqueue_iteration)
	{
		Ctx &c = ctx();
		cudaStream_t st = c.stream;
		int snap = 0, pending[2] = { 0, 0 };
		bool finished = false;
		const bool replay = gexec != nullptr && (int) (c.p2p.h_seq & 1ull) == parity;
		while (!finished) {
			if (replay) { STAPLE_CUDA_CHECK(cudaGraphLaunch(gexec, st)); c.launches += launches_per_batch; }
			else for (int b = 0; b < batch; b++) enqueue_iteration();
			STAPLE_CUDA_CHECK(cudaMemcpyAsync(&g_h_cg[snap], g_d_cg, sizeof(CgCtl), cudaMemcpyDeviceToHost, st));
			STAPLE_CUDA_CHECK(cudaEventRecord(g_ev_cgsnap[snap], st));
			pending[snap] = 1;
			const int look = lag ? snap ^ 1 : snap;
			if (pending[look]) {
				STAPLE_CUDA_CHECK(cudaEventSynchronize(g_ev_cgsnap[look]));
				pending[look] = 0;
				if (g_h_cg[look].done) finished = true;
			}
			snap ^= 1;
		}
		STAPLE_CUDA_CHECK(cudaMemcpyAsync(&g_h_cg[0], g_d_cg, sizeof(CgCtl), cudaMemcpyDeviceToHost, st));
		STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));
		return g_h_cg[0];
	}
	~CgBatchRunner() { if (gexec) cudaGraphExecDestroy(gexec); }
};

static void cg_upload_ctl(const CgCtl &h)
{
	g_h_cg[2] = h;
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(g_d_cg, &g_h_cg[2], sizeof(CgCtl), cudaMemcpyHostToDevice, ctx().stream));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
}

// restarted CG (inverter_full.c:19-132).  The iteration loop (:78-101) runs on the device: per iteration Doe, Deo (mass term,
// alpha = Re(p,s) and omega fused), ONE update kernel (solution, r, lambda, gammag, loop condition), p update -- four
// launches, no host synchronisation; the host only sequences the restarts (:66-77) and the final check (:104-118).
// With NCCL all-reduces (no peer mailboxes) the sums over ranks are library calls between kernels: cg_impl_hostloop.
template <typename T>
static int cg_impl_hostloop(const cplx_t<T> *u, ferm_param *pars, cplx_t<T> *solution, const cplx_t<T> *in, double res,
														cplx_t<T> *loc_r, cplx_t<T> *loc_h, cplx_t<T> *loc_s, cplx_t<T> *loc_p, const int max_cg,
														double shift, int *cg_return);

template <typename T>
static int cg_impl(const cplx_t<T> *u, ferm_param *pars, cplx_t<T> *solution, const cplx_t<T> *in, double res,
									 cplx_t<T> *loc_r, cplx_t<T> *loc_h, cplx_t<T> *loc_s, cplx_t<T> *loc_p, const int max_cg,
									 double shift, int *cg_return)
{
	if (!cg_device_resident()) return cg_impl_hostloop<T>(u, pars, solution, in, res, loc_r, loc_h, loc_s, loc_p, max_cg, shift, cg_return);
	Ctx &c = ctx();
	ensure_cg_ctl();
	const T *ph = PhasesOf<T>::get(pars);
	const double m2 = pars->ferm_mass * pars->ferm_mass + shift_as_seen<T>(shift);
	cudaStream_t st = c.stream;
	const bool fuse_red = c.nranks > 1 && !c.loopback;
	const RedView red = fuse_red ? make_redview() : single_rank_redview();
	int cg = 0;
	double lambda = 0;
	const double source_norm = reduce_global<T>(RED_L2NORM2, in, nullptr).re;
	auto iteration = [&]() {
		c.cg_hook = g_d_cg; c.cgm_hook_red = red;
		apply_mdagm<T>(u, loc_s, loc_p, loc_h, ph, m2, SLOT_ALPHA, &g_d_cg->done);
		c.cg_hook = nullptr;
		launch_cg_update<T>(g_d_cg, solution, loc_r, loc_p, loc_s, red, st);
		launch_cg_pupdate<T>(g_d_cg, loc_p, loc_r, st);
	};
	CgBatchRunner runner;
	runner.prepare(iteration);
	do {
		apply_mdagm<T>(u, loc_s, solution, loc_h, ph, m2, -1, nullptr);
		blas<T>(OP_IN1_MINUS_IN2, loc_r, in, loc_s, nullptr, 0.0);
		blas<T>(OP_ASSIGN, loc_p, loc_r, nullptr, nullptr, 0.0);
		const double delta = reduce_global<T>(RED_L2NORM2, loc_r, nullptr).re;
		if (verbosity_lv > 3 && 0 == c.myrank) printf("STARTING CG:\nCG\tR\n");
		CgCtl h;
		memset(&h, 0, sizeof(h));
		h.delta = delta; h.source_norm = source_norm; h.res = res; h.stop_factor = kSafetyMargin;
		h.cg = cg; h.cg_restarted = 0; h.restarting_every = inverter_tricks.restartingEvery; h.max_cg = max_cg;
		h.no_touch = 1;
		cg_upload_ctl(h);
		const CgCtl f = runner.run(iteration);
		cg = f.cg; lambda = f.lambda;
	} while ((sqrt(lambda / source_norm) > res) && cg < max_cg);

	apply_mdagm<T>(u, loc_s, solution, loc_h, ph, m2, -1, nullptr);
	blas<T>(OP_IN1_MINUS_IN2, loc_h, in, loc_s, nullptr, 0.0);
	const double current_res = reduce_global<T>(RED_L2NORM2, loc_h, nullptr).re / source_norm;
	if (verbosity_lv > 1 && 0 == c.myrank) {
		printf("Terminated invert after   %d    iterations", cg);
		printf("[res/stop_res=  %e , stop_res=%e ]\n", sqrt(current_res) / res, res);
	}
	if (cg == max_cg && 0 == c.myrank) printf("WARNING: maximum number of iterations reached in invert\n");
	*cg_return = cg;
	return sqrt(current_res) <= res ? INVERTER_SUCCESS : INVERTER_FAILURE;
}

// host-driven form: scalars are read back twice per iteration
template <typename T>
static int cg_impl_hostloop(const cplx_t<T> *u, ferm_param *pars, cplx_t<T> *solution, const cplx_t<T> *in, double res,
									 cplx_t<T> *loc_r, cplx_t<T> *loc_h, cplx_t<T> *loc_s, cplx_t<T> *loc_p, const int max_cg,
									 double shift, int *cg_return)
{
	Ctx &c = ctx();
	const T *ph = PhasesOf<T>::get(pars);
	const double m2 = pars->ferm_mass * pars->ferm_mass + shift_as_seen<T>(shift);
	int cg = 0;
	double delta, alpha, lambda = 0, omega, gammag;
	const double source_norm = reduce_global<T>(RED_L2NORM2, in, nullptr).re;
	do {
		apply_mdagm<T>(u, loc_s, solution, loc_h, ph, m2, -1, nullptr);
		blas<T>(OP_IN1_MINUS_IN2, loc_r, in, loc_s, nullptr, 0.0);
		blas<T>(OP_ASSIGN, loc_p, loc_r, nullptr, nullptr, 0.0);
		delta = reduce_global<T>(RED_L2NORM2, loc_r, nullptr).re;
		if (verbosity_lv > 3 && 0 == c.myrank) printf("STARTING CG:\nCG\tR\n");
		int cg_restarted = 0;
		do {
			cg++; cg_restarted++;
			apply_mdagm<T>(u, loc_s, loc_p, loc_h, ph, m2, SLOT_ALPHA, nullptr);
			fetch_results(SLOT_ALPHA, 1, &alpha);
			omega = delta / alpha;
			blas<T>(OP_IN1XFACTOR_PLUS_IN2, solution, loc_p, solution, nullptr, omega);
			blas<T>(OP_IN1XFACTOR_PLUS_IN2, loc_r, loc_s, loc_r, nullptr, -omega);
			lambda = reduce_global<T>(RED_L2NORM2, loc_r, nullptr).re;
			gammag = lambda / delta;
			delta = lambda;
			blas<T>(OP_IN1XFACTOR_PLUS_IN2, loc_p, loc_p, loc_r, nullptr, gammag);
			if (verbosity_lv > 3 && cg % 100 == 0 && 0 == c.myrank) printf("%d\t%1.1e\n", cg, sqrt(lambda / source_norm) / res);
		} while ((sqrt(lambda / source_norm) > res * kSafetyMargin) && cg_restarted < inverter_tricks.restartingEvery);
	} while ((sqrt(lambda / source_norm) > res) && cg < max_cg);

	apply_mdagm<T>(u, loc_s, solution, loc_h, ph, m2, -1, nullptr);
	blas<T>(OP_IN1_MINUS_IN2, loc_h, in, loc_s, nullptr, 0.0);
	const double current_res = reduce_global<T>(RED_L2NORM2, loc_h, nullptr).re / source_norm;
	if (verbosity_lv > 1 && 0 == c.myrank) {
		printf("Terminated invert after   %d    iterations", cg);
		printf("[res/stop_res=  %e , stop_res=%e ]\n", sqrt(current_res) / res, res);
	}
	if (cg == max_cg && 0 == c.myrank) printf("WARNING: maximum number of iterations reached in invert\n");
	*cg_return = cg;
	return sqrt(current_res) <= res ? INVERTER_SUCCESS : INVERTER_FAILURE;
}

}   // namespace staple

using namespace staple;

#define DD(p) ((double2 *) dev(p, #p))
#define DF(p) ((float2 *) dev(p, #p))
#define CDD(p) ((const double2 *) dev(p, #p))
#define CDF(p) ((const float2 *) dev(p, #p))

static vec3_soa_f *g_aux1_f = nullptr, *g_ferm_shiftmulti_acc_f = nullptr;   // staple_set_sp_globals
static int g_last_refinement_iterations = 0;
extern "C" {
__attribute__((weak)) vec3_soa_f *aux1_f = nullptr;                  // alloc_vars.h globals; the host program's definitions win
__attribute__((weak)) vec3_soa_f *ferm_shiftmulti_acc_f = nullptr;
}

extern "C" {

int multishift_invert(const su3_soa *u, ferm_param *pars, RationalApprox *approx, vec3_soa *out, const vec3_soa *in,
											double residuo, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_s, vec3_soa *loc_p,
											vec3_soa *shiftferm, const int max_cg, int *cg_return)
{
	require_init("multishift_invert");
	return multishift_impl<double>(CDD(u), pars, approx, DD(out), CDD(in), residuo, DD(loc_r), DD(loc_h), DD(loc_s),
																 DD(loc_p), DD(shiftferm), max_cg, cg_return);
}
int multishift_invert_f(const su3_soa_f *u, ferm_param *pars, RationalApprox *approx, vec3_soa_f *out,
												const vec3_soa_f *in, double residuo, vec3_soa_f *loc_r, vec3_soa_f *loc_h, vec3_soa_f *loc_s,
												vec3_soa_f *loc_p, vec3_soa_f *shiftferm, const int max_cg, int *cg_return)
{
	require_init("multishift_invert_f");
	return multishift_impl<float>(CDF(u), pars, approx, DF(out), CDF(in), residuo, DF(loc_r), DF(loc_h), DF(loc_s),
																DF(loc_p), DF(shiftferm), max_cg, cg_return);
}

int ker_invert_openacc(const su3_soa *u, ferm_param *pars, vec3_soa *solution, const vec3_soa *in, double res,
											 vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_s, vec3_soa *loc_p, const int max_cg, double shift,
											 int *cg_return)
{
	require_init("ker_invert_openacc");
	return cg_impl<double>(CDD(u), pars, DD(solution), CDD(in), res, DD(loc_r), DD(loc_h), DD(loc_s), DD(loc_p), max_cg,
												 shift, cg_return);
}
int ker_invert_openacc_f(const su3_soa_f *u, ferm_param *pars, vec3_soa_f *solution, const vec3_soa_f *in, double res,
												 vec3_soa_f *loc_r, vec3_soa_f *loc_h, vec3_soa_f *loc_s, vec3_soa_f *loc_p, const int max_cg,
												 double shift, int *cg_return)
{
	require_init("ker_invert_openacc_f");
	return cg_impl<float>(CDF(u), pars, DF(solution), CDF(in), res, DF(loc_r), DF(loc_h), DF(loc_s), DF(loc_p), max_cg,
												shift, cg_return);
}

// inverter_mixedp.c:41-181 -- FP32 CG with FP64 "magic touch" reliable updates, SAFETY_MARGIN 0.9
int inverter_mixed_precision(inverter_package ip, ferm_param *pars, vec3_soa *solution_h, const vec3_soa *in_h, double res,
														 const int max_cg, double shift, int *cg_return)
{
	require_init("inverter_mixed_precision");
	Ctx &c = ctx();
	const double2 *u = CDD(ip.u); const float2 *u_f = CDF(ip.u_f);
	double2 *solution = DD(solution_h); const double2 *in = CDD(in_h);
	double2 *d_r = DD(ip.loc_r), *d_h = DD(ip.loc_h), *d_s = DD(ip.loc_s);
	float2 *loc_r = DF(ip.loc_r_f), *loc_h = DF(ip.loc_h_f), *loc_s = DF(ip.loc_s_f), *loc_p = DF(ip.loc_p_f), *out = DF(ip.out_f);
	const double *ph = (const double *) dev(pars->phases, "pars->phases");
	const float *ph_f = (const float *) dev(pars->phases_f, "pars->phases_f");
	const double m2 = pars->ferm_mass * pars->ferm_mass + shift;
	const double m2_f = pars->ferm_mass * pars->ferm_mass + shift_as_seen<float>(shift);   // inverter_mixedp.c:101
	const long n = c.g.sizeh;
	int cg = 0, magicTouchCount = 0;
	double delta, alpha, lambda, omega, gammag, lastMaxResNorm = 0;
	const double source_norm = reduce_global<double>(RED_L2NORM2, in, nullptr).re;
	if (cg_device_resident()) {
		// The iteration loop (:99-141) on the device: FP32 Doe, Deo (alpha, omega fused), one update kernel (out, r, lambda, gammag,
		// loop condition and the magic-touch decision for the next iteration), p update; the double-precision kernels of a magic touch
		// (solution += out, M^+M solution, r = in - s in FP64 -> FP32) are part of every iteration and return at once unless the
		// control block says "touch".  No host synchronisation inside the loop.
		const Geom &g = c.g;
		ensure_cg_ctl();
		const long lo = g.r1_lo, cnt = g.r1_hi - g.r1_lo;
		const unsigned int grid = (unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), grid_all = (unsigned int) ((n + kBlasBlock - 1) / kBlasBlock);
		cudaStream_t st = c.stream;
		const bool fuse_red = c.nranks > 1 && !c.loopback;
		const RedView red = fuse_red ? make_redview() : single_rank_redview();
		apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
		blas<double>(OP_IN1_MINUS_IN2, d_r, in, d_s, nullptr, 0.0);
		convert_double_to_float_vec3_soa((const vec3_soa *) d_r, (vec3_soa_f *) loc_r);
		blas<float>(OP_ASSIGN, loc_p, loc_r, nullptr, nullptr, 0.0);
		delta = reduce_global<float>(RED_L2NORM2, loc_r, nullptr).re;
		blas<float>(OP_ZERO, out, nullptr, nullptr, nullptr, 0.0);
		if (verbosity_lv > 3 && 0 == c.myrank) printf("STARTING CG:\nCG\tR - mixed precision\n");
		CgCtl h;
		memset(&h, 0, sizeof(h));
		h.delta = delta; h.source_norm = source_norm; h.res = res; h.stop_factor = 0.9; h.mixed = 1;
		h.mixed_delta = inverter_tricks.mixedPrecisionDelta; h.max_cg = max_cg; h.restarting_every = 0;
		// the decision of the FIRST iteration (:112-114 with lastMaxResNorm = 0): lastMaxResNorm <- delta, touch iff delta < mpd * delta
		h.last_max_res_norm = delta; h.touch = delta < h.mixed_delta * delta ? 1 : 0;
		if (h.touch) h.last_max_res_norm = 0.0;
		h.no_touch = 1;
		if (h.touch) { h.touch = 0; h.touch_next = 1; h.done = 1; h.paused = 1; }      // (cannot happen for mixedPrecisionDelta < 1)
		cg_upload_ctl(h);
		// a plain iteration (captured in the batches) and a magic-touch iteration (enqueued by the host when the device asks for it:
		// `paused`).  The double-precision kernels are deliberately NOT part of the batches: as no-ops they would still dispatch
		// ~10^5 empty blocks per iteration on a 48^3 x 96 lattice (measured: +0.37 ms on a 1.0 ms iteration).
		auto head = [&]() {
			c.cg_hook = g_d_cg; c.cgm_hook_red = red;
			apply_mdagm<float>(u_f, loc_s, loc_p, loc_h, ph_f, m2_f, SLOT_ALPHA, &g_d_cg->done);
			c.cg_hook = nullptr;
			launch_cg_update<float>(g_d_cg, out, loc_r, loc_p, loc_s, red, st);
		};
		auto tail = [&]() {
			launch_cg_pupdate<float>(g_d_cg, loc_p, loc_r, st);
			cg_promote_kernel<<<1, 1, 0, st>>>(g_d_cg);
			count_launch();
		};
		auto iteration = [&]() { head(); tail(); };
		CgBatchRunner runner;
		runner.prepare(iteration);
		CgCtl f;
		for (;;) {
			f = runner.run(iteration);
			if (!f.paused) break;
			cg_resume_kernel<<<1, 1, 0, st>>>(g_d_cg);           // done = paused = 0, touch = 1
			head();                                              // FP32 M^+M, alpha, omega, out += omega p   (r untouched)
			mp_accumulate_kernel<<<grid_all, kBlasBlock, 0, st>>>(g_d_cg, solution, out, g.r1_lo, g.r1_hi, n);
			count_launch(2);
			apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
			mp_refresh_kernel<<<grid, kBlasBlock, 0, st>>>(g_d_cg, in, d_s, d_r, loc_r, lo, cnt, n, g.r0_lo, g.r0_hi, partials(SLOT_LAMBDA),
																										 ticket(SLOT_LAMBDA), result(SLOT_LAMBDA), red);
			count_launch();
			tail();
		}
		cg = f.cg; magicTouchCount = f.magic_touches;
		combine_add_in2_into_in1_mixed_precision((vec3_soa *) solution, (const vec3_soa_f *) out);
		apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
		blas<double>(OP_IN1_MINUS_IN2, d_h, in, d_s, nullptr, 0.0);
		const double giustoono = reduce_global<double>(RED_L2NORM2, d_h, nullptr).re / source_norm;
		if (verbosity_lv > 1 && 0 == c.myrank) {
			printf("Terminated invert after   %d    iterations", cg);
			printf("[res/stop_res=  %e , stop_res=%e ] (%d magic touches)\n", sqrt(giustoono) / res, res, magicTouchCount);
		}
		if (cg == max_cg && 0 == c.myrank) printf("WARNING: maximum number of iterations reached in invert\n");
		*cg_return = cg;
		return sqrt(giustoono) <= res ? INVERTER_SUCCESS : INVERTER_FAILURE;
	}

	apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
	blas<double>(OP_IN1_MINUS_IN2, d_r, in, d_s, nullptr, 0.0);
	convert_double_to_float_vec3_soa((const vec3_soa *) d_r, (vec3_soa_f *) loc_r);
	blas<float>(OP_ASSIGN, loc_p, loc_r, nullptr, nullptr, 0.0);
	delta = reduce_global<float>(RED_L2NORM2, loc_r, nullptr).re;
	blas<float>(OP_ZERO, out, nullptr, nullptr, nullptr, 0.0);
	if (verbosity_lv > 3 && 0 == c.myrank) printf("STARTING CG:\nCG\tR - mixed precision\n");
	do {
		cg++;
		apply_mdagm<float>(u_f, loc_s, loc_p, loc_h, ph_f, m2_f, SLOT_ALPHA, nullptr);
		fetch_results(SLOT_ALPHA, 1, &alpha);
		omega = delta / alpha;
		blas<float>(OP_IN1XFACTOR_PLUS_IN2, out, loc_p, out, nullptr, omega);
		if (lastMaxResNorm < delta) lastMaxResNorm = delta;
		if (delta < inverter_tricks.mixedPrecisionDelta * lastMaxResNorm) {
			combine_add_in2_into_in1_mixed_precision((vec3_soa *) solution, (const vec3_soa_f *) out);
			apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
			blas<double>(OP_IN1_MINUS_IN2, d_r, in, d_s, nullptr, 0.0);
			convert_double_to_float_vec3_soa((const vec3_soa *) d_r, (vec3_soa_f *) loc_r);
			blas<float>(OP_ZERO, out, nullptr, nullptr, nullptr, 0.0);
			lastMaxResNorm = 0;
			magicTouchCount++;
		} else blas<float>(OP_IN1XFACTOR_PLUS_IN2, loc_r, loc_s, loc_r, nullptr, -omega);
		lambda = reduce_global<float>(RED_L2NORM2, loc_r, nullptr).re;
		gammag = lambda / delta;
		delta = lambda;
		blas<float>(OP_IN1XFACTOR_PLUS_IN2, loc_p, loc_p, loc_r, nullptr, gammag);
	} while ((sqrt(lambda / source_norm) > res * 0.9) && cg < max_cg);
	combine_add_in2_into_in1_mixed_precision((vec3_soa *) solution, (const vec3_soa_f *) out);

	apply_mdagm<double>(u, d_s, solution, d_h, ph, m2, -1, nullptr);
	blas<double>(OP_IN1_MINUS_IN2, d_h, in, d_s, nullptr, 0.0);
	const double giustoono = reduce_global<double>(RED_L2NORM2, d_h, nullptr).re / source_norm;
	if (verbosity_lv > 1 && 0 == c.myrank) {
		printf("Terminated invert after   %d    iterations", cg);
		printf("[res/stop_res=  %e , stop_res=%e ] (%d magic touches)\n", sqrt(giustoono) / res, res, magicTouchCount);
	}
	if (cg == max_cg && 0 == c.myrank) printf("WARNING: maximum number of iterations reached in invert\n");
	(void) n;
	*cg_return = cg;
	return sqrt(giustoono) <= res ? INVERTER_SUCCESS : INVERTER_FAILURE;
}

// inverter_package.c:18-72 (including the aliasing check)
static void check_aliases(void **p, int nptrs)
{
	for (int i = 0; i < nptrs; i++)
		for (int j = i + 1; j < nptrs; j++)
			if (p[i] == p[j] && 0 != p[i]) {
				printf("BAD SETUP OF INVERTER PACKAGE! (%s:%d)\n", __FILE__, __LINE__);
				printf("Pointer %p used twice (%d == %d).\n", p[i], i, j);
				exit(1);
			}
}
void setup_inverter_package_dp(inverter_package *ip, su3_soa *u, vec3_soa *ferm_shift_temp, int nshifts, vec3_soa *loc_r,
															 vec3_soa *loc_h, vec3_soa *loc_s, vec3_soa *loc_p)
{
	ip->u = u; ip->ferm_shift_temp = ferm_shift_temp; ip->nshifts = nshifts;
	ip->loc_r = loc_r; ip->loc_h = loc_h; ip->loc_s = loc_s; ip->loc_p = loc_p;
	void *all[] = { loc_r, loc_h, loc_s, loc_p, ferm_shift_temp };
	check_aliases(all, 5);
}
void setup_inverter_package_sp(inverter_package *ip, su3_soa_f *u_f, vec3_soa_f *ferm_shift_temp_f, int nshifts,
															 vec3_soa_f *loc_r_f, vec3_soa_f *loc_h_f, vec3_soa_f *loc_s_f, vec3_soa_f *loc_p_f,
															 vec3_soa_f *out_f)
{
	ip->u_f = u_f; ip->ferm_shift_temp_f = ferm_shift_temp_f; ip->nshifts = nshifts;
	ip->loc_r_f = loc_r_f; ip->loc_h_f = loc_h_f; ip->loc_s_f = loc_s_f; ip->loc_p_f = loc_p_f; ip->out_f = out_f;
	void *all[] = { loc_r_f, loc_h_f, loc_s_f, loc_p_f, ferm_shift_temp_f, out_f };
	check_aliases(all, 6);
}

// inverter_wrappers.c:24-39
void convergence_messages(int conv_importance, int inverter_status)
{
	if (INVERTER_FAILURE == inverter_status) {
		if (CONVERGENCE_CRITICAL == conv_importance) {
			if (0 == ctx().myrank)
				printf("\n\n\t\tERROR : inverter failed to converge in a critical region of the code. Exiting now!!\n\n");
			exit(1);
		}
		if (0 == ctx().myrank) printf("\n\t\tWARNING : inverter failed to converge.\n");
	}
}

int staple_last_refinement_iterations(void) { return g_last_refinement_iterations; }

void staple_set_sp_globals(vec3_soa_f *aux1_f, vec3_soa_f *ferm_shiftmulti_acc_f)
{
	g_aux1_f = aux1_f; g_ferm_shiftmulti_acc_f = ferm_shiftmulti_acc_f;
}

// inverter_wrappers.c:117-159
int inverter_wrapper(inverter_package ip, ferm_param *pars, vec3_soa *out, const vec3_soa *in, double res, int max_cg,
										 double shift, int convergence_importance)
{
	int total_iterations = 0, cg_return = 0, temp_conv_check;
	if (inverter_tricks.useMixedPrecision)
		temp_conv_check = inverter_mixed_precision(ip, pars, out, in, res, max_cg, shift, &cg_return);
	else
		temp_conv_check = ker_invert_openacc(ip.u, pars, out, in, res, ip.loc_r, ip.loc_h, ip.loc_s, ip.loc_p, max_cg, shift,
																				 &cg_return);
	convergence_messages(convergence_importance, temp_conv_check);
	total_iterations += cg_return;
	return total_iterations;
}

// inverter_wrappers.c:45-115
int inverter_multishift_wrapper(inverter_package ip, ferm_param *pars, RationalApprox *approx, vec3_soa *out,
																const vec3_soa *in, double res, int max_cg, int convergence_importance)
{
	require_init("inverter_multishift_wrapper");
	int total_iterations = 0, cg_return = 0, temp_conv_check;
	if (inverter_tricks.singlePInvAccelMultiInv) {
		// the reference reads the alloc_vars globals aux1_f / ferm_shiftmulti_acc_f (inverter_wrappers.c:60-71): a host program that
		// defines them (they are weak here) needs no staple_set_sp_globals() call
		vec3_soa_f *const sp_in = g_aux1_f ? g_aux1_f : aux1_f;
		vec3_soa_f *const sp_out = g_ferm_shiftmulti_acc_f ? g_ferm_shiftmulti_acc_f : ferm_shiftmulti_acc_f;
		if (!sp_in || !sp_out) {
			fprintf(stderr, "inverter_multishift_wrapper: singlePInvAccelMultiInv needs staple_set_sp_globals(aux1_f, ferm_shiftmulti_acc_f)\n");
			exit(1);
		}
		convert_double_to_float_vec3_soa(in, sp_in);
		float singlePMultiInvTargetRes = 8e-7f * sqrtf((float) ctx().g.sizeh);
		if (singlePMultiInvTargetRes < res) singlePMultiInvTargetRes = res;
		if (0 == ctx().myrank && verbosity_lv > 3)
			printf("Multishift inverter, single precision, target res %e\n", singlePMultiInvTargetRes);
		temp_conv_check = multishift_invert_f(ip.u_f, pars, approx, sp_out, sp_in, singlePMultiInvTargetRes,
																					ip.loc_r_f, ip.loc_h_f, ip.loc_s_f, ip.loc_p_f, ip.ferm_shift_temp_f, max_cg,
																					&cg_return);
		convergence_messages(convergence_importance, temp_conv_check);
		total_iterations += cg_return;
		const size_t vbytes_f = sizeof(float2) * 3 * ctx().g.sizeh, vbytes_d = sizeof(double2) * 3 * ctx().g.sizeh;
		g_last_refinement_iterations = 0;
		for (int ishift = 0; ishift < approx->approx_order; ishift++) {
			const double bshift = approx->RA_b[ishift];
			printf("Shift %d, %f\n", ishift, bshift);
			vec3_soa_f *src = (vec3_soa_f *) ((char *) sp_out + ishift * vbytes_f);
			vec3_soa *dst = (vec3_soa *) ((char *) out + ishift * vbytes_d);
			convert_float_to_double_vec3_soa(src, dst);
			// literally the reference (inverter_wrappers.c:88-95): inverter_wrapper's return value -- an iteration count -- is what
			// convergence_messages gets as a status, and the count that is accumulated is the multishift solve's stale cg_return;
			// the number of refinement iterations actually spent is kept for staple_last_refinement_iterations()
			temp_conv_check = inverter_wrapper(ip, pars, dst, in, res, max_cg, bshift, convergence_importance);
			g_last_refinement_iterations += temp_conv_check;
			convergence_messages(convergence_importance, temp_conv_check);
			total_iterations += cg_return;
		}
	} else {
		if (0 == ctx().myrank && verbosity_lv > 3) printf("Multishift inverter, DOUBLE precision, target res %e\n", res);
		temp_conv_check = multishift_invert(ip.u, pars, approx, out, in, res, ip.loc_r, ip.loc_h, ip.loc_s, ip.loc_p,
																				ip.ferm_shift_temp, max_cg, &cg_return);
		convergence_messages(convergence_importance, temp_conv_check);
		total_iterations += cg_return;
	}
	return total_iterations;
}

// find_min_max.c:21-60
double ker_find_max_eigenvalue_openacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_p)
{
	require_init("ker_find_max_eigenvalue_openacc");
	int loop_count = 0;
	double norm, inorm, old_norm;
	norm = sqrt(l2norm2_global(loc_p));
	do {
		inorm = 1.0 / norm;
		multiply_fermion_x_doublefactor(loc_p, inorm);
		assign_in_to_out(loc_p, loc_r);
		old_norm = norm;
		fermion_matrix_multiplication(u, loc_p, loc_r, loc_h, pars);
		norm = sqrt(l2norm2_global(loc_p));
		old_norm = fabs(old_norm - norm);
		old_norm /= norm;
		loop_count++;
	} while (old_norm > 1.0e-5);
	return norm;
}

// find_min_max.c:62-98, literally: power iteration on (max - m^2) - Deo Doe (loc_p = start vector); returns max - norm.
// (Its top eigenvalue is ~2 max - 2 m^2, so the value is ~ m^2 - max rather than lambda_min; the reference's
// find_min_max_eigenvalue_soloopenacc does not call it and takes m^2 as the lower bound.  Same numbers as the reference.)
double ker_find_min_eigenvalue_openacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_p, double max)
{
	require_init("ker_find_min_eigenvalue_openacc");
	int loop_count = 0;
	double norm, inorm, old_norm;
	const double m2 = pars->ferm_mass * pars->ferm_mass;
	const double delta = max - m2;
	norm = sqrt(l2norm2_global(loc_p));
	do {
		inorm = 1.0 / norm;
		multiply_fermion_x_doublefactor(loc_p, inorm);
		assign_in_to_out(loc_p, loc_r);
		old_norm = norm;
		// literally the reference's call (:87): (m2 + delta - m2) r - Deo Doe r
		fermion_matrix_multiplication_shifted(u, loc_p, loc_r, loc_h, pars, delta - m2);
		norm = sqrt(l2norm2_global(loc_p));
		old_norm = fabs(old_norm - norm);
		old_norm /= norm;
		loop_count++;
	} while (old_norm > 1.0e-5);
	return max - norm;
}

// find_min_max.c:100-117
void find_min_max_eigenvalue_soloopenacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_p1,
																				 vec3_soa *loc_p2, double *minmax)
{
	(void) loc_p2;
	minmax[0] = pars->ferm_mass * pars->ferm_mass;
	minmax[1] = ker_find_max_eigenvalue_openacc(u, pars, loc_r, loc_h, loc_p1);
}

}   // extern "C"

// Hand-written sm_100a kernels of the staggered hot path and their C entry points:
//   Dirac operator  <- OpenAcc/fermion_matrix.c:47-746 + OpenAcc/matvecmul.h:88-172
//   BLAS-1          <- OpenAcc/fermionic_utilities.c:32-455
//   conversions     <- OpenAcc/float_double_conv.c:9-150
// One thread per output half-lattice site; every SoA stream (6 link entries x 8 links, 8 phases,
// 3 colours of 8 neighbour spinors) is read with 128-bit (FP64) / 64-bit (FP32) loads that are
// contiguous across the warp because idxh is the fastest index of every array.
#include "staple_internal.cuh"
#include <cstring>
#include <ctime>

namespace staple {

// tuning knobs (scripts/tune_dslash.py sweeps them; the defaults are the measured optimum, profiles/)
#ifndef STAPLE_DSLASH_MINBLOCKS
#define STAPLE_DSLASH_MINBLOCKS 7       // 72 registers, 28 warps/SM: +8% over the unconstrained build (profiles/r01_tune_dslash_*.txt)
#endif
// segmented (D3-slab) instantiations exist twice: 6 CTAs/SM (80 registers, no spills to speak of) when the launch has bulk
// slices, 7 CTAs/SM (72 registers, 20-150 B of spills in the face path) when it consists of face slices only -- measured on
// 2 GPUs (profiles/r02e_halo_probe_n2_sentinel.jsonl): 64^3 x 2 per GPU 55.2 vs 57.2 us, 64^3 x 16 per GPU 328 vs 320 us
#ifndef STAPLE_DSLASH_MINBLOCKS_MR
#define STAPLE_DSLASH_MINBLOCKS_MR 6
#endif
#ifndef STAPLE_DSLASH_MINBLOCKS_FACES
#define STAPLE_DSLASH_MINBLOCKS_FACES 7
#endif
#ifndef STAPLE_LINK_LOAD
#define STAPLE_LINK_LOAD 0       // 0: ld.global.cs (evict-first streaming)  1: ld.global.nc  2: ld.global.lu  3: plain
#endif
constexpr int kBlock = kDslashBlock;      // dslash CTA size (staple_internal.cuh)
constexpr int kBlasBlock = 256;

// ------------------------------------------------------------------ small complex helpers
template <typename T> __device__ __forceinline__ cplx_t<T> mk(T x, T y);
template <> __device__ __forceinline__ double2 mk<double>(double x, double y) { return make_double2(x, y); }
template <> __device__ __forceinline__ float2 mk<float>(float x, float y) { return make_float2(x, y); }

// streaming (read-once) loads for links/phases, cached read-only loads for spinors
template <typename V> __device__ __forceinline__ V ld_stream(const V *p)
{
#if STAPLE_LINK_LOAD == 0
	return __ldcs(p);
#elif STAPLE_LINK_LOAD == 1
	return __ldg(p);
#elif STAPLE_LINK_LOAD == 2
	return __ldlu(p);
#else
	return *p;
#endif
}
template <typename V> __device__ __forceinline__ V ld_cached(const V *p) { return __ldg(p); }

template <typename C> __device__ __forceinline__ C cmul(C a, C b)   // a*b
{ C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> __device__ __forceinline__ void cfma(C &acc, C a, C b)   // acc += a*b
{ acc.x += a.x * b.x; acc.x -= a.y * b.y; acc.y += a.x * b.y; acc.y += a.y * b.x; }
template <typename C> __device__ __forceinline__ void cfma_ca(C &acc, C a, C b)   // acc += conj(a)*b
{ acc.x += a.x * b.x; acc.x += a.y * b.y; acc.y += a.x * b.y; acc.y -= a.y * b.x; }
template <typename C> __device__ __forceinline__ C cross(C a, C b, C c, C d)   // a*b - c*d
{ C r; r.x = a.x * b.x - a.y * b.y - (c.x * d.x - c.y * d.y); r.y = a.x * b.y + a.y * b.x - (c.x * d.y + c.y * d.x); return r; }

// sin/cos of a link phase angle (matvecmul.h:96-97 `cos(arg)+I*sin(arg)`).  calc_u1_phases folds every
// angle into (-pi, pi] (backfield.c:163-187), so a two-term Cody-Waite reduction by pi/2 (exact for
// |k| <= 2 with FMA) plus the classic degree-13/14 minimax kernels on [-pi/4, pi/4] is enough: <= 1 ulp
// from glibc's sin/cos over the whole range (checked on 2e7 samples), sin(fl(pi)) = 1.2246e-16 reproduced.
// Unlike sincos() there is no Payne-Hanek slow-path CALL, so the operator stays one basic block and the
// scheduler can overlap the next hop's loads with this hop's arithmetic.
__device__ __forceinline__ void sincos_t(double x, double *s, double *c)
{
	const double kd = rint(x * 0.63661977236758134308);
	const int k = (int) kd;
	double r = fma(-kd, 1.57079632679489655800e+00, x);
	r = fma(-kd, 6.12323399573676603587e-17, r);
	const double z = r * r;
	double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
	ps = fma(z, ps, 2.75573137070700676789e-06);
	ps = fma(z, ps, -1.98412698298579493134e-04);
	ps = fma(z, ps, 8.33333333332248946124e-03);
	const double ks = fma(z * r, fma(z, ps, -1.66666666666666324348e-01), r);
	double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
	pc = fma(z, pc, -2.75573143513906633035e-07);
	pc = fma(z, pc, 2.48015872894767294178e-05);
	pc = fma(z, pc, -1.38888888888741095749e-03);
	pc = fma(z, pc, 4.16666666666666019037e-02);
	const double kc = 1.0 - fma(0.5, z, -(z * (z * pc)));
	const double a = (k & 1) ? kc : ks, b = (k & 1) ? ks : kc;
	*s = (k & 2) ? -a : a;
	*c = ((k + 1) & 2) ? -b : b;
}
__device__ __forceinline__ void sincos_t(float x, float *s, float *c)
{
	const float kd = rintf(x * 0.63661977236758134308f);
	const int k = (int) kd;
	float r = fmaf(-kd, 1.5707963705062866f, x);
	r = fmaf(-kd, -4.371139000186243e-08f, r);
	const float z = r * r;
	float ps = fmaf(z, 2.7183114939898219064e-6f, -1.98393348360966317347e-4f);
	ps = fmaf(z, ps, 8.3333293858894631756e-3f);
	ps = fmaf(z, ps, -0.166666666416265235595f);
	const float ks = fmaf(z * r, ps, r);
	float pc = fmaf(z, 2.43904487962774090654e-5f, -1.38867637746099294692e-3f);
	pc = fmaf(z, pc, 4.16666233237390631894e-2f);
	pc = fmaf(z, pc, -0.499999997251031003120f);
	const float kc = fmaf(z, pc, 1.0f);
	const float a = (k & 1) ? kc : ks, b = (k & 1) ? ks : kc;
	*s = (k & 2) ? -a : a;
	*c = ((k + 1) & 2) ? -b : b;
}

// One hop: acc += U(im) e^{i th(im)} v(iv)            (DAG=false, matvecmul.h:88-126)
//          acc -= U(im)^+ e^{-i th(im)} v(iv)         (DAG=true,  matvecmul.h:129-172)
// uk -> u[k].r0.c0, phk -> backfield[k].d ; third row rebuilt as conj(r0 x r1).
template <typename T, bool DAG>
__device__ __forceinline__ void hop(cplx_t<T> acc[3], const cplx_t<T> *__restrict__ uk, const T *__restrict__ phk,
																		unsigned int im, const cplx_t<T> *__restrict__ in, unsigned int iv, long n)
{
	using C = cplx_t<T>;
	const T th = ld_stream(phk + im);
	const C m00 = ld_stream(uk + im), m01 = ld_stream(uk + n + im), m02 = ld_stream(uk + 2 * n + im);
	const C m10 = ld_stream(uk + 3 * n + im), m11 = ld_stream(uk + 4 * n + im), m12 = ld_stream(uk + 5 * n + im);
	const C v0 = ld_cached(in + iv), v1 = ld_cached(in + n + iv), v2 = ld_cached(in + 2 * n + iv);
	T s, c;
	sincos_t(th, &s, &c);
	// x = r0 x r1 (unconjugated); third row of U is conj(x)
	const C x0 = cross(m01, m12, m02, m11);
	const C x1 = cross(m02, m10, m00, m12);
	const C x2 = cross(m00, m11, m01, m10);
	if (!DAG) {
		const C p = mk<T>(c, s);
		const C w0 = cmul(v0, p), w1 = cmul(v1, p), w2 = cmul(v2, p);
		cfma(acc[0], m00, w0); cfma(acc[0], m01, w1); cfma(acc[0], m02, w2);
		cfma(acc[1], m10, w0); cfma(acc[1], m11, w1); cfma(acc[1], m12, w2);
		cfma_ca(acc[2], x0, w0); cfma_ca(acc[2], x1, w1); cfma_ca(acc[2], x2, w2);
	} else {
		const C p = mk<T>(-c, s);   // -conj(phase): the backward hops enter with a minus sign
		const C w0 = cmul(v0, p), w1 = cmul(v1, p), w2 = cmul(v2, p);
		cfma_ca(acc[0], m00, w0); cfma_ca(acc[0], m10, w1); cfma(acc[0], x0, w2);
		cfma_ca(acc[1], m01, w0); cfma_ca(acc[1], m11, w1); cfma(acc[1], x1, w2);
		cfma_ca(acc[2], m02, w0); cfma_ca(acc[2], m12, w1); cfma(acc[2], x2, w2);
	}
}

// ------------------------------------------------------------------ deterministic grid reduction
// Block sums go to partials[index]; the last block to arrive (ticket) adds all of them in a fixed
// order, so the result does not depend on block scheduling.  NV = number of values (1 or 2).
template <int NV>
__device__ __forceinline__ void block_sum(double v[NV], double *sm)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
	for (int k = 0; k < NV; k++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
	}
	__syncthreads();   // sm may still be in use by a previous call
	if (lane == 0)
		for (int k = 0; k < NV; k++) sm[k * 32 + warp] = v[k];
	__syncthreads();
	if (warp == 0) {
#pragma unroll
		for (int k = 0; k < NV; k++) {
			double x = lane < nwarp ? sm[k * 32 + lane] : 0.0;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
			v[k] = x;
		}
	}
}

// Returns true (block-uniform) in the block that completed the sum; result[] is then final and, after the
// __syncthreads() the caller issues, visible to every thread of that block.
template <int NV>
__device__ __forceinline__ bool grid_sum_finalize(double v[NV], double *partials, long pstride, unsigned int *ticket,
																									double *result, unsigned int target, unsigned int index)
{
	__shared__ double sm[NV * 32];
	__shared__ bool last;
	block_sum<NV>(v, sm);
	if (threadIdx.x == 0) {
		for (int k = 0; k < NV; k++) partials[k * pstride + index] = v[k];
		__threadfence();
		const unsigned int t = atomicAdd(ticket, 1u);
		last = (t == target - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		// fixed summation order (thread t adds partials t, t+B, t+2B, ... in that order), but the loads are
		// issued 8 at a time: a dependent-latency loop here costs ~0.4 us per trip on the kernel's tail
		double s[NV];
		for (int k = 0; k < NV; k++) {
			s[k] = 0.0;
			for (unsigned int i0 = threadIdx.x; i0 < target; i0 += 8 * blockDim.x) {
				double x[8];
#pragma unroll
				for (int j = 0; j < 8; j++) {
					const unsigned int i = i0 + j * blockDim.x;
					x[j] = i < target ? __ldcg(partials + k * pstride + i) : 0.0;
				}
#pragma unroll
				for (int j = 0; j < 8; j++) s[k] += x[j];
			}
		}
		block_sum<NV>(s, sm);
		if (threadIdx.x == 0) {
			for (int k = 0; k < NV; k++) result[k] = s[k];
			*ticket = 0u;
		}
	}
	return last;
}

// ------------------------------------------------------------------ peer-memory halo channel (device side)
// push both interior faces of a vector (3 colour arrays) into the neighbours' staging slots (standalone
// communicate_fermion_borders; the operator's face blocks do this themselves).  grid = 2*nfb CTAs of kDslashBlock threads:
// [0,nfb) TOP interior slice -> rank R's slot 0, [nfb,2nfb) BOTTOM interior slice -> rank L's slot 1
template <typename C>
__global__ void __launch_bounds__(kDslashBlock) p2p_push_kernel(const C *src, long n, long top_lo, long bot_lo, unsigned int vol3h,
																																 unsigned int nfb, C *peer_top, C *peer_bot, long parity_stride,
																																 const unsigned long long seq)
{
	const bool bot = blockIdx.x >= nfb;
	const unsigned int j = bot ? blockIdx.x - nfb : blockIdx.x, t = j * kDslashBlock + threadIdx.x;
	C *peer = (bot ? peer_bot : peer_top) + (seq & 1ull) * parity_stride;
	const long lo = bot ? bot_lo : top_lo;
	if (t < vol3h) {
		C v[3];
#pragma unroll
		for (int c = 0; c < 3; c++) v[c] = src[c * n + lo + t];
#pragma unroll
		for (int c = 0; c < 3; c++) peer[(long) c * vol3h + t] = stageable(v[c]);
	}
}

// copy of one staged slice (3 colour arrays) into a halo slice: grid-stride, 4 sites x 3 colours = 12 independent 16-byte
// elements in flight per thread, each taken as soon as it has arrived.  Not inlined in the operator: its registers must not
// weigh on the register budget of the hops.
template <typename C>
__device__ __noinline__ void unpack_slice(C *dst, long n, C *src, unsigned int vol3h, unsigned int first, unsigned int stride)
{
	for (unsigned int t0 = first; t0 < vol3h; t0 += 4 * stride) {
		C v[4][3];
		bool ok[4][3];
		// all twelve loads first (normally everything has landed long ago: one L2 round trip instead of twelve in a row) ...
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const unsigned int tt = t0 + k * stride;
#pragma unroll
			for (int c = 0; c < 3; c++) {
				ok[k][c] = true;
				if (tt < vol3h) v[k][c] = peek_staged(src + (long) c * vol3h + tt, &ok[k][c]);
			}
		}
		// ... then the stragglers one by one, the resting pattern back into the staging area, and the halo slice
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const unsigned int tt = t0 + k * stride;
			if (tt < vol3h) {
#pragma unroll
				for (int c = 0; c < 3; c++) {
					C *p = src + (long) c * vol3h + tt;
					if (!ok[k][c]) v[k][c] = take_staged(p);
					else reset_staged(p);
					dst[c * n + tt] = v[k][c];
				}
			}
		}
	}
}

// standalone unpack: copy both staging slots of exchange `seq` into the halo slices as the neighbours' data lands.
// grid = 2*nb CTAs of kDslashBlock threads.
template <typename C>
__global__ void __launch_bounds__(kDslashBlock) p2p_unpack_kernel(C *dst, long n, long lower_lo, long upper_lo, unsigned int vol3h,
																																	 C *slot0, C *slot1, long parity_stride, const unsigned long long seq, const int *skip)
{
	if (skip != nullptr && *skip != 0) return;       // the producers skipped this exchange too (same flag on every rank)
	const unsigned int nb = gridDim.x / 2;
	const bool hi = blockIdx.x >= nb;                // first half of the grid: lower halo, second half: upper
	unpack_slice<C>(dst + (hi ? upper_lo : lower_lo), n, (hi ? slot1 : slot0) + (seq & 1ull) * parity_stride, vol3h,
									(hi ? blockIdx.x - nb : blockIdx.x) * kDslashBlock + threadIdx.x, nb * kDslashBlock);
}

// unpack blocks per halo: at most 2 per SM (grid-stride copy with 12 elements in flight per thread).  Thousands of
// 128-thread blocks only add scheduling time to the tail; 74 were measured too few to cover the HBM latency.
static inline unsigned int unpack_blocks_for(unsigned int face_blocks) { return face_blocks < 296u ? face_blocks : 296u; }

template <typename C>
static void p2p_unpack_t(void *base, cudaStream_t s, const int *skip)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	P2P &p = c.p2p;
	const long lower_lo = (long) (g.d3_halo - 1) * g.vol3h, upper_lo = (long) (g.d3_halo + g.loc_n3) * g.vol3h;
	const unsigned int nb = unpack_blocks_for((unsigned int) p.nfb);
	p2p_unpack_kernel<C><<<2 * nb, kDslashBlock, 0, s>>>((C *) base, g.sizeh, lower_lo, upper_lo, (unsigned int) g.vol3h,
		(C *) p.stage, (C *) (p.stage + p.slot_bytes), (long) (2 * p.slot_bytes / sizeof(C)), p.h_seq + 1, skip);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	p.h_seq += 1;                                    // the exchange whose faces were pushed before this call is complete
	count_launch();
}
void p2p_unpack(void *base, size_t elem_bytes, cudaStream_t s, const int *skip)
{
	if (elem_bytes == 16) p2p_unpack_t<double2>(base, s, skip);
	else p2p_unpack_t<float2>(base, s, skip);
}

template <typename C>
static void p2p_exchange_t(void *base, cudaStream_t s)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	P2P &p = c.p2p;
	const long top_lo = (long) (g.d3_halo + g.loc_n3 - 1) * g.vol3h, bot_lo = (long) g.d3_halo * g.vol3h;
	const long ps = (long) (2 * p.slot_bytes / sizeof(C));
	// top interior slice -> rank R's lower halo (its slot 0); bottom interior slice -> rank L's upper halo (slot 1)
	p2p_push_kernel<C><<<2 * (unsigned int) p.nfb, kDslashBlock, 0, s>>>((const C *) base, g.sizeh, top_lo, bot_lo, (unsigned int) g.vol3h,
		(unsigned int) p.nfb, (C *) p.stage_R, (C *) (p.stage_L + p.slot_bytes), ps, p.h_seq + 1);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
	p2p_unpack(base, sizeof(C), s, nullptr);
}
void p2p_exchange_fermion(void *base, size_t elem_bytes, cudaStream_t s)
{
	if (elem_bytes == 16) p2p_exchange_t<double2>(base, s);
	else p2p_exchange_t<float2>(base, s);
}

// ---- all-reduce of one or two doubles through the mailboxes (device side: staple_internal.cuh)
__global__ void p2p_allreduce_kernel(double *vals, int nd, RedView v) { p2p_allreduce_warp(vals, nd, v); }

RedView make_redview()
{
	Ctx &c = ctx();
	P2P &p = c.p2p;
	RedView v;
	v.nranks = c.loopback ? 1 : c.nranks; v.myrank = c.myrank; v.q = p.d_redq;
	for (int r = 0; r < kMaxRanks; r++) v.box[r] = nullptr;
	for (int r = 0; r < v.nranks; r++) v.box[r] = (double *) (p.peer_mailbox[r] + kMailboxRedBox);
	return v;
}
void p2p_allreduce(double *vals, int ndoubles, cudaStream_t s)
{
	p2p_allreduce_kernel<<<1, 32, 0, s>>>(vals, ndoubles, make_redview());
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}

void set_spin_timeout_kernels(unsigned long long ns)
{
	STAPLE_CUDA_CHECK(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof(ns)));
}

bool halo_lazy_ok()
{
	const Ctx &c = ctx();
	return c.nranks > 1 && c.p2p.on && c.p2p_single_launch && c.p2p_unpack_in_kernel && c.p2p_lazy;
}
PushView make_pushview(bool on)
{
	const Ctx &c = ctx();
	const Geom &g = c.g;
	const P2P &p = c.p2p;
	PushView v;
	v.on = on ? 1 : 0;
	v.peer_top = p.stage_R; v.peer_bot = p.stage_L ? p.stage_L + p.slot_bytes : nullptr;
	v.seq = p.h_seq + 1; v.parity_bytes = (long) (2 * p.slot_bytes);
	v.top_lo = (long) (g.d3_halo + g.loc_n3 - 1) * g.vol3h; v.bot_lo = (long) g.d3_halo * g.vol3h; v.vol3h = g.vol3h;
	return v;
}
// both interior faces of a vector into the neighbours' staging slots as a new exchange, without unpacking what arrives here:
// the next operator consumes it staged
void p2p_push_faces(const void *base, size_t elem_bytes, cudaStream_t s)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	P2P &p = c.p2p;
	const long top_lo = (long) (g.d3_halo + g.loc_n3 - 1) * g.vol3h, bot_lo = (long) g.d3_halo * g.vol3h;
	if (elem_bytes == 16)
		p2p_push_kernel<double2><<<2 * (unsigned int) p.nfb, kDslashBlock, 0, s>>>((const double2 *) base, g.sizeh, top_lo, bot_lo, (unsigned int) g.vol3h,
			(unsigned int) p.nfb, (double2 *) p.stage_R, (double2 *) (p.stage_L + p.slot_bytes), (long) (2 * p.slot_bytes / 16), p.h_seq + 1);
	else
		p2p_push_kernel<float2><<<2 * (unsigned int) p.nfb, kDslashBlock, 0, s>>>((const float2 *) base, g.sizeh, top_lo, bot_lo, (unsigned int) g.vol3h,
			(unsigned int) p.nfb, (float2 *) p.stage_R, (float2 *) (p.stage_L + p.slot_bytes), (long) (2 * p.slot_bytes / 8), p.h_seq + 1);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	p.h_seq += 1;
	count_launch();
}

// ------------------------------------------------------------------ Dirac operator kernel
// Fused Re(in0 . out): block partial -> deterministic grid sum; inside a CG-M solve the block that completes the
// sum (alpha) also runs the recurrences that consume it, so no one-warp kernel sits between M^+M and the update
template <typename T>
__device__ __forceinline__ void dslash_finish_dot(const DslashArgs<T> &a, double dot)
{
	double v[1] = { dot };
	const bool last = grid_sum_finalize<1>(v, a.partials, 0, a.ticket, a.result, a.ticket_target, a.partial_offset + blockIdx.x);
	if (last && (a.cgm != nullptr || a.cg != nullptr)) {
		__syncthreads();
		if (threadIdx.x < 32) {
			if (a.cgm != nullptr) cgm_after_alpha_warp(a.cgm, a.result, a.cgm_red);
			else cg_after_alpha_warp(a.cg, a.result, a.cgm_red);
		}
	}
}

// One d3 hop.  The spinor comes from (vin, colour stride vn, index iv), loaded past L1 (neighbours in d3 are vol3h sites away,
// L1 has nothing to offer them) -- or, STAGED, from the local staging area that a peer GPU fills over NVLink: each element is
// taken as soon as it has arrived (take_staged: the data is its own flag), which is the only wait of the whole exchange.
template <typename T, bool DAG, bool MAYBE_STAGED>
__device__ __forceinline__ void hop_d3(cplx_t<T> acc[3], const cplx_t<T> *__restrict__ uk, const T *__restrict__ phk,
																			 unsigned int im, cplx_t<T> *vin, unsigned int iv, long n, long vn, const bool STAGED)
{
	using C = cplx_t<T>;
	const T th = ld_stream(phk + im);
	const C m00 = ld_stream(uk + im), m01 = ld_stream(uk + n + im), m02 = ld_stream(uk + 2 * n + im);
	const C m10 = ld_stream(uk + 3 * n + im), m11 = ld_stream(uk + 4 * n + im), m12 = ld_stream(uk + 5 * n + im);
	C v0, v1, v2;
	if (MAYBE_STAGED && STAGED) {
		bool a0, a1, a2;           // three loads in flight (normally everything landed long ago); else the slow path, one by one
		v0 = peek_staged(vin + iv, &a0); v1 = peek_staged(vin + vn + iv, &a1); v2 = peek_staged(vin + 2 * vn + iv, &a2);
		if (a0 && a1 && a2) { reset_staged(vin + iv); reset_staged(vin + vn + iv); reset_staged(vin + 2 * vn + iv); }
		else { v0 = take_staged(vin + iv); v1 = take_staged(vin + vn + iv); v2 = take_staged(vin + 2 * vn + iv); }
	}
	else { v0 = __ldcg(vin + iv); v1 = __ldcg(vin + vn + iv); v2 = __ldcg(vin + 2 * vn + iv); }
	T s, c;
	sincos_t(th, &s, &c);
	const C x0 = cross(m01, m12, m02, m11);
	const C x1 = cross(m02, m10, m00, m12);
	const C x2 = cross(m00, m11, m01, m10);
	if (!DAG) {
		const C p = mk<T>(c, s);
		const C w0 = cmul(v0, p), w1 = cmul(v1, p), w2 = cmul(v2, p);
		cfma(acc[0], m00, w0); cfma(acc[0], m01, w1); cfma(acc[0], m02, w2);
		cfma(acc[1], m10, w0); cfma(acc[1], m11, w1); cfma(acc[1], m12, w2);
		cfma_ca(acc[2], x0, w0); cfma_ca(acc[2], x1, w1); cfma_ca(acc[2], x2, w2);
	} else {
		const C p = mk<T>(-c, s);
		const C w0 = cmul(v0, p), w1 = cmul(v1, p), w2 = cmul(v2, p);
		cfma_ca(acc[0], m00, w0); cfma_ca(acc[0], m10, w1); cfma(acc[0], x0, w2);
		cfma_ca(acc[1], m01, w0); cfma_ca(acc[1], m11, w1); cfma(acc[1], x1, w2);
		cfma_ca(acc[2], m02, w0); cfma_ca(acc[2], m12, w1); cfma(acc[2], x2, w2);
	}
}

// Hop order: the six hops in directions 0,1,2 first, the two d3 hops last.  On D3 slabs those are the only hops that can
// need a neighbour rank's data, so a face site whose input halo is staged has three quarters of its work done before it
// looks at the staging area.  (The reference adds backward 0..3 then forward 0..3, fermion_matrix.c:74-89; the different
// summation order moves results by O(1e-16) relative, like FMA contraction does.)
// FACE = false: site lo + t of a plain range -- the single-GPU kernel and the bulk blocks of a segmented launch
// FACE = true : face site; side 1 = TOP slice (result also stored into rank R's staging slot; with staged input its FORWARD d3
//               neighbour is staged), side 2 = BOTTOM slice (result also to rank L's slot; staged input: BACKWARD neighbour)
// returns the site's contribution to the fused Re(in0 . out)
template <typename T, int PAR, int EPI, bool FACE>
__device__ __forceinline__ double dslash_site(const DslashArgs<T> &a, const unsigned int lo, const unsigned int t,
																							cplx_t<T> *peer, cplx_t<T> *stage, const int side, const bool in_staged)
{
	using C = cplx_t<T>;
	const long n = a.sizeh;
	const long un = 9 * n;
	const unsigned int idx = lo + t;
	const unsigned int nd0h = a.nd0h, nd1 = a.nd1, nd2 = a.nd2, nd3 = a.nd3;
	const unsigned int hd0 = idx % nd0h;
	unsigned int q = idx / nd0h;
	const unsigned int d1 = q % nd1; q /= nd1;
	const unsigned int d2 = q % nd2;
	const unsigned int d3 = q / nd2;
	// d0 = 2*hd0 + rp  (fermion_matrix.c:64, :120)
	const unsigned int rp = (d1 + d2 + d3 + PAR) & 1u;
	const unsigned int s1 = nd0h, s2 = nd0h * nd1, s3 = (unsigned int) a.vol3h;
	const unsigned int i0m = rp ? idx : (hd0 == 0 ? idx + (nd0h - 1) : idx - 1);
	const unsigned int i0p = rp ? (hd0 == nd0h - 1 ? idx - (nd0h - 1) : idx + 1) : idx;
	const unsigned int i1m = d1 == 0 ? idx + s1 * (nd1 - 1) : idx - s1;
	const unsigned int i1p = d1 == nd1 - 1 ? idx - s1 * (nd1 - 1) : idx + s1;
	const unsigned int i2m = d2 == 0 ? idx + s2 * (nd2 - 1) : idx - s2;
	const unsigned int i2p = d2 == nd2 - 1 ? idx - s2 * (nd2 - 1) : idx + s2;
	C acc[3];
	acc[0] = mk<T>(0, 0); acc[1] = mk<T>(0, 0); acc[2] = mk<T>(0, 0);
	// backward hops: link and phase of the OTHER parity at the neighbour index (:74-77, :130-133)
	hop<T, true>(acc, a.u + (1 - PAR) * un, a.ph + (1 - PAR) * n, i0m, a.in, i0m, n);
	hop<T, true>(acc, a.u + (3 - PAR) * un, a.ph + (3 - PAR) * n, i1m, a.in, i1m, n);
	hop<T, true>(acc, a.u + (5 - PAR) * un, a.ph + (5 - PAR) * n, i2m, a.in, i2m, n);
	// forward hops: link and phase of this parity at the own index (:86-89, :144-147)
	hop<T, false>(acc, a.u + (0 + PAR) * un, a.ph + (0 + PAR) * n, idx, a.in, i0p, n);
	hop<T, false>(acc, a.u + (2 + PAR) * un, a.ph + (2 + PAR) * n, idx, a.in, i1p, n);
	hop<T, false>(acc, a.u + (4 + PAR) * un, a.ph + (4 + PAR) * n, idx, a.in, i2p, n);
	const unsigned int i3m = d3 == 0 ? idx + s3 * (nd3 - 1) : idx - s3;
	const unsigned int i3p = d3 == nd3 - 1 ? idx - s3 * (nd3 - 1) : idx + s3;
	C *const vec = const_cast<C *>(a.in);
	const bool sm = FACE && in_staged && side == 2, sp = FACE && in_staged && side == 1;
	hop_d3<T, true, FACE>(acc, a.u + (7 - PAR) * un, a.ph + (7 - PAR) * n, i3m, sm ? stage : vec, sm ? t : i3m, n, sm ? (long) s3 : n, sm);
	hop_d3<T, false, FACE>(acc, a.u + (6 + PAR) * un, a.ph + (6 + PAR) * n, idx, sp ? stage : vec, sp ? t : i3p, n, sp ? (long) s3 : n, sp);
	double dot = 0.0;
#pragma unroll
	for (int c = 0; c < 3; c++) {
		C o = mk<T>(acc[c].x * (T) 0.5, acc[c].y * (T) 0.5);   // :94-96
		if (EPI != EPI_NONE) {
			// fused combine_in1xferm_mass_minus_in2 (fermionic_utilities.c:261-272): double factor
			const C x = ld_cached(a.in0 + c * n + idx);
			o = mk<T>((T) ((double) x.x * a.m2 - (double) o.x), (T) ((double) x.y * a.m2 - (double) o.y));
			if (EPI == EPI_MASS_DOT) dot += (double) x.x * (double) o.x + (double) x.y * (double) o.y;
		}
		a.out[c * n + idx] = o;
		if (!FACE && EPI == EPI_NONE && a.out_host != nullptr) a.out_host[c * n + idx] = o;   // posted PCIe writes, 512 contiguous bytes per warp
		if (FACE && peer != nullptr) peer[(long) c * s3 + t] = stageable(o);     // posted NVLink store into the neighbour's staging slot
	}
	return dot;
}

// MR = false: plain launch over [site_lo, site_lo + nsites) -- the single-GPU kernel.
// MR = true : segmented launch on D3 slabs, block-uniform roles by block index (DslashArgs):
//             [top face][bottom face][bulk][unpack]
// No block ever waits for another block of the same launch, and no block has a tail: the exchange number arrives by value
// (a per-block fence + ticket to advance a device counter was measured to cost 12 % of a launch on the ~16k bulk blocks, and
// after NVLink stores the fence waits for the remote acknowledgements).
template <typename T, int PAR, int EPI, bool MR, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) dslash_kernel(const DslashArgs<T> a)
{
	using C = cplx_t<T>;
	if (a.skip != nullptr && *a.skip != 0) return;
	// sizeh < 2^31 is checked at staple_init_geometry: site indices are 32-bit (no 64-bit divisions),
	// only the array bases (k*9*sizeh) are 64-bit
	double dot = 0.0;
	if (!MR) {
		const unsigned int t = blockIdx.x * kBlock + threadIdx.x;
		if (t < (unsigned int) a.nsites) dot = dslash_site<T, PAR, EPI, false>(a, (unsigned int) a.site_lo, t, nullptr, nullptr, 0, false);
	} else {
		// block order: [top face][bottom face][unpack, if early][bulk][unpack, if late]
		const unsigned int b = blockIdx.x, b2 = a.nb_top, b3 = b2 + a.nb_bot, bk = b3 + (a.unpack_early ? 2 * a.nb_unpack : 0), b4 = bk + a.nb_bulk;
		if (b >= bk && b < b4) {
			const unsigned int t = (b - bk) * kBlock + threadIdx.x;
			if (t < (unsigned int) a.nsites) dot = dslash_site<T, PAR, EPI, false>(a, (unsigned int) a.site_lo, t, nullptr, nullptr, 0, false);
		} else {
			const unsigned long long cur = a.cur;           // this launch consumes exchange `cur` (staged input halos) and produces cur + 1
			const unsigned int vol3h = (unsigned int) a.vol3h;
			if (b < b3) {
				// TOP face   : -> rank R's slot 0; staged input: forward neighbour slice = upper halo = local slot 1
				// BOTTOM face: -> rank L's slot 1; staged input: backward neighbour slice = lower halo = local slot 0
				const bool bot = b >= b2;
				const unsigned int t = (bot ? b - b2 : b) * kBlock + threadIdx.x;
				C *peer = a.peer_top == nullptr ? nullptr : (bot ? a.peer_bot : a.peer_top) + ((cur + 1) & 1ull) * a.parity_stride;
				C *stage = (bot ? a.stage_lo : a.stage_hi) + (cur & 1ull) * a.parity_stride;
				if (t < vol3h) dot = dslash_site<T, PAR, EPI, true>(a, (unsigned int) (bot ? a.bot_lo : a.top_lo), t, peer, stage, bot ? 2 : 1, a.in_staged != 0);
			} else {
				// ---- unpack blocks: the halos of THIS exchange, element by element as they land.  With bulk slices to hide behind, a FEW
				// blocks placed right after the faces (they sit out the neighbours' face computation, then copy while the bulk runs: no
				// tail); without, many blocks at the end of the launch
				const unsigned int ub = a.nb_unpack, k = a.unpack_early ? b - b3 : b - b4;
				const bool hi = k >= ub;                         // first ub blocks: lower halo (slot 0, from rank L); then upper (slot 1, from R)
				unpack_slice<C>(a.out + (hi ? a.upper_lo : a.lower_lo), a.sizeh, (hi ? a.stage_hi : a.stage_lo) + ((cur + 1) & 1ull) * a.parity_stride,
												vol3h, (hi ? k - ub : k) * kBlock + threadIdx.x, ub * kBlock);
			}
		}
	}
	if (EPI == EPI_MASS_DOT && a.partials != nullptr) dslash_finish_dot(a, dot);   // blocks without sites add a zero partial: the count stays gridDim.x
}

unsigned int dslash_blocks(int d3lo, int d3hi)
{
	const long nsites = (long) (d3hi - d3lo) * ctx().g.vol3h;
	return (unsigned int) ((nsites + kBlock - 1) / kBlock);
}

template <typename T>
void launch_dslash(int par, int epi, const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, const T *ph,
									 const cplx_t<T> *in0, double m2, int d3lo, int d3hi, int dot_slot,
									 unsigned int ticket_target, unsigned int partial_offset, const int *skip, cudaStream_t s,
									 int face, int halo)
{
	using C = cplx_t<T>;
	const Geom &g = ctx().g;
	if (d3hi <= d3lo) return;
	DslashArgs<T> a;
	memset(&a, 0, sizeof(a));
	a.u = u; a.out = out; a.in = in; a.ph = ph; a.in0 = in0; a.m2 = m2;
	a.out_host = (C *) ctx().out_host_hook;
	a.partials = dot_slot >= 0 ? partials(dot_slot) : nullptr;
	a.ticket = dot_slot >= 0 ? ticket(dot_slot) : nullptr;
	a.result = dot_slot >= 0 ? result(dot_slot) : nullptr;
	a.ticket_target = ticket_target; a.partial_offset = partial_offset; a.skip = skip;
	a.cgm = (epi == EPI_MASS_DOT) ? ctx().cgm_hook : nullptr;
	a.cg = (epi == EPI_MASS_DOT) ? ctx().cg_hook : nullptr;
	a.cgm_red = ctx().cgm_hook_red;
	a.site_lo = (long) d3lo * g.vol3h; a.nsites = (long) (d3hi - d3lo) * g.vol3h;
	a.nd0h = g.nd0h; a.nd1 = g.nd1; a.nd2 = g.nd2; a.nd3 = g.nd3; a.vol3h = g.vol3h; a.sizeh = g.sizeh;
	unsigned int grid = dslash_blocks(d3lo, d3hi);
	if (face != FACE_NONE) {
		// top interior slice -> rank R's slot 0 (its lower halo); bottom interior slice -> rank L's slot 1
		P2P &p = ctx().p2p;
		const unsigned int nfb = (unsigned int) p.nfb;
		a.mr = 1;
		a.cur = p.h_seq; a.parity_stride = (long) (2 * p.slot_bytes / sizeof(C));
		a.peer_top = (C *) p.stage_R; a.peer_bot = (C *) (p.stage_L + p.slot_bytes);
		a.stage_lo = (C *) p.stage; a.stage_hi = (C *) (p.stage + p.slot_bytes);
		a.lower_lo = (long) (g.d3_halo - 1) * g.vol3h; a.upper_lo = (long) (g.d3_halo + g.loc_n3) * g.vol3h;
		a.top_lo = (long) (d3hi - 1) * g.vol3h; a.bot_lo = (long) d3lo * g.vol3h;
		if (face == FACE_TOP) a.nb_top = nfb;                        // single-slice launches of the three-queue form
		else if (face == FACE_BOTTOM) a.nb_bot = nfb;
		else {
			a.nb_top = a.nb_bot = nfb;
			a.nb_bulk = dslash_blocks(d3lo + 1, d3hi - 1);
			a.site_lo = (long) (d3lo + 1) * g.vol3h; a.nsites = (long) (d3hi - d3lo - 2) * g.vol3h;   // bulk
			if (face == FACE_BOTH_UNPACK) {
				// enough bulk (>= 4 slices) to cover the neighbours' face computation and the copy: one unpack CTA per halo and two SMs
				a.unpack_early = (d3hi - d3lo) >= 6 ? 1 : 0;
				a.nb_unpack = a.unpack_early ? (nfb < 74u ? nfb : 74u) : unpack_blocks_for(nfb);
			}
		}
		a.in_staged = (halo & HALO_IN_STAGED) ? 1 : 0;
		if (halo & HALO_NO_PUSH) a.peer_top = a.peer_bot = nullptr;
		if (halo & HALO_ADVANCE) p.h_seq += 1;       // (a.cur was taken above) both faces of exchange cur + 1 are pushed by this launch
		grid = a.nb_top + a.nb_bot + a.nb_bulk + 2 * a.nb_unpack;
	}
	if (dot_slot >= 0 && (long) partial_offset + grid > ctx().max_partials) {
		fprintf(stderr, "libstaple_b200: FATAL: reduction scratch too small (%u + %u partials > %ld)\n", partial_offset, grid, ctx().max_partials);
		exit(1);
	}
#define STAPLE_LAUNCH(P, E) do { \
		if (!a.mr) dslash_kernel<T, P, E, false, STAPLE_DSLASH_MINBLOCKS><<<grid, kBlock, 0, s>>>(a); \
		else if (a.nb_bulk != 0) dslash_kernel<T, P, E, true, STAPLE_DSLASH_MINBLOCKS_MR><<<grid, kBlock, 0, s>>>(a); \
		else dslash_kernel<T, P, E, true, STAPLE_DSLASH_MINBLOCKS_FACES><<<grid, kBlock, 0, s>>>(a); } while (0)
	// the mass epilogue without the dot product runs the EPI_MASS_DOT instantiation with the reduction switched off
	// (a.partials == nullptr): one code path less, and that instantiation fits the 72-register budget without spills
	if (par == 0) {
		if (epi == EPI_NONE) STAPLE_LAUNCH(0, EPI_NONE);
		else STAPLE_LAUNCH(0, EPI_MASS_DOT);
	} else {
		if (epi == EPI_NONE) STAPLE_LAUNCH(1, EPI_NONE);
		else STAPLE_LAUNCH(1, EPI_MASS_DOT);
	}
#undef STAPLE_LAUNCH
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}

// acc_Deo / acc_Doe (fermion_matrix.c:159-268): operator on the local interior followed by the exchange
// of the first/last interior slice of `out` into the neighbours' halos.
//   single rank      : one launch over all d3
//   peer memory      : ONE segmented launch (faces, bulk, unpack), see DslashArgs
//   async_comm (NCCL): d3p (queue 2), d3m (queue 3), bulk (queue 1) ; halo exchange after the two
//                      surface slices, overlapped with the bulk; join (:165-183)
//   otherwise        : one launch, then blocking exchange (:194-205)
template <typename T>
void apply_dslash(int par, int epi, const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, const T *ph,
									const cplx_t<T> *in0, double m2, int dot_slot, const int *skip, int halo)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	const int lo = g.d3_halo, hi = g.d3_halo + g.loc_n3;
	if (c.nranks <= 1) {
		launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, dslash_blocks(lo, hi), 0, skip, c.stream);
		return;
	}
	if (!c.async_comm_fermion && !c.p2p.on) {
		launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, dslash_blocks(lo, hi), 0, skip, c.stream);
		// dirac_times (fermion_matrix.c:196-205, :252-259): rank 0 times the blocking exchange with the wall clock and counts it
		cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
		if (cudaStreamIsCapturing(c.stream, &cap) != cudaSuccess) cudaGetLastError();
		const bool measure = c.myrank == 0 && cap == cudaStreamCaptureStatusNone;
		struct timespec t0, t1;
		if (measure) { STAPLE_CUDA_CHECK(cudaStreamSynchronize(c.stream)); clock_gettime(CLOCK_MONOTONIC, &t0); }
		exchange_slices(out, sizeof(cplx_t<T>), g.sizeh, 3, 1, c.stream);
		if (measure) {
			STAPLE_CUDA_CHECK(cudaStreamSynchronize(c.stream)); clock_gettime(CLOCK_MONOTONIC, &t1);
			dirac_times.totTransferTime += (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
			++dirac_times.count;
		}
		return;
	}
	const unsigned int bs = dslash_blocks(0, 1), bb = dslash_blocks(lo + 1, hi - 1);
	if (c.p2p.on && c.p2p_single_launch) {
		const unsigned int target = 2 * bs + bb;
		// ONE kernel: the face blocks (first in block order) store their sites into the neighbours' staging slots over NVLink
		// while the rest of the launch runs; the received halos are either copied into `out` by the last blocks of the same
		// launch (the API's contract: `out` leaves with valid halos), by a separate kernel, or -- inside the solvers -- left in
		// the staging area for the next kernel to consume.  No stream fork/join, no events, no wait on a block of this launch.
		const bool lazy = halo_lazy_ok();
		const int in_staged = lazy ? (halo & HALO_IN_STAGED) : 0;
		if (lazy && (halo & HALO_NO_PUSH))
			launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, target, 0, skip, c.stream, FACE_BOTH, in_staged | HALO_NO_PUSH);
		else if (lazy && (halo & HALO_OUT_STAGED))
			launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, target, 0, skip, c.stream, FACE_BOTH, in_staged | HALO_ADVANCE);
		else if (c.p2p_unpack_in_kernel)
			launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, target + 2 * ((hi - lo) >= 6 ? (bs < 74u ? bs : 74u) : unpack_blocks_for(bs)), 0, skip, c.stream,
											 FACE_BOTH_UNPACK, in_staged | HALO_ADVANCE);
		else {
			launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, hi, dot_slot, target, 0, skip, c.stream, FACE_BOTH, 0);
			p2p_unpack(out, sizeof(cplx_t<T>), c.stream, skip);
		}
		return;
	}
	// three-queue form (the reference's structure).  Peer-memory channel: the two surface kernels store their slice into the
	// neighbours' staging slots themselves (compute + transfer in one kernel).  A solver's `skip` flag (set in the same
	// iteration on every rank, because the all-reduced scalars are bit-identical) silences producers and consumer alike.
	const bool p2p = c.p2p.on;
	const unsigned int target = 2 * bs + bb;
	STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, c.stream));
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.s_p, c.ev_fork, 0));
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.s_m, c.ev_fork, 0));
	launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, hi - 1, hi, dot_slot, target, 0, skip, c.s_p, p2p ? FACE_TOP : FACE_NONE);      // d3p
	launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo, lo + 1, dot_slot, target, bs, skip, c.s_m, p2p ? FACE_BOTTOM : FACE_NONE);  // d3m
	STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_p, c.s_p));
	STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_m, c.s_m));
	launch_dslash<T>(par, epi, u, out, in, ph, in0, m2, lo + 1, hi - 1, dot_slot, target, 2 * bs, skip, c.stream);   // bulk
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.s_comm, c.ev_p, 0));
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.s_comm, c.ev_m, 0));
	if (p2p) p2p_unpack(out, sizeof(cplx_t<T>), c.s_comm, skip);
	else exchange_slices(out, sizeof(cplx_t<T>), g.sizeh, 3, 1, c.s_comm);
	STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_comm, c.s_comm));
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_comm, 0));
}

// fermion_matrix_multiplication[_shifted] (fermion_matrix.c:723-746) with the mass term (and, for the
// solvers, Re(in . out)) fused into the Deo epilogue.  On D3 slabs over peer memory the halos of tmp = Doe in go from
// the neighbours' face blocks to the staging area and from there straight into the d3 hops of the Deo face blocks.
template <typename T>
void apply_mdagm(const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, cplx_t<T> *tmp, const T *ph,
								 double m2, int dot_slot, const int *skip, bool cgm_interior, bool tmp_eager)
{
	const bool lazy = halo_lazy_ok() && !tmp_eager;
	const bool interior = lazy && cgm_interior;
	apply_dslash<T>(1, EPI_NONE, u, tmp, in, ph, nullptr, 0.0, -1, skip, (lazy ? HALO_OUT_STAGED : HALO_EAGER) | (interior ? HALO_IN_STAGED : 0));
	apply_dslash<T>(0, dot_slot >= 0 ? EPI_MASS_DOT : EPI_MASS, u, out, tmp, ph, in, m2, dot_slot, skip,
									(lazy ? HALO_IN_STAGED : 0) | (interior ? HALO_NO_PUSH : 0));
}

template void apply_dslash<double>(int, int, const double2 *, double2 *, const double2 *, const double *,
																	 const double2 *, double, int, const int *, int);
template void apply_dslash<float>(int, int, const float2 *, float2 *, const float2 *, const float *,
																	const float2 *, double, int, const int *, int);
template void apply_mdagm<double>(const double2 *, double2 *, const double2 *, double2 *, const double *,
																	double, int, const int *, bool, bool);
template void apply_mdagm<float>(const float2 *, float2 *, const float2 *, float2 *, const float *, double,
																 int, const int *, bool, bool);
template void launch_dslash<double>(int, int, const double2 *, double2 *, const double2 *, const double *,
																		const double2 *, double, int, int, int, unsigned int, unsigned int, const int *,
																		cudaStream_t, int, int);
template void launch_dslash<float>(int, int, const float2 *, float2 *, const float2 *, const float *,
																	 const float2 *, double, int, int, int, unsigned int, unsigned int, const int *,
																	 cudaStream_t, int, int);

// ------------------------------------------------------------------ operator "with a field" (magnetic susceptibility)
// field_times_fermion_matrix.c:77-196 with matvecmul.h:176-260: the same stencil with every link's U(1) phase multiplied by a
// complex per-link field (field_re + i field_im)[k][idx_mat] -- "not in U(1) anymore".  FP64 only (the reference generates
// no _f twin).  Algorithmic bytes per site: 928 + 8 x 16 (the field) = 1056.  A measurement-side sibling of the hot kernel,
// kept out of dslash_kernel so that the hot instantiations stay exactly as tuned.
template <bool DAG>
__device__ __forceinline__ void hop_wf(double2 acc[3], const double2 *__restrict__ uk, const double *__restrict__ phk,
																			 const double *__restrict__ frk, const double *__restrict__ fik, unsigned int im,
																			 const double2 *__restrict__ in, unsigned int iv, long n)
{
	using C = double2;
	const double th = ld_stream(phk + im), fr = ld_stream(frk + im), fi = ld_stream(fik + im);
	const C m00 = ld_stream(uk + im), m01 = ld_stream(uk + n + im), m02 = ld_stream(uk + 2 * n + im);
	const C m10 = ld_stream(uk + 3 * n + im), m11 = ld_stream(uk + 4 * n + im), m12 = ld_stream(uk + 5 * n + im);
	const C v0 = ld_cached(in + iv), v1 = ld_cached(in + n + iv), v2 = ld_cached(in + 2 * n + iv);
	double s, c;
	sincos_t(th, &s, &c);
	const C ph = cmul(mk<double>(c, s), mk<double>(fr, fi));      // matvecmul.h:188-189
	const C x0 = cross(m01, m12, m02, m11);
	const C x1 = cross(m02, m10, m00, m12);
	const C x2 = cross(m00, m11, m01, m10);
	if (!DAG) {
		const C w0 = cmul(v0, ph), w1 = cmul(v1, ph), w2 = cmul(v2, ph);
		cfma(acc[0], m00, w0); cfma(acc[0], m01, w1); cfma(acc[0], m02, w2);
		cfma(acc[1], m10, w0); cfma(acc[1], m11, w1); cfma(acc[1], m12, w2);
		cfma_ca(acc[2], x0, w0); cfma_ca(acc[2], x1, w1); cfma_ca(acc[2], x2, w2);
	} else {
		const C p = mk<double>(-ph.x, ph.y);                        // -conj(phase): backward hops are subtracted
		const C w0 = cmul(v0, p), w1 = cmul(v1, p), w2 = cmul(v2, p);
		cfma_ca(acc[0], m00, w0); cfma_ca(acc[0], m10, w1); cfma(acc[0], x0, w2);
		cfma_ca(acc[1], m01, w0); cfma_ca(acc[1], m11, w1); cfma(acc[1], x1, w2);
		cfma_ca(acc[2], m02, w0); cfma_ca(acc[2], m12, w1); cfma(acc[2], x2, w2);
	}
}

struct DslashWfArgs {
	const double2 *u, *in; double2 *out;
	const double *ph, *fre, *fim;
	long site_lo, nsites, sizeh, vol3h;
	int nd0h, nd1, nd2, nd3;
};

template <int PAR>
__global__ void __launch_bounds__(kBlock) dslash_wf_kernel(const DslashWfArgs a)
{
	using C = double2;
	const unsigned int t = blockIdx.x * kBlock + threadIdx.x;
	if (t >= (unsigned int) a.nsites) return;
	const unsigned int idx = (unsigned int) a.site_lo + t;
	const long n = a.sizeh;
	const unsigned int nd0h = a.nd0h, nd1 = a.nd1, nd2 = a.nd2, nd3 = a.nd3;
	const unsigned int hd0 = idx % nd0h;
	unsigned int q = idx / nd0h;
	const unsigned int d1 = q % nd1; q /= nd1;
	const unsigned int d2 = q % nd2;
	const unsigned int d3 = q / nd2;
	const unsigned int rp = (d1 + d2 + d3 + PAR) & 1u;
	const unsigned int s1 = nd0h, s2 = nd0h * nd1, s3 = (unsigned int) a.vol3h;
	const unsigned int i0m = rp ? idx : (hd0 == 0 ? idx + (nd0h - 1) : idx - 1);
	const unsigned int i0p = rp ? (hd0 == nd0h - 1 ? idx - (nd0h - 1) : idx + 1) : idx;
	const unsigned int i1m = d1 == 0 ? idx + s1 * (nd1 - 1) : idx - s1;
	const unsigned int i1p = d1 == nd1 - 1 ? idx - s1 * (nd1 - 1) : idx + s1;
	const unsigned int i2m = d2 == 0 ? idx + s2 * (nd2 - 1) : idx - s2;
	const unsigned int i2p = d2 == nd2 - 1 ? idx - s2 * (nd2 - 1) : idx + s2;
	const unsigned int i3m = d3 == 0 ? idx + s3 * (nd3 - 1) : idx - s3;
	const unsigned int i3p = d3 == nd3 - 1 ? idx - s3 * (nd3 - 1) : idx + s3;
	C acc[3];
	acc[0] = mk<double>(0, 0); acc[1] = mk<double>(0, 0); acc[2] = mk<double>(0, 0);
	const long un = 9 * n;
#define STAPLE_WF(DAG, K, IM, IV) hop_wf<DAG>(acc, a.u + (K) * un, a.ph + (K) * n, a.fre + (K) * n, a.fim + (K) * n, IM, a.in, IV, n)
	STAPLE_WF(true, 1 - PAR, i0m, i0m); STAPLE_WF(true, 3 - PAR, i1m, i1m);        // field_times_fermion_matrix.c:104-111
	STAPLE_WF(true, 5 - PAR, i2m, i2m); STAPLE_WF(true, 7 - PAR, i3m, i3m);
	STAPLE_WF(false, 0 + PAR, idx, i0p); STAPLE_WF(false, 2 + PAR, idx, i1p);      // :117-124
	STAPLE_WF(false, 4 + PAR, idx, i2p); STAPLE_WF(false, 6 + PAR, idx, i3p);
#undef STAPLE_WF
#pragma unroll
	for (int c = 0; c < 3; c++) a.out[c * n + idx] = mk<double>(acc[c].x * 0.5, acc[c].y * 0.5);   // :128-130
}

static void launch_dslash_wf(int par, const double2 *u, double2 *out, const double2 *in, const double *ph, const double *fre,
														 const double *fim)
{
	const Geom &g = ctx().g;
	DslashWfArgs a;
	a.u = u; a.in = in; a.out = out; a.ph = ph; a.fre = fre; a.fim = fim;
	a.site_lo = (long) g.d3_halo * g.vol3h; a.nsites = (long) g.loc_n3 * g.vol3h; a.sizeh = g.sizeh; a.vol3h = g.vol3h;
	a.nd0h = g.nd0h; a.nd1 = g.nd1; a.nd2 = g.nd2; a.nd3 = g.nd3;
	const unsigned int grid = dslash_blocks(g.d3_halo, g.d3_halo + g.loc_n3);
	if (par == 0) dslash_wf_kernel<0><<<grid, kBlock, 0, ctx().stream>>>(a);
	else dslash_wf_kernel<1><<<grid, kBlock, 0, ctx().stream>>>(a);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}

// ------------------------------------------------------------------ BLAS-1 element-wise kernels
// All arithmetic in double with double factors, stored back in T: this is what the reference's FP32
// twin does too (float complex * double promotes; double_to_single_transformer.py leaves
// fermionic_utilities.c out of filesALLDtoF).  Operands may alias element-wise (the reference calls
// combine_in1xfactor_plus_in2(p, g, r, p)), hence no __restrict__.
template <typename T, int OP>
__global__ void __launch_bounds__(kBlasBlock) blas_kernel(cplx_t<T> *out, const cplx_t<T> *a, const cplx_t<T> *b,
																													const cplx_t<T> *c, double f1, cplx_t<T> *out2, long lo,
																													long cnt, long n)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
#pragma unroll
	for (int col = 0; col < 3; col++) {
		const long j = col * n + lo + t;
		double rx = 0, ry = 0;
		if (OP == OP_IN1XFACTOR_PLUS_IN2) { rx = a[j].x * f1 + b[j].x; ry = a[j].y * f1 + b[j].y; }
		else if (OP == OP_SCALE) { rx = f1 * out[j].x; ry = f1 * out[j].y; }
		else if (OP == OP_ADD_FACTOR_X_IN2) { rx = out[j].x + f1 * a[j].x; ry = out[j].y + f1 * a[j].y; }
		else if (OP == OP_IN1XMASS2_MINUS_IN2_MINUS_IN3) {
			rx = a[j].x * f1 - b[j].x - c[j].x; ry = a[j].y * f1 - b[j].y - c[j].y;
		}
		else if (OP == OP_IN1XMASS_MINUS_IN2) { rx = a[j].x * f1 - out[j].x; ry = a[j].y * f1 - out[j].y; }
		else if (OP == OP_IN1_MINUS_IN2) { rx = (double) a[j].x - b[j].x; ry = (double) a[j].y - b[j].y; }
		else if (OP == OP_ASSIGN) { rx = a[j].x; ry = a[j].y; }
		else if (OP == OP_ZERO) { rx = 0; ry = 0; }
		else if (OP == OP_FACT1_MINUS_IN2) { rx = f1 * a[j].x - out[j].x; ry = f1 * a[j].y - out[j].y; }
		else if (OP == OP_IN1_MINUS_IN2_ALLXFACT) {
			rx = f1 * ((double) a[j].x - b[j].x); ry = f1 * ((double) a[j].y - b[j].y);
		}
		else if (OP == OP_INSIDE_LOOP) {   // out += omega*p ; r -= omega*s  (out2 = r, a = s, b = p)
			rx = out[j].x + b[j].x * f1; ry = out[j].y + b[j].y * f1;
			out2[j] = mk<T>((T) (out2[j].x - a[j].x * f1), (T) (out2[j].y - a[j].y * f1));
		}
		out[j] = mk<T>((T) rx, (T) ry);
	}
}

template <typename T>
void blas(BlasOp op, cplx_t<T> *out, const cplx_t<T> *a, const cplx_t<T> *b, const cplx_t<T> *c, double f1,
					cplx_t<T> *out2)
{
	Ctx &cx = ctx();
	const Geom &g = cx.g;
	long lo = g.r1_lo, cnt = g.r1_hi - g.r1_lo;
	if (op == OP_ZERO) { lo = 0; cnt = g.sizeh; }   // set_vec3_soa_to_zero covers all sizeh (:309)
	const unsigned int grid = (unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock);
#define STAPLE_BLAS(O) case O: blas_kernel<T, O><<<grid, kBlasBlock, 0, cx.stream>>>(out, a, b, c, f1, out2, lo, cnt, g.sizeh); break;
	switch (op) {
		STAPLE_BLAS(OP_IN1XFACTOR_PLUS_IN2) STAPLE_BLAS(OP_SCALE) STAPLE_BLAS(OP_ADD_FACTOR_X_IN2)
		STAPLE_BLAS(OP_IN1XMASS2_MINUS_IN2_MINUS_IN3) STAPLE_BLAS(OP_IN1XMASS_MINUS_IN2) STAPLE_BLAS(OP_IN1_MINUS_IN2)
		STAPLE_BLAS(OP_ASSIGN) STAPLE_BLAS(OP_ZERO) STAPLE_BLAS(OP_FACT1_MINUS_IN2) STAPLE_BLAS(OP_IN1_MINUS_IN2_ALLXFACT)
		STAPLE_BLAS(OP_INSIDE_LOOP)
	}
#undef STAPLE_BLAS
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}
template void blas<double>(BlasOp, double2 *, const double2 *, const double2 *, const double2 *, double, double2 *);
template void blas<float>(BlasOp, float2 *, const float2 *, const float2 *, const float2 *, double, float2 *);

// multi-vector updates of the reference API (fermionic_utilities.c:315-378): small host arrays by value
struct MultiArgs { int flag[MAX_APPROX_ORDER]; double f1[MAX_APPROX_ORDER]; double f2[MAX_APPROX_ORDER]; int maxiter; };

template <typename T, int WHICH>   // 0: out[ia] -= f1*in[ia]   1: in1[ia] = f1*in1[ia] + f2*in2
__global__ void __launch_bounds__(kBlasBlock) multi_kernel(cplx_t<T> *x, const cplx_t<T> *y, MultiArgs m, long lo,
																													 long cnt, long n)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	for (int col = 0; col < 3; col++) {
		const long j = col * n + lo + t;
		cplx_t<T> r2 = mk<T>(0, 0);
		if (WHICH == 1) r2 = y[j];
		for (int ia = 0; ia < m.maxiter; ia++) {
			if (m.flag[ia] != 1) continue;
			const long k = (long) ia * 3 * n + j;
			if (WHICH == 0) x[k] = mk<T>((T) (x[k].x - m.f1[ia] * y[k].x), (T) (x[k].y - m.f1[ia] * y[k].y));
			else x[k] = mk<T>((T) (m.f1[ia] * x[k].x + m.f2[ia] * r2.x), (T) (m.f1[ia] * x[k].y + m.f2[ia] * r2.y));
		}
	}
}

// calc_new_trialsol_for_inversion_in_force (fermionic_utilities.c:417-455)
template <typename T>
__global__ void __launch_bounds__(kBlasBlock) trialsol_kernel(cplx_t<T> *v, int halfLen, int odd, long lo, long cnt, long n)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	for (int iv = 0; iv < halfLen; iv++)
		for (int col = 0; col < 3; col++) {
			const long a = (long) iv * 3 * n + col * n + lo + t, b = (long) (iv + halfLen) * 3 * n + col * n + lo + t;
			if (odd) v[b] = mk<T>((T) (2 * v[a].x - v[b].x), (T) (2 * v[a].y - v[b].y));
			else v[a] = mk<T>((T) (2 * v[b].x - v[a].x), (T) (2 * v[b].y - v[a].y));
		}
}

// ------------------------------------------------------------------ reductions (fermionic_utilities.c:32-175)
template <typename T, int OP>
__global__ void __launch_bounds__(kBlasBlock) reduce_kernel(const cplx_t<T> *a, const cplx_t<T> *b, long lo, long cnt,
																														 long n, double *partials, long pstride, unsigned int *ticket,
																														 double *result)
{
	double v[2] = { 0.0, 0.0 };
	for (long t = (long) blockIdx.x * kBlasBlock + threadIdx.x; t < cnt; t += (long) gridDim.x * kBlasBlock) {
		const long i = lo + t;
		double sr = 0, si = 0;
#pragma unroll
		for (int col = 0; col < 3; col++) {
			const cplx_t<T> x = a[col * n + i];
			if (OP == RED_L2NORM2) sr += (double) x.x * x.x + (double) x.y * x.y;
			else {
				const cplx_t<T> y = b[col * n + i];
				sr += (double) x.x * y.x + (double) x.y * y.y;
				if (OP == RED_CPLX_DOT) si += (double) x.x * y.y - (double) x.y * y.x;
			}
		}
		v[0] += sr; v[1] += si;
	}
	grid_sum_finalize<2>(v, partials, pstride, ticket, result, gridDim.x, blockIdx.x);
}

template <typename T>
void reduce_local(RedOp op, const cplx_t<T> *a, const cplx_t<T> *b, int slot)
{
	Ctx &c = ctx();
	const Geom &g = c.g;
	const long lo = g.r0_lo, cnt = g.r0_hi - g.r0_lo;
	long want = (cnt + kBlasBlock - 1) / kBlasBlock;
	const unsigned int grid = (unsigned int) (want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
	double *p = partials(slot);
#define STAPLE_RED(O) reduce_kernel<T, O><<<grid, kBlasBlock, 0, c.stream>>>(a, b, lo, cnt, g.sizeh, p, c.max_partials, ticket(slot), result(slot))
	if (op == RED_L2NORM2) STAPLE_RED(RED_L2NORM2);
	else if (op == RED_REAL_DOT) STAPLE_RED(RED_REAL_DOT);
	else STAPLE_RED(RED_CPLX_DOT);
#undef STAPLE_RED
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}
template void reduce_local<double>(RedOp, const double2 *, const double2 *, int);
template void reduce_local<float>(RedOp, const float2 *, const float2 *, int);

template <typename T>
staple_dcomplex reduce_global(RedOp op, const cplx_t<T> *a, const cplx_t<T> *b)
{
	reduce_local<T>(op, a, b, 0);
	double h[2];
	fetch_results(0, 2, h);
	staple_dcomplex r = { h[0], h[1] };
	return r;
}
template staple_dcomplex reduce_global<double>(RedOp, const double2 *, const double2 *);
template staple_dcomplex reduce_global<float>(RedOp, const float2 *, const float2 *);

// ------------------------------------------------------------------ conversions / recombine
template <typename TI, typename TO>
__global__ void __launch_bounds__(kBlasBlock) convert_kernel(const TI *in, TO *out, long cnt)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t < cnt) out[t] = (TO) in[t];
}
template <typename TI, typename TO>
static void convert(const TI *in, TO *out, long cnt)
{
	convert_kernel<TI, TO><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(in, out, cnt);
	STAPLE_CUDA_CHECK(cudaGetLastError());
	count_launch();
}

// combine_add_in2_into_in1_mixed_precision (inverter_mixedp.c:24-34)
__global__ void __launch_bounds__(kBlasBlock) add_mixed_kernel(double2 *x, const float2 *y, long lo, long cnt, long n)
{
	const long t = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (t >= cnt) return;
	for (int col = 0; col < 3; col++) {
		const long j = col * n + lo + t;
		x[j] = make_double2(x[j].x + (double) y[j].x, x[j].y + (double) y[j].y);
	}
}

// recombine_shifted_vec3_to_vec3 (inverter_multishift_full.c:254-282): all sizeh, rounding to T after
// every accumulation step exactly as the reference's in-memory accumulation does.
struct RecombArgs { double a0; double a[MAX_APPROX_ORDER]; int order; };
template <typename T>
__global__ void __launch_bounds__(kBlasBlock) recombine_kernel(const cplx_t<T> *sh, const cplx_t<T> *in, cplx_t<T> *out,
																															 RecombArgs r, long n3)
{
	const long j = (long) blockIdx.x * kBlasBlock + threadIdx.x;
	if (j >= n3) return;
	cplx_t<T> o = mk<T>((T) (in[j].x * r.a0), (T) (in[j].y * r.a0));
	for (int i = 0; i < r.order; i++) {
		const cplx_t<T> s = sh[(long) i * n3 + j];
		o = mk<T>((T) (o.x + r.a[i] * s.x), (T) (o.y + r.a[i] * s.y));
	}
	out[j] = o;
}

}   // namespace staple

using namespace staple;

// ====================================================================== C ABI
#define DD(p) ((double2 *) dev(p, #p))
#define DF(p) ((float2 *) dev(p, #p))
#define CDD(p) ((const double2 *) dev(p, #p))
#define CDF(p) ((const float2 *) dev(p, #p))

// captured schedules of staple_acc_Doe_Deo_streamed, keyed by buffers + chunking (geometry is baked into their kernel nodes:
// staple_init_geometry and staple_shutdown flush them)
struct StreamedGraph { const void *u, *out, *in, *tmp, *ph, *d_in, *d_out; int cs, mode, par; long sizeh; cudaGraphExec_t exec; unsigned long long launches; };
static StreamedGraph g_streamed_cache[4] = {};
static int g_streamed_next = 0;
constexpr int kMaxChunks = 128;
static cudaEvent_t g_ev_up[kMaxChunks], g_ev_deo[kMaxChunks];            // chunk uploaded / Deo chunk computed
static cudaEvent_t g_tev0, g_tev_up[kMaxChunks], g_tev_deo[kMaxChunks], g_tev_dn[kMaxChunks];   // STAPLE_STREAMED_TRACE timeline
static bool g_have_events = false, g_have_tev = false;
namespace staple {
// cached graphs AND the events of the chunk schedule: they belong to the device that was current when they were created
// (staple_shutdown / a re-initialisation on another device must not leave them behind)
void release_streamed_state()
{
	for (auto &e : g_streamed_cache) {
		if (e.exec) cudaGraphExecDestroy(e.exec);
		e = StreamedGraph{};
	}
	g_streamed_next = 0;
	if (g_have_events) {
		for (int k = 0; k < kMaxChunks; k++) { cudaEventDestroy(g_ev_up[k]); cudaEventDestroy(g_ev_deo[k]); }
		g_have_events = false;
	}
	if (g_have_tev) {
		cudaEventDestroy(g_tev0);
		for (int k = 0; k < kMaxChunks; k++) { cudaEventDestroy(g_tev_up[k]); cudaEventDestroy(g_tev_deo[k]); cudaEventDestroy(g_tev_dn[k]); }
		g_have_tev = false;
	}
}
}   // namespace staple

extern "C" {

// ---- operator variants.  `unsafe`: whole local interior, no exchange; bulk/d3p/d3m/d3c: d3 sub-ranges
#define STAPLE_DSLASH_DEF(NAME, PAR, LO, HI, STREAM, FULL)                                                        \
	void NAME(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *backfield)                     \
	{                                                                                                               \
		require_init(#NAME);                                                                                          \
		Ctx &c = ctx(); const Geom &g = c.g; (void) g;                                                                \
		if (FULL) apply_dslash<double>(PAR, EPI_NONE, CDD(u), DD(out), CDD(in), (const double *) dev(backfield, "backfield"), nullptr, 0.0, -1, nullptr); \
		else launch_dslash<double>(PAR, EPI_NONE, CDD(u), DD(out), CDD(in), (const double *) dev(backfield, "backfield"), nullptr, 0.0, LO, HI, -1, 0, 0, nullptr, STREAM); \
	}                                                                                                               \
	void NAME##_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *backfield)            \
	{                                                                                                               \
		require_init(#NAME "_f");                                                                                     \
		Ctx &c = ctx(); const Geom &g = c.g; (void) g;                                                                \
		if (FULL) apply_dslash<float>(PAR, EPI_NONE, CDF(u), DF(out), CDF(in), (const float *) dev(backfield, "backfield"), nullptr, 0.0, -1, nullptr); \
		else launch_dslash<float>(PAR, EPI_NONE, CDF(u), DF(out), CDF(in), (const float *) dev(backfield, "backfield"), nullptr, 0.0, LO, HI, -1, 0, 0, nullptr, STREAM); \
	}
STAPLE_DSLASH_DEF(acc_Deo, 0, 0, 0, c.stream, true)
STAPLE_DSLASH_DEF(acc_Doe, 1, 0, 0, c.stream, true)
STAPLE_DSLASH_DEF(acc_Deo_unsafe, 0, g.d3_halo, g.d3_halo + g.loc_n3, c.stream, false)
STAPLE_DSLASH_DEF(acc_Doe_unsafe, 1, g.d3_halo, g.d3_halo + g.loc_n3, c.stream, false)
STAPLE_DSLASH_DEF(acc_Deo_bulk, 0, g.d3_halo + 1, g.d3_halo + 1 + g.loc_n3 - 2, c.stream, false)
STAPLE_DSLASH_DEF(acc_Doe_bulk, 1, g.d3_halo + 1, g.d3_halo + 1 + g.loc_n3 - 2, c.stream, false)
STAPLE_DSLASH_DEF(acc_Deo_d3p, 0, g.nd3 - g.d3_halo - 1, g.nd3 - g.d3_halo, c.stream, false)
STAPLE_DSLASH_DEF(acc_Doe_d3p, 1, g.nd3 - g.d3_halo - 1, g.nd3 - g.d3_halo, c.stream, false)
STAPLE_DSLASH_DEF(acc_Deo_d3m, 0, g.d3_halo, g.d3_halo + 1, c.stream, false)
STAPLE_DSLASH_DEF(acc_Doe_d3m, 1, g.d3_halo, g.d3_halo + 1, c.stream, false)

void acc_Deo_d3c(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *bf, int off3, int thick3)
{
	require_init("acc_Deo_d3c");
	launch_dslash<double>(0, EPI_NONE, CDD(u), DD(out), CDD(in), (const double *) dev(bf, "backfield"), nullptr, 0.0, off3, off3 + thick3, -1, 0, 0, nullptr, ctx().stream);
}
void acc_Doe_d3c(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *bf, int off3, int thick3)
{
	require_init("acc_Doe_d3c");
	launch_dslash<double>(1, EPI_NONE, CDD(u), DD(out), CDD(in), (const double *) dev(bf, "backfield"), nullptr, 0.0, off3, off3 + thick3, -1, 0, 0, nullptr, ctx().stream);
}
void acc_Deo_d3c_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *bf, int off3, int thick3)
{
	require_init("acc_Deo_d3c_f");
	launch_dslash<float>(0, EPI_NONE, CDF(u), DF(out), CDF(in), (const float *) dev(bf, "backfield"), nullptr, 0.0, off3, off3 + thick3, -1, 0, 0, nullptr, ctx().stream);
}
void acc_Doe_d3c_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *bf, int off3, int thick3)
{
	require_init("acc_Doe_d3c_f");
	launch_dslash<float>(1, EPI_NONE, CDF(u), DF(out), CDF(in), (const float *) dev(bf, "backfield"), nullptr, 0.0, off3, off3 + thick3, -1, 0, 0, nullptr, ctx().stream);
}

// field_times_fermion_matrix.c:77-232
#define STAPLE_WF_DEF(NAME, PAR, EXCHANGE)                                                                                    \
	void NAME(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *phases, const double_soa *field_re,        \
						const double_soa *field_im)                                                                                       \
	{                                                                                                                           \
		require_init(#NAME);                                                                                                      \
		launch_dslash_wf(PAR, CDD(u), DD(out), CDD(in), (const double *) dev(phases, "phases"),                                   \
										 (const double *) dev(field_re, "field_re"), (const double *) dev(field_im, "field_im"));                 \
		if (EXCHANGE && ctx().nranks > 1) communicate_fermion_borders(out);                                                       \
	}
STAPLE_WF_DEF(acc_Deo_wf_unsafe, 0, false)
STAPLE_WF_DEF(acc_Doe_wf_unsafe, 1, false)
STAPLE_WF_DEF(acc_Deo_wf, 0, true)
STAPLE_WF_DEF(acc_Doe_wf, 1, true)

// deo_doe_test.c's host round trip (update device; acc_Doe; acc_Deo; update host) pipelined over d3 chunks.
// Chunk k of Doe reads `in` chunks k-1,k,k+1 (periodic), chunk k of Deo reads the Doe output of k-1,k,k+1: both
// are issued on the compute stream as soon as the last chunk they depend on has been uploaded / computed, and
// every finished Deo chunk is downloaded on its own copy stream.  Same kernels, same per-site arithmetic as the
// whole-lattice launch => bit-identical results.
void staple_acc_Doe_Deo_streamed(const su3_soa *u, vec3_soa *out_h, const vec3_soa *in_h, vec3_soa *tmp,
																 const double_soa *backfield, int chunk_slices)
{
	require_init("staple_acc_Doe_Deo_streamed");
	Ctx &c = ctx();
	const Geom &g = c.g;
	const double2 *d_u = CDD(u); const double *d_ph = (const double *) dev(backfield, "backfield");
	double2 *d_in = (double2 *) dev(in_h, "in"), *d_out = DD(out_h), *d_tmp = DD(tmp);
	const bool in_host = (const void *) d_in != (const void *) in_h, out_host = (void *) d_out != (void *) out_h;
	const size_t vbytes = sizeof(double2) * 3 * g.sizeh;
	const bool slabs = c.nranks > 1;
	if ((slabs && !(c.p2p.on && c.p2p_single_launch && g.nd3 <= 4096)) || !in_host || !out_host) {
		if (in_host) staple_acc_update_device(in_h, vbytes);
		apply_dslash<double>(1, EPI_NONE, d_u, d_tmp, d_in, d_ph, nullptr, 0.0, -1, nullptr);
		apply_dslash<double>(0, EPI_NONE, d_u, d_out, d_tmp, d_ph, nullptr, 0.0, -1, nullptr);
		if (out_host) staple_acc_update_host(out_h, vbytes);
		else STAPLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
		return;
	}
	cudaEvent_t *const ev_up = g_ev_up, *const ev_deo = g_ev_deo;
	if (!g_have_events) {
		for (int k = 0; k < kMaxChunks; k++) {
			STAPLE_CUDA_CHECK(cudaEventCreateWithFlags(&ev_up[k], cudaEventDisableTiming));
			STAPLE_CUDA_CHECK(cudaEventCreateWithFlags(&ev_deo[k], cudaEventDisableTiming));
		}
		g_have_events = true;
	}
	// default: 16 chunks -- measured optimum on PCIe Gen5 (profiles/r01_streamed_trace.txt): smaller chunks lose copy
	// efficiency (3 strided pieces per chunk and direction), bigger ones lengthen the 5-chunk pipeline head and tail
	int cs = chunk_slices > 0 ? chunk_slices : (g.nd3 >= 16 ? g.nd3 / 16 : 1);
	if (cs > g.nd3) cs = g.nd3;
	if (slabs) { while ((g.nd3 + cs - 1) / cs + 4 > kMaxChunks) cs++; }
	else while (g.nd3 % cs != 0 || g.nd3 / cs > kMaxChunks) cs++;   // cs = nd3 always qualifies
	const int nc = g.nd3 / cs;
	const size_t pitch = sizeof(double2) * g.sizeh, width = sizeof(double2) * g.vol3h * cs;
	cudaStream_t s_up = c.s_p, s_dn = c.s_m, st = c.stream;
	// mode 0 (default): chunk downloads by the copy engine after every Deo chunk.
	// mode 1: the Deo chunk kernels store their result straight into the pinned host buffer (UVA: the host address is
	// valid on the device) -- the download is fused into the operator's epilogue and runs on its own stream next to
	// the Doe chunks.  Measured: SM stores over PCIe reach ~31 GB/s against ~45 GB/s for the copy engine, so this is
	// an option, not the default.
	const int mode = c.streamed_mode;
	// STAPLE_STREAMED_TRACE=1: direct issue with timing events after every chunk upload / Deo chunk / chunk download;
	// the timeline (ms since the call started) is printed on stderr.  Diagnostic only.
	static const bool trace = getenv("STAPLE_STREAMED_TRACE") != nullptr;
	cudaEvent_t &tev0 = g_tev0, *const tev_up = g_tev_up, *const tev_deo = g_tev_deo, *const tev_dn = g_tev_dn;
	if (trace && !g_have_tev) {
		cudaEventCreate(&tev0);
		for (int k = 0; k < kMaxChunks; k++) { cudaEventCreate(&tev_up[k]); cudaEventCreate(&tev_deo[k]); cudaEventCreate(&tev_dn[k]); }
		g_have_tev = true;
	}
	auto enqueue = [&]() {
		// the side streams may not touch `in`/`out` on the device before earlier work of the compute stream is done
		if (trace) cudaEventRecord(tev0, st);
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, st));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_up, c.ev_fork, 0));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, c.ev_fork, 0));
		bool doe_done[kMaxChunks] = {}, deo_done[kMaxChunks] = {};
		auto deps_ok = [&](const bool *have, int k, int upto) {   // chunks k-1,k,k+1 (mod nc) all present
			for (int d = -1; d <= 1; d++) {
				const int j = (k + d + nc) % nc;
				if (have ? !have[j] : j > upto) return false;
			}
			return true;
		};
		for (int j = 0; j < nc; j++) {
			const size_t off = (size_t) j * cs * g.vol3h;
			STAPLE_CUDA_CHECK(cudaMemcpy2DAsync(d_in + off, pitch, (const double2 *) in_h + off, pitch, width, 3,
																					cudaMemcpyHostToDevice, s_up));
			STAPLE_CUDA_CHECK(cudaEventRecord(ev_up[j], s_up));
			if (trace) cudaEventRecord(tev_up[j], s_up);
			bool waited = false;
			for (int k = 0; k < nc; k++) {
				if (doe_done[k] || !deps_ok(nullptr, k, j)) continue;
				if (!waited) { STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, ev_up[j], 0)); waited = true; }
				launch_dslash<double>(1, EPI_NONE, d_u, d_tmp, d_in, d_ph, nullptr, 0.0, k * cs, (k + 1) * cs, -1, 0, 0, nullptr, st);
				doe_done[k] = true;
			}
			bool doe_marked = false;
			for (int k = 0; k < nc; k++) {
				if (deo_done[k] || !deps_ok(doe_done, k, 0)) continue;
				const size_t o2 = (size_t) k * cs * g.vol3h;
				if (mode == 1) {
					// Deo chunks on the second stream, after every Doe chunk issued so far (in order on `st`)
					if (!doe_marked) {
						STAPLE_CUDA_CHECK(cudaEventRecord(ev_deo[j], st));
						STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, ev_deo[j], 0));
						doe_marked = true;
					}
					c.out_host_hook = out_h;
					launch_dslash<double>(0, EPI_NONE, d_u, d_out, d_tmp, d_ph, nullptr, 0.0, k * cs, (k + 1) * cs, -1, 0, 0, nullptr, s_dn);
					c.out_host_hook = nullptr;
					if (trace) cudaEventRecord(tev_dn[k], s_dn);
				} else {
					launch_dslash<double>(0, EPI_NONE, d_u, d_out, d_tmp, d_ph, nullptr, 0.0, k * cs, (k + 1) * cs, -1, 0, 0, nullptr, st);
					STAPLE_CUDA_CHECK(cudaEventRecord(ev_deo[k], st));
					STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, ev_deo[k], 0));
					if (trace) cudaEventRecord(tev_deo[k], st);
					STAPLE_CUDA_CHECK(cudaMemcpy2DAsync((double2 *) out_h + o2, pitch, d_out + o2, pitch, width, 3,
																							cudaMemcpyDeviceToHost, s_dn));
					if (trace) cudaEventRecord(tev_dn[k], s_dn);
				}
				deo_done[k] = true;
			}
		}
		// join: everything (last download / host store included) is ordered before whatever follows on the compute stream
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_p, s_up));
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_m, s_dn));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, c.ev_p, 0));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, c.ev_m, 0));
	};
	// ---- D3 slabs (peer-memory transport): the same pipeline per rank, boundaries first.  The local+halo box of `in` comes
	// from the host WITH its halo slices, so Doe needs no exchange on its input and -- unlike the single-rank lattice, which is
	// periodic in d3 -- nothing wraps around: only the two face slices couple to the neighbour ranks.
	//   uploads   : head [0, lo+2), tail [hi-2, nd3), then the middle in chunks                       (copy stream 1)
	//   Doe       : bottom and top face slice FIRST (they push tmp's faces to the neighbours: exchange s+1), then bulk chunks
	//   Deo       : face slices with the halo of tmp read straight from the staging area (no unpack of tmp), pushing the
	//               faces of `out` (exchange s+2); bulk chunks; every finished piece is downloaded       (copy stream 2)
	//   finally   : unpack of the received halos of `out` and their download
	// The exchange counter is device resident (advanced by one-thread kernels / the unpack kernel), so the captured graph replays.
	auto enqueue_slabs = [&]() {
		const int lo = g.d3_halo, hi = g.d3_halo + g.loc_n3, nd3 = g.nd3;
		const size_t slice = (size_t) g.vol3h;
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, st));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_up, c.ev_fork, 0));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, c.ev_fork, 0));
		auto copy_slices = [&](bool up, int a, int b, cudaStream_t cst) {
			if (b <= a) return;
			const size_t off = (size_t) a * slice, w = sizeof(double2) * slice * (size_t) (b - a);
			if (up) STAPLE_CUDA_CHECK(cudaMemcpy2DAsync(d_in + off, pitch, (const double2 *) in_h + off, pitch, w, 3, cudaMemcpyHostToDevice, cst));
			else STAPLE_CUDA_CHECK(cudaMemcpy2DAsync((double2 *) out_h + off, pitch, d_out + off, pitch, w, 3, cudaMemcpyDeviceToHost, cst));
		};
		// slices of `out` that no kernel writes (outer halo, HALO_WIDTH 2): `update host` copies the whole array, so do we
		copy_slices(false, 0, lo - 1, s_dn); copy_slices(false, hi + 1, nd3, s_dn);
		struct Unit { int a, b, face; bool doe, deo; };
		Unit units[kMaxChunks]; int nu = 0;
		units[nu++] = Unit{ lo, lo + 1, FACE_BOTTOM, false, false };
		units[nu++] = Unit{ hi - 1, hi, FACE_TOP, false, false };
		for (int a = lo + 1; a < hi - 1; a += cs) units[nu++] = Unit{ a, a + cs < hi - 1 ? a + cs : hi - 1, FACE_NONE, false, false };
		int piece_of[4096]; bool doe_done[4096] = {};
		for (int d = 0; d < nd3; d++) piece_of[d] = -1;
		int npieces = 0, waited = -1, ndn = 0, doe_faces = 0, deo_faces = 0;
		bool advanced = false, unpacked = false;
		auto need_piece = [&](int a, int b) { int m = -1; for (int d = a; d < b; d++) { if (piece_of[d] < 0) return -2; if (piece_of[d] > m) m = piece_of[d]; } return m; };
		auto download = [&](int a, int b) {
			STAPLE_CUDA_CHECK(cudaEventRecord(ev_deo[ndn], st));
			STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, ev_deo[ndn], 0));
			ndn++;
			copy_slices(false, a, b, s_dn);
		};
		auto progress = [&]() {
			bool again = true;
			while (again) {
				again = false;
				for (int k = 0; k < nu; k++) {
					Unit &un = units[k];
					if (!un.doe) {
						const int m = need_piece(un.a - 1, un.b + 1);
						if (m == -2) continue;
						if (m > waited) { STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, ev_up[m], 0)); waited = m; }
						launch_dslash<double>(1, EPI_NONE, d_u, d_tmp, d_in, d_ph, nullptr, 0.0, un.a, un.b, -1, 0, 0, nullptr, st, un.face, 0);
						un.doe = true; again = true;
						for (int d = un.a; d < un.b; d++) doe_done[d] = true;
						if (un.face != FACE_NONE) doe_faces++;
					}
				}
				if (doe_faces == 2 && !advanced) {       // both faces of tmp are on their way: exchange s+1 is what the Deo faces consume
					c.p2p.h_seq += 1; advanced = true; again = true;
				}
				for (int k = 0; k < nu && advanced; k++) {
					Unit &un = units[k];
					if (un.deo) continue;
					bool ok = true;                         // Doe output of the interior slices around the unit; halo slices of tmp are staged
					for (int d = un.a - 1; d < un.b + 1; d++) if (d >= lo && d < hi && !doe_done[d]) ok = false;
					if (!ok) continue;
					launch_dslash<double>(0, EPI_NONE, d_u, d_out, d_tmp, d_ph, nullptr, 0.0, un.a, un.b, -1, 0, 0, nullptr, st, un.face,
																un.face != FACE_NONE ? HALO_IN_STAGED : 0);
					un.deo = true; again = true;
					download(un.a, un.b);
					if (un.face != FACE_NONE) deo_faces++;
				}
				if (deo_faces == 2 && !unpacked) {
					p2p_unpack(d_out, sizeof(double2), st, nullptr);      // takes the neighbours' data of exchange s+2 as it lands; the exchange is complete
					unpacked = true;
					STAPLE_CUDA_CHECK(cudaEventRecord(ev_deo[ndn], st));
					STAPLE_CUDA_CHECK(cudaStreamWaitEvent(s_dn, ev_deo[ndn], 0));
					ndn++;
					copy_slices(false, lo - 1, lo, s_dn); copy_slices(false, hi, hi + 1, s_dn);
				}
			}
		};
		auto upload = [&](int a, int b) {
			if (b <= a) return;
			copy_slices(true, a, b, s_up);
			STAPLE_CUDA_CHECK(cudaEventRecord(ev_up[npieces], s_up));
			for (int d = a; d < b; d++) piece_of[d] = npieces;
			npieces++;
			progress();
		};
		const int head = lo + 2 < nd3 ? lo + 2 : nd3, tail = hi - 2 > head ? hi - 2 : head;
		upload(0, head);
		upload(tail, nd3);
		for (int a = head; a < tail; a += cs) upload(a, a + cs < tail ? a + cs : tail);
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_p, s_up));
		STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_m, s_dn));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, c.ev_p, 0));
		STAPLE_CUDA_CHECK(cudaStreamWaitEvent(st, c.ev_m, 0));
	};
	// The schedule is ~8 API calls per chunk: on the host that costs more than the PCIe time it hides.  It only
	// depends on the pointers and the chunking, so it is captured ONCE into a CUDA graph (three streams, copy
	// nodes included) and replayed with a single launch.  The legacy default stream cannot be captured: direct
	// issue there.
	using Cached = StreamedGraph;
	Cached (&cache)[4] = g_streamed_cache;
	int &cache_next = g_streamed_next;
	Cached *hit = nullptr;
	const int par = slabs ? (int) (c.p2p.h_seq & 1ull) : 0;      // the staging parity the schedule starts from
	if (st != nullptr && c.use_graphs && !trace) {
		for (auto &e : cache)
			if (e.exec && e.u == u && e.out == out_h && e.in == in_h && e.tmp == tmp && e.ph == backfield && e.d_in == d_in && e.d_out == d_out && e.cs == cs && e.mode == mode && e.par == par && e.sizeh == g.sizeh) hit = &e;
		if (!hit) {
			const unsigned long long before = c.launches, seq_before = c.p2p.h_seq;
			cudaGraph_t graph = nullptr;
			cudaGraphExec_t exec = nullptr;
			if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
				if (slabs) enqueue_slabs(); else enqueue();
				// (two exchanges per call on D3 slabs: the frozen staging parities replay consistently from the same starting parity)
				if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph != nullptr && ((c.p2p.h_seq - seq_before) & 1ull) == 0 &&
						cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
					Cached &e = cache[cache_next]; cache_next = (cache_next + 1) % 4;
					if (e.exec) cudaGraphExecDestroy(e.exec);
					e = Cached{ u, out_h, in_h, tmp, backfield, d_in, d_out, cs, mode, par, g.sizeh, exec, c.launches - before };
					hit = &e;
				}
				if (graph) cudaGraphDestroy(graph);
			}
			cudaGetLastError();
			c.launches = before;
			c.p2p.h_seq = seq_before;      // nothing was executed
		}
	}
	if (hit) { STAPLE_CUDA_CHECK(cudaGraphLaunch(hit->exec, st)); c.launches += hit->launches; }
	else if (slabs) enqueue_slabs();
	else enqueue();
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));
	if (trace && !slabs) {
		fprintf(stderr, "streamed trace: mode %d, %d chunks of %d slices\n chunk   upload_done   deo_done   download_done [ms]\n", mode, nc, cs);
		for (int k = 0; k < nc; k++) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, tev0, tev_up[k]);
			if (mode == 0) cudaEventElapsedTime(&b, tev0, tev_deo[k]);
			cudaEventElapsedTime(&d, tev0, tev_dn[k]);
			fprintf(stderr, " %3d   %9.4f   %9.4f   %9.4f\n", k, a, b, d);
		}
	}
}

void fermion_matrix_multiplication(const su3_soa *u, vec3_soa *out, const vec3_soa *in, vec3_soa *temp1, ferm_param *pars)
{
	require_init("fermion_matrix_multiplication");
	apply_mdagm<double>(CDD(u), DD(out), CDD(in), DD(temp1), (const double *) dev(pars->phases, "pars->phases"),
											pars->ferm_mass * pars->ferm_mass, -1, nullptr, false, true);
}
void fermion_matrix_multiplication_shifted(const su3_soa *u, vec3_soa *out, const vec3_soa *in, vec3_soa *temp1,
																					 ferm_param *pars, double shift)
{
	require_init("fermion_matrix_multiplication_shifted");
	apply_mdagm<double>(CDD(u), DD(out), CDD(in), DD(temp1), (const double *) dev(pars->phases, "pars->phases"),
											pars->ferm_mass * pars->ferm_mass + shift, -1, nullptr, false, true);
}
void fermion_matrix_multiplication_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, vec3_soa_f *temp1, ferm_param *pars)
{
	require_init("fermion_matrix_multiplication_f");
	apply_mdagm<float>(CDF(u), DF(out), CDF(in), DF(temp1), (const float *) dev(pars->phases_f, "pars->phases_f"),
										 pars->ferm_mass * pars->ferm_mass, -1, nullptr, false, true);
}
void fermion_matrix_multiplication_shifted_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, vec3_soa_f *temp1,
																						 ferm_param *pars, float shift)
{
	require_init("fermion_matrix_multiplication_shifted_f");
	apply_mdagm<float>(CDF(u), DF(out), CDF(in), DF(temp1), (const float *) dev(pars->phases_f, "pars->phases_f"),
										 pars->ferm_mass * pars->ferm_mass + shift, -1, nullptr, false, true);
}

// ---- BLAS-1
#define STAPLE_BLAS_DEF(V, S, T, D, CD)                                                                              \
	staple_dcomplex scal_prod_global##S(const V *a, const V *b) { require_init("scal_prod_global"); return reduce_global<T>(RED_CPLX_DOT, CD(a), CD(b)); } \
	double real_scal_prod_global##S(const V *a, const V *b) { require_init("real_scal_prod_global"); return reduce_global<T>(RED_REAL_DOT, CD(a), CD(b)).re; } \
	double l2norm2_global##S(const V *a) { require_init("l2norm2_global"); return reduce_global<T>(RED_L2NORM2, CD(a), nullptr).re; } \
	void combine_in1xfactor_plus_in2##S(const V *in_vect1, const double factor, const V *in_vect2, V *out)             \
	{ require_init("combine_in1xfactor_plus_in2"); blas<T>(OP_IN1XFACTOR_PLUS_IN2, D(out), CD(in_vect1), CD(in_vect2), nullptr, factor); } \
	void multiply_fermion_x_doublefactor##S(V *in1, const double factor)                                               \
	{ require_init("multiply_fermion_x_doublefactor"); blas<T>(OP_SCALE, D(in1), nullptr, nullptr, nullptr, factor); } \
	void combine_add_factor_x_in2_to_in1##S(V *in1, const V *in2, double factor)                                       \
	{ require_init("combine_add_factor_x_in2_to_in1"); blas<T>(OP_ADD_FACTOR_X_IN2, D(in1), CD(in2), nullptr, nullptr, factor); } \
	void combine_in1xferm_mass2_minus_in2_minus_in3##S(const V *in_vect1, double ferm_mass, const V *in_vect2, const V *in_vect3, V *out) \
	{ require_init("combine_in1xferm_mass2_minus_in2_minus_in3"); blas<T>(OP_IN1XMASS2_MINUS_IN2_MINUS_IN3, D(out), CD(in_vect1), CD(in_vect2), CD(in_vect3), ferm_mass); } \
	void combine_inside_loop##S(V *vect_out, V *vect_r, const V *vect_s, const V *vect_p, const double omega)          \
	{ require_init("combine_inside_loop"); blas<T>(OP_INSIDE_LOOP, D(vect_out), CD(vect_s), CD(vect_p), nullptr, omega, D(vect_r)); } \
	void combine_in1xferm_mass_minus_in2##S(const V *in_vect1, double ferm_mass2, V *in_vect2)                         \
	{ require_init("combine_in1xferm_mass_minus_in2"); blas<T>(OP_IN1XMASS_MINUS_IN2, D(in_vect2), CD(in_vect1), nullptr, nullptr, ferm_mass2); } \
	void combine_in1_minus_in2##S(const V *in_vect1, const V *in_vect2, V *out)                                        \
	{ require_init("combine_in1_minus_in2"); blas<T>(OP_IN1_MINUS_IN2, D(out), CD(in_vect1), CD(in_vect2), nullptr, 0.0); } \
	void assign_in_to_out##S(const V *in_vect1, V *out)                                                                \
	{ require_init("assign_in_to_out"); blas<T>(OP_ASSIGN, D(out), CD(in_vect1), nullptr, nullptr, 0.0); }            \
	void set_vec3_soa_to_zero##S(V *fermion)                                                                           \
	{ require_init("set_vec3_soa_to_zero"); blas<T>(OP_ZERO, D(fermion), nullptr, nullptr, nullptr, 0.0); }           \
	void combine_in1_x_fact1_minus_in2_back_into_in2##S(const V *in1, double fact1, V *in2)                            \
	{ require_init("combine_in1_x_fact1_minus_in2_back_into_in2"); blas<T>(OP_FACT1_MINUS_IN2, D(in2), CD(in1), nullptr, nullptr, fact1); } \
	void combine_in1_minus_in2_allxfact##S(const V *in1, const V *in2, double fact, V *out)                            \
	{ require_init("combine_in1_minus_in2_allxfact"); blas<T>(OP_IN1_MINUS_IN2_ALLXFACT, D(out), CD(in1), CD(in2), nullptr, fact); } \
	void multiple_combine_in1_minus_in2x_factor_back_into_in1##S(V *out, const V *in, const int maxiter, const int *flag, const double *omegas) \
	{                                                                                                                  \
		require_init("multiple_combine_in1_minus_in2x_factor_back_into_in1");                                           \
		const Geom &g = ctx().g; MultiArgs m; m.maxiter = maxiter;                                                      \
		for (int i = 0; i < maxiter; i++) { m.flag[i] = flag[i]; m.f1[i] = omegas[i]; m.f2[i] = 0; }                    \
		const long cnt = g.r1_hi - g.r1_lo;                                                                             \
		multi_kernel<T, 0><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(D(out), CD(in), m, g.r1_lo, cnt, g.sizeh); \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                          \
	}                                                                                                                  \
	void multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1##S(V *in1, int maxiter, const int *flag, const double *gammas, const V *in2, const double *zeta_iii) \
	{                                                                                                                  \
		require_init("multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1");                                   \
		const Geom &g = ctx().g; MultiArgs m; m.maxiter = maxiter;                                                      \
		for (int i = 0; i < maxiter; i++) { m.flag[i] = flag[i]; m.f1[i] = gammas[i]; m.f2[i] = zeta_iii[i]; }          \
		const long cnt = g.r1_hi - g.r1_lo;                                                                             \
		multi_kernel<T, 1><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(D(in1), CD(in2), m, g.r1_lo, cnt, g.sizeh); \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                          \
	}                                                                                                                  \
	void calc_new_trialsol_for_inversion_in_force##S(int halfLen, V *inout, int nPrecCalculations)                     \
	{                                                                                                                  \
		require_init("calc_new_trialsol_for_inversion_in_force");                                                       \
		const Geom &g = ctx().g; const long cnt = g.r1_hi - g.r1_lo;                                                    \
		trialsol_kernel<T><<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(D(inout), halfLen, nPrecCalculations % 2, g.r1_lo, cnt, g.sizeh); \
		STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();                                                          \
	}
STAPLE_BLAS_DEF(vec3_soa, , double, DD, CDD)
STAPLE_BLAS_DEF(vec3_soa_f, _f, float, DF, CDF)

// ---- conversions (over all sizeh; su3: rows r0,r1,r2 of ONE su3_soa)
void convert_float_to_double_vec3_soa(const vec3_soa_f *f, vec3_soa *d)
{ require_init("convert_float_to_double_vec3_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), 6 * ctx().g.sizeh); }
void convert_double_to_float_vec3_soa(const vec3_soa *d, vec3_soa_f *f)
{ require_init("convert_double_to_float_vec3_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), 6 * ctx().g.sizeh); }
void convert_float_to_double_su3_soa(const su3_soa_f *f, su3_soa *d)
{ require_init("convert_float_to_double_su3_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), 8 * 18 * ctx().g.sizeh); }
void convert_double_to_float_su3_soa(const su3_soa *d, su3_soa_f *f)
{ require_init("convert_double_to_float_su3_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), 8 * 18 * ctx().g.sizeh); }
void convert_float_to_double_real_soa(const float_soa *f, double_soa *d)
{ require_init("convert_float_to_double_real_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), ctx().g.sizeh); }
void convert_double_to_float_real_soa(const double_soa *d, float_soa *f)
{ require_init("convert_double_to_float_real_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), ctx().g.sizeh); }

// tamat_soa[8] / thmat_soa[8]: three complex and two real arrays per link = 8 reals per site and link, all 8 links per call
// (float_double_conv.c:150-228); dcomplex_soa: one complex array (:48-70)
void convert_float_to_double_tamat_soa(const tamat_soa_f *f, tamat_soa *d)
{ require_init("convert_float_to_double_tamat_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), 8 * 8 * ctx().g.sizeh); }
void convert_double_to_float_tamat_soa(const tamat_soa *d, tamat_soa_f *f)
{ require_init("convert_double_to_float_tamat_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), 8 * 8 * ctx().g.sizeh); }
void convert_float_to_double_thmat_soa(const thmat_soa_f *f, thmat_soa *d)
{ require_init("convert_float_to_double_thmat_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), 8 * 8 * ctx().g.sizeh); }
void convert_double_to_float_thmat_soa(const thmat_soa *d, thmat_soa_f *f)
{ require_init("convert_double_to_float_thmat_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), 8 * 8 * ctx().g.sizeh); }
void convert_float_to_double_complex_soa(const fcomplex_soa *f, dcomplex_soa *d)
{ require_init("convert_float_to_double_complex_soa"); convert<float, double>((const float *) dev(f, "f_var"), (double *) dev(d, "d_var"), 2 * ctx().g.sizeh); }
void convert_double_to_float_complex_soa(const dcomplex_soa *d, fcomplex_soa *f)
{ require_init("convert_double_to_float_complex_soa"); convert<double, float>((const double *) dev(d, "d_var"), (float *) dev(f, "f_var"), 2 * ctx().g.sizeh); }

// one colour vector held by value on the HOST (struct_c_def.h:29-33): plain casts, no device work (float_double_conv.c:34-47)
void convert_float_to_double_vec3(const vec3_f *f, vec3 *d)
{
	const float *s = (const float *) f; double *o = (double *) d;
	for (int i = 0; i < 6; i++) o[i] = (double) s[i];
}
void convert_double_to_float_vec3(const vec3 *d, vec3_f *f)
{
	const double *s = (const double *) d; float *o = (float *) f;
	for (int i = 0; i < 6; i++) o[i] = (float) s[i];
}

void combine_add_in2_into_in1_mixed_precision(vec3_soa *in1, const vec3_soa_f *in2)
{
	require_init("combine_add_in2_into_in1_mixed_precision");
	const Geom &g = ctx().g; const long cnt = g.r1_hi - g.r1_lo;
	add_mixed_kernel<<<(unsigned int) ((cnt + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(DD(in1), CDF(in2), g.r1_lo, cnt, g.sizeh);
	STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();
}

void recombine_shifted_vec3_to_vec3(const vec3_soa *in_shifted, const vec3_soa *in, vec3_soa *out, const RationalApprox *approx)
{
	require_init("recombine_shifted_vec3_to_vec3");
	RecombArgs r; r.a0 = approx->RA_a0; r.order = approx->approx_order;
	for (int i = 0; i < r.order; i++) r.a[i] = approx->RA_a[i];
	const long n3 = 3 * ctx().g.sizeh;
	recombine_kernel<double><<<(unsigned int) ((n3 + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(CDD(in_shifted), CDD(in), DD(out), r, n3);
	STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();
}
void recombine_shifted_vec3_to_vec3_f(const vec3_soa_f *in_shifted, const vec3_soa_f *in, vec3_soa_f *out, const RationalApprox *approx)
{
	require_init("recombine_shifted_vec3_to_vec3_f");
	RecombArgs r; r.a0 = approx->RA_a0; r.order = approx->approx_order;
	for (int i = 0; i < r.order; i++) r.a[i] = approx->RA_a[i];
	const long n3 = 3 * ctx().g.sizeh;
	recombine_kernel<float><<<(unsigned int) ((n3 + kBlasBlock - 1) / kBlasBlock), kBlasBlock, 0, ctx().stream>>>(CDF(in_shifted), CDF(in), DF(out), r, n3);
	STAPLE_CUDA_CHECK(cudaGetLastError()); count_launch();
}

}   // extern "C"

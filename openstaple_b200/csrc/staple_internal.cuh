// Internal declarations shared by the translation units of libstaple_b200.so.
// The public boundary is include/staple_b200.h; nothing here is exported.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../../include/staple_b200.h"

#define STAPLE_CUDA_CHECK(x)                                                                     \
	do {                                                                                           \
		cudaError_t e_ = (x);                                                                        \
		if (e_ != cudaSuccess) {                                                                     \
			fprintf(stderr, "libstaple_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_),   \
							__FILE__, __LINE__, #x);                                                           \
			exit(1);                                                                                   \
		}                                                                                            \
	} while (0)

namespace staple {

// Run-time copy of the reference's compile-time geometry (geometry.h:12-29, geometry_multidev.h:6-148).
struct Geom {
	int nd0h, nd0, nd1, nd2, nd3;   // local+halo box, nd0h = nd0/2
	int loc_n3;                     // LOC_N3
	int nranks;                     // NRANKS_D3
	int halo_width;                 // HALO_WIDTH
	int d3_halo, d3_fhalo;          // D3_HALO, D3_FERMION_HALO
	long vol3h;                     // nd0*nd1*nd2/2 : half-sites per d3 slice
	long sizeh;                     // vol3h*nd3
	long r0_lo, r0_hi;              // reduction range  (fermionic_utilities.c:41)
	long r1_lo, r1_hi;              // update range     (fermionic_utilities.c:188)
};

constexpr int kResultSlots = 16;        // device scalars produced by reductions

struct Comm;   // NCCL state, staple_core.cu
struct CgmCtl; // CG-M control block, below
struct CgCtl;  // CG / mixed-precision CG control block, below

// Peer-memory channels over NVLink (CUDA IPC between the one-process-per-GPU ranks).  Every rank owns ONE
// shared "mailbox" allocation:
//   header (reduction boxes[2 parities][kMaxRanks][2 doubles]) | halo staging stage[2 parities][2 slots][3 colours x vol3h x 16 B]
// Halo slot 0 receives the data of this rank's LOWER fermion halo (stored by rank L's top-face blocks), slot 1 the UPPER
// halo (stored by rank R's bottom-face blocks).
// THE DATA IS ITS OWN ARRIVAL FLAG.  Every 8-byte (FP32: 4-byte) word of the staging area and of the reduction boxes rests at a
// reserved bit pattern (all ones: a NaN no arithmetic produces); a producer just stores its values into the neighbour's memory --
// posted NVLink writes, no fence, no flag -- and a consumer spins on the very word it needs until it differs from the pattern,
// uses it and puts the pattern back.  8-byte accesses are single-copy atomic, so a word is either old (pattern) or new (value);
// a value that happens to equal the pattern is nudged to another NaN by the producer.  Measured motivation
// (profiles/r02c_halo_probe_*.jsonl, r02d_*): one fence.sys after remote stores costs 7-15 us on B200/NVSwitch, whoever
// issues it; with per-chunk flags that was 25-30 us per operator launch.  Exchange number s uses staging parity s&1: the
// neighbour may already deliver exchange s+1 while exchange s is being consumed here, never s+2 (it needs our s+1 first).
// The exchange number is a HOST counter handed to the kernels by value (only its parity matters on the device): no device
// counter to read, no ticket to decide who advances it, no kernel tail.  A captured CUDA graph freezes the parities of its
// launches, which is consistent as long as it contains an EVEN number of exchanges (every solver batch does: two per
// iteration; checked at capture) and is replayed from the parity it was captured at.  The reduction counter d_redq stays on
// the device (one warp per reduction).
constexpr int kMaxRanks = 16;
constexpr size_t kMailboxRedBox = 64, kMailboxHeader = 1024;
constexpr unsigned long long kSentinel64 = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned int kSentinel32 = 0xFFFFFFFFu;
struct P2P {
	bool on = false;
	char *mailbox = nullptr;                  // local, cudaMalloc (IPC exported)
	char *peer_mailbox[kMaxRanks] = {};       // every rank's mailbox as mapped here ([myrank] = local)
	char *stage = nullptr;                    // local staging = mailbox + kMailboxHeader
	char *stage_L = nullptr, *stage_R = nullptr;
	unsigned long long h_seq = 0;             // number of halo exchanges ENQUEUED so far (same on every rank: same call sequence)
	unsigned long long *d_redq = nullptr;     // local, device: number of completed reductions
	size_t slot_bytes = 0;                    // 3 * vol3h * 16
	long nfb = 0;                             // operator CTAs per face slice
	long vol3h = 0;                           // the geometry the mailbox was built for
};
// by-value kernel argument of the peer-memory all-reduce
struct RedView {
	int nranks, myrank;
	unsigned long long *q;
	double *box[kMaxRanks];                   // reduction boxes of every rank (parity 0 base)
};
// by-value kernel argument: a kernel other than the operator that produces a vector AND pushes its two interior face slices
// into the neighbours' staging slots (CG-M: p = r + gamma p)
struct PushView {
	int on;
	char *peer_top, *peer_bot;                // parity-0 staging slot in rank R's (slot 0) / rank L's (slot 1) memory
	unsigned long long seq;                   // the exchange this kernel produces
	long parity_bytes;                        // bytes between the parity-0 and parity-1 staging
	long top_lo, bot_lo, vol3h;               // first idxh of the top / bottom interior slice
};

struct Ctx {
	bool inited = false;
	Geom g{};
	int device = 0;
	cudaStream_t own_stream = nullptr;   // created by the library
	cudaStream_t stream = nullptr;       // stream all entry points enqueue on
	cudaStream_t s_p = nullptr, s_m = nullptr, s_comm = nullptr;   // OpenACC queues 2,3 + comm
	cudaEvent_t ev_fork = nullptr, ev_p = nullptr, ev_m = nullptr, ev_comm = nullptr, ev_misc = nullptr;
	double *d_partials = nullptr;        // [kResultSlots][2][max_partials] per-block partial sums
	long max_partials = 0;
	unsigned int *d_tickets = nullptr;   // [kResultSlots]
	double *d_results = nullptr;         // [kResultSlots][2]
	double *h_results = nullptr;         // pinned mirror
	unsigned long long launches = 0;
	// rank layer (multidev.h:10-41)
	int myrank = 0, nranks = 1, rank_L = 0, rank_R = 0, async_comm_fermion = 0;
	bool loopback = false;           // staple_init_loopback: D3-slab layout and protocol with THIS rank as both neighbours
	Comm *comm = nullptr;
	P2P p2p;
	bool use_graphs = true;          // CG-M iteration batches as CUDA graphs (single GPU, non-default stream)
	bool p2p_unpack_in_kernel = true; // ... and the unpack blocks ride in the same launch
	bool p2p_single_launch = true;   // acc_Deo/acc_Doe as one kernel + unpack (false: d3p/d3m/bulk on three streams)
	bool p2p_lazy = true;            // solvers leave intermediate halos in the staging area and consume them there
	// set by the CG-M solver around its iteration batches: fuse the after-alpha recurrences into the Deo tail
	CgmCtl *cgm_hook = nullptr;
	CgCtl *cg_hook = nullptr;        // same for the single-system CG / mixed-precision CG (cg_after_alpha_warp)
	void *out_host_hook = nullptr;   // staple_acc_Doe_Deo_streamed: the Deo chunk kernels also store their result in host memory
	int streamed_mode = 0;           // 0: chunk downloads by the copy engine  1: stores over PCIe from the Deo kernels
	bool cgm_fuse_tail = true;       // false: one-warp kernels of their own (staple_set_cgm_fuse_tail, A/B tests)
	bool cg_device_loops = true;     // false: ker_invert_openacc / inverter_mixed_precision read their scalars back every iteration (A/B)
	RedView cgm_hook_red{};
	// last multishift statistics
	int last_iterations = 0;
	long long last_active = 0;
	double last_loop_ms = 0;
};

Ctx &ctx();
void require_init(const char *fn);

// host->device pointer translation (OpenACC "present" semantics); aborts if not present
void *resolve_raw(const void *p, const char *what);
template <typename T>
inline T *dev(const T *p, const char *what) { return static_cast<T *>(resolve_raw(p, what)); }

inline double *partials(int slot) { return ctx().d_partials + (size_t) slot * 2 * ctx().max_partials; }
inline unsigned int *ticket(int slot) { return ctx().d_tickets + slot; }
inline double *result(int slot) { return ctx().d_results + 2 * slot; }

// rank layer
void allreduce_results(int slot, int ndoubles, cudaStream_t s);   // in-stream sum over ranks of d_results[slot]
void exchange_slices(void *base, size_t elem_bytes, long stride_elems, int narrays, int thickness,
										 cudaStream_t s);                              // communications.c:34-104 on device memory
// peer-memory variant for one vector (3 colour arrays, thickness 1): push both faces + unpack both halos
void p2p_exchange_fermion(void *base, size_t elem_bytes, cudaStream_t s);
// unpack only (the faces were pushed by the surface kernels themselves); advances the exchange counter
void p2p_unpack(void *base, size_t elem_bytes, cudaStream_t s, const int *skip);
// in-place sum over ranks of `ndoubles` (1 or 2) doubles through the peer mailboxes, fixed rank order
void p2p_allreduce(double *vals, int ndoubles, cudaStream_t s);
RedView make_redview();
inline RedView single_rank_redview() { RedView v; v.nranks = 1; v.myrank = 0; v.q = nullptr; return v; }

// operator CTA size (scripts/tune_dslash.py sweeps it; 128 is the measured optimum)
#ifndef STAPLE_DSLASH_BLOCK
#define STAPLE_DSLASH_BLOCK 128
#endif
constexpr int kDslashBlock = STAPLE_DSLASH_BLOCK;

#ifdef __CUDACC__
// Every in-kernel wait is for a word that a PEER GPU stores (never for a block of the same launch), so forward progress
// does not depend on block scheduling order; it does depend on the peer being alive.  The spin is bounded: after
// g_spin_timeout_ns (staple_set_spin_timeout, default 60 s; 0 = unbounded, MPI_Wait semantics) the kernel reports what it
// was waiting for and traps, which surfaces on the host as a CUDA error instead of a hang.
// (no relocatable device code in this build: the variable and the cold path exist once per translation unit that waits --
// staple_kernels.cu and staple_solvers.cu -- and staple_set_spin_timeout sets both copies)
static __device__ unsigned long long g_spin_timeout_ns = 60ull * 1000000000ull;
static __device__ __noinline__ void spin_timeout_trap(int what)
{
	printf("libstaple_b200: FATAL: block %u thread %u waited more than %llu s for a peer GPU (%s) -- is a rank dead, or did the "
				 "ranks issue different call sequences?  (staple_set_spin_timeout changes the limit)\n",
				 blockIdx.x, threadIdx.x, g_spin_timeout_ns / 1000000000ull, what == 1 ? "halo data" : "reduction mailbox");
	__trap();
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
struct SpinGuard {
	unsigned int spins = 0;
	unsigned long long t0 = 0;
	__device__ __forceinline__ void pause(int what)
	{
		__nanosleep(spins < 32 ? 20 : 200);
		if ((++spins & 0x3fffu) == 0) {            // look at the clock every ~3 ms
			const unsigned long long now = globaltimer_ns(), lim = *(volatile unsigned long long *) &g_spin_timeout_ns;
			if (t0 == 0) t0 = now;
			else if (lim != 0 && now - t0 > lim) spin_timeout_trap(what);
		}
	}
};
// consume one staged element: spin until both words have arrived, put the resting pattern back (system-scope relaxed
// accesses: always served by L2, where the peer's NVLink writes land)
__device__ __forceinline__ double2 take_staged(double2 *p)
{
	unsigned long long x, y;
	SpinGuard g;
	for (;;) {
		asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
		if (x != kSentinel64 && y != kSentinel64) break;
		g.pause(1);
	}
	asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(kSentinel64), "l"(kSentinel64) : "memory");
	return make_double2(__longlong_as_double((long long) x), __longlong_as_double((long long) y));
}
__device__ __forceinline__ float2 take_staged(float2 *p)
{
	unsigned int x, y;
	SpinGuard g;
	for (;;) {
		asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "l"(p) : "memory");
		if (x != kSentinel32 && y != kSentinel32) break;
		g.pause(1);
	}
	asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(kSentinel32), "r"(kSentinel32) : "memory");
	return make_float2(__uint_as_float(x), __uint_as_float(y));
}
// the same in two steps, for consumers that want many elements in flight: peek_staged issues the load and returns at once
// (`arrived` tells whether both words were there), take_staged(p) is the slow path for the stragglers; the caller puts the
// resting pattern back with reset_staged once it holds the value
__device__ __forceinline__ double2 peek_staged(double2 *p, bool *arrived)
{
	unsigned long long x, y;
	asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
	*arrived = x != kSentinel64 && y != kSentinel64;
	return make_double2(__longlong_as_double((long long) x), __longlong_as_double((long long) y));
}
__device__ __forceinline__ float2 peek_staged(float2 *p, bool *arrived)
{
	unsigned int x, y;
	asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "l"(p) : "memory");
	*arrived = x != kSentinel32 && y != kSentinel32;
	return make_float2(__uint_as_float(x), __uint_as_float(y));
}
__device__ __forceinline__ void reset_staged(double2 *p)
{
	asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(kSentinel64), "l"(kSentinel64) : "memory");
}
__device__ __forceinline__ void reset_staged(float2 *p)
{
	asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(kSentinel32), "r"(kSentinel32) : "memory");
}

// producer side: a value that equals the resting pattern (a NaN with all payload bits set) becomes the canonical NaN
__device__ __forceinline__ double2 stageable(double2 v)
{
	if ((unsigned long long) __double_as_longlong(v.x) == kSentinel64) v.x = __longlong_as_double(0x7ff8000000000000ll);
	if ((unsigned long long) __double_as_longlong(v.y) == kSentinel64) v.y = __longlong_as_double(0x7ff8000000000000ll);
	return v;
}
__device__ __forceinline__ float2 stageable(float2 v)
{
	if (__float_as_uint(v.x) == kSentinel32) v.x = __uint_as_float(0x7fc00000u);
	if (__float_as_uint(v.y) == kSentinel32) v.y = __uint_as_float(0x7fc00000u);
	return v;
}

// One warp: every rank stores its value into every rank's box (lane = destination rank), waits for all
// contributions to its own box and adds them in rank order -- bit-identical results everywhere, about one
// NVLink round trip, and usable as the prologue of a kernel that consumes the sum (no separate launch).
// Same protocol as the halos: the doubles are their own arrival flags, no fence.
__device__ __forceinline__ void p2p_allreduce_warp(double *vals, int nd, const RedView &v)
{
	const int lane = threadIdx.x & 31;
	const unsigned long long q = *v.q + 1;
	const int par = (int) (q & 1ull);
	double s0 = 0.0, s1 = 0.0;
	if (lane < v.nranks) {
		double2 mine = stageable(make_double2(vals[0], nd > 1 ? vals[1] : 0.0));
		double *b = v.box[lane] + ((size_t) par * kMaxRanks + v.myrank) * 2;
		asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(b), "d"(mine.x), "d"(mine.y) : "memory");
		const double2 got = take_staged((double2 *) (v.box[v.myrank] + ((size_t) par * kMaxRanks + lane) * 2));
		s0 = got.x; s1 = got.y;
	}
	// rank-ordered sum: lane r holds rank r's contribution
	double t0 = 0.0, t1 = 0.0;
	for (int r = 0; r < v.nranks; r++) { t0 += __shfl_sync(0xffffffffu, s0, r); t1 += __shfl_sync(0xffffffffu, s1, r); }
	if (lane == 0) {
		vals[0] = t0;
		if (nd > 1) vals[1] = t1;
		*v.q = q;
		__threadfence();
	}
	__syncwarp();
}
#endif

// ---- CG-M control block (inverter_multishift_full.c:53-58 host arrays, here device resident)
constexpr double kSafetyMargin = 0.95;   // inverter_multishift_full.c:18, inverter_full.c:17
struct CgmCtl {
	double alpha, delta, lambda, omega, omega_save, gammag, source_norm, residuo;
	double zeta_i[MAX_APPROX_ORDER], zeta_ii[MAX_APPROX_ORDER], zeta_iii[MAX_APPROX_ORDER];
	double omegas[MAX_APPROX_ORDER], gammas[MAX_APPROX_ORDER], shifts[MAX_APPROX_ORDER];
	// coefficients of the search-direction update ps_i = pgam_i ps_i + pzeta_i r that the reference performs at
	// the end of an iteration (:152-157); here it is carried into the next iteration's single pass over ps_i
	double pgam[MAX_APPROX_ORDER], pzeta[MAX_APPROX_ORDER];
	int flag[MAX_APPROX_ORDER];    // current flags (inverter_multishift_full.c:58)
	int order, maxiter, cg, max_cg;
	int pending;     // 1 once a ps_i update is waiting (every iteration but the first)
	int done;        // set when maxiter==0 or cg==max_cg: every later kernel is a no-op
	long long active_sum;
};

#ifdef __CUDACC__
// The scalar recurrences of one CG-M iteration, executed by ONE full warp (lane = shift index).  They run either
// as one-warp kernels of their own or -- fused -- as the tail of the kernel whose grid reduction produced the
// scalar (the last block to arrive, once the sum is final).  Multi-rank: the sum over ranks through the peer
// mailboxes is the first thing the warp does.
// after alpha = Re(p, s)  (inverter_multishift_full.c:122-137)
__device__ __forceinline__ void cgm_after_alpha_warp(CgmCtl *c, double *alpha_slot, const RedView &red)
{
	if (red.nranks > 1) p2p_allreduce_warp(alpha_slot, 1, red);
	const int i = threadIdx.x & 31;
	const double alpha = *(volatile double *) alpha_slot;
	const double omega_save = c->omega, delta = c->delta, gammag = c->gammag;
	const int maxiter = c->maxiter, cg = c->cg;
	const double omega = -delta / alpha;
	if (i < maxiter && c->flag[i] == 1) {
		const double zi = c->zeta_i[i], zii = c->zeta_ii[i];
		const double ziii = (zi * zii * omega_save) /
			(omega * gammag * (zi - zii) + zi * omega_save * (1.0 - c->shifts[i] * omega));
		c->zeta_iii[i] = ziii;
		c->omegas[i] = omega * ziii / zii;
	}
	__syncwarp();
	if (i == 0) { c->alpha = alpha; c->omega_save = omega_save; c->omega = omega; c->cg = cg + 1; }
}
// after lambda = (r, r)  (:143-171): gammas, convergence flags, zeta rotation, delta <- lambda
__device__ __forceinline__ void cgm_after_lambda_warp(CgmCtl *c, double *lambda_slot, const RedView &red)
{
	if (red.nranks > 1) p2p_allreduce_warp(lambda_slot, 1, red);
	const int i = threadIdx.x & 31;
	const double lambda = *(volatile double *) lambda_slot;
	const double delta = c->delta, omega = c->omega, source_norm = c->source_norm, residuo = c->residuo;
	const int order = c->order, cg = c->cg, max_cg = c->max_cg;
	const double gammag = lambda / delta;
	int active = 0, was = 0;
	if (i < order) {
		was = c->flag[i];
		if (was == 1) {
			const double zii = c->zeta_ii[i], ziii = c->zeta_iii[i];
			const double gi = gammag * ziii * c->omegas[i] / (zii * omega);
			c->gammas[i] = gi; c->pgam[i] = gi; c->pzeta[i] = ziii;
			const double fact = sqrt(delta * zii * zii / source_norm);
			if (fact < residuo * kSafetyMargin) c->flag[i] = 0;
			else active = 1;
			c->zeta_i[i] = zii;
			c->zeta_ii[i] = ziii;
		}
	}
	const unsigned int wasm = __ballot_sync(0xffffffffu, was == 1);
	const unsigned int act = __ballot_sync(0xffffffffu, active);
	if (i == 0) {
		const int maxiter = act ? 32 - __clz(act) : 0;   // highest still-active shift + 1
		c->maxiter = maxiter; c->pending = 1;
		c->lambda = lambda; c->gammag = gammag; c->delta = lambda;
		c->active_sum += __popc(wasm);
		if (maxiter == 0 || cg >= max_cg) c->done = 1;
	}
}
#endif

// ---- control block of the restarted CG (inverter_full.c:19-132) and of the mixed-precision CG (inverter_mixedp.c:41-181):
// the host arrays/scalars of the reference, device resident, advanced by the kernel that completes the reduction they wait for
struct CgCtl {
	double alpha, delta, lambda, omega, gammag, source_norm, res;
	double stop_factor;        // SAFETY_MARGIN: 0.95 (inverter_full.c:17) / 0.9 (inverter_mixedp.c:141)
	double last_max_res_norm, mixed_delta;     // inverter_mixedp.c:112-114
	int cg, cg_restarted, restarting_every, max_cg;
	int done;                  // the iteration loop has ended: every later kernel of the batch is a no-op
	int mixed;                 // 1: inverter_mixed_precision
	int touch;                 // mixed: THIS iteration refreshes the residual in double precision ("magic touch")
	int no_touch;              // = !touch: skip flag of the double-precision kernels of a mixed iteration
	int touch_next;            // the decision for the next iteration (taken in the lambda tail, acted upon by cg_promote_kernel)
	int paused;                // done was set because the next iteration is a magic touch (enqueued by the host), not because the loop ended
	int magic_touches;
};

#ifdef __CUDACC__
// after alpha = Re(p, s)  (inverter_full.c:82-84, inverter_mixedp.c:104-108): omega, iteration counters
__device__ __forceinline__ void cg_after_alpha_warp(CgCtl *c, double *alpha_slot, const RedView &red)
{
	if (red.nranks > 1) p2p_allreduce_warp(alpha_slot, 1, red);
	if ((threadIdx.x & 31) == 0) {
		const double alpha = *(volatile double *) alpha_slot;
		c->alpha = alpha; c->omega = c->delta / alpha;
		c->cg += 1; c->cg_restarted += 1;
	}
}
// after lambda = (r, r)  (inverter_full.c:92-100, inverter_mixedp.c:133-141): gammag, delta <- lambda, loop condition; mixed:
// the "magic touch" decision of the NEXT iteration, which the reference takes at its start from the same delta (:112-114)
__device__ __forceinline__ void cg_after_lambda_warp(CgCtl *c, double *lambda_slot, const RedView &red)
{
	if (red.nranks > 1) p2p_allreduce_warp(lambda_slot, 1, red);
	if ((threadIdx.x & 31) == 0) {
		const double lambda = *(volatile double *) lambda_slot;
		c->lambda = lambda; c->gammag = lambda / c->delta; c->delta = lambda;
		const bool above = sqrt(lambda / c->source_norm) > c->res * c->stop_factor;
		const bool more = c->mixed ? c->cg < c->max_cg : c->cg_restarted < c->restarting_every;
		const int done = (above && more) ? 0 : 1;
		c->done = done;
		if (c->mixed) {
			if (c->touch) c->magic_touches += 1;
			double lm = c->last_max_res_norm;
			if (lm < lambda) lm = lambda;
			const int touch = lambda < c->mixed_delta * lm ? 1 : 0;
			if (touch) lm = 0.0;
			c->last_max_res_norm = lm; c->touch_next = touch;
		}
	}
}
#endif

// ---- precision traits -------------------------------------------------------------------
template <typename T> struct Prec;
template <> struct Prec<double> { using cplx = double2; };
template <> struct Prec<float> { using cplx = float2; };
template <typename T> using cplx_t = typename Prec<T>::cplx;

// operator launch description (fermion_matrix.c:47-157, :271-718)
template <typename T>
struct DslashArgs {
	const cplx_t<T> *u;       // u[8] : k*9*sizeh + (3r+c)*sizeh + idxh
	cplx_t<T> *out;
	cplx_t<T> *out_host;      // optional second copy of the result, stored straight into (mapped, pinned) HOST memory
	const cplx_t<T> *in;
	const T *ph;              // backfield[8] : k*sizeh + idxh
	const cplx_t<T> *in0;     // epilogue operand (M^+M) or null
	double m2;                // mass^2 + shift for the fused epilogue
	double *partials;         // fused Re(in0 . out) reduction (or null)
	unsigned int *ticket;
	double *result;
	unsigned int ticket_target;   // total blocks contributing to this reduction
	unsigned int partial_offset;  // first partial index of this launch
	const int *skip;          // device flag: nonzero -> kernel is a no-op (solver overrun)
	CgmCtl *cgm;              // EPI_MASS_DOT only: the block that completes the alpha sum also advances the CG-M recurrences
	CgCtl *cg;                // ... or those of the single-system CG
	RedView cgm_red;
	long site_lo, nsites;     // idxh range [site_lo, site_lo+nsites) of a plain launch / of the bulk segment
	int nd0h, nd1, nd2, nd3;
	long vol3h, sizeh;
	// ---- D3 slabs over NVLink peer memory (mr != 0): the launch is segmented by block index into
	//   [nb_top blocks: TOP interior slice -> rank R's slot 0] [nb_bot: BOTTOM interior slice -> rank L's slot 1]
	//   [nb_bulk: the slices in between] [2*nb_unpack: copy of the staged halos of THIS exchange into `out`; before the bulk if unpack_early]
	// any segment may be empty.  A face block stores its sites into `out` AND into the neighbour's staging slot (posted NVLink
	// writes; the data is its own arrival flag, see P2P); faces come first in block order, so the transfer overlaps the rest.
	int mr;
	unsigned int nb_top, nb_bot, nb_bulk, nb_unpack;
	int unpack_early;                            // the unpack blocks come right after the faces (bulk to hide behind) instead of last
	long top_lo, bot_lo;                         // first idxh of the two surface slices
	cplx_t<T> *peer_top, *peer_bot;              // parity-0 staging slot in the neighbour's memory
	long parity_stride;                          // elements between the parity-0 and parity-1 staging areas
	unsigned long long cur;                      // exchanges enqueued before this launch: it consumes `cur` (staged input) and produces cur + 1
	// consumer side: the halo slices of `in` were left in the local staging area by exchange *seq_rw (in_staged), and/or
	// the staged halos of the exchange this launch produces are copied into `out` by the unpack blocks
	int in_staged;
	cplx_t<T> *stage_lo, *stage_hi;              // local slot 0 (lower halo) / slot 1 (upper halo), parity-0 base
	long lower_lo, upper_lo;                     // first idxh of the lower / upper halo slice
};

enum Epilogue { EPI_NONE = 0, EPI_MASS = 1, EPI_MASS_DOT = 2 };

// launches on stream s the operator for output parity `par` over d3 in [d3lo, d3hi)
// face: FACE_NONE   plain launch over the range
//       FACE_TOP    the range is the TOP interior slice, pushed to rank R (slot 0)      } three-queue form
//       FACE_BOTTOM the range is the BOTTOM interior slice, pushed to rank L (slot 1)  }
//       FACE_BOTH   whole local interior in one segmented launch, both faces pushed
//       FACE_BOTH_UNPACK  ... and the staged halos of this exchange copied into `out` by the last blocks of the launch
// halo: HALO_IN_STAGED  the halo slices of `in` are in the staging area (FACE_BOTH* only)
//       HALO_ADVANCE    this launch completes an exchange (both faces pushed): the host counter advances
enum { FACE_NONE = 0, FACE_TOP = 1, FACE_BOTTOM = 2, FACE_BOTH = 3, FACE_BOTH_UNPACK = 4 };
enum { HALO_EAGER = 0, HALO_OUT_STAGED = 1, HALO_IN_STAGED = 2, HALO_ADVANCE = 4, HALO_NO_PUSH = 8 };
template <typename T>
void launch_dslash(int par, int epi, const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, const T *ph,
									 const cplx_t<T> *in0, double m2, int d3lo, int d3hi, int dot_slot,
									 unsigned int ticket_target, unsigned int partial_offset, const int *skip, cudaStream_t s,
									 int face = FACE_NONE, int halo = 0);
unsigned int dslash_blocks(int d3lo, int d3hi);

// full operator with halo handling (acc_Deo/acc_Doe, fermion_matrix.c:159-268); epilogue as above.
// halo (solvers only; honoured when the peer-memory single-launch transport is active, see halo_lazy_ok()):
//   HALO_OUT_STAGED  do not copy the received halos into `out`: the next kernel consumes them from the staging area
//   HALO_IN_STAGED   the halos of `in` are in the staging area (left there by the previous HALO_OUT_STAGED operator)
//   HALO_NO_PUSH     `out` is not exchanged at all (its halo slices are never read: CG-M's s = M^+M p)
template <typename T>
void apply_dslash(int par, int epi, const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, const T *ph,
									const cplx_t<T> *in0, double m2, int dot_slot, const int *skip, int halo = HALO_EAGER);
// out = (mass^2+shift) in - Deo Doe in; optionally leaves Re(in.out) (local, not yet all-reduced) in result(dot_slot).
// The halos of tmp are never unpacked when halo_lazy_ok(); out_staged: nor are those of `out` (CG-M consumes them staged)
// cgm_interior: CG-M on D3 slabs with staged halos -- the halos of `in` are staged (pushed by the kernel that made it), those
// of tmp go through the staging area, `out` is not exchanged
template <typename T>
void apply_mdagm(const cplx_t<T> *u, cplx_t<T> *out, const cplx_t<T> *in, cplx_t<T> *tmp, const T *ph,
								 double m2, int dot_slot, const int *skip, bool cgm_interior = false, bool tmp_eager = false);
void set_spin_timeout_kernels(unsigned long long ns);   // one copy of g_spin_timeout_ns per translation unit
void set_spin_timeout_solvers(unsigned long long ns);
bool halo_lazy_ok();                  // peer-memory transport, single-launch operator, lazy halos enabled
PushView make_pushview(bool on);
void p2p_push_faces(const void *base, size_t elem_bytes, cudaStream_t s);   // push + advance the counter, no unpack

// BLAS-1 (device pointers)
enum BlasOp {
	OP_IN1XFACTOR_PLUS_IN2, OP_SCALE, OP_ADD_FACTOR_X_IN2, OP_IN1XMASS2_MINUS_IN2_MINUS_IN3,
	OP_IN1XMASS_MINUS_IN2, OP_IN1_MINUS_IN2, OP_ASSIGN, OP_ZERO, OP_FACT1_MINUS_IN2, OP_IN1_MINUS_IN2_ALLXFACT,
	OP_INSIDE_LOOP
};
template <typename T>
void blas(BlasOp op, cplx_t<T> *out, const cplx_t<T> *a, const cplx_t<T> *b, const cplx_t<T> *c, double f1,
					cplx_t<T> *out2 = nullptr);
enum RedOp { RED_L2NORM2, RED_REAL_DOT, RED_CPLX_DOT };
// enqueue local reduction into result(slot) (no all-reduce, no host sync)
template <typename T>
void reduce_local(RedOp op, const cplx_t<T> *a, const cplx_t<T> *b, int slot);
// full reduction: local + all-reduce + copy to host + sync; returns {re, im}
template <typename T>
staple_dcomplex reduce_global(RedOp op, const cplx_t<T> *a, const cplx_t<T> *b);
void fetch_results(int slot, int ndoubles, double *host_out);   // all-reduce + D2H + sync

void count_launch(int n = 1);
void release_solver_state();      // staple_solvers.cu: CG-M control block, snapshot buffers, timing events
void release_streamed_state();    // staple_kernels.cu: cached graphs of staple_acc_Doe_Deo_streamed
void blocking_point();            // staple_set_blocking(1): wait for the device here (no-op otherwise and during graph capture)

}   // namespace staple

// Context, geometry, memory boundary and the D3 rank/halo layer of libstaple_b200.so.
//   geometry        <- geometry.h:12-29, Mpi/geometry_multidev.h:6-148 (compile-time macros there)
//   memory boundary <- Include/memory_wrapper.c:14-57 + alloc_vars.c `#pragma acc enter data create`
//   rank layer      <- Mpi/multidev.c:20-108, Mpi/communications.c:34-332 (MPI there, NCCL/NVLink here)
#include "staple_internal.cuh"
#include <dlfcn.h>
#include <cstring>
#include <map>
#include <mutex>

// ---- globals the path reads; weak so that the host program's own definitions win at link time
extern "C" {
__attribute__((weak)) int verbosity_lv = 0;                                // common_defines.h:88
__attribute__((weak)) inv_tricks inverter_tricks = { 0, 0, 0.1, 10000 };   // inverter_tricks.h:4-11
int multishift_invert_iterations = 0;                                     // inverter_wrappers.c:43
__attribute__((weak)) diracTimeContainer dirac_times = { 0.0, 0u };       // tests_and_benchmarks/test_and_benchmarks.c:30
}

namespace staple {

static Ctx g_ctx;
Ctx &ctx() { return g_ctx; }

void require_init(const char *fn)
{
	if (!g_ctx.inited) {
		fprintf(stderr, "libstaple_b200: %s called before staple_init_geometry()\n", fn);
		exit(1);
	}
}
// Blocking mode (staple_set_blocking): every entry point returns with its device work finished, which is what the
// reference's synchronous OpenACC `kernels` regions give a host program that reads results (managed memory, gettimeofday
// timers) right after a call.  Launches that are being captured into a CUDA graph are left alone: the solver that captures
// them synchronises when it replays.
static bool g_blocking = false;
void blocking_point()
{
	if (!g_blocking) return;
	cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
	if (cudaStreamIsCapturing(g_ctx.stream, &st) != cudaSuccess) { cudaGetLastError(); return; }
	if (st == cudaStreamCaptureStatusNone) STAPLE_CUDA_CHECK(cudaDeviceSynchronize());
}
void count_launch(int n) { g_ctx.launches += n; blocking_point(); }

// ------------------------------------------------------------------ present table
struct Entry { size_t bytes; char *dptr; bool owned_host; };
static std::map<uintptr_t, Entry> g_present;   // keyed by host base address
static std::mutex g_present_mu;
// small cache of pointers already classified as device memory
struct DevRange { uintptr_t p; };
static uintptr_t g_devcache[64];

void *resolve_raw(const void *p, const char *what)
{
	if (p == nullptr) return nullptr;
	uintptr_t a = (uintptr_t) p;
	if (!g_present.empty()) {
		auto it = g_present.upper_bound(a);
		if (it != g_present.begin()) {
			--it;
			if (a < it->first + it->second.bytes) return it->second.dptr + (a - it->first);
		}
	}
	unsigned h = (unsigned) ((a >> 8) * 2654435761u) & 63u;
	if (g_devcache[h] == a) return const_cast<void *>(p);
	cudaPointerAttributes at;
	cudaError_t e = cudaPointerGetAttributes(&at, p);
	if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
		g_devcache[h] = a;
		return const_cast<void *>(p);
	}
	cudaGetLastError();
	fprintf(stderr,
					"libstaple_b200: FATAL: argument '%s' (%p) is not present on the device.\n"
					"  Pass a device pointer, or make the host array present with staple_posix_memalign()/\n"
					"  staple_acc_enter_data() (OpenACC `enter data create`).  There is no CPU fallback.\n",
					what, p);
	exit(1);
}

// ------------------------------------------------------------------ NCCL (resolved at run time)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclChar = 0, ncclDouble = 8 };   // nccl.h ncclDataType_t: ncclInt8=0 ... ncclFloat64=8
enum { ncclSum = 0 };

struct Comm {
	void *lib = nullptr;
	ncclComm_t comm = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static Comm *load_nccl()
{
	static Comm c;
	if (c.lib) return &c;
	// prefer the copy already mapped into the process (torch's bundled libnccl.so.2)
	const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };
	for (int i = 0; names[i] && !c.lib; i++) c.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!c.lib) {
		fprintf(stderr, "libstaple_b200: FATAL: cannot load libnccl.so.2 (%s)\n", dlerror());
		exit(1);
	}
#define LOADSYM(field, name)                                                      \
	*(void **) (&c.field) = dlsym(c.lib, name);                                     \
	if (!c.field) { fprintf(stderr, "libstaple_b200: FATAL: NCCL symbol %s missing\n", name); exit(1); }
	LOADSYM(GetUniqueId, "ncclGetUniqueId")
	LOADSYM(CommInitRank, "ncclCommInitRank")
	LOADSYM(CommDestroy, "ncclCommDestroy")
	LOADSYM(Send, "ncclSend")
	LOADSYM(Recv, "ncclRecv")
	LOADSYM(AllReduce, "ncclAllReduce")
	LOADSYM(AllGather, "ncclAllGather")
	LOADSYM(GroupStart, "ncclGroupStart")
	LOADSYM(GroupEnd, "ncclGroupEnd")
	LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
	return &c;
}

#define STAPLE_NCCL_CHECK(c, x)                                                                       \
	do {                                                                                                \
		ncclResult_t r_ = (x);                                                                            \
		if (r_ != 0) {                                                                                    \
			fprintf(stderr, "libstaple_b200: NCCL error %s at %s:%d\n", (c)->GetErrorString(r_), __FILE__,  \
							__LINE__);                                                                              \
			exit(1);                                                                                        \
		}                                                                                                 \
	} while (0)

void allreduce_results(int slot, int ndoubles, cudaStream_t s)
{
	Ctx &c = ctx();
	if (c.nranks <= 1 || c.loopback) return;
	double *p = result(slot);
	if (c.p2p.on && c.p2p.d_redq != nullptr && ndoubles <= 2) { p2p_allreduce(p, ndoubles, s); return; }
	STAPLE_NCCL_CHECK(c.comm, c.comm->AllReduce(p, p, (size_t) ndoubles, ncclDouble, ncclSum, c.comm->comm, s));
}

// communications.c:34-104: for each of `narrays` arrays (stride_elems apart) send the first interior
// `thickness` slices to rank L (received there in the top halo) and the last interior slices to rank R
// (received in the bottom halo).  Offsets follow the reference literally:
//   offset_size = vol3h*HALO_WIDTH ; slab = vol3h*thickness
//   send [offset_size, +slab) -> L      recv [sizeh-offset_size, +slab) <- R
//   send [sizeh-offset_size-slab, +slab) -> R   recv [offset_size-slab, +slab) <- L
void exchange_slices(void *base, size_t elem_bytes, long stride_elems, int narrays, int thickness,
										 cudaStream_t s)
{
	Ctx &c = ctx();
	if (c.nranks <= 1) return;
	const Geom &g = c.g;
	const size_t slab = (size_t) g.vol3h * thickness * elem_bytes;
	const size_t off = (size_t) g.vol3h * g.halo_width * elem_bytes;
	const size_t total = (size_t) g.sizeh * elem_bytes;
	if (c.loopback) {      // this rank is its own L and R neighbour: the same four slab moves as device-to-device copies
		for (int a = 0; a < narrays; a++) {
			char *p = (char *) base + (size_t) a * stride_elems * elem_bytes;
			STAPLE_CUDA_CHECK(cudaMemcpyAsync(p + total - off, p + off, slab, cudaMemcpyDeviceToDevice, s));
			STAPLE_CUDA_CHECK(cudaMemcpyAsync(p + off - slab, p + total - off - slab, slab, cudaMemcpyDeviceToDevice, s));
		}
		return;
	}
	Comm *n = c.comm;
	STAPLE_NCCL_CHECK(n, n->GroupStart());
	for (int a = 0; a < narrays; a++) {
		char *p = (char *) base + (size_t) a * stride_elems * elem_bytes;
		STAPLE_NCCL_CHECK(n, n->Send(p + off, slab, ncclChar, c.rank_L, n->comm, s));
		STAPLE_NCCL_CHECK(n, n->Recv(p + total - off, slab, ncclChar, c.rank_R, n->comm, s));
		STAPLE_NCCL_CHECK(n, n->Send(p + total - off - slab, slab, ncclChar, c.rank_R, n->comm, s));
		STAPLE_NCCL_CHECK(n, n->Recv(p + off - slab, slab, ncclChar, c.rank_L, n->comm, s));
	}
	STAPLE_NCCL_CHECK(n, n->GroupEnd());
}

void fetch_results(int slot, int ndoubles, double *host_out)
{
	Ctx &c = ctx();
	allreduce_results(slot, ndoubles, c.stream);
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(c.h_results + 2 * slot, result(slot), sizeof(double) * ndoubles,
																		cudaMemcpyDeviceToHost, c.stream));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
	for (int i = 0; i < ndoubles; i++) host_out[i] = c.h_results[2 * slot + i];
}


// geometry.h:12-29, Mpi/geometry_multidev.h:6-148 as run-time arithmetic (no CUDA calls: also used by
// staple_geometry_plan, which must work on a host without a GPU)
static bool geom_ok(int n0, int n1, int n2, int n3, int nranks_d3, int halo_width)
{
	if (n0 < 2 || n1 < 1 || n2 < 1 || n3 < 2 || (n0 & 1) || nranks_d3 < 1 || halo_width < 1 || halo_width > 2) return false;
	// the reference flips parities for odd LOC_N3 / odd halo in multi-rank runs (io.c:595-597); not supported
	if (nranks_d3 > 1 && ((n3 & 1) || (halo_width & 1))) return false;
	// kernels use 32-bit site indices (the reference's own idxh is an int, geometry_multidev.h:219)
	if ((double) n0 * n1 * n2 * (n3 + 4) / 2 >= 2147483648.0) return false;
	return true;
}
static void fill_geom(Geom &g, int n0, int n1, int n2, int n3, int nranks_d3, int halo_width)
{
	g.nranks = nranks_d3; g.halo_width = halo_width;
	g.d3_halo = nranks_d3 > 1 ? halo_width : 0;
	g.d3_fhalo = nranks_d3 > 1 ? 1 : 0;
	g.nd0 = n0; g.nd0h = n0 / 2; g.nd1 = n1; g.nd2 = n2; g.loc_n3 = n3; g.nd3 = n3 + 2 * g.d3_halo;
	g.vol3h = (long) n0 * n1 * n2 / 2;
	g.sizeh = g.vol3h * g.nd3;
	const long loc_sizeh = g.vol3h * n3;
	g.r0_lo = nranks_d3 > 1 ? (g.sizeh - loc_sizeh) / 2 : 0;
	g.r0_hi = nranks_d3 > 1 ? (g.sizeh + loc_sizeh) / 2 : g.sizeh;
	g.r1_lo = g.vol3h * (g.d3_halo - g.d3_fhalo);
	g.r1_hi = g.sizeh - g.r1_lo;
}

// the peer-memory mailbox belongs to ONE geometry (staging slots are sized by vol3h): collective release
static void release_p2p()
{
	Ctx &c = ctx();
	if (!c.p2p.mailbox) { c.p2p = P2P(); return; }
	cudaDeviceSynchronize();
	if (c.comm && c.comm->comm) {                          // peers are done with our memory
		double *p = result(kResultSlots - 1);
		STAPLE_NCCL_CHECK(c.comm, c.comm->AllReduce(p, p, 1, ncclDouble, ncclSum, c.comm->comm, c.s_comm));
		STAPLE_CUDA_CHECK(cudaStreamSynchronize(c.s_comm));
	}
	for (int r = 0; r < c.nranks && !c.loopback; r++)
		if (r != c.myrank && c.p2p.peer_mailbox[r]) cudaIpcCloseMemHandle(c.p2p.peer_mailbox[r]);
	cudaFree(c.p2p.mailbox); cudaFree(c.p2p.d_redq);
	c.p2p = P2P();
}

// local part of the peer-memory set-up: mailbox (reduction boxes, staging) and the reduction counter.  Every word of the mailbox
// starts at the resting pattern (all ones): "nothing has arrived"
static void p2p_alloc_local()
{
	Ctx &c = ctx();
	P2P &p = c.p2p;
	const Geom &g = c.g;
	p.vol3h = g.vol3h;
	p.nfb = (g.vol3h + kDslashBlock - 1) / kDslashBlock;
	p.slot_bytes = (size_t) 3 * g.vol3h * 16;
	const size_t mb_bytes = kMailboxHeader + 4 * p.slot_bytes;
	STAPLE_CUDA_CHECK(cudaMalloc((void **) &p.mailbox, mb_bytes));
	STAPLE_CUDA_CHECK(cudaMemset(p.mailbox, 0xFF, mb_bytes));
	STAPLE_CUDA_CHECK(cudaMalloc((void **) &p.d_redq, sizeof(unsigned long long)));
	STAPLE_CUDA_CHECK(cudaMemset(p.d_redq, 0, sizeof(unsigned long long)));
	p.h_seq = 0;
	STAPLE_CUDA_CHECK(cudaDeviceSynchronize());
}
static void p2p_bind_neighbours()
{
	Ctx &c = ctx();
	P2P &p = c.p2p;
	p.stage = p.mailbox + kMailboxHeader;
	p.stage_L = p.peer_mailbox[c.rank_L] + kMailboxHeader; p.stage_R = p.peer_mailbox[c.rank_R] + kMailboxHeader;
	p.on = true;
}

}   // namespace staple

using namespace staple;

// ====================================================================== C ABI
extern "C" {

const char *staple_version(void) { return "staple_b200 0.2 (sm_100a)"; }

static void shutdown_rank_layer(void);
int staple_init_geometry(int n0, int n1, int n2, int n3, int nranks_d3, int halo_width, int device)
{
	Ctx &c = ctx();
	if (!geom_ok(n0, n1, n2, n3, nranks_d3, halo_width)) {
		fprintf(stderr, "libstaple_b200: invalid geometry %dx%dx%dx%d ranks %d halo %d (LOC_N0 even; multi-rank needs even "
						"LOC_N3 and HALO_WIDTH 2)\n", n0, n1, n2, n3, nranks_d3, halo_width);
		return 1;
	}
	if (device >= 0) STAPLE_CUDA_CHECK(cudaSetDevice(device));
	STAPLE_CUDA_CHECK(cudaGetDevice(&c.device));
	Geom &g = c.g;
	{	// a new geometry invalidates everything that was sized or captured for the old one: the peer-memory mailbox (staging
		// slots of 3*vol3h elements -- a larger lattice would write past them in the NEIGHBOURS' memory) and, when the number of
		// ranks changes, the communicator.  Collective, like the staple_init_geometry calls of an SPMD host program are.
		Geom ng;
		fill_geom(ng, n0, n1, n2, n3, nranks_d3, halo_width);
		if (c.inited && c.p2p.mailbox && (ng.vol3h != c.p2p.vol3h || ng.nranks != c.nranks)) release_p2p();
		if (c.inited && (c.comm || c.loopback) && ng.nranks != c.nranks) shutdown_rank_layer();
	}
	fill_geom(g, n0, n1, n2, n3, nranks_d3, halo_width);
	release_streamed_state();      // cached schedules carry the previous geometry in their kernel arguments
	if (!c.own_stream) {
		STAPLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
		STAPLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.s_p, cudaStreamNonBlocking));
		STAPLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.s_m, cudaStreamNonBlocking));
		STAPLE_CUDA_CHECK(cudaStreamCreateWithFlags(&c.s_comm, cudaStreamNonBlocking));
		cudaEvent_t *evs[] = { &c.ev_fork, &c.ev_p, &c.ev_m, &c.ev_comm, &c.ev_misc };
		for (auto e : evs) STAPLE_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		STAPLE_CUDA_CHECK(cudaMalloc(&c.d_tickets, sizeof(unsigned int) * kResultSlots));
		STAPLE_CUDA_CHECK(cudaMemset(c.d_tickets, 0, sizeof(unsigned int) * kResultSlots));
		STAPLE_CUDA_CHECK(cudaMalloc(&c.d_results, sizeof(double) * 2 * kResultSlots));
		STAPLE_CUDA_CHECK(cudaMemset(c.d_results, 0, sizeof(double) * 2 * kResultSlots));
		STAPLE_CUDA_CHECK(cudaHostAlloc(&c.h_results, sizeof(double) * 2 * kResultSlots, cudaHostAllocDefault));
		c.stream = c.own_stream;
	}
	{	// one partial per operator CTA of the largest fused-reduction launch (+ face and unpack blocks of a segmented launch)
		const long need = g.sizeh / kDslashBlock + 4096;
		if (need > c.max_partials) {
			if (c.d_partials) STAPLE_CUDA_CHECK(cudaFree(c.d_partials));
			c.max_partials = need;
			STAPLE_CUDA_CHECK(cudaMalloc(&c.d_partials, sizeof(double) * kResultSlots * 2 * c.max_partials));
		}
	}
	if (nranks_d3 == 1) { c.myrank = 0; c.nranks = 1; c.rank_L = c.rank_R = 0; }
	c.inited = true;
	return 0;
}

// Pure host arithmetic (no GPU needed): everything a host program has to know to shard a lattice in D3
// slabs the way the reference does.  out[0..3] = nd0..nd3, [4] = sizeh, [5] = vol3h (half-sites per d3
// slice), [6..7] = reduction range R0, [8..9] = update range R1 (fermionic_utilities.c:41,188);
// fermion halo exchange in ELEMENTS of each colour array, thickness 1 (communications.c:51-96):
// [10] send-to-L offset, [11] recv-from-R offset, [12] send-to-R offset, [13] recv-from-L offset,
// [14] slab length; [15] = D3_HALO.
int staple_geometry_plan(const int loc_n[4], int nranks_d3, int halo_width, long out[16])
{
	if (!geom_ok(loc_n[0], loc_n[1], loc_n[2], loc_n[3], nranks_d3, halo_width)) return 1;
	Geom g;
	fill_geom(g, loc_n[0], loc_n[1], loc_n[2], loc_n[3], nranks_d3, halo_width);
	out[0] = g.nd0; out[1] = g.nd1; out[2] = g.nd2; out[3] = g.nd3; out[4] = g.sizeh; out[5] = g.vol3h;
	out[6] = g.r0_lo; out[7] = g.r0_hi; out[8] = g.r1_lo; out[9] = g.r1_hi;
	const long off = g.vol3h * g.halo_width, slab = g.vol3h;
	out[10] = off; out[11] = g.sizeh - off; out[12] = g.sizeh - off - slab; out[13] = off - slab; out[14] = slab;
	out[15] = g.d3_halo;
	return 0;
}

// Releases everything the library owns: rank layer, streams, events, reduction scratch, solver control blocks, cached
// graphs.  The present table (device mirrors of host arrays) belongs to the caller's arrays and goes with staple_free /
// staple_acc_exit_data.  staple_init_geometry may be called again afterwards.
void staple_shutdown(void)
{
	Ctx &c = ctx();
	if (!c.inited) return;
	cudaDeviceSynchronize();
	shutdown_rank_layer();
	release_solver_state();
	release_streamed_state();
	cudaStream_t *streams[] = { &c.own_stream, &c.s_p, &c.s_m, &c.s_comm };
	for (auto s : streams) { if (*s) cudaStreamDestroy(*s); *s = nullptr; }
	cudaEvent_t *evs[] = { &c.ev_fork, &c.ev_p, &c.ev_m, &c.ev_comm, &c.ev_misc };
	for (auto e : evs) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
	cudaFree(c.d_tickets); cudaFree(c.d_results); cudaFree(c.d_partials); cudaFreeHost(c.h_results);
	c.d_tickets = nullptr; c.d_results = nullptr; c.d_partials = nullptr; c.h_results = nullptr; c.max_partials = 0;
	c.stream = nullptr; c.cgm_hook = nullptr; c.out_host_hook = nullptr;
	for (auto &e : g_devcache) e = 0;
	cudaGetLastError();
	c.inited = false;
}

long staple_sizeh(void) { require_init("staple_sizeh"); return ctx().g.sizeh; }

void staple_geometry(int nd[4], long ranges[4])
{
	require_init("staple_geometry");
	const Geom &g = ctx().g;
	nd[0] = g.nd0; nd[1] = g.nd1; nd[2] = g.nd2; nd[3] = g.nd3;
	ranges[0] = g.r0_lo; ranges[1] = g.r0_hi; ranges[2] = g.r1_lo; ranges[3] = g.r1_hi;
}

// The handle is used as is: NULL is CUDA's legacy default stream (what a plain C host program and
// torch's default stream use), not "the library stream" -- kernels must be ordered with the caller's
// own copies and allocations on that stream.
void staple_set_stream(void *s)
{
	require_init("staple_set_stream");
	ctx().stream = (cudaStream_t) s;
}
void staple_set_use_graphs(int on) { ctx().use_graphs = on != 0; }
void staple_set_cgm_fuse_tail(int on) { ctx().cgm_fuse_tail = on != 0; }
void staple_set_cg_device_loops(int on) { ctx().cg_device_loops = on != 0; }
void staple_set_streamed_mode(int mode) { ctx().streamed_mode = mode; }
void staple_use_library_stream(void)
{
	require_init("staple_use_library_stream");
	ctx().stream = ctx().own_stream;
}
void *staple_get_stream(void) { return (void *) ctx().stream; }
void staple_synchronize(void) { STAPLE_CUDA_CHECK(cudaStreamSynchronize(ctx().stream)); }
unsigned long long staple_kernel_launches(void) { return ctx().launches; }

// ------------------------------------------------------------------ memory boundary
void staple_acc_enter_data(const void *host, size_t bytes)
{
	std::lock_guard<std::mutex> lk(g_present_mu);
	uintptr_t a = (uintptr_t) host;
	auto it = g_present.find(a);
	if (it != g_present.end()) {
		if (it->second.bytes >= bytes) return;
		cudaFree(it->second.dptr);
		g_present.erase(it);
	}
	Entry e; e.bytes = bytes; e.owned_host = false;
	STAPLE_CUDA_CHECK(cudaMalloc((void **) &e.dptr, bytes));
	g_present[a] = e;
}

void staple_acc_exit_data(const void *host)
{
	std::lock_guard<std::mutex> lk(g_present_mu);
	auto it = g_present.find((uintptr_t) host);
	if (it == g_present.end()) return;
	cudaFree(it->second.dptr);
	g_present.erase(it);
}

void *staple_acc_deviceptr(const void *host) { return resolve_raw(host, "staple_acc_deviceptr"); }

void staple_acc_update_device(const void *host, size_t bytes)
{
	void *d = resolve_raw(host, "staple_acc_update_device");
	if (d == host) return;
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
}

void staple_acc_update_host(void *host, size_t bytes)
{
	void *d = resolve_raw(host, "staple_acc_update_host");
	if (d == host) return;
	STAPLE_CUDA_CHECK(cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx().stream));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
}

void staple_set_blocking(int on) { g_blocking = on != 0; }

// SURVEY 8b option (i): one address valid on host and device.  The reference's `#pragma acc update host/device` (no-ops
// under gcc) become page migrations, so an UNMODIFIED host program runs against the library with only this allocator
// behind posix_memalign_wrapper -- together with staple_set_blocking(1).
int staple_posix_memalign_managed(void **memptr, size_t alignment, size_t size)
{
	(void) alignment;   // cudaMallocManaged returns memory aligned to at least 256 bytes (the reference asks for 128)
	void *p = nullptr;
	if (cudaMallocManaged(&p, size ? size : 1, cudaMemAttachGlobal) != cudaSuccess) { cudaGetLastError(); return 12; /* ENOMEM */ }
	*memptr = p;
	return 0;
}

int staple_posix_memalign(void **memptr, size_t alignment, size_t size)
{
	(void) alignment;   // cudaHostAlloc returns page-aligned memory (>= the reference's ALIGN 128)
	void *h = nullptr;
	if (cudaHostAlloc(&h, size, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return 12; /* ENOMEM */ }
	staple_acc_enter_data(h, size);
	{
		std::lock_guard<std::mutex> lk(g_present_mu);
		g_present[(uintptr_t) h].owned_host = true;
	}
	*memptr = h;
	return 0;
}

void staple_free(void *memptr)
{
	if (!memptr) return;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, memptr) == cudaSuccess && at.type == cudaMemoryTypeManaged) {   // staple_posix_memalign_managed
		cudaDeviceSynchronize(); cudaFree(memptr);
		for (auto &e : g_devcache) e = 0;            // cached "is a device address" answers may lie inside the freed block
		return;
	}
	cudaGetLastError();
	bool owned = false;
	{
		std::lock_guard<std::mutex> lk(g_present_mu);
		auto it = g_present.find((uintptr_t) memptr);
		if (it != g_present.end()) owned = it->second.owned_host;
	}
	staple_acc_exit_data(memptr);
	if (owned) cudaFreeHost(memptr);
}

// ------------------------------------------------------------------ rank layer
int staple_nccl_unique_id(void *id128)
{
	Comm *n = load_nccl();
	ncclUniqueId id;
	STAPLE_NCCL_CHECK(n, n->GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return 0;
}

int staple_init_multidev1D(int myrank, int nranks, const void *id128, int async_comm_fermion)
{
	require_init("staple_init_multidev1D");
	Ctx &c = ctx();
	if (nranks != c.g.nranks) {
		// multidev.c:40-44
		fprintf(stderr, "MPI%02d - NRANKS_D3 is different from nranks: no salamino? Exiting now\n", myrank);
		fprintf(stderr, "MPI%02d - NRANKS_D3 = %d, nranks = %d\n", myrank, c.g.nranks, nranks);
		exit(1);
	}
	if (c.comm || c.loopback) shutdown_rank_layer();   // a second call replaces the rank layer (mailbox, communicator) instead of leaking it
	c.myrank = myrank; c.nranks = nranks;
	c.rank_L = (myrank + (nranks - 1)) % nranks;   // multidev.c:60-61 (SALAMINO ring)
	c.rank_R = (myrank + 1) % nranks;
	c.async_comm_fermion = async_comm_fermion;
	if (nranks > 1) {
		Comm *n = load_nccl();
		ncclUniqueId id;
		memcpy(&id, id128, sizeof(id));
		STAPLE_NCCL_CHECK(n, n->CommInitRank(&n->comm, nranks, id, myrank));
		c.comm = n;
	}
	return 0;
}

// Peer-memory channels (collective: every rank calls it after staple_init_multidev1D).  Allocates the mailbox
// (halo staging + reduction boxes), exports it with CUDA IPC, all-gathers the handles over the NCCL
// communicator (no MPI needed) and maps every rank's mailbox.  Returns 1 if the channels are active.
static void nccl_barrier(cudaStream_t st)
{
	Ctx &c = ctx();
	double *p = result(kResultSlots - 1);
	STAPLE_NCCL_CHECK(c.comm, c.comm->AllReduce(p, p, 1, ncclDouble, ncclSum, c.comm->comm, st));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));
}

int staple_enable_p2p(int on)
{
	require_init("staple_enable_p2p");
	Ctx &c = ctx();
	P2P &p = c.p2p;
	if (!on || c.nranks <= 1) { p.on = false; return 0; }
	c.p2p_single_launch = (on != 2);                   // 2: keep the reference's d3p/d3m/bulk three-queue structure
	c.p2p_unpack_in_kernel = (on != 3);                // 3: single operator launch + separate unpack kernel
	c.p2p_lazy = (on == 1);                            // 1: the solvers consume intermediate halos from the staging area
	const Geom &g = c.g;
	if (p.mailbox && p.vol3h != g.vol3h) release_p2p();  // built for another geometry (staple_init_geometry does this too)
	if (p.stage_L) { p.on = true; return 1; }          // already mapped
	if (c.loopback) {                                  // single process, own memory as both neighbours' memory
		p2p_alloc_local();
		for (int r = 0; r < kMaxRanks; r++) p.peer_mailbox[r] = p.mailbox;
		p2p_bind_neighbours();
		return 1;
	}
	if (!c.comm) { fprintf(stderr, "libstaple_b200: staple_enable_p2p before staple_init_multidev1D\n"); exit(1); }
	if (c.nranks > kMaxRanks) { fprintf(stderr, "libstaple_b200: peer-memory channels support up to %d ranks\n", kMaxRanks); return 0; }
	p2p_alloc_local();
	cudaIpcMemHandle_t mine;
	STAPLE_CUDA_CHECK(cudaIpcGetMemHandle(&mine, p.mailbox));
	cudaIpcMemHandle_t *d = nullptr, all[kMaxRanks];
	STAPLE_CUDA_CHECK(cudaMalloc((void **) &d, (size_t) (c.nranks + 1) * sizeof(mine)));
	STAPLE_CUDA_CHECK(cudaMemcpy(d + c.nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice));
	Comm *n = c.comm;
	cudaStream_t st = c.s_comm;
	STAPLE_NCCL_CHECK(n, n->AllGather(d + c.nranks, d, sizeof(mine), ncclChar, n->comm, st));
	STAPLE_CUDA_CHECK(cudaStreamSynchronize(st));
	STAPLE_CUDA_CHECK(cudaMemcpy(all, d, (size_t) c.nranks * sizeof(mine), cudaMemcpyDeviceToHost));
	STAPLE_CUDA_CHECK(cudaFree(d));
	bool ok = true;
	cudaError_t err = cudaSuccess;
	for (int r = 0; r < c.nranks; r++) {
		if (r == c.myrank) { p.peer_mailbox[r] = p.mailbox; continue; }
		void *ptr = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { cudaGetLastError(); ok = false; err = e; ptr = nullptr; }
		p.peer_mailbox[r] = (char *) ptr;
	}
	// every rank must agree (a partial set-up would deadlock the first exchange): sum of failure flags
	{
		double bad = ok ? 0.0 : 1.0, *slot = result(kResultSlots - 1);
		STAPLE_CUDA_CHECK(cudaMemcpy(slot, &bad, sizeof(double), cudaMemcpyHostToDevice));
		nccl_barrier(st);
		STAPLE_CUDA_CHECK(cudaMemcpy(&bad, slot, sizeof(double), cudaMemcpyDeviceToHost));
		if (bad != 0.0) {
			if (!ok) fprintf(stderr, "MPI%02d - libstaple_b200: CUDA IPC mapping of a peer mailbox failed (%s); halos and "
											 "reductions stay on NCCL\n", c.myrank, cudaGetErrorString(err));
			for (int r = 0; r < c.nranks; r++)
				if (r != c.myrank && p.peer_mailbox[r]) cudaIpcCloseMemHandle(p.peer_mailbox[r]);
			cudaFree(p.mailbox); cudaFree(p.d_redq);
			p = P2P();
			return 0;
		}
	}
	p2p_bind_neighbours();
	return 1;
}

// The D3-slab code path on ONE GPU, one process: the geometry must have been initialised with NRANKS_D3 > 1 (local+halo box,
// ranges R0/R1), this rank becomes its own L and R neighbour and its own memory stands in for the peers' mailboxes.  The lattice
// that results is the single-rank LOC_N0..3 lattice (periodic in d3 with period LOC_N3) stored with halos; every kernel, staging slot
// and counter of the peer-memory transport runs exactly as on N GPUs, minus NVLink.  Used by the single-GPU parity tests and
// for profiling the segmented operator kernel (a multi-rank run cannot be replayed by ncu).
int staple_init_loopback(int p2p_mode)
{
	require_init("staple_init_loopback");
	Ctx &c = ctx();
	if (c.g.nranks <= 1) { fprintf(stderr, "libstaple_b200: staple_init_loopback needs a geometry with NRANKS_D3 > 1\n"); return 1; }
	if (c.comm || c.p2p.mailbox) shutdown_rank_layer();
	c.myrank = 0; c.nranks = c.g.nranks; c.rank_L = c.rank_R = 0; c.async_comm_fermion = 1; c.loopback = true;
	return staple_enable_p2p(p2p_mode) == (p2p_mode != 0) ? 0 : 1;
}

void staple_set_spin_timeout(double seconds)
{
	require_init("staple_set_spin_timeout");
	const unsigned long long ns = seconds <= 0 ? 0ull : (unsigned long long) (seconds * 1e9);
	set_spin_timeout_kernels(ns);
	set_spin_timeout_solvers(ns);
}

// internal name: a host that swaps in host/multidev_staple.c defines shutdown_multidev itself (with MPI_Finalize in it), and the
// executable's definition would pre-empt the library's own calls to the exported symbol
static void shutdown_rank_layer(void)
{
	Ctx &c = ctx();
	release_p2p();
	c.loopback = false;
	if (c.comm && c.comm->comm) {
		cudaDeviceSynchronize();
		c.comm->CommDestroy(c.comm->comm);
		c.comm->comm = nullptr;
	}
	c.comm = nullptr;
}

void shutdown_multidev(void) { shutdown_rank_layer(); }
void staple_shutdown_multidev(void) { shutdown_rank_layer(); }   // for a host that defines shutdown_multidev itself (host/multidev_staple.c)
int staple_rank_layer_ready(void) { return ctx().inited && (ctx().g.nranks <= 1 || ctx().comm != nullptr || ctx().loopback) ? 1 : 0; }
int staple_myrank(void) { return ctx().myrank; }

// fermion borders: 3 colour arrays, thickness FERMION_HALO = 1 (communications.c:158-167)
static void fermion_borders(void *v, size_t elem_bytes, cudaStream_t s)
{
	if (ctx().nranks > 1 && ctx().p2p.on) p2p_exchange_fermion(v, elem_bytes, s);
	else exchange_slices(v, elem_bytes, ctx().g.sizeh, 3, 1, s);
}
// gauge borders: 8 link arrays x rows r0,r1 x 3 columns (communications.c:306-318); r2 is not sent
static void su3_borders(void *u, size_t elem_bytes, int thickness, cudaStream_t s)
{
	const long n = ctx().g.sizeh;
	for (int k = 0; k < 8; k++)
		exchange_slices((char *) u + (size_t) k * 9 * n * elem_bytes, elem_bytes, n, 6, thickness, s);
}

void communicate_fermion_borders(vec3_soa *f)
{
	require_init("communicate_fermion_borders");
	fermion_borders(dev(f, "lnh_fermion"), 16, ctx().stream);
}
void communicate_fermion_borders_hostonly(vec3_soa *f) { communicate_fermion_borders(f); }
void communicate_fermion_borders_f(vec3_soa_f *f)
{
	require_init("communicate_fermion_borders_f");
	fermion_borders(dev(f, "lnh_fermion"), 8, ctx().stream);
}
void communicate_su3_borders(su3_soa *u, int thickness)
{
	require_init("communicate_su3_borders");
	su3_borders(dev(u, "lnh_conf"), 16, thickness, ctx().stream);
}
void communicate_su3_borders_hostonly(su3_soa *u, int thickness) { communicate_su3_borders(u, thickness); }
void communicate_su3_borders_f(su3_soa_f *u, int thickness)
{
	require_init("communicate_su3_borders_f");
	su3_borders(dev(u, "lnh_conf"), 8, thickness, ctx().stream);
}

static void fork_comm()
{
	Ctx &c = ctx();
	STAPLE_CUDA_CHECK(cudaEventRecord(c.ev_fork, c.stream));
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(c.s_comm, c.ev_fork, 0));
}
void communicate_fermion_borders_async(vec3_soa *f, void *, void *)
{
	require_init("communicate_fermion_borders_async");
	fork_comm();
	fermion_borders(dev(f, "lnh_fermion"), 16, ctx().s_comm);
	STAPLE_CUDA_CHECK(cudaEventRecord(ctx().ev_comm, ctx().s_comm));
}
void communicate_su3_borders_async(su3_soa *u, int thickness, void *, void *)
{
	require_init("communicate_su3_borders_async");
	fork_comm();
	su3_borders(dev(u, "lnh_conf"), 16, thickness, ctx().s_comm);
	STAPLE_CUDA_CHECK(cudaEventRecord(ctx().ev_comm, ctx().s_comm));
}
// generated Mpi/sp_communications.c twins
void communicate_fermion_borders_hostonly_f(vec3_soa_f *f) { communicate_fermion_borders_f(f); }
void communicate_su3_borders_hostonly_f(su3_soa_f *u, int thickness) { communicate_su3_borders_f(u, thickness); }
void communicate_fermion_borders_async_f(vec3_soa_f *f, void *, void *)
{
	require_init("communicate_fermion_borders_async_f");
	fork_comm();
	fermion_borders(dev(f, "lnh_fermion"), 8, ctx().s_comm);
	STAPLE_CUDA_CHECK(cudaEventRecord(ctx().ev_comm, ctx().s_comm));
}
void communicate_su3_borders_async_f(su3_soa_f *u, int thickness, void *, void *)
{
	require_init("communicate_su3_borders_async_f");
	fork_comm();
	su3_borders(dev(u, "lnh_conf"), 8, thickness, ctx().s_comm);
	STAPLE_CUDA_CHECK(cudaEventRecord(ctx().ev_comm, ctx().s_comm));
}
void staple_wait_borders(void)
{
	STAPLE_CUDA_CHECK(cudaStreamWaitEvent(ctx().stream, ctx().ev_comm, 0));
}

void staple_last_solve_stats(int *iterations, long long *active, double *loop_ms)
{
	if (iterations) *iterations = ctx().last_iterations;
	if (active) *active = ctx().last_active;
	if (loop_ms) *loop_ms = ctx().last_loop_ms;
}

}   // extern "C"

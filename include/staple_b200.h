/* staple_b200.h -- C ABI of libstaple_b200.so
 *
 * B200-native (CUDA sm_100a) replacement for OpenStaPLE's fermion-solver hot path:
 *   src/OpenAcc/fermion_matrix.[ch]          even/odd staggered Dirac operator
 *   src/OpenAcc/fermionic_utilities.[ch]     BLAS-1 updates and reductions
 *   src/OpenAcc/inverter_multishift_full.*   CG-M (multishift CG)
 *   src/OpenAcc/inverter_full.*              restarted CG
 *   src/OpenAcc/inverter_mixedp.*            FP32 inner / FP64 outer CG
 *   src/OpenAcc/inverter_wrappers.*, inverter_package.*, float_double_conv.*
 *   src/Mpi/communications.c:34-332, multidev.c   D3-slab halo layer
 *
 * Every entry point below keeps the NAME, ARGUMENT ORDER, ARGUMENT MEANING and RETURN
 * CONVENTION of the reference function cited next to it ("ref:" = path:line under the
 * OpenStaPLE source tree).  Lattice arrays are passed as plain pointers to the
 * reference's SoA layouts (struct_c_def.h:16-42):
 *
 *   vec3_soa   : complex c0[sizeh], c1[sizeh], c2[sizeh]           (48*sizeh bytes FP64)
 *   su3_soa    : vec3_soa r0, r1, r2  (r2 stored, never read)      (144*sizeh bytes)
 *   double_soa : double d[sizeh]
 *   u[8], backfield[8] : index k = 2*dir + parity ; arrays of vec3_soa are contiguous
 *   idxh = snum_acc(d0,d1,d2,d3) = (d0 + nd0*(d1 + nd1*(d2 + nd2*d3)))/2      (geometry_multidev.h:219)
 *
 * The reference fixes the geometry at compile time (geom_defines.txt -> -DLOC_N0.. macros).
 * This library takes the same numbers once at run time (staple_init_geometry) so that a
 * single .so serves every lattice; the struct tags are declared incomplete here, and are
 * layout-compatible with the reference's complete types of the same name when both headers
 * are visible.
 *
 * Pointer semantics (ref: OpenACC present table, alloc_vars.c:94-144).  An array argument
 * may be (a) a DEVICE pointer (cudaMalloc / torch tensor) -- used as is; or (b) a HOST
 * pointer previously made "present" with staple_acc_enter_data()/staple_posix_memalign()
 * -- translated to its device mirror, which the caller synchronises with
 * staple_acc_update_device()/staple_acc_update_host() exactly where the reference has
 * `#pragma acc update device/host`.  Anything else aborts with a message (no CPU fallback).
 */
#ifndef STAPLE_B200_H_
#define STAPLE_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ types (layout = reference) */
typedef struct vec3_soa_t vec3_soa;         /* ref: OpenAcc/struct_c_def.h:16-20 */
typedef struct su3_soa_t su3_soa;           /* ref: OpenAcc/struct_c_def.h:35-42 */
typedef struct double_soa_t double_soa;     /* ref: OpenAcc/struct_c_def.h:25-27 */
typedef struct vec3_soa_f_t vec3_soa_f;     /* generated sp_struct_c_def.h */
typedef struct su3_soa_f_t su3_soa_f;
typedef struct float_soa_t float_soa;
typedef struct tamat_soa_t tamat_soa;       /* ref: OpenAcc/struct_c_def.h:45-51  {c01,c02,c12 complex[sizeh]; ic00,ic11 double[sizeh]} */
typedef struct tamat_soa_f_t tamat_soa_f;
typedef struct thmat_soa_t thmat_soa;       /* ref: OpenAcc/struct_c_def.h:52-58  {c01,c02,c12 complex[sizeh]; rc00,rc11 double[sizeh]} */
typedef struct thmat_soa_f_t thmat_soa_f;
typedef struct vec3_t vec3;                 /* ref: OpenAcc/struct_c_def.h:29-33  one colour vector by value {complex c0,c1,c2}, host side */
typedef struct vec3_f_t vec3_f;
typedef struct dcomplex_soa_t dcomplex_soa; /* ref: OpenAcc/struct_c_def.h:21-23  {complex c[sizeh]} */
typedef struct fcomplex_soa_t fcomplex_soa;

#ifndef MAX_APPROX_ORDER
#define MAX_APPROX_ORDER 25                 /* ref: RationalApprox/rationalapprox.h:8 */
#endif
#ifndef RATIONAL_APPROX_H_
typedef struct RationalApprox_t {           /* ref: RationalApprox/rationalapprox.h:15-26 */
	int exponent_num;
	int exponent_den;
	int approx_order;
	double lambda_min;
	double lambda_max;
	int gmp_remez_precision;
	double error;
	double RA_a0;
	double RA_a[MAX_APPROX_ORDER];
	double RA_b[MAX_APPROX_ORDER];
} RationalApprox;
#endif

#ifndef FERMION_PARAMETERS_H
typedef struct ferm_param_t {               /* ref: Include/fermion_parameters.h:9-41 */
	double ferm_mass;
	int degeneracy;
	int number_of_ps;
	char name[30];
	double ferm_charge;
	double ferm_im_chem_pot;
	int index_of_the_first_ps;
	int index_of_the_first_shift;
	double_soa *phases;        /* read by the path */
	double_soa *mag_re;
	double_soa *mag_im;
	int printed_bf_dbg_info;
	float_soa *phases_f;       /* read by the _f path */
	RationalApprox approx_fi_mother, approx_md_mother, approx_li_mother;
	RationalApprox approx_fi, approx_md, approx_li;
} ferm_param;
#endif

#ifndef INVERTER_PACKAGE_H_
typedef struct inverter_package_t {         /* ref: OpenAcc/inverter_package.h:12-29 (passed BY VALUE) */
	const su3_soa *u;
	const su3_soa_f *u_f;
	vec3_soa *ferm_shift_temp;
	vec3_soa_f *ferm_shift_temp_f;
	int nshifts;
	vec3_soa *loc_r, *loc_h, *loc_s, *loc_p;
	vec3_soa_f *loc_r_f, *loc_h_f, *loc_s_f, *loc_p_f;
	vec3_soa_f *out_f;
} inverter_package;
#endif

#ifndef INVERTER_TRICKS_H_
typedef struct inv_tricks_t {               /* ref: Include/inverter_tricks.h:4-11 */
	int singlePInvAccelMultiInv;
	int useMixedPrecision;
	double mixedPrecisionDelta;
	int restartingEvery;
} inv_tricks;
extern inv_tricks inverter_tricks;          /* weak in the library; the host's definition wins */
#endif

/* complex return values: {re, im} pairs, ABI-identical to C99 `double complex` (x86-64 SysV: both travel in xmm0:xmm1).
 * A C translation unit that has <complex.h> in scope (every reference source has, through struct_c_def.h) sees the
 * reference's own return type, so this header and the reference's fermionic_utilities.h can be included together. */
typedef struct { double re, im; } staple_dcomplex;
#if !defined(__cplusplus) && defined(_Complex_I)
typedef double _Complex staple_dcomplex_ret;
#else
typedef staple_dcomplex staple_dcomplex_ret;
#endif
/* MPI_Request arrays of the reference's *_async exchanges (ignored here: the requests are CUDA events owned by the
 * library); typed as the host's MPI_Request when <mpi.h> is in scope so that communications.h can be included too */
#if defined(MPI_VERSION) || defined(MPI_COMM_WORLD)
typedef MPI_Request staple_request;
#else
typedef void staple_request;
#endif

#ifndef INVERTER_SUCCESS
#define INVERTER_SUCCESS 1                  /* ref: OpenAcc/inverter_full.h:12-13 */
#define INVERTER_FAILURE 0
#endif
#ifndef CONVERGENCE_CRITICAL
#define CONVERGENCE_CRITICAL 1              /* ref: OpenAcc/inverter_wrappers.h:11-12 */
#define CONVERGENCE_NONCRITICAL 0
#endif

#ifndef TESTS_AND_BENCHMARK_H_
typedef struct diracTimes_t {               /* ref: tests_and_benchmarks/test_and_benchmarks.h:37-42 */
	double totTransferTime;                   /* wall time rank 0 spent in the BLOCKING fermion halo exchange of acc_Deo/acc_Doe */
	unsigned int count;                       /* (fermion_matrix.c:196-205, :252-259; "makes sense only without async communications") */
} diracTimeContainer;
extern diracTimeContainer dirac_times;      /* weak in the library; the host's definition (test_and_benchmarks.c:30) wins */
#endif

extern int verbosity_lv;                    /* ref: Include/common_defines.h:88 (weak in the library) */
extern int multishift_invert_iterations;    /* ref: OpenAcc/inverter_wrappers.c:43 */

/* ------------------------------------------------------------------ library set-up (new; replaces compile-time macros) */
/* ref: geom_defines.txt / configure.ac:70-132 -> LOC_N0..3, NRANKS_D3; geometry_multidev.h:6-12 HALO_WIDTH
 * (2 for ACTION_TYPE TLSM, 1 for WILSON).  Returns 0 on success.  Selects the CUDA device
 * `device` (ref: deviceinit.c / main.c:231) ; pass -1 to keep the current device. */
int staple_init_geometry(int loc_n0, int loc_n1, int loc_n2, int loc_n3, int nranks_d3,
												 int halo_width, int device);
void staple_shutdown(void);
/* geometry queries: sizeh, nd[4], reduction range R0 and update range R1 (fermionic_utilities.c:41,188) */
long staple_sizeh(void);
void staple_geometry(int nd[4], long ranges[4]);
/* Same numbers without a GPU or an initialised library (host-side sharding logic; ref:
 * geometry_multidev.h:120-148, fermionic_utilities.c:41,188, communications.c:51-96):
 * out[0..3]=nd0..3, [4]=sizeh, [5]=half-sites per d3 slice, [6,7]=R0, [8,9]=R1, fermion halo exchange in
 * elements of one colour array: [10] send->L, [11] recv<-R, [12] send->R, [13] recv<-L, [14] slab length,
 * [15]=D3_HALO.  Returns 0 on success, 1 for an unsupported geometry. */
int staple_geometry_plan(const int loc_n[4], int nranks_d3, int halo_width, long out[16]);
/* Stream all entry points enqueue on.  After staple_init_geometry it is a library-owned non-blocking
 * stream; every host-visible result (reductions, solver returns, staple_acc_update_host) synchronises
 * it.  staple_set_stream() switches to the caller's stream, used as is: NULL means CUDA's legacy default
 * stream (e.g. torch's default stream).  staple_use_library_stream() switches back. */
void staple_set_stream(void *cuda_stream);
void staple_use_library_stream(void);
/* CG-M replays its iteration batches as CUDA graphs on a single GPU (default on; 0 = direct launches). */
void staple_set_use_graphs(int on);
/* CG-M: run the scalar recurrences (ref: inverter_multishift_full.c:122-137, :143-171) in the tail of the kernel whose
 * grid reduction produces alpha / lambda (default on; 0 = one-warp kernels of their own, for A/B comparisons). */
void staple_set_cgm_fuse_tail(int on);
/* ker_invert_openacc / inverter_mixed_precision: iteration loop on the device (control block, recurrences in the tails of the
 * reduction kernels, CUDA-graph batches; default on), or 0 = scalars read back by the host twice per iteration (A/B comparisons;
 * also what runs when the sums over ranks go through NCCL instead of the peer mailboxes). */
void staple_set_cg_device_loops(int on);
void *staple_get_stream(void);
void staple_synchronize(void);
/* number of CUDA kernels launched by this library since start (bench.py "gpu_launches") */
unsigned long long staple_kernel_launches(void);
const char *staple_version(void);

/* ------------------------------------------------------------------ memory boundary */
/* ref: Include/memory_wrapper.c:14-31 posix_memalign_wrapper + alloc_vars.c `#pragma acc enter data create`:
 * pinned host allocation with a device mirror. */
int staple_posix_memalign(void **memptr, size_t alignment, size_t size);
void staple_free(void *memptr);                                   /* ref: memory_wrapper.c:33-57 free_wrapper */
/* The same choke point with CUDA managed memory (SURVEY 8b option i): ONE address valid on host and device, so a host
 * program built with gcc -- where every `#pragma acc update` is a no-op -- needs no other change than this allocator behind
 * posix_memalign_wrapper and staple_set_blocking(1).  openstaple_b200/host/memory_wrapper_staple.c does exactly that for the reference's own
 * deo_doe_test / inverter_multishift_test programs (tests/test_gpu_reference_host.py). */
int staple_posix_memalign_managed(void **memptr, size_t alignment, size_t size);
/* on != 0: every entry point returns with its device work complete (the reference's OpenACC regions are synchronous);
 * default 0: vector kernels are stream-ordered and only host-visible results synchronise. */
void staple_set_blocking(int on);
void staple_acc_enter_data(const void *host, size_t bytes);       /* #pragma acc enter data create(...) */
void staple_acc_exit_data(const void *host);                      /* #pragma acc exit data delete(...)  */
void staple_acc_update_device(const void *host, size_t bytes);    /* #pragma acc update device(...)     */
void staple_acc_update_host(void *host, size_t bytes);            /* #pragma acc update host(...)       */
void *staple_acc_deviceptr(const void *host);                     /* acc_deviceptr()                    */
/* The host-buffer round trip of the stock operator test in ONE call (ref: tests_and_benchmarks/deo_doe_test.c:
 * `#pragma acc update device(in)`; acc_Doe(u,tmp,in); acc_Deo(u,out,tmp); `#pragma acc update host(out)`),
 * software-pipelined over d3 chunks of `chunk_slices` slices (0 = library default): chunk uploads on a copy
 * stream, acc_Doe_d3c / acc_Deo_d3c launches as soon as the +-1 neighbour chunks they read have landed, chunk
 * downloads on a second copy stream -- PCIe runs in both directions while the operator computes.  `in` and `out`
 * are present host arrays (staple_posix_memalign), `tmp` an odd-site scratch vector; returns after `out` is valid
 * on the host.  Results are bit-identical to the unpipelined sequence.  On D3 slabs over the peer-memory transport the same
 * pipeline runs per rank, face slices first (`in` arrives with its halo slices, the faces of `tmp` and `out` are exchanged
 * through the staging area); with NCCL halos the plain sequence incl. the exchanges is executed. */
void staple_acc_Doe_Deo_streamed(const su3_soa *u, vec3_soa *out, const vec3_soa *in, vec3_soa *tmp,
																 const double_soa *backfield, int chunk_slices);
/* how the result of the call above reaches the host: 0 (default) = chunk downloads by the copy engine, 1 = the Deo
 * chunk kernels store it straight into the pinned host buffer (download fused into the operator epilogue). */
void staple_set_streamed_mode(int mode);

/* ------------------------------------------------------------------ rank / halo layer */
/* ref: Mpi/multidev.c:20-108 pre_init_multidev1D + init_multidev1D.  MPI is replaced by NCCL over
 * NVLink: rank 0 creates an id (128 bytes) with staple_nccl_unique_id(), the host program
 * distributes it (MPI_Bcast / torch.distributed.broadcast / file), every rank calls
 * staple_init_multidev1D().  async_comm_fermion mirrors devinfo.async_comm_fermion. */
int staple_nccl_unique_id(void *id128);
int staple_init_multidev1D(int myrank, int nranks, const void *id128, int async_comm_fermion);
/* Optional (collective, after staple_init_multidev1D): fermion halos through NVLink peer memory instead of
 * ncclSend/Recv -- the face blocks of acc_Deo/acc_Doe store their sites straight into the neighbour's staging area (CUDA
 * IPC; posted writes, the stored words are their own arrival flags: no fence, no flag); returns 1
 * if active, 0 if it fell back to NCCL.
 * on = 1: acc_Deo/acc_Doe with their exchange are ONE kernel (face blocks first, bulk, unpack blocks last), and the
 *         solvers leave the halos of their intermediate vectors in the staging area, where the next kernel consumes them;
 * on = 2: the reference's three-queue structure (d3p, d3m, bulk on separate streams) with peer stores;
 * on = 3: one operator kernel + a separate unpack kernel;  on = 4: like 1 without the staged halos inside the solvers.
 * Global sums use the same mailboxes. */
int staple_enable_p2p(int on);
/* Every in-kernel wait of the peer-memory channels is for data that a PEER GPU stores; it is bounded: after `seconds`
 * (default 60; 0 = wait for ever, which is what MPI_Wait does) the waiting kernel prints what it was waiting for and
 * traps, so a dead rank surfaces as a CUDA error on the survivors instead of a hang. */
void staple_set_spin_timeout(double seconds);
/* The D3-slab code path on ONE GPU in one process (tests, profiling): after staple_init_geometry(..., nranks_d3 > 1, ...) this
 * rank becomes its own L and R neighbour and its own memory stands in for the peers' mailboxes; p2p_mode as staple_enable_p2p
 * (0: slab moves as device-to-device copies).  The lattice is the single-rank LOC lattice stored with halos.  Returns 0 on success. */
int staple_init_loopback(int p2p_mode);
void shutdown_multidev(void);                                     /* ref: Mpi/multidev.c:110-114 */
void staple_shutdown_multidev(void);                              /* the same under a name a host that defines shutdown_multidev itself can call */
int staple_rank_layer_ready(void);                                /* 1 once geometry and (for NRANKS_D3 > 1) the rank layer are initialised */
int staple_myrank(void);

void communicate_fermion_borders(vec3_soa *lnh_fermion);          /* ref: Mpi/communications.c:158-167 */
void communicate_fermion_borders_hostonly(vec3_soa *lnh_fermion); /* ref: :171-182 (same exchange, device resident) */
void communicate_su3_borders(su3_soa *lnh_conf, int thickness);   /* ref: :306-318 */
void communicate_su3_borders_hostonly(su3_soa *lnh_conf, int thickness); /* ref: :319-332 */
void communicate_fermion_borders_f(vec3_soa_f *lnh_fermion);      /* generated Mpi/sp_communications.c */
void communicate_su3_borders_f(su3_soa_f *lnh_conf, int thickness);
void communicate_fermion_borders_hostonly_f(vec3_soa_f *lnh_fermion);
void communicate_su3_borders_hostonly_f(su3_soa_f *lnh_conf, int thickness);
/* ref: :257-271 / :273-303 take MPI_Request arrays; here the requests are CUDA events owned by the
 * library: *_async starts the exchange on the comm stream, staple_wait_borders() joins it. */
void communicate_fermion_borders_async(vec3_soa *lnh_fermion, staple_request *unused_send_req, staple_request *unused_recv_req);
void communicate_su3_borders_async(su3_soa *lnh_conf, int thickness, staple_request *unused_send_req, staple_request *unused_recv_req);
void communicate_fermion_borders_async_f(vec3_soa_f *lnh_fermion, staple_request *unused_send_req, staple_request *unused_recv_req);
void communicate_su3_borders_async_f(su3_soa_f *lnh_conf, int thickness, staple_request *unused_send_req, staple_request *unused_recv_req);
void staple_wait_borders(void);

/* ------------------------------------------------------------------ Dirac operator  (ref: OpenAcc/fermion_matrix.h:20-106) */
#define STAPLE_DSLASH_DECL(name) \
	void name(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *backfield); \
	void name##_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *backfield);
STAPLE_DSLASH_DECL(acc_Deo)          /* ref: fermion_matrix.c:159-212 */
STAPLE_DSLASH_DECL(acc_Doe)          /* ref: :214-268 */
STAPLE_DSLASH_DECL(acc_Deo_unsafe)   /* ref: :47-99  */
STAPLE_DSLASH_DECL(acc_Doe_unsafe)   /* ref: :101-157 */
STAPLE_DSLASH_DECL(acc_Deo_bulk)     /* ref: :271-325 */
STAPLE_DSLASH_DECL(acc_Doe_bulk)     /* ref: :327-382 */
STAPLE_DSLASH_DECL(acc_Deo_d3p)      /* ref: :496-550 */
STAPLE_DSLASH_DECL(acc_Doe_d3p)      /* ref: :552-607 */
STAPLE_DSLASH_DECL(acc_Deo_d3m)      /* ref: :609-662 */
STAPLE_DSLASH_DECL(acc_Doe_d3m)      /* ref: :664-718 */
void acc_Deo_d3c(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *backfield, int off3, int thick3);   /* ref: :386-439 */
void acc_Doe_d3c(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *backfield, int off3, int thick3);   /* ref: :441-494 */
void acc_Deo_d3c_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *backfield, int off3, int thick3);
void acc_Doe_d3c_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, const float_soa *backfield, int off3, int thick3);

/* The operator "with a field" of the magnetic-susceptibility measurement (ref: OpenAcc/field_times_fermion_matrix.c:77-232,
 * matvecmul.h:176-260): every link's phase e^{i backfield} is multiplied by the complex per-link field field_re + i field_im
 * (double_soa[8] each, same k = 2*dir+parity indexing).  FP64 only, as in the reference. */
void acc_Deo_wf_unsafe(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *phases, const double_soa *field_re, const double_soa *field_im);
void acc_Doe_wf_unsafe(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *phases, const double_soa *field_re, const double_soa *field_im);
void acc_Deo_wf(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *phases, const double_soa *field_re, const double_soa *field_im);
void acc_Doe_wf(const su3_soa *u, vec3_soa *out, const vec3_soa *in, const double_soa *phases, const double_soa *field_re, const double_soa *field_im);

/* out = (m^2 [+shift]) in - Deo Doe in ; temp1 = odd-site scratch.  ref: fermion_matrix.c:723-746 */
void fermion_matrix_multiplication(const su3_soa *u, vec3_soa *out, const vec3_soa *in, vec3_soa *temp1, ferm_param *pars);
void fermion_matrix_multiplication_shifted(const su3_soa *u, vec3_soa *out, const vec3_soa *in, vec3_soa *temp1, ferm_param *pars, double shift);
void fermion_matrix_multiplication_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, vec3_soa_f *temp1, ferm_param *pars);
void fermion_matrix_multiplication_shifted_f(const su3_soa_f *u, vec3_soa_f *out, const vec3_soa_f *in, vec3_soa_f *temp1, ferm_param *pars, float shift);   /* float: generated sp_fermion_matrix.h */

/* ------------------------------------------------------------------ BLAS-1  (ref: OpenAcc/fermionic_utilities.h:38-123) */
#define STAPLE_BLAS_DECL(V, S) \
	staple_dcomplex_ret scal_prod_global##S(const V *in_vect1, const V *in_vect2);            /* ref: fermionic_utilities.c:85-100,126-146 */ \
	double real_scal_prod_global##S(const V *in_vect1, const V *in_vect2);                /* ref: :101-110,147-161 */ \
	double l2norm2_global##S(const V *in_vect1);                                          /* ref: :111-120,162-175 */ \
	void combine_in1xfactor_plus_in2##S(const V *in_vect1, const double factor, const V *in_vect2, V *out);       /* ref: :180-193 */ \
	void multiply_fermion_x_doublefactor##S(V *in1, const double factor);                 /* ref: :196-207 */ \
	void combine_add_factor_x_in2_to_in1##S(V *in1, const V *in2, double factor);         /* ref: :209-220 */ \
	void combine_in1xferm_mass2_minus_in2_minus_in3##S(const V *in_vect1, double ferm_mass, const V *in_vect2, const V *in_vect3, V *out); /* ref: :225-238 */ \
	void combine_inside_loop##S(V *vect_out, V *vect_r, const V *vect_s, const V *vect_p, const double omega);   /* ref: :241-259 */ \
	void combine_in1xferm_mass_minus_in2##S(const V *in_vect1, double ferm_mass2, V *in_vect2);                  /* ref: :261-272 */ \
	void combine_in1_minus_in2##S(const V *in_vect1, const V *in_vect2, V *out);          /* ref: :274-285 */ \
	void assign_in_to_out##S(const V *in_vect1, V *out);                                  /* ref: :287-299 */ \
	void set_vec3_soa_to_zero##S(V *fermion);                                             /* ref: :302-314 */ \
	void multiple_combine_in1_minus_in2x_factor_back_into_in1##S(V *out, const V *in, const int maxiter, const int *flag, const double *omegas); /* ref: :315-340 */ \
	void multiple1_combine_in1_x_fact1_plus_in2_x_fact2_back_into_in1##S(V *in1, int maxiter, const int *flag, const double *gammas, const V *in2, const double *zeta_iii); /* ref: :342-378 */ \
	void combine_in1_x_fact1_minus_in2_back_into_in2##S(const V *in1, double fact1, V *in2);                     /* ref: :379-400 */ \
	void combine_in1_minus_in2_allxfact##S(const V *in1, const V *in2, double fact, V *out);                     /* ref: :401-415 */ \
	void calc_new_trialsol_for_inversion_in_force##S(int halfLen, V *inout, int nPrecCalculations);              /* ref: :417-455 */
STAPLE_BLAS_DECL(vec3_soa, )
STAPLE_BLAS_DECL(vec3_soa_f, _f)

/* ------------------------------------------------------------------ precision conversion (ref: OpenAcc/float_double_conv.c:9-150) */
void convert_float_to_double_vec3_soa(const vec3_soa_f *f_var, vec3_soa *d_var);
void convert_double_to_float_vec3_soa(const vec3_soa *d_var, vec3_soa_f *f_var);
void convert_float_to_double_su3_soa(const su3_soa_f *f_var, su3_soa *d_var);   /* all 8 links of a conf, rows r0,r1,r2 (ref: :94-150) */
void convert_double_to_float_su3_soa(const su3_soa *d_var, su3_soa_f *f_var);
void convert_float_to_double_tamat_soa(const tamat_soa_f *f_var, tamat_soa *d_var);   /* all 8 links (ref: :150-185) */
void convert_double_to_float_tamat_soa(const tamat_soa *d_var, tamat_soa_f *f_var);
void convert_float_to_double_thmat_soa(const thmat_soa_f *f_var, thmat_soa *d_var);   /* ref: :187-228 */
void convert_double_to_float_thmat_soa(const thmat_soa *d_var, thmat_soa_f *f_var);
void convert_float_to_double_complex_soa(const fcomplex_soa *f_var, dcomplex_soa *d_var);   /* ref: :48-70 */
void convert_double_to_float_complex_soa(const dcomplex_soa *d_var, fcomplex_soa *f_var);
void convert_float_to_double_vec3(const vec3_f *f_var, vec3 *d_var);                  /* host structs, ref: :34-47 */
void convert_double_to_float_vec3(const vec3 *d_var, vec3_f *f_var);
void convert_float_to_double_real_soa(const float_soa *f_var, double_soa *d_var);
void convert_double_to_float_real_soa(const double_soa *d_var, float_soa *f_var);

/* ------------------------------------------------------------------ solvers */
/* CG-M for (M^+M + RA_b[i]) out[i] = in.  ref: OpenAcc/inverter_multishift_full.c:23-252 */
int multishift_invert(const su3_soa *u, ferm_param *pars, RationalApprox *approx, vec3_soa *out,
											const vec3_soa *in, double residuo, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_s,
											vec3_soa *loc_p, vec3_soa *shiftferm, const int max_cg, int *cg_return);
int multishift_invert_f(const su3_soa_f *u, ferm_param *pars, RationalApprox *approx, vec3_soa_f *out,
												const vec3_soa_f *in, double residuo, vec3_soa_f *loc_r, vec3_soa_f *loc_h,
												vec3_soa_f *loc_s, vec3_soa_f *loc_p, vec3_soa_f *shiftferm, const int max_cg, int *cg_return);
/* out = RA_a0 in + sum_i RA_a[i] in_shifted[i].  ref: inverter_multishift_full.c:254-282 */
void recombine_shifted_vec3_to_vec3(const vec3_soa *in_shifted, const vec3_soa *in, vec3_soa *out, const RationalApprox *approx);
void recombine_shifted_vec3_to_vec3_f(const vec3_soa_f *in_shifted, const vec3_soa_f *in, vec3_soa_f *out, const RationalApprox *approx);
/* restarted CG on (M^+M + shift).  ref: OpenAcc/inverter_full.c:19-132 */
int ker_invert_openacc(const su3_soa *u, ferm_param *pars, vec3_soa *solution, const vec3_soa *in, double res,
											 vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_s, vec3_soa *loc_p, const int max_cg,
											 double shift, int *cg_return);
int ker_invert_openacc_f(const su3_soa_f *u, ferm_param *pars, vec3_soa_f *solution, const vec3_soa_f *in, double res,
												 vec3_soa_f *loc_r, vec3_soa_f *loc_h, vec3_soa_f *loc_s, vec3_soa_f *loc_p,
												 const int max_cg, double shift, int *cg_return);
/* FP32 inner CG with FP64 reliable updates.  ref: OpenAcc/inverter_mixedp.c:24-181 */
void combine_add_in2_into_in1_mixed_precision(vec3_soa *in1, const vec3_soa_f *in2);
int inverter_mixed_precision(inverter_package ip, ferm_param *pars, vec3_soa *solution, const vec3_soa *in,
														 double res, const int max_cg, double shift, int *cg_return);
/* ref: OpenAcc/inverter_package.c:18-72 */
void setup_inverter_package_dp(inverter_package *ip, su3_soa *u, vec3_soa *ferm_shift_temp, int nshifts,
															 vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_s, vec3_soa *loc_p);
void setup_inverter_package_sp(inverter_package *ip, su3_soa_f *u_f, vec3_soa_f *ferm_shift_temp_f, int nshifts,
															 vec3_soa_f *loc_r_f, vec3_soa_f *loc_h_f, vec3_soa_f *loc_s_f, vec3_soa_f *loc_p_f,
															 vec3_soa_f *out_f);
/* ref: OpenAcc/inverter_wrappers.c:24-159.  The reference's FP32-accelerated branch uses the globals
 * ferm_shiftmulti_acc_f and aux1_f (alloc_vars); supply them with staple_set_sp_globals(). */
void convergence_messages(int conv_importance, int inverter_status);
int inverter_multishift_wrapper(inverter_package ip, ferm_param *pars, RationalApprox *approx, vec3_soa *out,
																const vec3_soa *in, double res, int max_cg, int convergence_importance);
int inverter_wrapper(inverter_package ip, ferm_param *pars, vec3_soa *out, const vec3_soa *in, double res,
										 int max_cg, double shift, int convergence_importance);
void staple_set_sp_globals(vec3_soa_f *aux1_f, vec3_soa_f *ferm_shiftmulti_acc_f);
/* The wrapper's return value follows the reference literally, stale counter included (inverter_wrappers.c:88-95: it adds the
 * FP32 multishift count once per shift); this is the number of refinement iterations the last accelerated call really spent. */
int staple_last_refinement_iterations(void);
extern vec3_soa_f *aux1_f, *ferm_shiftmulti_acc_f;   /* weak in the library: a host program's own definitions (alloc_vars.c) are used when
                                                        staple_set_sp_globals() was not called */

/* "next" row N1.  ref: OpenAcc/find_min_max.c:21-117 */
double ker_find_max_eigenvalue_openacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_p);
double ker_find_min_eigenvalue_openacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h, vec3_soa *loc_p, double max);   /* ref: :62-98 */
void find_min_max_eigenvalue_soloopenacc(su3_soa *u, ferm_param *pars, vec3_soa *loc_r, vec3_soa *loc_h,
																				 vec3_soa *loc_p1, vec3_soa *loc_p2, double *minmax);

/* "next" row N2: fermion-force outer products, the step after every MD multishift solve.
 * ref: OpenAcc/fermion_force_utilities.c:17-201, fermion_force_utilities.h:16-216; set_su3_soa_to_zero: su3_utilities.c.
 * aux_u / auxmat / pseudo_ipdot are gl(3) fields in the su3_soa[8] layout (all three rows meaningful), ipdot is
 * tamat_soa[8]; all loops cover the local interior d3 in [D3_HALO, nd3-D3_HALO) like the reference's.
 * ker_openacc_compute_fermion_force: for every shift of tpars->approx_md, aux_u += RA_a[i] * outer products of
 * in_shiftmulti[i] and acc_Doe(in_shiftmulti[i]); loc_s / loc_h are scratch (left as the reference leaves them). */
#define STAPLE_FORCE_DECL(S, SU3, VEC3, TAMAT) \
	void set_tamat_soa_to_zero##S(TAMAT *matrix);                                               /* ref: fermion_force_utilities.c:17-29 */ \
	void set_su3_soa_to_zero##S(SU3 *matrix);                                                   /* ref: su3_utilities.c (all 8 links, all sizeh) */ \
	void direct_product_of_fermions_into_auxmat##S(const VEC3 *loc_s, const VEC3 *loc_h, SU3 *aux_u, \
																								 const RationalApprox *approx, int iter);     /* ref: :31-95 */ \
	void multiply_conf_times_force_and_take_ta_nophase##S(const SU3 *u, const SU3 *auxmat, TAMAT *ipdot); /* ref: :97-121 */ \
	void multiply_backfield_times_force##S(ferm_param *tpars, const SU3 *auxmat, SU3 *pseudo_ipdot);       /* ref: :123-153 */ \
	void accumulate_gl3soa_into_gl3soa##S(const SU3 *auxmat, SU3 *pseudo_ipdot);                /* ref: :155-180 */ \
	void ker_openacc_compute_fermion_force##S(const SU3 *u, SU3 *aux_u, const VEC3 *in_shiftmulti, VEC3 *loc_s, \
																						VEC3 *loc_h, ferm_param *tpars);                  /* ref: :183-201 */
STAPLE_FORCE_DECL(, su3_soa, vec3_soa, tamat_soa)
STAPLE_FORCE_DECL(_f, su3_soa_f, vec3_soa_f, tamat_soa_f)

/* "next" row N4: isotropic stout smearing, the producer of the links the operator reads.
 * ref: OpenAcc/stouting.c:27-167, plaquettes.c:196-255 (staples), su3_utilities.c:210-237 (rho TA(U S)),
 * cayley_hamilton.h:24-180 (exp).  ACTION_TYPE TLSM (C_ZERO = 5/3).  stout_wrapper reads the reference's globals
 * act_params.{stout_steps,topo_action,topo_stout_steps}, gl_stout_rho / gl_topo_rho (action.h:6-23, main.c:276-277)
 * and the parking arrays auxbis_conf_acc, glocal_staples, gipdot (+_f) of alloc_vars.h:23,54-55: all are WEAK
 * definitions in the library, so the host program's own definitions win at link time; a host that binds the
 * library dynamically sets them through these symbols.  tstout_conf_acc_arr holds stout_steps consecutive su3_soa[8]. */
#ifndef ACTION_H_
typedef struct action_param_t {             /* ref: OpenAcc/action.h:6-18 */
	double beta; int stout_steps; double stout_rho;
	int topo_action; double barrier; double width; char topo_file_path[20]; int topo_stout_steps; double topo_rho;
} action_param;
extern action_param act_params;
extern double gl_stout_rho, gl_topo_rho;
#endif
extern su3_soa *auxbis_conf_acc, *glocal_staples;
extern tamat_soa *gipdot;
extern su3_soa_f *auxbis_conf_acc_f, *glocal_staples_f;
extern tamat_soa_f *gipdot_f;
#define STAPLE_STOUT_DECL(S, SU3, TAMAT) \
	void calc_loc_staples_nnptrick_all_onlyferms##S(const SU3 *u, SU3 *loc_stap);                /* ref: plaquettes.c:196-255 */ \
	void RHO_times_conf_times_staples_ta_part##S(const SU3 *u, const SU3 *loc_stap, TAMAT *tipdot, int istopo); /* ref: su3_utilities.c:210-237 */ \
	void exp_minus_QA_times_conf##S(const SU3 *tu, const TAMAT *QA, SU3 *tu_out, SU3 *exp_aux);  /* ref: stouting.c:136-167 */ \
	void stout_isotropic##S(const SU3 *u, SU3 *uprime, SU3 *local_staples, SU3 *auxiliary, TAMAT *tipdot, const int istopo); /* ref: stouting.c:74-100 */ \
	void stout_wrapper##S(const SU3 *tconf_acc, SU3 *tstout_conf_acc_arr, const int istopo);     /* ref: stouting.c:27-72 */
STAPLE_STOUT_DECL(, su3_soa, tamat_soa)
STAPLE_STOUT_DECL(_f, su3_soa_f, tamat_soa_f)

/* N4, force side: Sigma' -> Sigma through one smearing level (the stouted fermion force).
 * ref: OpenAcc/stouting.c:171-548 (compute_lambda), :550-1305 (compute_sigma), OpenAcc/fermion_force.c:52-163 (the chain);
 * border exchanges of gl(3), tamat and thmat fields: Mpi/communications.c (thickness 1 inside the chain). */
#define STAPLE_SF_DECL(S, SU3, TAMAT, THMAT) \
	void compute_lambda##S(THMAT *L, const SU3 *SP, const SU3 *U, const TAMAT *QA, SU3 *TMP);                 /* ref: stouting.c:516-548 */ \
	void compute_sigma##S(const THMAT *L, const SU3 *U, SU3 *S_, const TAMAT *QA, SU3 *TMP, const int istopo); /* ref: stouting.c:1175-1305 */ \
	void communicate_gl3_borders##S(SU3 *lnh_conf, int thickness); \
	void communicate_tamat_soa_borders##S(TAMAT *lnh_ipdot, int thickness); \
	void communicate_thmat_soa_borders##S(THMAT *lnh_ipdot, int thickness); \
	void compute_sigma_from_sigma_prime_backinto_sigma_prime##S(SU3 *Sigma, THMAT *Lambda, TAMAT *QA, const SU3 *U, SU3 *TMP, \
																															 const int istopo);                          /* ref: fermion_force.c:52-163 */
STAPLE_SF_DECL(, su3_soa, tamat_soa, thmat_soa)
STAPLE_SF_DECL(_f, su3_soa_f, tamat_soa_f, thmat_soa_f)

/* ------------------------------------------------------------------ callers of the path: whole fermion force, even/odd inversion */
/* The MD fermion force from thin links to the momenta's time derivative (ref: OpenAcc/fermion_force.c:166-357, called by
 * md_integrator.c:536-715): stout_wrapper -> for every flavour and pseudofermion inverter_multishift_wrapper on approx_md +
 * ker_openacc_compute_fermion_force, multiply_backfield_times_force per flavour -> Sigma' -> Sigma through every stout
 * level -> multiply_conf_times_force_and_take_ta_nophase.  Argument list = the reference's with STOUT_FERMIONS defined
 * (common_defines.h:58).  Globals read like the reference: act_params.stout_steps, inverter_tricks, md_parameters.
 * recycleInvsForce (the reference's recycle branch exits with "not implemented correctly", so does this), the parking
 * arrays aux_th / aux_ta (alloc_vars.h:56-57) and conf_acc_f (FP32 links when singlePInvAccelMultiInv is set);
 * nMdInversionPerformed is incremented.  All are WEAK in the library.  debug_settings diagnostics / dbg prints
 * (fermion_force.c:325-354) are not part of the path and are not written.  The _f twin (generated sp_fermion_force.c:
 * 158-300) calls multishift_invert_f directly and takes `float res`. */
#ifndef MD_PARAMETERS_H
typedef struct md_param_t {                 /* ref: OpenAcc/md_parameters.h:6-19 */
	int no_md; int gauge_scale; double t; double residue_metro; double expected_max_eigenvalue; int singlePrecMD;
	double residue_md; int max_cg_iterations; int recycleInvsForce; int extrapolateInvsForce;
} md_param;
extern md_param md_parameters;
extern int nMdInversionPerformed;
#endif
extern thmat_soa *aux_th;
extern tamat_soa *aux_ta;
extern thmat_soa_f *aux_th_f;
extern tamat_soa_f *aux_ta_f;
extern su3_soa_f *conf_acc_f;
void fermion_force_soloopenacc(su3_soa *tconf_acc, su3_soa *tstout_conf_acc_arr, su3_soa *gl3_aux, tamat_soa *tipdot_acc,
															 ferm_param *tfermion_parameters, int tNDiffFlavs, const vec3_soa *ferm_in_acc, double res,
															 su3_soa *taux_conf_acc, vec3_soa *tferm_shiftmulti_acc, inverter_package ipt, const int max_cg);
void fermion_force_soloopenacc_f(su3_soa_f *tconf_acc, su3_soa_f *tstout_conf_acc_arr, su3_soa_f *gl3_aux, tamat_soa_f *tipdot_acc,
																 ferm_param *tfermion_parameters, int tNDiffFlavs, const vec3_soa_f *ferm_in_acc, float res,
																 su3_soa_f *taux_conf_acc, vec3_soa_f *tferm_shiftmulti_acc, inverter_package ipt, const int max_cg);
/* Full-lattice solve (D + m) out = in from the even/odd solver (ref: Meas/ferm_meas.c:50-72):
 * phi_e = m in_e - Deo in_o ; out_e = (M^+M)^-1 phi_e (inverter_wrapper, CONVERGENCE_CRITICAL) ; out_o = (in_o - Doe out_e)/m.
 * phi_e / phi_o are parking vectors. */
void eo_inversion(inverter_package ip, ferm_param *tfermions_parameters, double res, int max_cg, vec3_soa *in_e, vec3_soa *in_o,
									vec3_soa *out_e, vec3_soa *out_o, vec3_soa *phi_e, vec3_soa *phi_o);

/* ------------------------------------------------------------------ introspection for benches/tests */
/* statistics of the last multishift_invert[_f] call: iterations, sum over iterations of active
 * shifts (for the algorithmic-bytes roofline figure), device time of the loop in ms. */
void staple_last_solve_stats(int *iterations, long long *active_shift_iterations, double *loop_ms);

#ifdef __cplusplus
}
#endif
#endif /* STAPLE_B200_H_ */

/* HOST-SIDE GLUE, compiled by the host program's own C compiler IN PLACE OF src/Include/memory_wrapper.c -- the one file a
 * maintainer swaps to run OpenStaPLE's programs on libstaple_b200.so without editing a single reference source:
 * memory_wrapper.c is the allocation choke point every lattice array of the reference goes through (alloc_vars.c), and the
 * first allocation happens after pre_init_multidev1D / init_multidev1D (main.c:113-260, deo_doe_test.c:100-163), so this is
 * where
 *   (a) the compile-time geometry (LOC_N0..3, NRANKS_D3 of geom_defines.txt) is handed to the library (staple_init_geometry),
 *   (b) on NRANKS_D3 > 1 the rank layer is joined: rank 0 makes the NCCL id, the host's own MPI broadcasts it, every rank
 *       calls staple_init_multidev1D + staple_enable_p2p (the host keeps src/Mpi/multidev.c -- MPI_Init, devinfo -- as it is),
 *   (c) arrays are allocated through the library as CUDA managed memory: one address valid on host and device, so the
 *       program's `#pragma acc update host/device` (no-ops under gcc) need no replacement, and
 *   (d) the reference's synchronous semantics are requested (staple_set_blocking).
 * Everything else in such a binary -- main(), the input-file parser, the dSFMT generators, backfield phases, IO, the gauge
 * sector -- is the reference's own object code; the hot path is libstaple_b200.so.  oracle/build_ref_host.sh builds the
 * reference's deo_doe_test, inverter_multishift_test and RHMC main exactly this way (1 and 2 ranks) and the GPU tests run them
 * against the pure-reference builds (tests/test_gpu_reference_host.py, test_gpu_zz_*.py). */
#include <stdio.h>
#include <stdlib.h>
#ifndef NRANKS_D3
#define NRANKS_D3 1
#endif
#if NRANKS_D3 > 1
#include "mpi.h"
#endif
#include "staple_b200.h"

/* statistics main.c prints (main.c:274,1245); the test programs never read them, so only the totals are kept */
struct memory_allocated_t;
struct memory_allocated_t *memory_allocated_base = NULL;
size_t memory_used = 0, max_memory_used = 0;

static void init_once(void)
{
	static int done = 0;
	if (done) return;
	done = 1;
	if (staple_rank_layer_ready()) {          /* host/multidev_staple.c (in place of src/Mpi/multidev.c) has done (a) and (b) already */
		staple_set_blocking(1);
		fprintf(stderr, "memory_wrapper_staple: hot path served by %s (rank layer joined in init_multidev1D)\n", staple_version());
		return;
	}
	const char *dev = getenv("STAPLE_DEVICE");
	int rank = 0, nranks = 1;
#if NRANKS_D3 > 1
	/* the reference has called pre_init_multidev1D / init_multidev1D (MPI_Init included) before its first allocation */
	MPI_Comm_rank(MPI_COMM_WORLD, &rank); MPI_Comm_size(MPI_COMM_WORLD, &nranks);
	if (nranks != NRANKS_D3) { fprintf(stderr, "memory_wrapper_staple: built for %d ranks, started with %d\n", NRANKS_D3, nranks); exit(1); }
#endif
	if (staple_init_geometry(LOC_N0, LOC_N1, LOC_N2, LOC_N3, nranks, 2 /* HALO_WIDTH, TLSM */, (dev ? atoi(dev) : 0) + rank) != 0) {
		fprintf(stderr, "memory_wrapper_staple: staple_init_geometry failed\n"); exit(1);
	}
#if NRANKS_D3 > 1
	{	/* rank 0 makes the NCCL id, MPI carries it, every rank joins the D3 ring */
		char id[128];
		if (rank == 0) staple_nccl_unique_id(id);
		MPI_Bcast(id, 128, MPI_CHAR, 0, MPI_COMM_WORLD);
		staple_init_multidev1D(rank, nranks, id, 1);
		staple_enable_p2p(getenv("STAPLE_P2P") ? atoi(getenv("STAPLE_P2P")) : 1);    /* NVLink peer memory; 0 keeps NCCL send/recv */
	}
#endif
	staple_set_blocking(1);
	fprintf(stderr, "memory_wrapper_staple: hot path served by %s\n", staple_version());
}

/* same signature as memory_wrapper.c:14; every lattice array of alloc_vars.c arrives here */
int posix_memalign_wrapper(void **memptr, size_t alignment, size_t size, const char *varname)
{
	(void) varname;
	init_once();
	if (staple_posix_memalign_managed(memptr, alignment, size) != 0) return 12;
	memory_used += size;
	if (max_memory_used < memory_used) max_memory_used = memory_used;
	return 0;
}

/* same signature as memory_wrapper.c:33 */
void free_wrapper(void *memptr) { staple_free(memptr); }

/* HOST-SIDE GLUE, compiled by the host program's own C compiler (against ITS mpi.h and the reference's headers) IN PLACE OF
 * src/Mpi/multidev.c.  It defines what that file defines -- `devinfo`, pre_init_multidev1D, init_multidev1D,
 * shutdown_multidev -- with the reference's behaviour (MPI_Init, the replica communicator, the D3 "salamino" ring, the
 * messages; ref: src/Mpi/multidev.c:20-114) and, at the end of init_multidev1D, joins the library's rank layer: geometry,
 * NCCL id broadcast over the host's MPI communicator, peer-memory channels.  These three entry points cannot live inside
 * libstaple_b200.so: `dev_info` contains an MPI_Comm and an MPI_MAX_PROCESSOR_NAME-sized array, whose size and meaning belong
 * to the host's MPI.  With this file and memory_wrapper_staple.c swapped in, no reference source is edited.
 * Optional: a host that keeps its own multidev.c needs memory_wrapper_staple.c only (it joins the rank layer lazily). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Mpi/multidev.h"          /* -I<reference>/src: dev_info, MPI_PRINTF*, geometry_multidev.h's rank <-> coordinates helpers */
#include "OpenAcc/geometry.h"      /* geom_par */
#include "Include/rep_info.h"
/* the six entry points of include/staple_b200.h used here (that header is not included: it shares type names with the reference's) */
int staple_init_geometry(int n0, int n1, int n2, int n3, int nranks_d3, int halo_width, int device);
int staple_nccl_unique_id(void *id128);
int staple_init_multidev1D(int myrank, int nranks, const void *id128, int async_comm_fermion);
int staple_enable_p2p(int on);
void staple_shutdown_multidev(void);

dev_info devinfo;

#ifdef MULTIDEVICE
#include <mpi.h>
extern int verbosity_lv;

static void die_if_mismatch(const char *what, const char *macro, int expected, const char *field, int got)
{
	if (expected == got) return;
	MPI_PRINTF1("%s. Exiting now\n", what);
	MPI_PRINTF1("%s = %d, %s = %d\n", macro, expected, field, got);
	exit(1);
}

/* ref: multidev.c:20-57 */
void pre_init_multidev1D(dev_info *mdi)
{
	MPI_Init(NULL, NULL);
	MPI_Comm_rank(MPI_COMM_WORLD, &mdi->myrank_world);
	MPI_Comm_size(MPI_COMM_WORLD, &mdi->nranks_world);
	MPI_Get_processor_name(mdi->processor_name, &mdi->namelen);
	/* one communicator per replica: NRANKS_D3 consecutive world ranks each */
	mdi->num_replicas = mdi->nranks_world / NRANKS_D3;
	if (mdi->num_replicas > 1) {
		mdi->replica_idx = mdi->myrank_world / NRANKS_D3;
		MPI_Comm_split(MPI_COMM_WORLD, mdi->replica_idx, mdi->myrank_world, &mdi->mpi_comm);
		MPI_Comm_rank(mdi->mpi_comm, &mdi->myrank);
		MPI_Comm_size(mdi->mpi_comm, &mdi->nranks);
	} else {
		mdi->replica_idx = 0;
		mdi->mpi_comm = MPI_COMM_WORLD;
		mdi->myrank = mdi->myrank_world;
		mdi->nranks = mdi->nranks_world;
	}
	die_if_mismatch("NRANKS_D3 is different from nranks: no salamino?", "NRANKS_D3", NRANKS_D3, "nranks", mdi->nranks);
	die_if_mismatch("NREPLICAS is different from devinfo.num_replicas", "NREPLICAS", NREPLICAS, "num_replicas", mdi->num_replicas);
	if (verbosity_lv > 2) MPI_PRINTF0("- Called MPI_Init\n");
}

/* ref: multidev.c:59-108, then the library's rank layer */
void init_multidev1D(dev_info *mdi)
{
	int dir, where[4];
	mdi->myrank_L = (mdi->myrank + mdi->nranks - 1) % mdi->nranks;      /* salamino ring */
	mdi->myrank_R = (mdi->myrank + 1) % mdi->nranks;
	mdi->node_subrank = mdi->myrank % mdi->proc_per_node;
	if (mdi->num_replicas > 1) sprintf(mdi->myrankstr, "MPI%02d", mdi->myrank);
	else sprintf(mdi->myrankstr, "MPI%02d:%02d", mdi->replica_idx, mdi->myrank);
	MPI_PRINTF1("of \"%02d\" tasks running on host \"%s\", replica index: %d, local rank: %d, rankL: %d, rankR: %d\n", mdi->nranks_world,
							mdi->processor_name, mdi->replica_idx, mdi->node_subrank, mdi->myrank_L, mdi->myrank_R);
	mdi->myrank4int = xyzt_rank(mdi->myrank);
	where[0] = mdi->myrank4int.d0; where[1] = mdi->myrank4int.d1; where[2] = mdi->myrank4int.d2; where[3] = mdi->myrank4int.d3;
	for (dir = 0; dir < 4; dir++) {
		const int n = geom_par.nranks[dir];
		mdi->nnranks[dir][0] = (where[dir] + n - 1) % n;
		mdi->nnranks[dir][1] = (where[dir] + 1) % n;
	}
	mdi->gl_loc_origin4int = gl_loc_origin_from_rank(mdi->myrank);
	mdi->halo_widths0123[0] = D0_HALO; mdi->halo_widths0123[1] = D1_HALO; mdi->halo_widths0123[2] = D2_HALO; mdi->halo_widths0123[3] = D3_HALO;
	mdi->origin_0123[0] = mdi->gl_loc_origin4int.d0; mdi->origin_0123[1] = mdi->gl_loc_origin4int.d1;
	mdi->origin_0123[2] = mdi->gl_loc_origin4int.d2; mdi->origin_0123[3] = mdi->gl_loc_origin4int.d3;
	if (verbosity_lv > 2) {
		MPI_PRINTF0("- Finished init_multidev1D\n");
		MPI_PRINTF1("- Origin(%d,%d,%d,%d)", mdi->origin_0123[0], mdi->origin_0123[1], mdi->origin_0123[2], mdi->origin_0123[3]);
	}
	/* ---- the library's rank layer: geometry of geom_defines.txt, device as main.c:231 picks it, NCCL id over the replica's
	 * communicator, NVLink peer memory for the fermion halos and the global sums */
	{
		const char *dev0 = getenv("STAPLE_DEVICE"), *p2p = getenv("STAPLE_P2P");
		const int ppn = mdi->proc_per_node > 0 ? mdi->proc_per_node : mdi->nranks;
		const int device = ((dev0 ? atoi(dev0) : mdi->single_dev_choice) + mdi->myrank_world) % ppn;
		char id[128];
		if (staple_init_geometry(LOC_N0, LOC_N1, LOC_N2, LOC_N3, NRANKS_D3, HALO_WIDTH, device) != 0) {
			MPI_PRINTF0("staple_init_geometry failed. Exiting now\n");
			exit(1);
		}
		memset(id, 0, sizeof(id));
		if (0 == mdi->myrank) staple_nccl_unique_id(id);
		MPI_Bcast(id, 128, MPI_CHAR, 0, mdi->mpi_comm);
		staple_init_multidev1D(mdi->myrank, mdi->nranks, id, mdi->async_comm_fermion);
		staple_enable_p2p(p2p ? atoi(p2p) : 1);
	}
}

/* ref: multidev.c:110-114; the host's definition is the one its main() calls, the library's own is reached by its other name */
void shutdown_multidev()
{
	staple_shutdown_multidev();
	MPI_PRINTF0("Finalizing...\n");
	MPI_Finalize();
}
#endif

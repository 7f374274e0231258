/* TEST INFRASTRUCTURE ONLY -- the handful of MPI entry points a single-rank (NRANKS_D3 = 1) build of the reference's test
 * programs still references (oracle/build_ref_host.sh); the image has no MPI. */
#include <stdlib.h>
#include "mpi.h"
int MPI_Init(int *a, char ***b) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int e) { exit(e ? e : 1); }
int MPI_Barrier(MPI_Comm c) { return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *n) { *n = 1; return 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int r, MPI_Comm c) { return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { memcpy(r, s, (size_t) n * (size_t) t); return 0; }

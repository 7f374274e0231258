/* TEST INFRASTRUCTURE ONLY -- a minimal MPI for running the reference's NRANKS_D3 > 1 programs as N local processes (the
 * image has no MPI): exactly the calls the reference makes (src/Mpi, main.c, HPT_utilities.c), over Unix-domain sockets.
 * Launch: oracle/mpi_mini/mpirun.py -n N prog args...  (sets MINI_MPI_RANK / MINI_MPI_SIZE / MINI_MPI_DIR). */
#ifndef MPI_MINI_H_
#define MPI_MINI_H_
#include <string.h>
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG; } MPI_Status;
#define MPI_VERSION 3
#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_FLOAT 3
#define MPI_DOUBLE 4
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MAX_PROCESSOR_NAME 64
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
int MPI_Init(int *, char ***); int MPI_Finalize(void); int MPI_Abort(MPI_Comm, int);
int MPI_Barrier(MPI_Comm); int MPI_Comm_rank(MPI_Comm, int *); int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *); int MPI_Get_processor_name(char *, int *);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Wait(MPI_Request *, MPI_Status *);
#endif

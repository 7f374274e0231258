/* Minimal single-process MPI stand-in used only to compile the reference's
 * hot-path sources with gcc for the CPU oracle. */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H
#include <string.h>
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Request;
typedef struct { int s; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_FLOAT 4
#define MPI_INT 5
#define MPI_CHAR 1
#define MPI_SUM 0
#define MPI_MAX_PROCESSOR_NAME 64
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
int MPI_Init(int*, char***); int MPI_Finalize(void); int MPI_Abort(MPI_Comm,int);
int MPI_Barrier(MPI_Comm); int MPI_Comm_rank(MPI_Comm,int*); int MPI_Comm_size(MPI_Comm,int*);
int MPI_Comm_split(MPI_Comm,int,int,MPI_Comm*); int MPI_Get_processor_name(char*,int*);
int MPI_Bcast(void*,int,MPI_Datatype,int,MPI_Comm);
int MPI_Allreduce(const void*,void*,int,MPI_Datatype,MPI_Op,MPI_Comm);
int MPI_Sendrecv(const void*,int,MPI_Datatype,int,int,void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Status*);
int MPI_Isend(const void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Request*);
int MPI_Irecv(void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Request*);
int MPI_Send(const void*,int,MPI_Datatype,int,int,MPI_Comm);
int MPI_Recv(void*,int,MPI_Datatype,int,int,MPI_Comm,MPI_Status*);
int MPI_Waitall(int,MPI_Request*,MPI_Status*);
int MPI_Wait(MPI_Request*,MPI_Status*);
#endif

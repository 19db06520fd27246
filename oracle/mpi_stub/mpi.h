/*
 * mpi.h -- the 15-call MPI subset the reference hot path uses, for building the
 * reference sources in place where no MPI installation exists.  TEST INFRASTRUCTURE (oracle/).
 *
 * Implemented by oracle/mini_mpi.c:
 *   - MINI_MPI_NP unset or 1 : a single rank, every collective is a local copy / no-op;
 *   - MINI_MPI_NP = P > 1    : MPI_Init() forks P-1 children that share one anonymous mmap;
 *     collectives go through that mapping, reductions are summed in rank order (deterministic).
 * Collectives complete at the I-call; MPI_Wait() is a no-op (legal: the standard lets a
 * non-blocking collective finish at any time between its start and its Wait).
 */
#ifndef ORACLE_MINI_MPI_H
#define ORACLE_MINI_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_SUM 1
#define MPI_SUCCESS 0
#define MPI_MAX_PROCESSOR_NAME 128
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Get_processor_name(char *name, int *len);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm comm);
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt,
               int root, MPI_Comm comm);
int MPI_Iallgatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, const int *rcounts,
                    const int *displs, MPI_Datatype rt, MPI_Comm comm, MPI_Request *req);
int MPI_Iallreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm,
                   MPI_Request *req);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
/* only referenced by the dead MPI_csr_spmv_async (matrix.c:450-492); abort if ever reached */
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitsome(int incount, MPI_Request *reqs, int *outcount, int *indices, MPI_Status *sts);

#ifdef __cplusplus
}
#endif
#endif

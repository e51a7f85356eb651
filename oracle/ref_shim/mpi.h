/*
 * mpi.h -- single-process stand-in for the seven MPI entry points the
 * reference MPI solver uses (/root/reference/code/MPI/cg.cc:50-142,
 * cg_main.cc:15-20,67), plus the three the libcgb200 binding of INTEGRATION.md B adds
 * (MPI_Allgather of the exchange blobs, MPI_Barrier, MPI_Abort; ref_shim/cg_cgb.cc).  There is no MPI in this image; this header lets the
 * reference sources compile UNMODIFIED.  Test infrastructure only (oracle/).
 *
 * Rank count: 1.  The definitions live in mpi_single.cc so that a multi-rank
 * shim could replace them without touching this header.
 */
#ifndef CGB_ORACLE_MPI_STUB_H
#define CGB_ORACLE_MPI_STUB_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_BYTE 1
#define MPI_SUM 1
#define MPI_THREAD_SINGLE 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void *)1)

#ifdef __cplusplus
extern "C" {
#endif

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op,
                  MPI_Comm comm);
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                   const int *recvcounts, const int *displs, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int *recvcounts, const int *displs, MPI_Datatype recvtype, int root,
                MPI_Comm comm);

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                  int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int errorcode);

#ifdef __cplusplus
}
#endif
#endif

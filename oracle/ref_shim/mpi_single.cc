// mpi_single.cc -- one-rank bodies for oracle/ref_shim/mpi.h.  Test infrastructure only.
//
// Side channel for the parity tests (never read by the reference code itself):
//   CGREF_XOUT=<path>  MPI_Gatherv writes the gathered x (cg.cc:140-142) as raw doubles.
#include "mpi.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {

int MPI_Init_thread(int *, char ***, int required, int *provided)
{
    if (provided) *provided = required;
    return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = 1; return MPI_SUCCESS; }

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op, MPI_Comm)
{
    if (sendbuf != MPI_IN_PLACE) std::memcpy(recvbuf, sendbuf, (size_t)count * (size_t)dt);
    return MPI_SUCCESS;
}

int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                   const int *, const int *displs, MPI_Datatype, MPI_Comm)
{
    char *dst = static_cast<char *>(recvbuf) + (size_t)displs[0] * (size_t)sendtype;
    std::memcpy(dst, sendbuf, (size_t)sendcount * (size_t)sendtype);
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int *, const int *displs, MPI_Datatype, int, MPI_Comm)
{
    char *dst = static_cast<char *>(recvbuf) + (size_t)displs[0] * (size_t)sendtype;
    std::memcpy(dst, sendbuf, (size_t)sendcount * (size_t)sendtype);
    if (const char *path = std::getenv("CGREF_XOUT")) {
        if (FILE *f = std::fopen(path, "wb")) {
            std::fwrite(recvbuf, (size_t)sendtype, (size_t)sendcount, f);
            std::fclose(f);
        }
    }
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int,
                  MPI_Datatype, MPI_Comm)
{
    std::memcpy(recvbuf, sendbuf, (size_t)sendcount * (size_t)sendtype);
    return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) { std::_Exit(errorcode); }

} // extern "C"

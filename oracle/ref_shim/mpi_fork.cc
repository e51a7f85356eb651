// mpi_fork.cc -- multi-rank bodies for oracle/ref_shim/mpi.h on ONE host: MPI_Init_thread forks
// CGREF_NP - 1 children and the seven MPI calls the reference uses run over an anonymous shared
// mapping + a process-shared barrier.  Lets the UNMODIFIED reference MPI solver run at P > 1
// ranks on the box's cores (there is no MPI in this image): a multi-rank CPU baseline, and a
// direct check that the oracle's emulated ranks follow the reference's partition_matrix /
// Allreduce / Allgatherv semantics.  Test infrastructure only (oracle/).
//
//   CGREF_NP=<P>       number of ranks (default 1)
//   CGREF_XOUT=<path>  rank 0 writes the gathered x (MPI_Gatherv, cg.cc:140-142) as raw doubles
//   CGREF_ALLREDUCE=<path>  rank 0 writes every single-value Allreduce result, raw doubles, in
//                      call order: [r.p (cg.cc:91-92), then per iteration p'Ap (:105-106) and
//                      r'r (:116-117)] -- the GLOBAL residual history of a multi-rank run
// MPI_SUM reductions add the ranks' contributions in rank order (0, 1, ..., P-1).
// Non-root ranks leave through _exit() in MPI_Finalize, so atexit side channels (history,
// timestamps -- cblas_provider.c) are written by rank 0 only.
#include "mpi.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <csignal>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

namespace {

constexpr size_t kMaxRanks = 64;
constexpr size_t kExchangeDoubles = 1u << 21;  // 16 MiB: Allgatherv / Gatherv of up to 2M doubles
constexpr size_t kReduceDoubles = 64;          // per-rank Allreduce payload

struct Shared {
    pthread_barrier_t barrier;
    double reduce[kMaxRanks][kReduceDoubles];
    double exchange[kExchangeDoubles];
};

Shared *g_sh = nullptr;
int g_rank = 0, g_size = 1;
std::vector<pid_t> g_children;
std::vector<double> g_allreduce_log;

void die(const char *msg)
{
    std::fprintf(stderr, "[mpi_fork] %s\n", msg);
    std::_Exit(70);
}

void barrier()
{
    if (g_size > 1) pthread_barrier_wait(&g_sh->barrier);
}

} // namespace

extern "C" {

int MPI_Init_thread(int *, char ***, int required, int *provided)
{
    if (provided) *provided = required;
    const char *np = std::getenv("CGREF_NP");
    g_size = np ? std::atoi(np) : 1;
    if (g_size < 1 || (size_t)g_size > kMaxRanks) die("CGREF_NP out of range");
    if (g_size == 1) return MPI_SUCCESS;
    void *mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) die("mmap failed");
    g_sh = static_cast<Shared *>(mem);
    pthread_barrierattr_t attr;
    pthread_barrierattr_init(&attr);
    pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
    if (pthread_barrier_init(&g_sh->barrier, &attr, (unsigned)g_size) != 0) die("barrier init failed");
    std::fflush(nullptr);
    for (int r = 1; r < g_size; ++r) {
        const pid_t pid = fork();
        if (pid < 0) die("fork failed");
        if (pid == 0) {
            g_rank = r;
            g_children.clear();
            return MPI_SUCCESS;
        }
        g_children.push_back(pid);
    }
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    barrier();
    if (g_rank != 0) {
        std::fflush(nullptr);
        std::_Exit(0);
    }
    if (const char *path = std::getenv("CGREF_ALLREDUCE")) {
        if (FILE *f = std::fopen(path, "wb")) {
            std::fwrite(g_allreduce_log.data(), sizeof(double), g_allreduce_log.size(), f);
            std::fclose(f);
        }
    }
    int bad = 0;
    for (pid_t pid : g_children) {
        int st = 0;
        if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
    }
    if (bad) die("a rank failed");
    return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = g_size; return MPI_SUCCESS; }

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm)
{
    if (dt != MPI_DOUBLE || op != MPI_SUM || (size_t)count > kReduceDoubles) die("unsupported Allreduce");
    const double *src = static_cast<const double *>(sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf);
    double *dst = static_cast<double *>(recvbuf);
    if (g_size == 1) {
        if (sendbuf != MPI_IN_PLACE) std::memcpy(dst, src, (size_t)count * sizeof(double));
    } else {
        std::memcpy(g_sh->reduce[g_rank], src, (size_t)count * sizeof(double));
        barrier();
        for (int i = 0; i < count; ++i) {
            double s = g_sh->reduce[0][i];
            for (int r = 1; r < g_size; ++r) s += g_sh->reduce[r][i];
            dst[i] = s;
        }
        barrier(); // nobody overwrites its slot before everyone has read it
    }
    if (g_rank == 0 && count == 1) g_allreduce_log.push_back(dst[0]);
    return MPI_SUCCESS;
}

static size_t total_count(const int *recvcounts, const int *displs)
{
    size_t total = 0;
    for (int r = 0; r < g_size; ++r) {
        const size_t end = (size_t)displs[r] + (size_t)recvcounts[r];
        if (end > total) total = end;
    }
    if (total > kExchangeDoubles) die("exchange buffer too small");
    return total;
}

int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                   const int *recvcounts, const int *displs, MPI_Datatype, MPI_Comm)
{
    if (sendtype != MPI_DOUBLE) die("unsupported Allgatherv type");
    double *dst = static_cast<double *>(recvbuf);
    if (g_size == 1) {
        std::memcpy(dst + displs[0], sendbuf, (size_t)sendcount * sizeof(double));
        return MPI_SUCCESS;
    }
    const size_t total = total_count(recvcounts, displs);
    std::memcpy(g_sh->exchange + displs[g_rank], sendbuf, (size_t)sendcount * sizeof(double));
    barrier();
    std::memcpy(dst, g_sh->exchange, total * sizeof(double));
    barrier();
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int *recvcounts, const int *displs, MPI_Datatype, int root, MPI_Comm)
{
    if (sendtype != MPI_DOUBLE) die("unsupported Gatherv type");
    size_t total = (size_t)sendcount;
    if (g_size == 1) {
        std::memcpy(static_cast<double *>(recvbuf) + displs[0], sendbuf, (size_t)sendcount * sizeof(double));
    } else {
        // recvcounts / displs are significant at the root only; the reference passes them on all ranks
        total = total_count(recvcounts, displs);
        std::memcpy(g_sh->exchange + displs[g_rank], sendbuf, (size_t)sendcount * sizeof(double));
        barrier();
        if (g_rank == root) std::memcpy(recvbuf, g_sh->exchange, total * sizeof(double));
        barrier();
    }
    if (g_rank == root) {
        if (const char *path = std::getenv("CGREF_XOUT")) {
            if (FILE *f = std::fopen(path, "wb")) {
                std::fwrite(recvbuf, sizeof(double), total, f);
                std::fclose(f);
            }
        }
    }
    return MPI_SUCCESS;
}

// the libcgb200 binding (ref_shim/cg_cgb.cc): exchange blobs, a barrier before teardown, abort
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int,
                  MPI_Datatype, MPI_Comm)
{
    const size_t bytes = (size_t)sendcount * (size_t)sendtype;
    if (g_size == 1) {
        std::memcpy(recvbuf, sendbuf, bytes);
        return MPI_SUCCESS;
    }
    if (bytes * (size_t)g_size > sizeof(g_sh->exchange)) die("exchange buffer too small");
    char *ex = reinterpret_cast<char *>(g_sh->exchange);
    std::memcpy(ex + bytes * (size_t)g_rank, sendbuf, bytes);
    barrier();
    std::memcpy(recvbuf, ex, bytes * (size_t)g_size);
    barrier();
    return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm)
{
    barrier();
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int errorcode)
{
    std::fflush(nullptr);
    if (g_rank == 0)
        for (pid_t pid : g_children) kill(pid, SIGTERM);
    std::_Exit(errorcode ? errorcode : 1);
}

} // extern "C"

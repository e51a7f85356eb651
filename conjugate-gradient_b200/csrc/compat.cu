// compat.cu -- the reference CUDA program's two mat-vec TOPOLOGIES re-created for sm_100a with
// NUM_THREADS and BLOCK_WIDTH honoured literally (option "compat" = 1; BASELINE.json config 5,
// SURVEY.md 8f-4).  They exist to show the reference's sweep curve (results/CUDA_T.txt) on a
// B200 next to the reference kernels -- the product mat-vec is gemv.cu.
//
//   column topology (`true`,  MatVecT, code/CUDA/cg.cu:63-110): thread <-> column, block.y <->
//       chunk of BLOCK_WIDTH rows; a warp reads 32 consecutive doubles of a row (coalesced).
//   row topology    (`false`, MatVec,  code/CUDA/cg.cu:14-61):  thread <-> row, block.x <->
//       chunk of BLOCK_WIDTH columns; a warp's loads are N doubles apart (uncoalesced).
//   grids and blocks exactly as cg.cu:196-210.
//
// What differs from the reference, on purpose: its atomicAdd(double) into a pre-zeroed Ap
// (cg.cu:58,107,239) makes the result depend on the order the chunks happen to finish (15
// distinct answers in 24 runs, profiles/r01/config5_sweep.md).  Here every chunk writes its
// partial to part[chunk][i] and a second kernel adds the chunks in ascending order --
// deterministic, and the same bits for both topologies because A is symmetric:
//   y_i = (((0 + s_0) + s_1) + ...),  s_c = fma-chain over k in chunk c of A[i][k] * p[k].
// oracle/cg_oracle.c restates that order (cgo_set_gemv_chunk).
#include "cgb_device.cuh"
#include "cgb_kernels.h"

namespace cgb {

namespace {

__global__ void compat_matvec_t_kernel(const double *A, const double *v, double *part, long long n,
                                       long long ld, int bw, const State *st)
{
    if (st->done) return;
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; // cg.cu:97
    if (col >= n) return;
    const long long rb = (long long)blockIdx.y * bw;                         // cg.cu:93
    const long long re = (rb + bw < n) ? rb + bw : n;
    double s = 0.0;
    for (long long r = rb; r < re; ++r) s = __fma_rn(A[r * ld + col], v[r], s); // cg.cu:100-104
    part[(long long)blockIdx.y * n + col] = s;
}

__global__ void compat_matvec_kernel(const double *A, const double *v, double *part, long long n,
                                     long long ld, int bw, const State *st)
{
    if (st->done) return;
    const long long row = (long long)blockIdx.y * blockDim.x + threadIdx.x;  // cg.cu:48
    if (row >= n) return;
    const long long cb = (long long)blockIdx.x * bw;                          // cg.cu:44
    const long long ce = (cb + bw < n) ? cb + bw : n;
    double s = 0.0;
    for (long long j = cb; j < ce; ++j) s = __fma_rn(A[row * ld + j], v[j], s); // cg.cu:51-55
    part[(long long)blockIdx.x * n + row] = s;
}

// Second level: rows of Ap = ordered sum of the chunk partials (+ the scalar bookkeeping the
// product mat-vec does).  p'Ap is reduced from the gathered rows like for every other mat-vec.
__global__ void __launch_bounds__(256) compat_reduce_kernel(const GemvArgs a, const double *part,
                                                            long long nchunk)
{
    if (a.st->done) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = gridDim.x, c = blockIdx.x;
    const long long r0 = (long long)c * a.rows / nblk;
    const long long r1 = (long long)(c + 1) * a.rows / nblk;
    if (a.advance && c == 0 && warp == 0) advance_state(a, lane);
    for (long long i = r0 + tid; i < r1; i += blockDim.x) {
        double y = 0.0;
        for (long long ch = 0; ch < nchunk; ++ch) y = __dadd_rn(y, part[ch * a.rows + i]);
        a.base[a.slot_off + i] = y;
    }
}

} // namespace

size_t compat_part_doubles(long long n, int block_width)
{
    const long long nchunk = (n + block_width - 1) / block_width;
    return (size_t)nchunk * (size_t)n;
}

cudaError_t launch_compat_matvec(const GemvArgs &a, int nblk, int num_threads, int block_width,
                                 int transposed, double *part, cudaStream_t s)
{
    const long long n = a.rows; // single GPU: the shard is the whole matrix
    const long long nchunk = (n + block_width - 1) / block_width;
    const long long nvec = (n + num_threads - 1) / num_threads;
    if (transposed) { // cg.cu:203-204
        if (nchunk > 65535) return cudaErrorInvalidConfiguration;
        dim3 grid((unsigned)nvec, (unsigned)nchunk);
        compat_matvec_t_kernel<<<grid, num_threads, 0, s>>>(a.A, a.v, part, n, a.ld, block_width, a.st);
    } else {          // cg.cu:207-209
        if (nvec > 65535) return cudaErrorInvalidConfiguration;
        dim3 grid((unsigned)nchunk, (unsigned)nvec);
        compat_matvec_kernel<<<grid, num_threads, 0, s>>>(a.A, a.v, part, n, a.ld, block_width, a.st);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    compat_reduce_kernel<<<nblk, 256, 0, s>>>(a, part, nchunk);
    return cudaGetLastError();
}

} // namespace cgb

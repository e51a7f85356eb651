// cgb_kernels.h -- host-callable launchers of the sm_100a kernels (gemv.cu, vec.cu).
#pragma once

#include "cgb_device.cuh"

#include <cstring>

#include <utility>

namespace cgb {

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute: the
// kernels of the CG loop are chained with programmatic dependent launches so that the next
// kernel is already resident (and the mat-vec already prefetching A) when its predecessor ends.
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem,
                                 cudaStream_t stream, bool pdl, Args &&...args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- mat-vec (gemv.cu) --------------------------------------------------------------
struct GemvVariant {
    const char *name;
    int ctas_per_sm; // grid = sm_count * ctas_per_sm persistent CTAs
    int threads;
    cudaError_t (*launch)(const GemvArgs &a, int nblk, cudaStream_t s);
    cudaError_t (*preload)(); // load the kernel now (lazy module loading), not at its first launch
};
int gemv_variant_count();
const GemvVariant &gemv_variant(int i);
cudaError_t launch_read_stream(const double *A, long long ndoubles, double *sink, int sm_count,
                               cudaStream_t s);

// ---- persistent schedule (persist.cu): the CG loop as one cooperative launch --------------
struct PersistVariant {
    const char *name; // = the gemv.cu variant with the same tile shape
    cudaError_t (*launch)(const PersistArgs &a, int nblk, cudaStream_t s);
    cudaError_t (*preload)();
    size_t (*smem_fixed)(); // shared memory without the size-dependent scratch (qs_n + scr_n doubles)
};
int persist_variant_count();
const PersistVariant &persist_variant(int i);
int persist_max_chunks(); // 256-element chunks of the vectors one CTA can own

// ---- reference-topology mat-vec (compat.cu): NUM_THREADS / BLOCK_WIDTH honoured literally ----
size_t compat_part_doubles(long long n, int block_width);
cudaError_t launch_compat_matvec(const GemvArgs &a, int nblk, int num_threads, int block_width,
                                 int transposed, double *part /* compat_part_doubles */, cudaStream_t s);

// ---- vector kernels (vec.cu) -------------------------------------------------------
// load every kernel of the solve path now instead of at first launch
cudaError_t preload_vec_kernels();
struct VecArgs {
    double *x, *r, *p;       // full-length (ld) replicated work vectors
    const double *b;
    const double *apx;       // gathered mat-vec result (see Gather)
    double *rrpart;          // nchunks chunk partials of r'r
    double *papart;          // nchunks chunk partials of p'Ap
    State *st;
    int *host_done;          // mapped pinned flag the host polls (nullable)
    double *hist;            // nullable
    Gather g;
    long long n;
    double tol;
    int pdl;                 // host side: chain update_xr / update_p with programmatic dependent launch
    Trace trace_xr, trace_p; // diagnostic timelines of update_xr / update_p (buf == nullptr: off)
};
// r = b - A x0 ; p = r ; rrpart = chunk partials of r.p             (cg.cc:77-92)
cudaError_t launch_init_residual(const VecArgs &a, cudaStream_t s);
// chunk partials of p'Ap from the gathered rows                      (cg.cc:105)
cudaError_t launch_pap_partials(const VecArgs &a, cudaStream_t s);
// alpha = rsold / max(p'Ap, rsold*NEARZERO) ; x += alpha p ; r -= alpha Ap ; r'r partials
//                                                                   (cg.cc:106-117)
cudaError_t launch_update_xr(const VecArgs &a, cudaStream_t s);
// rsnew = sum(partials) ; stop test ; beta = rsnew/rsold ; p = r + beta p   (cg.cc:120-129)
cudaError_t launch_update_p(const VecArgs &a, cudaStream_t s);
// after the last iteration: rsold = rsnew, iter += 1 when the loop did not break (cg.cc:132)
cudaError_t launch_finalize(const VecArgs &a, cudaStream_t s);
// DEBUG block (cg.cc:144-154): out[0] = ||x||, out[1] = ||Ax-b|| / ||b||, given apx = A x
cudaError_t launch_debug_norms(const VecArgs &a, double *scratch /* 3*nchunks */, double *out,
                               cudaStream_t s);
// generic deterministic dot (tests): out[0] = a.b
cudaError_t launch_dot(const double *a, const double *b, long long n, double *scratch, double *out,
                       cudaStream_t s);
// hooks: chunk partials of v.(A v) from the plain gather buffer into papart, out[0] = their det_sum
cudaError_t launch_pap_plain(const double *v, const double *apx, const Gather &g, long long n, double *papart,
                             double *out, cudaStream_t s);

// fused mode, hooks only: consume the running exchange into the plain gather buffer `apx`
// (the kernels of the iteration read the LL entries directly)
cudaError_t launch_exchange_collect(double *apx, const Gather &g, cudaStream_t s);

// ---- matrix construction (vec.cu) --------------------------------------------------
// generate_lap2d_matrix (cg.cc:159-188) into the shard
cudaError_t launch_generate_lap2d(double *A, long long n, long long ld, long long row0,
                                  long long rows, cudaStream_t s);
// Matrix::read's densification (matrix.cc:12-21) on the device, "later entries overwrite" resolved
// with an atomicMax of the entry index per cell (win: nz bytes of scratch, bad: 1 int, both device)
cudaError_t launch_scatter_coo(double *A, long long n, long long ld, long long row0, long long rows,
                               const int *irn, const int *jcn, const double *val, long long nz,
                               int symmetric, unsigned char *win, int *bad, cudaStream_t s);

} // namespace cgb

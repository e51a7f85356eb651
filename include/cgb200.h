/*
 * cgb200.h -- C ABI of the B200-native conjugate-gradient hot path.
 *
 * The reference (federicobetti99/Conjugate-Gradient) has no plugin / FFI surface: its only
 * stable interfaces are the `cgsolver` command line + results file and the C++ class
 * `CGSolver` (code/MPI/cg.hh:11-57, code/CUDA/cg.hh:13-45).  This ABI sits directly UNDER
 * `CGSolver::solve`; every entry point below names the reference code it replaces.  The host
 * program `cgsolver` (conjugate-gradient_b200/host/) and the Python mirror bind exactly these
 * symbols; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns a
 * cgb_status (0 = ok) and never throws or exits; cgb_last_error() gives the message of the
 * last failure on the calling thread.  Host buffers belong to the caller, device buffers to
 * the library.  One cgb_ctx = one rank = one GPU holding one contiguous row shard of A
 * (reference rule partition_matrix, code/MPI/cg.cc:236-268); a ctx is not thread-safe, but
 * different ctxs may be driven from different threads or processes.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with CGB_ERR_NO_DEVICE.
 */
#ifndef CGB200_H
#define CGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_ABI_VERSION 1

typedef enum cgb_status {
    CGB_OK = 0,
    CGB_ERR_INVALID = 1,   /* bad argument */
    CGB_ERR_STATE = 2,     /* call out of order (e.g. solve before the matrix is set) */
    CGB_ERR_NO_DEVICE = 3, /* no CUDA device / driver */
    CGB_ERR_CUDA = 4,      /* a CUDA runtime call failed */
    CGB_ERR_NCCL = 5,      /* NCCL missing or a NCCL call failed */
    CGB_ERR_NOMEM = 6,
    CGB_ERR_TIMEOUT = 7    /* a device-side wait on another rank exceeded "spin_timeout_ms" */
} cgb_status;

typedef struct cgb_ctx cgb_ctx;

/* ---- library ------------------------------------------------------------------------- */
int cgb_abi_version(void);
const char *cgb_last_error(void);
int cgb_device_count(int *count);

/* CGSolver::partition_matrix (code/MPI/cg.cc:236-268) with 64-bit indices: ranks 0..P-2 get
 * N/P rows, the last rank the remainder. */
int cgb_partition(int64_t n, int psize, int64_t *start_rows, int64_t *num_rows);

/* ---- context ------------------------------------------------------------------------- */
/* One rank of `world` on CUDA device `device`, for an n x n system.  Allocates the row shard
 * (rows x ld doubles, ld = n rounded up to 16) and the work vectors -- the
 * std::vector / cudaMallocManaged temporaries of solve (cg.cc:55-75, cg.cu:179-193). */
int cgb_create(int64_t n, int rank, int world, int device, cgb_ctx **out);
int cgb_destroy(cgb_ctx *ctx);

/* Replaces MPI_Init_thread / MPI_COMM_WORLD (code/MPI/cg_main.cc:15-20): rank 0 obtains an
 * id with cgb_comm_unique_id, ships it to the other ranks by any means, and every rank calls
 * cgb_comm_init (collective; wraps ncclCommInitRank).  Not needed when world == 1. */
#define CGB_UNIQUE_ID_BYTES 128
int cgb_comm_unique_id(void *id_out);
int cgb_comm_init(cgb_ctx *ctx, const void *id);

/* Fused exchange (the B200-native replacement of MPI_Allgatherv, cg.cc:135-136): the mat-vec
 * kernel stores its rows directly into every rank's gather buffer over NVLink and raises a
 * flag; no collective kernel runs.  Every rank exports a blob describing its buffer, the blobs
 * are shipped to all ranks by any means (world x CGB_EXCHANGE_BLOB_BYTES, rank order) and
 * imported; ranks may be threads of one process (peer access) or separate processes (CUDA
 * IPC).  After the import the option "exchange" is 1 (fused); 0 selects ncclAllGather, which
 * needs cgb_comm_init instead.  Like a collective, every rank must issue the same sequence of
 * solves / mat-vec hooks. */
#define CGB_EXCHANGE_BLOB_BYTES 128
int cgb_exchange_export(cgb_ctx *ctx, void *blob_out);
int cgb_exchange_import(cgb_ctx *ctx, const void *blobs);

/* ---- inputs -------------------------------------------------------------------------- */
/* CGSolver::generate_lap2d_matrix (code/MPI/cg.cc:159-188), written straight into this
 * rank's device shard (values exactly 0, -1, 4; inc = floor(sqrt(n))). */
int cgb_generate_lap2d(cgb_ctx *ctx);

/* Dense rows from the host (what Matrix::read, matrix.cc:6-22, produces): `nrows` rows
 * starting at GLOBAL row `first_row`, row-major with leading dimension `ld_host`; the part
 * that falls inside this rank's shard is uploaded, the rest ignored. */
int cgb_set_matrix_rows(cgb_ctx *ctx, const double *rows_host, int64_t first_row, int64_t nrows,
                        int64_t ld_host);

/* Matrix::read's densification (matrix.cc:12-21) done on the device: zero the shard, then
 * scatter the COO triples (0-based) in file order, mirroring (j,i) when `symmetric`.
 * Later duplicates win, as in the reference. */
int cgb_set_matrix_coo(cgb_ctx *ctx, int64_t nz, const int32_t *irn, const int32_t *jcn,
                       const double *val, int symmetric);

/* Read rows of the shard back (tests). */
int cgb_get_matrix_rows(cgb_ctx *ctx, double *rows_host, int64_t first_row, int64_t nrows,
                        int64_t ld_host);

/* CGSolver::init_source_term (code/MPI/cg.cc:218-234, code/CUDA/cg.cu:324-340):
 * b_i = -2 i pi^2 sin(10 pi i h)^2 into a caller-owned HOST buffer of n doubles.  Like the
 * reference this is a host loop over libm sin (the values feed cgb_set_rhs); it is the one
 * entry point that needs no GPU. */
int cgb_init_source_term(int64_t n, double h, double *b_host);

/* Right-hand side m_b (filled on the host by init_source_term, cg.cc:218-234, so that libm
 * sin is shared with the reference), full length n. */
int cgb_set_rhs(cgb_ctx *ctx, const double *b_host);

/* ---- tuning -------------------------------------------------------------------------- */
/* Integer options.  "gemv_variant": index into the table cgb_gemv_variant_name() lists;
 * "num_threads"/"block_width": the NUM_THREADS / BLOCK_WIDTH command-line knobs of the
 * reference CUDA program (code/CUDA/cg_main.cc:21-25), mapped onto threads per CTA and
 * column-tile width of the mat-vec; "graph": 0/1 CUDA-graph replay of the iteration;
 * "poll_every" / "graph_unroll": graph schedule only -- iterations between two looks at the stop flag and
 * iterations per graph (the instantiated graph holds min of the two; "graph_replays", read-only,
 * counts its launches);
 * "pdl": 0/1 programmatic dependent launch between the kernels of the iteration (the next
 * mat-vec prefetches A while the vector updates still run); "l2_prefetch": pipeline steps of A
 * the mat-vec additionally pulls into L2 before that wait (0 = off); "l2_ramp": persistent
 * schedule only -- further steps pulled into L2 the moment p arrives, so that HBM keeps
 * streaming while the steps already on chip are consumed (default 2, 0 = off);
 * "exchange": 0 ncclAllGather / 1 fused peer stores (see cgb_exchange_import);
 * "schedule": 1 (default) = the whole loop of cgb_iterate as ONE persistent cooperative kernel
 * (one CTA per SM stays resident, the iteration's dependencies are data-flow waits inside the
 * kernel, A streams across iteration boundaries), 0 = a CUDA graph of four kernels per
 * iteration; bitwise identical results; "schedule_in_use" (read-only) tells which one the
 * current configuration gets (the persistent kernel needs the fused exchange, a 1-CTA-per-SM
 * variant, and is not used with "compat" / "profile");
 * "spin_timeout_ms": bound of the device-side waits on other ranks (default 20000); when it
 * expires the call returns CGB_ERR_TIMEOUT (the context is then unusable);
 * "balance": 1 (default) = the persistent kernel re-partitions the rows between its CTAs from the
 * mat-vec times they measure, when a CTA has at least 64 rows (no result bit depends on the partition);
 * "trace": launches kept by the diagnostic timeline (cgb_trace_read; 0 = off);
 * "loopback": profiling aid -- a rank of world > 1 with no peers aims every peer pointer of the
 * fused exchange at its own buffer, so ONE GPU runs one rank's shard under the production
 * schedule (timing and traffic are real, the numbers are not a CG solve);
 * "transposed": the reference's true/false kernel switch (accepted, A is symmetric);
 * "compat": 1 = run the mat-vec in the reference CUDA program's own topologies (MatVecT /
 * MatVec, code/CUDA/cg.cu:14-110) with num_threads / block_width / transposed honoured
 * literally but a deterministic chunk reduction instead of atomicAdd -- single GPU, for the
 * NUM_THREADS x BLOCK_WIDTH sweep of BASELINE.json config 5; set the three knobs first. */
int cgb_set_option(cgb_ctx *ctx, const char *key, int64_t value);
int cgb_get_option(cgb_ctx *ctx, const char *key, int64_t *value);
int cgb_gemv_variant_count(void);
const char *cgb_gemv_variant_name(int variant);

/* One-off choice of the mat-vec tile shape ("gemv_variant") for this rank's shard shape on this
 * GPU: every candidate shape runs `iters` (<= 0: 32) loop bodies of the schedule in use on the
 * resident matrix, the fastest becomes the configured variant.  Call it after the matrix is
 * set and OUTSIDE any timed region (the reference's timer brackets solve() only,
 * code/MPI/cg_main.cc:53-55).  A rank of world > 1 tunes alone (its exchange looped back to
 * itself on scratch buffers); peers need not take part or wait.  All shapes share one summation order:
 * the choice never changes a result bit.  us_per_iter (nullable, cgb_gemv_variant_count()
 * floats) receives the time of every candidate, < 0 for shapes that were not candidates. */
int cgb_autotune(cgb_ctx *ctx, int iters, int *chosen, float *us_per_iter);

typedef struct cgb_layout {
    int64_t n, ld, rows, row0; /* shard geometry */
    int rank, world, device;
    int nblk;                  /* mat-vec grid (CTAs); no reduction depends on it */
    int sm_count;
    int64_t nchunks;           /* 256-element chunks of the r'r reduction */
} cgb_layout;
int cgb_get_layout(cgb_ctx *ctx, cgb_layout *out);

/* ---- the hot path: CGSolver::solve (code/MPI/cg.cc:38-156, code/CUDA/cg.cu:166-305) ----- */
typedef struct cgb_solve_info {
    int64_t k;          /* the k of "[STEP k]": loop index at break, or max_iter */
    int converged;      /* loop left through sqrt(rsnew) < tol (cg.cc:120-121) */
    double rsold;       /* rsold at exit -- stale by one iteration after a break, as printed */
    double rsnew;       /* last r'r computed */
    double seconds;     /* device time of the iteration loop (CUDA events) */
    int64_t iterations; /* loop bodies executed on the device */
} cgb_solve_info;

/* Whole solve, host buffers in and out: x_host holds x0 on entry (n doubles), the solution
 * on return (every rank receives the full x -- MPI_Gatherv, cg.cc:140-142, delivered it to
 * rank 0 only).  resid_hist (nullable, capacity max_iter) receives r'r of every executed loop
 * index.  tol is the absolute tolerance on sqrt(r'r) (m_tolerance, cg.hh:56). */
int cgb_solve(cgb_ctx *ctx, double *x_host, int64_t max_iter, double tol, double *resid_hist,
              cgb_solve_info *info);

/* The same solve in three steps, for callers that keep the state on the device between
 * batches of iterations (bench.py's device-resident timing, the CLI's convergence polling):
 *   begin   : upload x0, r = b - A x0, p = r, rsold = r.p      (cg.cc:77-92)
 *   iterate : up to `iters` loop bodies (stops early on convergence); *ms = device time
 *   end     : download x, fill info                             (cg.cc:140-142) */
int cgb_solve_begin(cgb_ctx *ctx, const double *x0_host, int64_t max_iter, double tol,
                    int keep_history);
int cgb_iterate(cgb_ctx *ctx, int64_t iters, float *ms);
int cgb_solve_end(cgb_ctx *ctx, double *x_host, double *resid_hist, cgb_solve_info *info);

/* The DEBUG block at the end of solve (cg.cc:144-154, cg.cu:272-296): ||x|| and
 * ||A x - b|| / ||b|| for the x held on the device after a solve. */
int cgb_residual_check(cgb_ctx *ctx, double *norm_x, double *rel_resid);

/* ---- kernel-level hooks (tests, ncu, roofline) ------------------------------------------ */
/* y = A_shard . v through the production mat-vec: v_host has n doubles; y_host (nullable)
 * receives this rank's `rows` results, chunk_partials (nullable, `nchunks` of cgb_get_layout) the
 * partial sums of v_i * (A v)_i over the 256-element chunks of the GLOBAL vector (all ranks'
 * rows, after the exchange), pAp (nullable) their deterministic total -- level 1 and 2 of the
 * reduction that replaces cblas_ddot + MPI_Allreduce (cg.cc:105-106). */
int cgb_gemv(cgb_ctx *ctx, const double *v_host, double *y_host, double *chunk_partials,
             double *pAp);
/* Deterministic two-level dot of two host vectors (n doubles each) on the device. */
int cgb_dot(cgb_ctx *ctx, const double *a_host, const double *b_host, double *result);
/* `reps` back-to-back launches of mat-vec variant `variant` (-1 = the configured one) on the
 * resident shard and the resident p; *ms_avg = average device time per launch. */
int cgb_bench_gemv(cgb_ctx *ctx, int variant, int reps, float *ms_avg);
/* Streams the shard once with a plain read-only kernel: the read-bandwidth ceiling. */
int cgb_bench_read(cgb_ctx *ctx, int reps, float *ms_avg);
/* Average device time of the mat-vec launches inside the last cgb_iterate that ran with the
 * option "profile" = 1 (per-launch CUDA events). */
int cgb_last_gemv_timing(cgb_ctx *ctx, float *ms_avg, int64_t *launches);
/* Kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int cgb_launch_count(cgb_ctx *ctx, int64_t *launches);
/* Diagnostic timeline of the iteration under its PRODUCTION schedule (CUDA graph + programmatic
 * dependent launch, where no CUDA event can sit between the kernels).  With the option
 * "trace" = L every CTA of the loop kernels stamps %globaltimer (ns) into a ring of the last L
 * launches; this call copies the ring of kernel `which` (0 = mat-vec, 1 = x/r update, 2 = p
 * update) to `out` as [L][*blocks][8] words and reports how many launches were seen in total.
 * mat-vec words: 0 entry, 1 ring pre-filled, 2 dependency met (p final), 3 first tile
 * consumed, 4 last tile issued, 5 rows done, 6 exit, 7 = smid | rows << 32.  vector kernels:
 * 0 entry, 1 dependency met, 2 scalar known, 3 exit.  The replacement for what the reference
 * measures with one chrono pair around solve (code/MPI/cg_main.cc:53-55). */
int cgb_trace_read(cgb_ctx *ctx, int which, uint64_t *out, int64_t capacity_words,
                   int64_t *launches, int64_t *blocks);

#ifdef __cplusplus
}
#endif
#endif /* CGB200_H */

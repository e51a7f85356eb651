/*
 * cg_oracle.h -- CPU restatement of the reference conjugate-gradient hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * `conjugate-gradient_b200/`, `include/`, the `cgsolver` host program) may
 * include, link or call this.  Allowed users: `tests/`, `__graft_entry__.smoke()`
 * and the `cpu_baseline` / `--impl reference` legs of `bench.py`.
 *
 * What it restates (all citations into /root/reference):
 *   code/MPI/cg.cc:38-156    CGSolver::solve      -> cgo_solve
 *   code/MPI/cg.cc:159-188   generate_lap2d_matrix-> cgo_generate_lap2d
 *   code/MPI/cg.cc:218-234   init_source_term     -> cgo_init_source_term
 *   code/MPI/cg.cc:236-268   partition_matrix     -> cgo_partition
 *   code/CUDA/cg.cu:216-270  sumVec forms (x = 1*x + a*p, r = 1*r + (-a)*Ap,
 *                            p = b*p + 1*r) cross-checked for the fused updates
 *
 * The arithmetic of the reference lives in un-vendored third-party BLAS
 * (OpenBLAS cblas_dgemv/ddot/daxpy, cuBLAS cublasDdot), whose
 * summation order is unspecified.  This restatement fixes ONE fully specified
 * order -- the order the sm_100a kernels use -- so GPU-vs-oracle is bitwise:
 *
 *   row dot (A_i . p), "lane order":  64 accumulators, acc[c % 64] =
 *       fma(A[i][c], p[c], acc[c % 64]) for c ascending (c%64 = 2*lane + sub,
 *       i.e. 32 lanes striding over 128-bit chunks, even/odd accumulators);
 *       lane[l] = acc[2l] + acc[2l+1]; butterfly l += l^16, ^8, ^4, ^2, ^1.
 *   det_sum(v, n): lane[l] = v[l] + v[l+32] + ... (ascending, plain adds,
 *       starting from +0.0), then the same butterfly.
 *   chunk256(v): perfect xor tree over 256 consecutive elements
 *       (offsets 16,8,4,2,1 inside each 32-group, then 128,64,32 across groups).
 *   p'Ap  = det_sum over chunk256 partials of p_i * Ap_i over the GLOBAL vector,
 *   r'r   = det_sum over chunk256 partials of r_i * r_i over the GLOBAL vector
 *       (the same two-level tree for both): no reduction depends on how rows are
 *       distributed over GPUs, CTAs or ranks, so 1, 2, 4 and 8 GPUs -- and any
 *       dynamic re-balancing of rows between SMs -- give the same bits.
 *   x_i = fma(alpha, p_i, x_i); r_i = fma(-alpha, Ap_i, r_i); p_i = fma(beta, p_i, r_i).
 *
 * Parity pin: validated here against the reference's own sources compiled
 * unmodified (oracle/_ref, see oracle/Makefile) -- iteration counts, residual
 * histories and final x -- and frozen as fixtures in tests/golden/.  The
 * reference ships no tests or golden vectors of its own.
 */
#ifndef CG_ORACLE_H
#define CG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cg.cc:159-188.  A is n*n row-major (ld = n), caller-allocated. */
void cgo_generate_lap2d(int64_t n, double *A);
/* same rule, rows [row0, row0+nrows) only, into a buffer with leading dim ld. */
void cgo_generate_lap2d_rows(int64_t n, int64_t row0, int64_t nrows, double *A, int64_t ld);

/* cg.cc:218-234 (CUDA: cg.cu:324-340): b_i = -2 i pi^2 sin(10 pi i h)^2 */
void cgo_init_source_term(int64_t n, double h, double *b);

/* cg.cc:236-268 with 64-bit indices */
void cgo_partition(int64_t n, int psize, int64_t *start_rows, int64_t *num_rows);

/* balanced contiguous split used for GEMV blocks: block c of nblk over `rows`
 * covers [c*rows/nblk, (c+1)*rows/nblk) (integer floor). */
void cgo_block_range(int64_t rows, int nblk, int c, int64_t *r0, int64_t *r1);

/* lane-order row dot over `n` columns */
double cgo_row_dot(const double *a_row, const double *p, int64_t n);
/* y[0..rows) = A[rows x n, ld] . p, lane order per row */
void cgo_gemv(int64_t rows, int64_t n, const double *A, int64_t ld, const double *p, double *y);

/* select the mat-vec order used by cgo_gemv / cgo_solve / cgo_residual_check: 0 = lane order
 * (the product kernels), bw > 0 = the reference-topology order with BLOCK_WIDTH = bw
 * (csrc/compat.cu; see cg_oracle.c).  Process-global; tests reset it to 0. */
void cgo_set_gemv_chunk(int block_width);

double cgo_det_sum(const double *v, int64_t n);
/* two-level dot of the global vectors a.b: chunk256 partials -> det_sum */
double cgo_dot(const double *a, const double *b, int64_t n);

typedef struct {
    int64_t k;            /* the k printed in "[STEP k]": loop index at break, or max_iter */
    int     converged;    /* 1 if the loop broke on sqrt(rsnew) < tol */
    double  rsold;        /* rsold at exit (stale by one iteration on convergence) */
    double  rsnew;        /* last rsnew computed */
    double  norm_x;       /* DEBUG block: ||x|| */
    double  rel_resid;    /* DEBUG block: ||Ax-b|| / ||b|| */
} cgo_info;

/* cg.cc:38-156 restated.  `nranks` / `nblk` (emulated ranks, mat-vec grid) no longer enter the
 * arithmetic -- see the reduction order above; kept in the signature.  A: n*n, ld.  x: in = x0,
 * out = solution.  hist (nullable): rsnew of every executed iteration
 * (hist[j] = rsnew computed in loop index j), capacity max_iter. */
void cgo_solve(int64_t n, const double *A, int64_t ld, const double *b, double *x,
               int64_t max_iter, double tol, int nranks, int nblk,
               double *hist, cgo_info *info);

/* the DEBUG block alone (cg.cc:144-154) */
void cgo_residual_check(int64_t n, const double *A, int64_t ld, const double *b,
                        const double *x, double *norm_x, double *rel_resid);

#ifdef __cplusplus
}
#endif
#endif

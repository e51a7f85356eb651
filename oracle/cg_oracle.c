/*
 * cg_oracle.c -- CPU restatement of the reference CG hot path (see cg_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off + explicit fma() keeps every rounding where the spec says.
 */
#include "cg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const double NEARZERO = 1.0e-14; /* cg.cc:8 */

/* ------------------------------------------------------------------ inputs */

/* cg.cc:159-188: zeros; inc = floor(sqrt(n)); the five writes in source order. */
void cgo_generate_lap2d_rows(int64_t n, int64_t row0, int64_t nrows, double *A, int64_t ld)
{
    int64_t inc = (int64_t)floor(sqrt((double)n));
    for (int64_t li = 0; li < nrows; ++li) {
        int64_t i = row0 + li;
        double *row = A + li * ld;
        for (int64_t j = 0; j < n; ++j) row[j] = 0.0;
        if (i > inc) row[i - 1 - inc] = -1.0;
        if (i > 0) row[i - 1] = -1.0;
        row[i] = 4.0;
        if (i < n - 1) row[i + 1] = -1.0;
        if (i < n - 1 - inc) row[i + 1 + inc] = -1.0;
    }
}

void cgo_generate_lap2d(int64_t n, double *A) { cgo_generate_lap2d_rows(n, 0, n, A, n); }

/* cg.cc:218-234: libm sin, evaluated exactly as the reference spells it. */
void cgo_init_source_term(int64_t n, double h, double *b)
{
    for (int64_t i = 0; i < n; i++) {
        b[i] = -2. * i * M_PI * M_PI * sin(10. * M_PI * i * h) * sin(10. * M_PI * i * h);
    }
}

/* cg.cc:236-268 */
void cgo_partition(int64_t n, int psize, int64_t *start_rows, int64_t *num_rows)
{
    if (psize == 1) {
        start_rows[0] = 0;
        num_rows[0] = n;
        return;
    }
    int64_t n_loc = n / psize;
    int64_t i0 = 0;
    for (int r = 0; r < psize - 1; ++r) {
        start_rows[r] = i0;
        num_rows[r] = n_loc;
        i0 += n_loc;
    }
    start_rows[psize - 1] = i0;
    num_rows[psize - 1] = n - i0;
}

void cgo_block_range(int64_t rows, int nblk, int c, int64_t *r0, int64_t *r1)
{
    *r0 = (int64_t)c * rows / nblk;
    *r1 = (int64_t)(c + 1) * rows / nblk;
}

/* --------------------------------------------------------- reduction order */

static inline double butterfly32(double *lane)
{
    for (int off = 16; off >= 1; off >>= 1)
        for (int l = 0; l < off; ++l) lane[l] = lane[l] + lane[l + off];
    return lane[0];
}

double cgo_row_dot(const double *a, const double *p, int64_t n)
{
    double acc[64];
    for (int j = 0; j < 64; ++j) acc[j] = 0.0;
    int64_t c = 0;
    for (; c + 64 <= n; c += 64)
        for (int j = 0; j < 64; ++j) acc[j] = fma(a[c + j], p[c + j], acc[j]);
    for (int j = 0; c + j < n; ++j) acc[j] = fma(a[c + j], p[c + j], acc[j]);
    double lane[32];
    for (int l = 0; l < 32; ++l) lane[l] = acc[2 * l] + acc[2 * l + 1];
    return butterfly32(lane);
}

/* Order of the reference-topology ("compat") mat-vec, conjugate-gradient_b200/csrc/compat.cu:
 * the row is cut into chunks of `bw` columns (MatVec: BLOCK_WIDTH columns per block, cg.cu:44-55;
 * MatVecT: BLOCK_WIDTH rows per block, cg.cu:93-104 -- the same products because A is symmetric),
 * each chunk a sequential fma chain from 0, the chunks added in ascending order from 0. */
static int g_gemv_chunk = 0; /* 0 = lane order */
void cgo_set_gemv_chunk(int block_width) { g_gemv_chunk = block_width > 0 ? block_width : 0; }

static double row_dot_chunked(const double *a, const double *p, int64_t n, int64_t bw)
{
    double y = 0.0;
    for (int64_t c0 = 0; c0 < n; c0 += bw) {
        int64_t c1 = (c0 + bw < n) ? c0 + bw : n;
        double s = 0.0;
        for (int64_t k = c0; k < c1; ++k) s = fma(a[k], p[k], s);
        y = y + s;
    }
    return y;
}

void cgo_gemv(int64_t rows, int64_t n, const double *A, int64_t ld, const double *p, double *y)
{
    const int64_t bw = g_gemv_chunk;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i)
        y[i] = bw > 0 ? row_dot_chunked(A + i * ld, p, n, bw) : cgo_row_dot(A + i * ld, p, n);
}

double cgo_det_sum(const double *v, int64_t n)
{
    double lane[32];
    for (int l = 0; l < 32; ++l) {
        double s = 0.0;
        for (int64_t t = l; t < n; t += 32) s = s + v[t];
        lane[l] = s;
    }
    return butterfly32(lane);
}

/* perfect xor tree over 256 products a_i*b_i starting at `base` (0 beyond n) */
static double chunk256(const double *a, const double *b, int64_t base, int64_t n)
{
    double w[8];
    for (int g = 0; g < 8; ++g) {
        double lane[32];
        for (int l = 0; l < 32; ++l) {
            int64_t i = base + 32 * g + l;
            lane[l] = (i < n) ? a[i] * b[i] : 0.0;
        }
        w[g] = butterfly32(lane);
    }
    for (int off = 4; off >= 1; off >>= 1)
        for (int g = 0; g < off; ++g) w[g] = w[g] + w[g + off];
    return w[0];
}

double cgo_dot(const double *a, const double *b, int64_t n)
{
    int64_t nch = (n + 255) / 256;
    double *part = (double *)malloc(sizeof(double) * (size_t)(nch > 0 ? nch : 1));
    for (int64_t c = 0; c < nch; ++c) part[c] = chunk256(a, b, c * 256, n);
    double s = cgo_det_sum(part, nch);
    free(part);
    return s;
}

/* ------------------------------------------------------------------- solve */

void cgo_residual_check(int64_t n, const double *A, int64_t ld, const double *b,
                        const double *x, double *norm_x, double *rel_resid)
{
    /* cg.cc:144-154: r = A x; r -= b; res = sqrt(r.r)/sqrt(b.b); nx = sqrt(x.x) */
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    cgo_gemv(n, n, A, ld, x, r);
    for (int64_t i = 0; i < n; ++i) r[i] = fma(-1.0, b[i], r[i]);
    *rel_resid = sqrt(cgo_dot(r, r, n)) / sqrt(cgo_dot(b, b, n));
    *norm_x = sqrt(cgo_dot(x, x, n));
    free(r);
}

void cgo_solve(int64_t n, const double *A, int64_t ld, const double *b, double *x,
               int64_t max_iter, double tol, int nranks, int nblk,
               double *hist, cgo_info *info)
{
    /* cg.cc:60-68 shards the rows over the ranks; every reduction below is defined on the GLOBAL
     * vectors (row results do not depend on who computes them), so neither the rank count nor
     * the mat-vec grid enters the arithmetic -- the parameters are kept for the callers. */
    (void)nranks;
    (void)nblk;

    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *p = (double *)malloc(sizeof(double) * (size_t)n);
    double *Ap = (double *)malloc(sizeof(double) * (size_t)n);

    /* cg.cc:77-82: r = b - A x  (row results do not depend on the sharding) */
    cgo_gemv(n, n, A, ld, x, Ap);
    for (int64_t i = 0; i < n; ++i) r[i] = fma(-1.0, Ap[i], b[i]);
    /* cg.cc:85-88: p = r (replicated) */
    memcpy(p, r, sizeof(double) * (size_t)n);
    /* cg.cc:91-92: rsold = r.p */
    double rsold = cgo_dot(r, p, n);
    double rsnew = rsold;
    int converged = 0;

    int64_t k = 0;
    for (; k < max_iter; ++k) { /* cg.cc:96 */
        cgo_gemv(n, n, A, ld, p, Ap);                                  /* :100-102 */
        double conj = cgo_dot(p, Ap, n);                               /* :105-106 */
        double clamp = rsold * NEARZERO;
        double alpha = rsold / ((conj < clamp) ? clamp : conj);        /* :107 std::max */
        for (int64_t i = 0; i < n; ++i) x[i] = fma(alpha, p[i], x[i]); /* :110 */
        double nalpha = -alpha;
        for (int64_t i = 0; i < n; ++i) r[i] = fma(nalpha, Ap[i], r[i]); /* :113 */
        rsnew = cgo_dot(r, r, n);                                      /* :116-117 */
        if (hist) hist[k] = rsnew;
        if (sqrt(rsnew) < tol) { converged = 1; break; }               /* :120-121 */
        double beta = rsnew / rsold;                                   /* :124 */
        for (int64_t i = 0; i < n; ++i) p[i] = fma(beta, p[i], r[i]);  /* :127-129 */
        rsold = rsnew;                                                 /* :132 */
    }

    if (info) {
        info->k = k;
        info->converged = converged;
        info->rsold = rsold;
        info->rsnew = rsnew;
        cgo_residual_check(n, A, ld, b, x, &info->norm_x, &info->rel_resid);
    }
    free(r); free(p); free(Ap);
}

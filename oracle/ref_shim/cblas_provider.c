/*
 * cblas_provider.c -- supplies cblas_dgemv / cblas_daxpy / cblas_ddot to the
 * UNMODIFIED reference MPI solver (call sites: /root/reference/code/MPI/cg.cc:80-151).
 * Test infrastructure only (oracle/): used to pin the oracle and as the CPU baseline.
 *
 * Provider selection (env CGREF_BLAS = auto | openblas | naive; default auto):
 *   openblas  forward to a real OpenBLAS found by dlopen: $CGREF_OPENBLAS_LIB, else the
 *             copy bundled in the opencv wheel of this image (OpenBLAS 0.3.15, LP64,
 *             plain cblas_* symbols), threaded by OPENBLAS_NUM_THREADS;
 *   naive     our own loops: left-to-right sums, OpenMP over the rows of dgemv.
 * The choice is printed once on stderr ("[cgref] blas = ...").
 *
 * Side channels for tests / bench (the reference never sees them):
 *   CGREF_HIST=<path>        every ddot(x, x) value (the r'r history; the last three are the
 *                            DEBUG block's r.r, b.b, x.x -- cg.cc:148-151), raw doubles.
 *   CGREF_GEMV_TIMES=<path>  CLOCK_MONOTONIC seconds at every dgemv entry and one final
 *                            stamp at exit, raw doubles (steady-state seconds/iteration).
 */
#define _GNU_SOURCE
#include "cblas.h"

#include <dlfcn.h>
#include <glob.h>
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void (*dgemv_fn)(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int, int, double, const double *,
                         int, const double *, int, double, double *, int);
typedef void (*daxpy_fn)(int, double, const double *, int, double *, int);
typedef double (*ddot_fn)(int, const double *, int, const double *, int);

static dgemv_fn real_dgemv;
static daxpy_fn real_daxpy;
static ddot_fn real_ddot;
static int initialised;

static double *hist_buf, *time_buf;
static size_t hist_n, hist_cap, time_n, time_cap;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void push(double **buf, size_t *n, size_t *cap, double v)
{
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 4096;
        *buf = (double *)realloc(*buf, *cap * sizeof(double));
    }
    (*buf)[(*n)++] = v;
}

static void dump(const char *env, const double *buf, size_t n)
{
    const char *path = getenv(env);
    if (!path) return;
    FILE *f = fopen(path, "wb");
    if (!f) return;
    fwrite(buf, sizeof(double), n, f);
    fclose(f);
}

static void at_exit_dump(void)
{
    if (getenv("CGREF_GEMV_TIMES")) push(&time_buf, &time_n, &time_cap, now_s());
    dump("CGREF_HIST", hist_buf, hist_n);
    dump("CGREF_GEMV_TIMES", time_buf, time_n);
}

static void *try_open(const char *path)
{
    /* the wheel copy needs its private libgfortran/libquadmath, which sit beside it */
    char dirbuf[4096];
    snprintf(dirbuf, sizeof dirbuf, "%s", path);
    const char *dir = dirname(dirbuf);
    const char *deps[] = {"libquadmath-*.so*", "libgfortran-*.so*"};
    for (int d = 0; d < 2; ++d) {
        char pat[4200];
        snprintf(pat, sizeof pat, "%s/%s", dir, deps[d]);
        glob_t g;
        if (glob(pat, 0, NULL, &g) == 0) {
            for (size_t i = 0; i < g.gl_pathc; ++i) dlopen(g.gl_pathv[i], RTLD_NOW | RTLD_GLOBAL);
            globfree(&g);
        }
    }
    return dlopen(path, RTLD_NOW | RTLD_LOCAL);
}

static void init_once(void)
{
    if (initialised) return;
    initialised = 1;
    atexit(at_exit_dump);
    const char *mode = getenv("CGREF_BLAS");
    if (!mode) mode = "auto";
    if (strcmp(mode, "naive") != 0) {
        void *h = NULL;
        char found[4096] = "";
        const char *explicit_path = getenv("CGREF_OPENBLAS_LIB");
        if (explicit_path) {
            h = try_open(explicit_path);
            snprintf(found, sizeof found, "%s", explicit_path);
        } else {
            const char *pats[] = {
                "/opt/prime-rl/.venv/lib/python3*/site-packages/opencv_python_headless.libs/"
                "libopenblasp-*.so",
                "/usr/lib/x86_64-linux-gnu/libopenblas.so*",
                "/usr/lib/x86_64-linux-gnu/openblas-pthread/libopenblas.so*",
            };
            for (size_t k = 0; !h && k < sizeof pats / sizeof *pats; ++k) {
                glob_t g;
                if (glob(pats[k], 0, NULL, &g) == 0) {
                    for (size_t i = 0; !h && i < g.gl_pathc; ++i) {
                        h = try_open(g.gl_pathv[i]);
                        if (h) snprintf(found, sizeof found, "%s", g.gl_pathv[i]);
                    }
                    globfree(&g);
                }
            }
        }
        if (h) {
            real_dgemv = (dgemv_fn)dlsym(h, "cblas_dgemv");
            real_daxpy = (daxpy_fn)dlsym(h, "cblas_daxpy");
            real_ddot = (ddot_fn)dlsym(h, "cblas_ddot");
            if (real_dgemv && real_daxpy && real_ddot) {
                typedef char *(*cfg_fn)(void);
                typedef int (*nt_fn)(void);
                cfg_fn cfg = (cfg_fn)dlsym(h, "openblas_get_config");
                nt_fn nt = (nt_fn)dlsym(h, "openblas_get_num_threads");
                fprintf(stderr, "[cgref] blas = openblas (%s; %s; threads=%d)\n", found,
                        cfg ? cfg() : "?", nt ? nt() : -1);
                return;
            }
            real_dgemv = NULL; real_daxpy = NULL; real_ddot = NULL;
        }
        if (strcmp(mode, "openblas") == 0) {
            fprintf(stderr, "[cgref] CGREF_BLAS=openblas but no OpenBLAS could be loaded\n");
            exit(3);
        }
    }
    fprintf(stderr, "[cgref] blas = naive (own loops, OpenMP over dgemv rows)\n");
}

void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, double alpha,
                 const double *a, int lda, const double *x, int incx, double beta, double *y,
                 int incy)
{
    init_once();
    if (getenv("CGREF_GEMV_TIMES")) push(&time_buf, &time_n, &time_cap, now_s());
    if (real_dgemv) {
        real_dgemv(order, trans, m, n, alpha, a, lda, x, incx, beta, y, incy);
        return;
    }
    if (order != CblasRowMajor || trans != CblasNoTrans) {
        fprintf(stderr, "[cgref] naive dgemv supports RowMajor/NoTrans only\n");
        exit(3);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        const double *row = a + (size_t)i * (size_t)lda;
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += row[j] * x[(size_t)j * incx];
        double *yi = y + (size_t)i * incy;
        *yi = (beta == 0.0) ? alpha * s : alpha * s + beta * *yi;
    }
}

void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy)
{
    init_once();
    if (real_daxpy) {
        real_daxpy(n, alpha, x, incx, y, incy);
        return;
    }
    for (int i = 0; i < n; ++i) y[(size_t)i * incy] += alpha * x[(size_t)i * incx];
}

double cblas_ddot(int n, const double *x, int incx, const double *y, int incy)
{
    init_once();
    double s;
    if (real_ddot) {
        s = real_ddot(n, x, incx, y, incy);
    } else {
        s = 0.0;
        for (int i = 0; i < n; ++i) s += x[(size_t)i * incx] * y[(size_t)i * incy];
    }
    if (x == y && getenv("CGREF_HIST")) push(&hist_buf, &hist_n, &hist_cap, s);
    return s;
}

/*
 * cblas.h -- prototypes for the three CBLAS routines the reference MPI solver
 * calls (/root/reference/code/MPI/cg.cc:80-151).  No system BLAS headers exist
 * in this image.  The symbols are provided by cblas_provider.c, which forwards
 * to a real OpenBLAS when one can be dlopen'ed and otherwise computes them
 * itself.  Test infrastructure only (oracle/).
 */
#ifndef CGB_ORACLE_CBLAS_H
#define CGB_ORACLE_CBLAS_H

#ifdef __cplusplus
extern "C" {
#endif

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, double alpha,
                 const double *a, int lda, const double *x, int incx, double beta, double *y,
                 int incy);
void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);

#ifdef __cplusplus
}
#endif
#endif

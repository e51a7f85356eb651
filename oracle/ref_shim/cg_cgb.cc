// cg_cgb.cc -- the reference-side binding of INTEGRATION.md section B, compiled for real.
//
// What a maintainer of federicobetti99/Conjugate-Gradient would add: a replacement for the body
// of CGSolver::solve (code/MPI/cg.cc:38-156) that hands the rank's rows of m_A and m_b to
// libcgb200.so (include/cgb200.h) -- one MPI rank per GPU, launched by the same mpirun.
// Everything else of the reference stays: cg_main.cc (main, timer, results row), cg.hh (this
// file includes it from the reference tree), the rest of cg.cc (generate_lap2d_matrix,
// init_source_term, partition_matrix, set_max_iter, read_matrix), Matrix, the reader.
//
// oracle/Makefile builds `_ref/cgsolver_ref_cgb` from the reference's own objects + this file:
// the reference's cg.o keeps its CGSolver::solve as a WEAK symbol (objcopy --weaken-symbol), so
// the definition below wins at link time, for direct calls and for the vtable slot alike.
// tests/test_gpu_cli.py runs it at P = 1 and P = 2 (fork shim, ref_shim/mpi_fork.cc) against
// the unmodified `cgsolver_ref`.  Test infrastructure: lives under oracle/, links the product.
#include "cg.hh" // the reference's own header (code/MPI/cg.hh), through -I

#include "cgb200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

namespace {
const bool DEBUG = true; // code/MPI/cg.cc:9

void die(const char *what)
{
    std::fprintf(stderr, "cgsolver (cgb200 binding): %s: %s\n", what, cgb_last_error());
    MPI_Abort(MPI_COMM_WORLD, 1);
    std::exit(1);
}
} // namespace

void CGSolver::solve(std::vector<double> & x) {
    int prank, psize;
    MPI_Comm_rank(MPI_COMM_WORLD, &prank);
    MPI_Comm_size(MPI_COMM_WORLD, &psize);

    int ndev = 0;
    if (cgb_device_count(&ndev) || ndev < 1) die("no CUDA device");
    cgb_ctx *ctx = nullptr;
    if (cgb_create(m_n, prank, psize, prank % ndev, &ctx)) die("cgb_create");

    if (psize > 1) {
        // replaces the communicator set-up (cg_main.cc:15-20): every rank learns every rank's
        // gather buffer; afterwards the mat-vec kernel stores its rows there directly
        char blob[CGB_EXCHANGE_BLOB_BYTES];
        std::vector<char> all((size_t)psize * CGB_EXCHANGE_BLOB_BYTES);
        if (cgb_exchange_export(ctx, blob)) die("cgb_exchange_export");
        MPI_Allgather(blob, CGB_EXCHANGE_BLOB_BYTES, MPI_BYTE, all.data(), CGB_EXCHANGE_BLOB_BYTES,
                      MPI_BYTE, MPI_COMM_WORLD);
        if (cgb_exchange_import(ctx, all.data())) die("cgb_exchange_import");
    }

    // cg.cc:80 reads m_A.data() + start_row * m_n; rows outside this rank's shard are ignored
    if (cgb_set_matrix_rows(ctx, m_A.data(), 0, m_m, m_n)) die("cgb_set_matrix_rows");
    if (cgb_set_rhs(ctx, m_b.data())) die("cgb_set_rhs");

    cgb_solve_info info;
    if (cgb_solve(ctx, x.data(), m_maxIter, m_tolerance, nullptr, &info)) die("cgb_solve"); // cg.cc:77-142

    double nx = 0., res = 0.;
    if (DEBUG) { // cg.cc:144-154; collective here (every rank's shard takes part in A x)
        if (cgb_residual_check(ctx, &nx, &res)) die("cgb_residual_check");
        if (prank == 0)
            std::cout << "\t[STEP " << info.k << "] residual = " << std::scientific
                      << std::sqrt(info.rsold) << ", ||x|| = " << nx << ", ||Ax - b||/||b|| = " << res
                      << std::endl;
    }
    if (psize > 1) MPI_Barrier(MPI_COMM_WORLD); // no rank unmaps its buffer while a peer may still write it
    cgb_destroy(ctx);
}

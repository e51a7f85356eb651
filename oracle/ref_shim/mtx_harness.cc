// mtx_harness.cc -- drives the UNMODIFIED reference MPI CGSolver on a Matrix Market file.
// Test infrastructure only (oracle/).
//
// Why it exists: the reference MPI main (/root/reference/code/MPI/cg_main.cc:31) only ever
// calls generate_lap2d_matrix, and CGSolver::read_matrix (cg.cc:191-202) does not set
// m_m / m_n / m_maxIter (they are private, cg.hh:41-50).  BASELINE.json config 1
// ("lap2D_5pt_n100.mtx via the reference MPI cgsolver, 1 rank") therefore needs this
// small main; everything it calls is the reference's own code.
//
// usage: cgsolver_ref_mtx file.mtx outfile [max_iter]
#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#define private public
#include "cg.hh"
#undef private

int main(int argc, char **argv)
{
    int provided;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_SINGLE, &provided);
    if (argc < 3) {
        std::cerr << "usage: " << argv[0] << " file.mtx outfile [max_iter]" << std::endl;
        return 1;
    }
    CGSolver solver;
    solver.read_matrix(argv[1]);
    int n = solver.m_A.n();
    solver.m_m = n;
    solver.m_n = n;
    solver.m_maxIter = n;
    if (argc >= 4) {
        int max_iter;
        std::stringstream s(argv[3]);
        s >> max_iter;
        solver.set_max_iter(max_iter);
    }
    solver.init_source_term(1. / n);
    std::vector<double> x(n, 0.);
    auto t1 = std::chrono::high_resolution_clock::now();
    solver.solve(x);
    std::chrono::duration<double> elapsed = std::chrono::high_resolution_clock::now() - t1;
    std::ofstream out(argv[2], std::ios_base::app);
    out << n << "," << 1 << "," << elapsed.count() << std::endl;
    MPI_Finalize();
    return 0;
}

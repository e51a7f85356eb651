// mtx_dump_harness.cc -- runs the UNMODIFIED reference reader (MatrixCOO::read,
// /root/reference/code/MPI/matrix_coo.cc:7-60 and Matrix::read, matrix.cc:6-22) and dumps what
// it parsed in the format of conjugate-gradient_b200/host/mtx_dump.cpp, so the tests can pin
// the product reader against the reference reader itself.  Test infrastructure only (oracle/).
//   mtx_dump_ref in.mtx out.bin           int32 m, n, nz, sym; int32 irn[nz]; int32 jcn[nz]; double a[nz]
//   mtx_dump_ref in.mtx out.bin --dense   int32 m, n; double dense[m * n]
#include <cstdio>
#include <cstring>

#include "matrix.hh"
#include "matrix_coo.hh"

int main(int argc, char **argv)
{
    if (argc < 3) return 64;
    if (argc >= 4 && std::strcmp(argv[3], "--dense") == 0) {
        Matrix A;
        A.read(argv[1]);
        std::FILE *f = std::fopen(argv[2], "wb");
        if (!f) return 65;
        const int head[2] = {A.m(), A.n()};
        std::fwrite(head, sizeof(int), 2, f);
        std::fwrite(A.data(), sizeof(double), (size_t)A.m() * (size_t)A.n(), f);
        std::fclose(f);
        return 0;
    }
    MatrixCOO coo;
    coo.read(argv[1]);
    std::FILE *f = std::fopen(argv[2], "wb");
    if (!f) return 65;
    const int head[4] = {coo.m(), coo.n(), coo.nz(), coo.is_sym()};
    std::fwrite(head, sizeof(int), 4, f);
    std::fwrite(coo.irn.data(), sizeof(int), coo.irn.size(), f);
    std::fwrite(coo.jcn.data(), sizeof(int), coo.jcn.size(), f);
    std::fwrite(coo.a.data(), sizeof(double), coo.a.size(), f);
    std::fclose(f);
    return 0;
}

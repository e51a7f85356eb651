// mtx_dump.cpp -- developer tool: runs the product Matrix Market reader and writes what it
// parsed as raw binary so the tests can compare it with the oracle's restatement.
//   mtx_dump in.mtx out.bin           int32 m, n, nz, sym; int32 irn[nz]; int32 jcn[nz]; double a[nz]
//   mtx_dump in.mtx out.bin --dense   int32 m, n; double dense[m * n]   (host Matrix::read)
#include "matrix.hpp"
#include "matrix_coo.hpp"

#include <cstdio>
#include <cstring>

int main(int argc, char **argv)
{
    if (argc < 3) return 64;
    std::FILE *f = nullptr;
    if (argc >= 4 && std::strcmp(argv[3], "--dense") == 0) {
        Matrix A;
        A.read(argv[1]);
        if (!(f = std::fopen(argv[2], "wb"))) return 65;
        const int head[2] = {static_cast<int>(A.m()), static_cast<int>(A.n())};
        std::fwrite(head, sizeof(int), 2, f);
        std::fwrite(A.data(), sizeof(double), static_cast<size_t>(A.m() * A.n()), f);
    } else {
        MatrixCOO coo;
        coo.read(argv[1]);
        if (!(f = std::fopen(argv[2], "wb"))) return 65;
        const int head[4] = {coo.m(), coo.n(), coo.nz(), coo.is_sym()};
        std::fwrite(head, sizeof(int), 4, f);
        std::fwrite(coo.irn.data(), sizeof(int), coo.irn.size(), f);
        std::fwrite(coo.jcn.data(), sizeof(int), coo.jcn.size(), f);
        std::fwrite(coo.a.data(), sizeof(double), coo.a.size(), f);
    }
    std::fclose(f);
    return 0;
}

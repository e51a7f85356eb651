// matrix.hpp -- dense row-major host matrix, the reference's `Matrix` container
// (/root/reference/code/MPI/matrix.hh:7-29) with 64-bit indexing (the reference's
// `int i * m_n + j` overflows at N >= 46341).  Only used for small inputs and tests: the
// solver never builds a dense matrix on the host (device-side generator / COO scatter).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "matrix_coo.hpp"

class Matrix {
public:
    Matrix(int64_t m = 0, int64_t n = 0) : m_m(m), m_n(n), m_a(static_cast<size_t>(m * n)) {}

    void resize(int64_t m, int64_t n)
    {
        m_m = m;
        m_n = n;
        m_a.assign(static_cast<size_t>(m * n), 0.0);
    }

    inline double &operator()(int64_t i, int64_t j) { return m_a[static_cast<size_t>(i * m_n + j)]; }

    inline int64_t m() const { return m_m; }
    inline int64_t n() const { return m_n; }
    inline double *data() { return m_a.data(); }

    /// densify a Matrix Market file (matrix.cc:6-22): later entries overwrite, symmetric
    /// banners mirror every entry.  The product path does this on the device
    /// (cgb_set_matrix_coo); this host version serves the reader tests (mtx_dump --dense).
    void read(const std::string &filename)
    {
        MatrixCOO coo;
        coo.read(filename);
        resize(coo.m(), coo.n());
        coo.scatter_dense(m_a.data(), m_n);
    }

private:
    int64_t m_m{0};
    int64_t m_n{0};
    std::vector<double> m_a;
};

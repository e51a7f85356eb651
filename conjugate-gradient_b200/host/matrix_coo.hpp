// matrix_coo.hpp -- Matrix Market coordinate reader of the host program.
// Mirrors the interface and behaviour of the reference's MatrixCOO
// (/root/reference/code/MPI/matrix_coo.hh:12-43, matrix_coo.cc:7-60) and of the three NIST
// mmio routines it reaches (mmio.c:96-179 mm_read_banner, :189-217 mm_read_mtx_crd_size,
// :455-511 mm_typecode_to_str): same accepted files, same messages, exit(1) on the same
// errors, 1-based -> 0-based indices, `%d %d %lg` triples.  Written from scratch.
#pragma once

#include <string>
#include <vector>

class MatrixCOO {
public:
    MatrixCOO() = default;

    inline int m() const { return m_m; }
    inline int n() const { return m_n; }
    inline int nz() const { return static_cast<int>(irn.size()); }
    inline int is_sym() const { return m_is_sym; }

    /// read a `%%MatrixMarket matrix coordinate <field> <symmetry>` file; exits like the
    /// reference when the file cannot be opened / is not a coordinate matrix
    void read(const std::string &filename);

    /// write the triples into a zero-initialised dense row-major buffer in file order
    /// (later duplicates win; a symmetric banner mirrors each entry) -- matrix.cc:12-21
    void scatter_dense(double *dense, long long ld) const;

    std::vector<int> irn;
    std::vector<int> jcn;
    std::vector<double> a;

private:
    int m_m{0};
    int m_n{0};
    bool m_is_sym{false};
};

namespace mm {
// Matrix Market banner, the subset of NIST mmio the reader needs.
struct TypeCode {
    char object = ' ';   // 'M' matrix
    char format = ' ';   // 'C' coordinate, 'A' array
    char field = ' ';    // 'R' real, 'C' complex, 'P' pattern, 'I' integer
    char symmetry = 'G'; // 'G' general, 'S' symmetric, 'H' hermitian, 'K' skew-symmetric
    bool is_matrix() const { return object == 'M'; }
    bool is_coordinate() const { return format == 'C'; }
    bool is_symmetric() const { return symmetry == 'S'; }
    std::string str() const; // "matrix coordinate real symmetric"
};
// return 0 on success, the NIST error code otherwise (12 premature EOF, 14 no header,
// 15 unsupported type)
int read_banner(std::FILE *f, TypeCode &tc);
int read_crd_size(std::FILE *f, int &m, int &n, int &nz);
} // namespace mm

// matrix_coo.cpp -- see matrix_coo.hpp.
#include "matrix_coo.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mm {

namespace {
constexpr int kMaxLine = 1025; // MM_MAX_LINE_LENGTH
constexpr int kPrematureEof = 12, kNoHeader = 14, kUnsupported = 15;

std::string lower(const char *s)
{
    std::string out(s);
    std::transform(out.begin(), out.end(), out.begin(), [](unsigned char c) { return std::tolower(c); });
    return out;
}
} // namespace

int read_banner(std::FILE *f, TypeCode &tc)
{
    tc = TypeCode();
    char line[kMaxLine];
    if (!std::fgets(line, kMaxLine, f)) return kPrematureEof;
    char banner[kMaxLine], obj[kMaxLine], fmt[kMaxLine], field[kMaxLine], sym[kMaxLine];
    if (std::sscanf(line, "%s %s %s %s %s", banner, obj, fmt, field, sym) != 5) return kPrematureEof;
    static const char kBanner[] = "%%MatrixMarket";
    if (std::strncmp(banner, kBanner, std::strlen(kBanner)) != 0) return kNoHeader;

    if (lower(obj) != "matrix") return kUnsupported;
    tc.object = 'M';

    const std::string sfmt = lower(fmt);
    if (sfmt == "coordinate") tc.format = 'C';
    else if (sfmt == "array") tc.format = 'A';
    else return kUnsupported;

    const std::string sfield = lower(field);
    if (sfield == "real") tc.field = 'R';
    else if (sfield == "complex") tc.field = 'C';
    else if (sfield == "pattern") tc.field = 'P';
    else if (sfield == "integer") tc.field = 'I';
    else return kUnsupported;

    const std::string ssym = lower(sym);
    if (ssym == "general") tc.symmetry = 'G';
    else if (ssym == "symmetric") tc.symmetry = 'S';
    else if (ssym == "hermitian") tc.symmetry = 'H';
    else if (ssym == "skew-symmetric") tc.symmetry = 'K';
    else return kUnsupported;
    return 0;
}

int read_crd_size(std::FILE *f, int &m, int &n, int &nz)
{
    m = n = nz = 0;
    char line[kMaxLine];
    do { // skip the comment block
        if (!std::fgets(line, kMaxLine, f)) return kPrematureEof;
    } while (line[0] == '%');
    if (std::sscanf(line, "%d %d %d", &m, &n, &nz) == 3) return 0;
    for (;;) { // blank line after the comments: take the next three integers
        const int got = std::fscanf(f, "%d %d %d", &m, &n, &nz);
        if (got == EOF) return kPrematureEof;
        if (got == 3) return 0;
    }
}

std::string TypeCode::str() const
{
    const char *fmt = format == 'C' ? "coordinate" : "array";
    const char *fld = field == 'R' ? "real" : field == 'C' ? "complex" : field == 'P' ? "pattern" : "integer";
    const char *sym = symmetry == 'G' ? "general"
                      : symmetry == 'S' ? "symmetric"
                      : symmetry == 'H' ? "hermitian"
                                        : "skew-symmetric";
    return std::string("matrix ") + fmt + " " + fld + " " + sym;
}

} // namespace mm

void MatrixCOO::read(const std::string &fn)
{
    std::FILE *f = std::fopen(fn.c_str(), "r");
    if (!f) { // matrix_coo.cc:13-16
        std::printf("Could not open matrix");
        std::exit(1);
    }
    mm::TypeCode tc;
    if (mm::read_banner(f, tc) != 0) { // matrix_coo.cc:18-21
        std::printf("Could not process Matrix Market banner.\n");
        std::exit(1);
    }
    if (!(tc.is_matrix() && tc.is_coordinate())) { // matrix_coo.cc:24-28
        std::printf("Sorry, this application does not support ");
        std::printf("Market Market type: [%s]\n", tc.str().c_str());
        std::exit(1);
    }
    int nz = 0;
    if (mm::read_crd_size(f, m_m, m_n, nz) != 0) std::exit(1); // matrix_coo.cc:30-32

    irn.assign(static_cast<size_t>(nz), 0);
    jcn.assign(static_cast<size_t>(nz), 0);
    a.assign(static_cast<size_t>(nz), 0.0);
    m_is_sym = tc.is_symmetric();
    for (int z = 0; z < nz; ++z) { // matrix_coo.cc:43-55: unchecked "%d %d %lg\n"
        int i = 0, j = 0;
        double v = 0.0;
        if (std::fscanf(f, "%d %d %lg\n", &i, &j, &v) != 3) { /* the reference ignores it too */ }
        irn[z] = i - 1;
        jcn[z] = j - 1;
        a[z] = v;
    }
    std::fclose(f);
}

void MatrixCOO::scatter_dense(double *dense, long long ld) const
{
    const size_t count = irn.size();
    for (size_t z = 0; z < count; ++z) {
        const long long row = irn[z], col = jcn[z];
        dense[row * ld + col] = a[z];
        if (m_is_sym) dense[col * ld + row] = a[z];
    }
}

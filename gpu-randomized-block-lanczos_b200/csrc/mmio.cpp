// Matrix Market coordinate reader (host only): the loader side of the reference's benchmark driver
// (Julia/benchmark.jl:3,21,28 `mmread("../Matrix/audi.mtx")` through MatrixMarket.jl).  Produces what rbl_create takes:
// the CSC arrays of the FULL symmetric matrix (Int64 colptr / rowval, Float64 nzval, sorted row indices), symmetric
// storage expanded, duplicate entries summed (as SparseArrays.sparse does).
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mmio.h"

namespace rbl {

namespace {
std::string lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
}  // namespace

bool read_matrix_market(const char* path, int index_base, MmMatrix& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<char> line(1 << 16);
    if (!std::fgets(line.data(), (int)line.size(), f)) { std::fclose(f); err = "empty file"; return false; }
    char banner[64], obj[64], fmt[64], field[64], sym[64];
    if (std::sscanf(line.data(), "%63s %63s %63s %63s %63s", banner, obj, fmt, field, sym) != 5 ||
        lower(banner) != "%%matrixmarket" || lower(obj) != "matrix") {
        std::fclose(f); err = "not a MatrixMarket matrix file"; return false;
    }
    const std::string sfmt = lower(fmt), sfield = lower(field), ssym = lower(sym);
    if (sfmt != "coordinate") { std::fclose(f); err = "only coordinate (sparse) MatrixMarket files are supported"; return false; }
    const bool pattern = sfield == "pattern";
    if (!(pattern || sfield == "real" || sfield == "integer" || sfield == "double")) {
        std::fclose(f); err = "unsupported MatrixMarket field '" + sfield + "' (need real, integer or pattern)"; return false;
    }
    const bool symmetric = ssym == "symmetric";
    const bool skew = ssym == "skew-symmetric";
    if (!(symmetric || skew || ssym == "general")) { std::fclose(f); err = "unsupported MatrixMarket symmetry '" + ssym + "'"; return false; }
    // size line (skip comments / blank lines)
    long long M = 0, N = 0, NZ = 0;
    for (;;) {
        if (!std::fgets(line.data(), (int)line.size(), f)) { std::fclose(f); err = "missing size line"; return false; }
        const char* p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '%' || *p == '\n' || *p == '\r' || *p == 0) continue;
        if (std::sscanf(p, "%lld %lld %lld", &M, &N, &NZ) != 3) { std::fclose(f); err = "bad size line"; return false; }
        break;
    }
    if (M != N) { std::fclose(f); err = "matrix is not square"; return false; }
    if (M <= 0 || NZ < 0) { std::fclose(f); err = "bad dimensions"; return false; }
    struct Ent { int64_t r, c; double v; };
    std::vector<Ent> ents;
    ents.reserve((size_t)NZ * ((symmetric || skew) ? 2 : 1));
    for (long long e = 0; e < NZ; ++e) {
        long long i = 0, j = 0;
        double v = 1.0;
        int got = pattern ? std::fscanf(f, "%lld %lld", &i, &j) : std::fscanf(f, "%lld %lld %lf", &i, &j, &v);
        if (got != (pattern ? 2 : 3)) { std::fclose(f); err = "truncated entry list at entry " + std::to_string(e); return false; }
        if (i < 1 || i > M || j < 1 || j > N) { std::fclose(f); err = "entry index out of range"; return false; }
        ents.push_back({i - 1, j - 1, v});
        if ((symmetric || skew) && i != j) ents.push_back({j - 1, i - 1, skew ? -v : v});
    }
    std::fclose(f);
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
    out.n = M;
    out.colptr.assign((size_t)M + 1, 0);
    out.rowval.clear();
    out.nzval.clear();
    out.rowval.reserve(ents.size());
    out.nzval.reserve(ents.size());
    for (size_t e = 0; e < ents.size();) {
        size_t e2 = e;
        double s = 0.0;
        while (e2 < ents.size() && ents[e2].c == ents[e].c && ents[e2].r == ents[e].r) s += ents[e2++].v;   // duplicates add up
        out.rowval.push_back(ents[e].r + index_base);
        out.nzval.push_back(s);
        out.colptr[(size_t)ents[e].c + 1] += 1;
        e = e2;
    }
    for (int64_t c = 0; c < M; ++c) out.colptr[(size_t)c + 1] += out.colptr[(size_t)c];
    for (auto& p : out.colptr) p += index_base;
    out.symmetric_storage = symmetric;
    return true;
}

}  // namespace rbl

// Matrix Market reader (host only) - see mmio.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rbl {

struct MmMatrix {
    int64_t n = 0;
    std::vector<int64_t> colptr, rowval;   // CSC of the full matrix, index_base as requested
    std::vector<double> nzval;
    bool symmetric_storage = false;
};

bool read_matrix_market(const char* path, int index_base, MmMatrix& out, std::string& err);

}  // namespace rbl

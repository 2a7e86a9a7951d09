#include "partition.h"

#include <algorithm>

namespace rbl {

void partition_rows(int64_t n, int world, int64_t* row_starts) {
    for (int p = 0; p <= world; ++p) row_starts[p] = (int64_t)(((__int128)n * p) / world);
}

bool halo_plan(int64_t n, int world, const int64_t* row_starts, int rank, int64_t nloc, int64_t nnz,
               const int64_t* rowptr, const int64_t* colidx_global, int index_base, HaloPlan& out) {
    (void)rowptr;
    const int64_t r0 = row_starts[rank], r1 = row_starts[rank + 1];
    if (r1 - r0 != nloc) return false;
    std::vector<int64_t> ext;
    ext.reserve(1024);
    for (int64_t p = 0; p < nnz; ++p) {
        const int64_t c = colidx_global[p] - index_base;
        if (c < 0 || c >= n) return false;
        if (c < r0 || c >= r1) ext.push_back(c);
    }
    std::sort(ext.begin(), ext.end());
    ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
    out.halo_cols = ext;
    out.halo_owner_ptr.assign(world + 1, 0);
    for (int p = 0; p < world; ++p) {
        // halo columns owned by rank p are those in [row_starts[p], row_starts[p+1])
        out.halo_owner_ptr[p + 1] =
            (int64_t)(std::lower_bound(ext.begin(), ext.end(), row_starts[p + 1]) - ext.begin());
    }
    out.colidx_local.resize(nnz);
    for (int64_t p = 0; p < nnz; ++p) {
        const int64_t c = colidx_global[p] - index_base;
        if (c >= r0 && c < r1) {
            out.colidx_local[p] = (int32_t)(c - r0);
        } else {
            const int64_t pos = (int64_t)(std::lower_bound(ext.begin(), ext.end(), c) - ext.begin());
            out.colidx_local[p] = (int32_t)(nloc + pos);
        }
    }
    return true;
}

}  // namespace rbl

// 1-D contiguous row partition and SpMM halo plan (integer work, host only).  New design: the
// reference is single-GPU.  Bit-exact twin: oracle/partition_oracle.py.
#pragma once
#include <cstdint>
#include <vector>

namespace rbl {

void partition_rows(int64_t n, int world, int64_t* row_starts /* world+1 */);

struct HaloPlan {
    std::vector<int64_t> halo_cols;       // sorted unique global columns outside the owned range
    std::vector<int64_t> halo_owner_ptr;  // world+1 offsets into halo_cols by owner rank
    std::vector<int32_t> colidx_local;    // nnz remapped column indices into [own rows | halo]
};

// colidx_global 0-based.  Returns false on an out-of-range index.
bool halo_plan(int64_t n, int world, const int64_t* row_starts, int rank, int64_t nloc, int64_t nnz,
               const int64_t* rowptr, const int64_t* colidx_global, int index_base, HaloPlan& out);

}  // namespace rbl

// Row schedule for stencil / banded matrices (host side): the rows of the matrix, grouped into compact patches of the grid
// the stencil offsets imply.  Used by the SpMM laboratory (spmm_lab.cu) - NOT by the solve path.
//
// Idea: rows of a block are whole 128-byte lines (B = 16), so visiting rows in any order costs nothing in coalescing.  A
// stencil matrix's offsets col-row cluster around a few strides (1, N, N^2 for the BASELINE Laplacians; 1, W for the image
// graph); rows are points of a 1-3 dimensional grid, and a CTA that processes a compact patch (say 10 x 5 x 5 points)
// finds most neighbours inside the patch: 1.86 distinct rows of Q per row instead of 5.06 for 32 consecutive rows.
//
// Measured (tools/spmm_lab.py, B200, config-2 matrix): the L2->SM traffic drops as modelled, the launch time does not
// (95 -> 99-103 us): ncu shows L2 at 39% and DRAM at 52% of their peaks with the L1 data pipe the busiest unit (67%),
// so saving L2 traffic buys nothing.  Two kernels built on the schedule were removed again: a cooperative-load variant
// (lane j of a row group loads entry j, shuffles broadcast it - shuffles go through the same L1 data pipe, 138 us) and
// the straight port of the gather kernel (101 us).  DESIGN.md section 4 has the table.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"

namespace rbl {

namespace {

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return (e && e[0]) ? std::atoi(e) : dflt;
}

}  // namespace

// Host: grid structure from a sample of the rows, then the patch-ordered schedule.  Returns false (order untouched) when
// the offsets do not look like a stencil; the caller then keeps the plain gather kernel.
bool spmm_plan_schedule(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx, int slots, std::vector<int>& order,
                        SpmmSchedule* info) {
    if (nrows < 16384 || slots < 32) return false;
    std::vector<int64_t> offs;
    const int64_t step = std::max<int64_t>(1, nrows / 16384);
    for (int64_t r = 0; r < nrows; r += step)
        for (int p = rowptr[r]; p < rowptr[r + 1]; ++p)
            if (colidx[p] < nown && colidx[p] != r) offs.push_back(std::llabs((int64_t)colidx[p] - r));
    if (offs.empty()) return false;
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    if (offs.size() > 256) return false;           // no diagonal structure (e.g. the Erdos-Renyi config)
    // clusters of offsets: members within `near` of their predecessor, `near` = the reach of the innermost cluster
    // (offsets 1..h around the diagonal) or 2
    int64_t h0 = 0;
    while (h0 < (int64_t)offs.size() && offs[h0] == h0 + 1) ++h0;     // offsets 1, 2, .., h0 present
    const int64_t near = std::max<int64_t>(2, h0);
    struct Cl { int64_t lo, hi; };
    std::vector<Cl> cl;
    for (size_t i = (size_t)h0; i < offs.size(); ++i) {
        if (!cl.empty() && offs[i] - cl.back().hi <= near) cl.back().hi = offs[i];
        else cl.push_back({offs[i], offs[i]});
    }
    if (cl.size() > 2) return false;               // more than three grid dimensions: not handled
    SpmmSchedule sc{};
    sc.dims = 1 + (int)cl.size();
    sc.stride[0] = 1;
    sc.halo[0] = (int)h0;
    for (size_t i = 0; i < cl.size(); ++i) {
        const int64_t c = (cl[i].lo + cl[i].hi) / 2;
        if (c < 8 * sc.stride[i] || c >= nrows) return false;
        sc.stride[i + 1] = c;
        sc.halo[i + 1] = 1;
        sc.halo[0] = (int)std::max<int64_t>(sc.halo[0], std::max(cl[i].hi - c, c - cl[i].lo));
    }
    // grid extents of the decomposition r = x + stride1 * y + stride2 * z (x < stride1, x + stride1*y < stride2)
    int64_t ext[3] = {1, 1, 1};
    if (sc.dims == 1) ext[0] = nrows;
    else if (sc.dims == 2) { ext[0] = sc.stride[1]; ext[1] = (nrows + sc.stride[1] - 1) / sc.stride[1]; }
    else { ext[0] = sc.stride[1]; ext[1] = (sc.stride[2] + sc.stride[1] - 1) / sc.stride[1]; ext[2] = (nrows + sc.stride[2] - 1) / sc.stride[2]; }
    for (int i = 0; i < 3; ++i) sc.ext[i] = ext[i];
    // alternative order (RBL_SPMM_ZTILE = tile height in y): strips of the x-y plane swept along z, so that the +-stride2
    // neighbours of a row were touched one strip-plane (not one whole plane) ago - for planes too large for the L2
    if (sc.dims == 3 && env_int("RBL_SPMM_ZTILE", 0) > 0) {
        const int64_t ty = env_int("RBL_SPMM_ZTILE", 0);
        std::vector<int> ord;
        ord.reserve((size_t)nrows + slots);
        const int64_t s1 = sc.stride[1], s2 = sc.stride[2];
        for (int64_t y0 = 0; y0 < ext[1]; y0 += ty)
            for (int64_t z = 0; z < ext[2]; ++z)
                for (int64_t y = y0; y < std::min(ext[1], y0 + ty); ++y)
                    for (int64_t x = 0; x < ext[0]; ++x) {
                        const int64_t rem = y * s1 + x, r = z * s2 + rem;
                        if (rem < s2 && r < nrows) ord.push_back((int)r);
                    }
        if ((int64_t)ord.size() != nrows) return false;
        while (ord.size() % slots) ord.push_back(-1);
        sc.patch[0] = (int)ext[0]; sc.patch[1] = (int)ty; sc.patch[2] = 1;
        sc.slots = slots;
        sc.npatch = (int64_t)ord.size() / slots;
        sc.fetch_model = 0.0;
        order.swap(ord);
        if (info) *info = sc;
        return true;
    }
    // patch shape: minimise rows of Q fetched per row, prod_i (1 + 2 halo_i / p_i), times the slot padding, over shapes
    // with p0*p1*p2 <= slots; x runs of at least 8 rows (1 KB contiguous pieces of every stream) when the grid allows
    double best = 1e300;
    int bp[3] = {1, 1, 1};
    const int p0min = (int)std::min<int64_t>(8, ext[0]);
    for (int p0 = p0min; p0 <= slots && p0 <= ext[0]; ++p0)
        for (int p1 = 1; p0 * p1 <= slots && p1 <= ext[1]; ++p1) {
            int p2 = (int)std::min<int64_t>(slots / (p0 * p1), ext[2]);
            if (p2 < 1) continue;
            if (sc.dims < 3) p2 = 1;
            if (sc.dims < 2 && p1 > 1) continue;
            const int pp[3] = {p0, p1, p2};
            double fetch = 1.0, pad = (double)slots / (double)(p0 * p1 * p2);
            for (int i = 0; i < sc.dims; ++i) {
                fetch *= 1.0 + 2.0 * sc.halo[i] / pp[i];
                const int64_t np = (ext[i] + pp[i] - 1) / pp[i];
                pad *= (double)(np * pp[i]) / (double)ext[i];
            }
            const double cost = fetch * (1.0 + 0.5 * (pad - 1.0));
            if (cost < best) { best = cost; bp[0] = p0; bp[1] = p1; bp[2] = p2; }
        }
    for (int i = 0; i < 3; ++i) sc.patch[i] = bp[i];
    sc.slots = slots;
    const int64_t np0 = (ext[0] + bp[0] - 1) / bp[0], np1 = (ext[1] + bp[1] - 1) / bp[1], np2 = (ext[2] + bp[2] - 1) / bp[2];
    const int64_t max_patches = np0 * np1 * np2;
    if ((double)max_patches * slots > 1.3 * (double)nrows + 65536 || max_patches * (int64_t)slots >= (int64_t)1 << 31) return false;
    std::vector<int> ord;
    ord.reserve((size_t)max_patches * slots);
    const int64_t s1 = sc.dims >= 2 ? sc.stride[1] : nrows, s2 = sc.dims >= 3 ? sc.stride[2] : nrows;
    int64_t placed = 0, npatch = 0;
    for (int64_t b2 = 0; b2 < np2; ++b2)
        for (int64_t b1 = 0; b1 < np1; ++b1)
            for (int64_t b0 = 0; b0 < np0; ++b0) {
                const size_t start = ord.size();
                int cnt = 0;
                for (int d2 = 0; d2 < bp[2]; ++d2)
                    for (int d1 = 0; d1 < bp[1]; ++d1)
                        for (int d0 = 0; d0 < bp[0]; ++d0) {
                            const int64_t x = b0 * bp[0] + d0, y = b1 * bp[1] + d1, z = b2 * bp[2] + d2;
                            int row = -1;
                            if (x < ext[0] && y < ext[1] && z < ext[2]) {
                                const int64_t rem = y * s1 + x;
                                const int64_t r = z * s2 + rem;
                                if ((sc.dims < 3 || rem < s2) && r < nrows) row = (int)r;
                            }
                            if (row >= 0) ++cnt;
                            ord.push_back(row);
                        }
                if (cnt == 0) { ord.resize(start); continue; }
                ord.resize(start + slots, -1);
                placed += cnt;
                ++npatch;
            }
    if (placed != nrows) return false;             // (cannot happen: the decomposition is a bijection)
    sc.npatch = npatch;
    sc.fetch_model = best;
    order.swap(ord);
    if (info) *info = sc;
    return true;
}

int spmm_sched_default_slots(int B) { return env_int("RBL_SPMM_PATCH", B == 32 ? 256 : 256); }

}  // namespace rbl

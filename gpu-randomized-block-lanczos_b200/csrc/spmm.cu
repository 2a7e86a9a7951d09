// K1, second generation: block SpMM for banded / stencil matrices with the b-wide Q block staged in shared memory by
// the TMA engine (cp.async.bulk + mbarrier) - BASELINE north_star (2).  Replaces `mul!(U,Ag,Qg_d)` (cuSPARSE SpMM,
// Julia/RBL_gpu.jl:152,176) like kernels.cu spmm_kernel, which stays the path for unstructured matrices.
//
// Why: the gather kernel reads every one of the nnz/row neighbour rows of Q through L1/L2 (ncu on config 2: 735 MB of
// L2->SM traffic for 343 MB of DRAM traffic, 80% of the L2 throughput the chip sustains, warps stalled on the dependent
// rowptr -> colidx -> Q loads).  Matrices of the BASELINE Laplacian configs keep their entries on a few diagonals
// bands: offsets col-row form a handful of clusters ("windows").  A persistent CTA walks a contiguous range of rows in
// tiles of R rows; for every window it keeps the rows of Q the current tile can reference in a shared-memory RING and
// asks the TMA engine for the R new rows of the next tile while it computes the current one.  Every row of Q then
// crosses L2->SM once per window (3 Q + A bytes on config 2 instead of ~5.7 Q + A), in large contiguous bulk copies.
//
// Per nonzero the device-side "relative index" array (built once per handle and block size from colidx) holds
// (window id, offset inside the window) so that the ring slot is head[w] + local row + offset (one conditional
// wrap); entries that fall outside every window (halo columns of a row-sharded matrix, stray entries) keep their
// column index and are gathered from global memory as before - correctness never depends on the window table.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"

namespace rbl {

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA engine, 1-D bulk copy global -> shared; completion is signalled on the mbarrier (bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace

// irregular entries are encoded as -1 - column
__global__ void spmm_build_rel_kernel(int64_t nrows, int64_t nown, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                      SpmmWindows wt, int* __restrict__ rel, unsigned long long* __restrict__ irregular) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad = 0;
    if (r < nrows) {
        for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) {
            const int c = colidx[p];
            const int64_t d = (int64_t)c - r;
            int enc = -1 - c;
            if (c < nown) {
                for (int w = 0; w < wt.nwin; ++w)
                    if (d >= wt.lo[w] && d <= wt.hi[w]) {
                        enc = (w << 24) | (int)(d - wt.lo[w]);
                        break;
                    }
            }
            if (enc < 0) ++bad;
            rel[p] = enc;
        }
    }
    for (int off = 16; off > 0; off >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, off);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(irregular, bad);
}

// Host: window table from a sample of rows.  Returns nwin = 0 when the matrix has no such structure.
SpmmWindows spmm_plan_windows(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx) {
    SpmmWindows wt{};
    constexpr int R = SpmmWindows::kTileRows;
    if (nrows < 4096) return wt;
    std::vector<int64_t> offs;
    const int64_t step = std::max<int64_t>(1, nrows / 8192);
    for (int64_t r = 0; r < nrows; r += step)
        for (int p = rowptr[r]; p < rowptr[r + 1]; ++p)
            if (colidx[p] < nown) offs.push_back((int64_t)colidx[p] - r);
    if (offs.empty()) return wt;
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    if (offs.size() > 4096) return wt;   // no diagonal structure
    // clusters: a window costs NB*R ring rows, a merged gap costs the gap: merge the small gaps
    static const int64_t gap_merge = [] { const char* e = std::getenv("RBL_SPMM_MERGE_GAP"); return e ? std::atoll(e) : 0ll; }();
    const int64_t merge = gap_merge > 0 ? gap_merge : 4 * R;
    std::vector<std::pair<int64_t, int64_t>> win;
    for (int64_t d : offs) {
        if (!win.empty() && d - win.back().second < merge) win.back().second = d;
        else win.push_back({d, d});
    }
    if ((int)win.size() > SpmmWindows::kMax) return wt;
    for (auto& w : win)
        if (w.second - w.first > (1 << 20)) return wt;
    int64_t mx = 0;
    for (int64_t g = 0; g < nrows; g += R) mx = std::max<int64_t>(mx, (int64_t)rowptr[std::min<int64_t>(nrows, g + R)] - rowptr[g]);
    if (mx > 4096) return wt;
    wt.max_tile_nnz = (int)mx;
    wt.nwin = (int)win.size();
    for (int w = 0; w < wt.nwin; ++w) {
        wt.lo[w] = (int)win[w].first;
        wt.hi[w] = (int)win[w].second;
    }
    wt.diag_w = -1;   // the diagonal (beta * Q[r,:]) comes out of a ring when some window holds offset 0
    for (int w = 0; w < wt.nwin; ++w)
        if (wt.lo[w] <= 0 && wt.hi[w] >= 0) wt.diag_w = w;
    return wt;
}

// shared-memory layout for NB stages: ring bases (rows) and total bytes.  Returns 0 stages when nothing fits.
namespace {
struct SlotShape {
    int rp_words, rel_words, val_words;
    size_t bytes;
};
__host__ __device__ inline SlotShape slot_shape(int max_tile_nnz) {
    SlotShape s;
    s.rp_words = SpmmWindows::kTileRows + 4;
    s.rel_words = (max_tile_nnz + 8 + 3) & ~3;
    s.val_words = (max_tile_nnz + 4 + 1) & ~1;
    s.bytes = 128 + (size_t)s.rp_words * 4 + (size_t)s.rel_words * 4 + (size_t)s.val_words * 8;
    return s;
}
}  // namespace
int spmm_window_stages(SpmmWindows& wt, int B, size_t* smem_bytes_out) {
    constexpr int R = SpmmWindows::kTileRows;
    const SlotShape ss = slot_shape(wt.max_tile_nnz);
    for (int NB : {8, 6, 5, 4, 3}) {
        size_t rows = 0;
        for (int w = 0; w < wt.nwin; ++w) rows += (size_t)(NB * R + (wt.hi[w] - wt.lo[w]));
        const size_t bytes = 128 + (size_t)NB * ss.bytes + rows * (size_t)B * 8;
        if (bytes <= (size_t)200 * 1024) {
            int base = 0;
            for (int w = 0; w < wt.nwin; ++w) {
                wt.base[w] = base;
                base += NB * R + (wt.hi[w] - wt.lo[w]);
            }
            *smem_bytes_out = bytes;
            return NB;
        }
    }
    return 0;
}

// Warp-specialised: warps 0..15 compute (one row per LPR-lane group and tile), warp 16 is the TMA producer.  Per tile the
// producer asks the TMA engine for (a) the R new rows of Q of every window ring, (b) the tile's slice of the CSR stream
// (row pointers, relative indices, values) - everything the consumers touch is in shared memory.  NB tiles are in flight
// (full / empty mbarrier pairs); ring capacity is NB*R + (hi - lo) rows so that the rows tile t brings in replace
// exactly the rows only tile t-NB could reference.
template <int B>
__global__ void __launch_bounds__(544, 1)
    spmm_window_kernel(int64_t nrows, int64_t nown, const int* __restrict__ rowptr, const int* __restrict__ rel,
                       const double* __restrict__ vals, const double* __restrict__ Q, double* U, SpmmCoef cf, const double* Z,
                       SpmmWindows wt, int64_t rows_per_cta, int NB, int max_tile_nnz) {
    constexpr int LPR = B / 2;            // lanes per row, two columns each
    constexpr int RPW = 32 / LPR;         // rows per warp and pass
    constexpr int R = SpmmWindows::kTileRows;
    constexpr int NPASS = R / (16 * RPW); // 1 (B = 16) or 2 (B = 32)
    extern __shared__ __align__(128) unsigned char smraw[];
    // layout: [full barriers | empty barriers] (128 B) | per-slot CSR slices | window rings
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smraw);
    unsigned long long* empty = full + SpmmWindows::kMaxStages;
    const SlotShape ss = slot_shape(max_tile_nnz);
    const int rp_words = ss.rp_words, rel_words = ss.rel_words;
    const size_t slot_bytes = ss.bytes;       // [ring heads, capacities, bases: 128 B | row pointers | relative indices | values]
    unsigned char* slots = smraw + 128;
    double* ring = reinterpret_cast<double*>(slots + (size_t)NB * slot_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t bend = min(nrows, a + rows_per_cta);
    if (a >= bend) return;
    const int ntiles = (int)((bend - a + R - 1) / R);
    const double2* __restrict__ Q2 = reinterpret_cast<const double2*>(Q);

    if (tid == 0) {
        for (int q = 0; q < NB; ++q) {
            mbar_init(&full[q], 1);
            mbar_init(&empty[q], 16);   // one arrival per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 16) {
        // ===== TMA producer =====
        int head[SpmmWindows::kMax];      // (t * R) mod ring capacity, kept incrementally (lane 0)
#pragma unroll
        for (int w = 0; w < SpmmWindows::kMax; ++w) head[w] = 0;
        // The CSR slice of a tile is bounded by two row pointers.  A dependent global load per tile in the issuing lane
        // would cap the producer at one tile per memory latency (measured: 1.9 us per 64-row tile): the 32 lanes fetch
        // the bounds of 32 tiles at once, lane 0 then issues the tiles with the bounds handed over by shuffles.
        for (int t0 = 0; t0 < ntiles; t0 += 32) {
            const int64_t rl = min(bend, a + (int64_t)(t0 + lane) * R);
            const int pl = __ldg(rowptr + rl);
            const int pnext = __ldg(rowptr + min(bend, a + (int64_t)(t0 + 32) * R));
          for (int j = 0; j < 32 && t0 + j < ntiles; ++j) {
            const int t = t0 + j;
            const int p0 = __shfl_sync(0xffffffffu, pl, j);
            const int p1s = __shfl_sync(0xffffffffu, pl, (j + 1) & 31);
            const int p1 = (j < 31) ? p1s : pnext;
            if (lane == 0) {
                const int slot = t % NB;
                if (t >= NB) mbar_wait(&empty[slot], (unsigned)(((t / NB) - 1) & 1));
                const int64_t r0 = a + (int64_t)t * R;
                const int p0r = p0 & ~3, p0v = p0 & ~1;                       // 16-byte aligned starts
                const unsigned rel_bytes = (unsigned)(((p1 - p0r) * 4 + 15) & ~15);
                const unsigned val_bytes = (unsigned)(((p1 - p0v) * 8 + 15) & ~15);
                const unsigned rp_bytes = (unsigned)(rp_words * 4);
                unsigned bytes = rp_bytes + rel_bytes + val_bytes;
                unsigned char* sl = slots + (size_t)slot * slot_bytes;
                int* hdr = reinterpret_cast<int*>(sl);
                // window rows: count first (the transaction count must be armed before the copies can complete it)
#pragma unroll
                for (int w = 0; w < SpmmWindows::kMax; ++w)
                    if (w < wt.nwin) {
                        int64_t g0 = (t == 0) ? r0 + wt.lo[w] : r0 + wt.hi[w];
                        int64_t g1 = r0 + R + wt.hi[w];
                        g0 = max(g0, (int64_t)0);
                        g1 = min(g1, nown);
                        if (g1 > g0) bytes += (unsigned)((g1 - g0) * B * 8);
                        hdr[w] = head[w];
                        hdr[8 + w] = NB * R + (wt.hi[w] - wt.lo[w]);
                        hdr[16 + w] = wt.base[w];
                    }
                mbar_expect_tx(&full[slot], bytes);      // (release: the heads written above are visible after the wait)
                bulk_g2s(sl + 128, rowptr + r0, rp_bytes, &full[slot]);
                bulk_g2s(sl + 128 + (size_t)rp_words * 4, rel + p0r, rel_bytes, &full[slot]);
                bulk_g2s(sl + 128 + (size_t)rp_words * 4 + (size_t)rel_words * 4, vals + p0v, val_bytes, &full[slot]);
#pragma unroll
                for (int w = 0; w < SpmmWindows::kMax; ++w)
                    if (w < wt.nwin) {
                        const int sw = wt.hi[w] - wt.lo[w];
                        const int cap = NB * R + sw;
                        const int64_t g0u = (t == 0) ? r0 + wt.lo[w] : r0 + wt.hi[w];
                        const int64_t g0 = max(g0u, (int64_t)0);
                        const int64_t g1 = min(r0 + R + wt.hi[w], nown);
                        if (g1 > g0) {
                            // ring slot of row g: (g - a - lo) mod cap = head + (t == 0 ? 0 : s) + clip, at most two wraps
                            int64_t pos = head[w] + (t == 0 ? 0 : sw) + (g0 - g0u);
                            while (pos >= cap) pos -= cap;
                            const int64_t cnt = g1 - g0;
                            const int64_t first = min(cnt, (int64_t)cap - pos);
                            bulk_g2s(ring + (size_t)(wt.base[w] + pos) * B, Q + (size_t)g0 * B, (unsigned)(first * B * 8), &full[slot]);
                            if (cnt > first)
                                bulk_g2s(ring + (size_t)wt.base[w] * B, Q + (size_t)(g0 + first) * B, (unsigned)((cnt - first) * B * 8), &full[slot]);
                        }
                        head[w] += R;
                        if (head[w] >= cap) head[w] -= cap;
                    }
            }
          }
        }
        return;
    }

    // ===== consumers =====
    const int sub = lane % LPR, rsel = lane / LPR;
    for (int t = 0; t < ntiles; ++t) {
        const int slot = t % NB;
        mbar_wait(&full[slot], (unsigned)((t / NB) & 1));
        const unsigned char* sl = slots + (size_t)slot * slot_bytes;
        const int* s_hdr = reinterpret_cast<const int*>(sl);     // [0..8) ring heads, [8..16) capacities, [16..24) bases
        const int* s_rp = reinterpret_cast<const int*>(sl + 128);
        const int p0 = s_rp[0];
        const int* s_rel = reinterpret_cast<const int*>(sl + 128 + (size_t)rp_words * 4) - (p0 & ~3);       // indexed with the global p
        const double* s_val = reinterpret_cast<const double*>(sl + 128 + (size_t)rp_words * 4 + (size_t)rel_words * 4) - (p0 & ~1);
        const int64_t trow = (int64_t)t * R;   // first row of the tile relative to `a`
        // ring address (in doubles) of the Q row an entry refers to
        auto ring_row = [&](int enc, int li) -> const double* {
            const int w = enc >> 24;
            int pos = s_hdr[w] + li + (enc & 0xffffff);
            const int cap = s_hdr[8 + w];
            if (pos >= cap) pos -= cap;
            return ring + (size_t)(s_hdr[16 + w] + pos) * B;
        };
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
            const int li = ps * 16 * RPW + warp * RPW + rsel;      // row inside the tile
            const int64_t row = a + trow + li;
            if (row >= bend) continue;
            const int pb = s_rp[li], pe = s_rp[li + 1];
            const int cnt = pe - pb;
            double2 acc = make_double2(0.0, 0.0);
            if (cnt > 0) {
                // the first EPT entries as one batch: all index loads, then all address computations, then all Q loads
                // are independent instructions (a serial loop over the entries chains three shared-memory latencies per entry)
                constexpr int EPT = 8;
                int enc[EPT];
                double v[EPT];
                double2 q[EPT];
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    const int p = min(pb + j, pe - 1);
                    enc[j] = s_rel[p];
                    v[j] = (j < cnt) ? s_val[p] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    if (enc[j] >= 0) q[j] = *reinterpret_cast<const double2*>(ring_row(enc[j], li) + 2 * sub);
                    else q[j] = __ldg(Q2 + (size_t)(-1 - enc[j]) * LPR + sub);
                }
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    acc.x = fma(v[j], q[j].x, acc.x);
                    acc.y = fma(v[j], q[j].y, acc.y);
                }
                for (int p = pb + EPT; p < pe; ++p) {                        // long rows
                    const int e = s_rel[p];
                    const double vv = s_val[p];
                    const double2 qq = (e >= 0) ? *reinterpret_cast<const double2*>(ring_row(e, li) + 2 * sub)
                                                : __ldg(Q2 + (size_t)(-1 - e) * LPR + sub);
                    acc.x = fma(vv, qq.x, acc.x);
                    acc.y = fma(vv, qq.y, acc.y);
                }
            }
            acc.x *= cf.alpha;
            acc.y *= cf.alpha;
            if (cf.beta != 0.0) {
                double2 q;
                if (wt.diag_w >= 0) q = *reinterpret_cast<const double2*>(ring_row((wt.diag_w << 24) | (-wt.lo[wt.diag_w]), li) + 2 * sub);
                else q = __ldg(Q2 + (size_t)row * LPR + sub);
                acc.x = fma(cf.beta, q.x, acc.x);
                acc.y = fma(cf.beta, q.y, acc.y);
            }
            if (cf.gamma != 0.0) {
                const double2 z = reinterpret_cast<const double2*>(Z)[(size_t)row * LPR + sub];
                acc.x = fma(cf.gamma, z.x, acc.x);
                acc.y = fma(cf.gamma, z.y, acc.y);
            }
            reinterpret_cast<double2*>(U)[(size_t)row * LPR + sub] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[slot])) : "memory");
        }
    }
}

void launch_spmm_build_rel(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx, const SpmmWindows& wt, int* rel,
                           unsigned long long* irregular, cudaStream_t st) {
    if (nrows <= 0) return;
    spmm_build_rel_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(nrows, nown, rowptr, colidx, wt, rel, irregular);
}

bool spmm_window_supported(int B) { return B == 16 || B == 32; }

void launch_spmm_window(int B, int64_t nrows, int64_t nown, const int* rowptr, const int* rel, const double* vals, const double* Q,
                        double* U, SpmmCoef cf, const double* Z, const SpmmWindows& wt_in, cudaStream_t st) {
    if (nrows <= 0) return;
    SpmmWindows wt = wt_in;
    size_t smem = 0;
    const int NB = spmm_window_stages(wt, B, &smem);
    if (NB == 0) { std::fprintf(stderr, "rbl: spmm window kernel launched without a fitting layout\n"); std::abort(); }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int R = SpmmWindows::kTileRows;
    int64_t grid = sms;
    int64_t rows_per_cta = ((nrows + grid - 1) / grid + R - 1) / R * R;
    grid = (nrows + rows_per_cta - 1) / rows_per_cta;
    if (B == 16) {
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(spmm_window_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        spmm_window_kernel<16><<<(unsigned)grid, 544, smem, st>>>(nrows, nown, rowptr, rel, vals, Q, U, cf, Z, wt, rows_per_cta, NB, wt.max_tile_nnz);
    } else {
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(spmm_window_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        spmm_window_kernel<32><<<(unsigned)grid, 544, smem, st>>>(nrows, nown, rowptr, rel, vals, Q, U, cf, Z, wt, rows_per_cta, NB, wt.max_tile_nnz);
    }
}

}  // namespace rbl

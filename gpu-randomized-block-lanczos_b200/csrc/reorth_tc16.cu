// K5, second tensor-core version: scaled two-term FP16 split on mma.sync m16n8k16 (fp32 accumulate).
//
//   x * S = h + l,  h = fp16(x S),  l = fp16(x S - h)        (S a power of two, |x| <= 1 for orthonormal columns)
//   x y S_x S_y ~= h_x h_y + h_x l_y + l_x h_y               (22 significant bits, like 3xTF32)
//
// Three k16 MMAs cover 16 reduction steps where the TF32 kernel needs six k8 MMAs, and ncu shows the
// TF32 kernel bound by the tensor pipe (profiles/r01_ncu_full_prof_gram.txt: tensor active 69%, DRAM 46%);
// f16 mma.sync runs at 2x the TF32 rate on B200 (553 vs 277 TFLOP/s measured, rbl_microbench 6 / 4).
//
// Operands that are reused by every CTA (the 2B target columns for the Gram pass, the coefficient
// blocks for the update pass) are converted once per re-orthogonalisation into packed f16x2 words
// whose two halves are the two reduction-dimension neighbours an MMA fragment register needs.
//
// Replaces hybrid_part_reorth! / part_reorth_gpu_async!  (Julia/RBL_gpu.jl:59-81, 29-47).
#include <cuda_fp16.h>

#include <cstdio>

#include "kernels.h"
#include "split16.h"

namespace rbl {

namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// four 8x8 b16 matrices from shared memory; every lane supplies one 16-byte row address (lanes 8i..8i+7: matrix i)
__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* smem_row) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* smem_row) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// Pre-split slab rows (split16.h) staged in shared memory: a row keeps its 4B contiguous bytes (B/4 chunks of
// 16 bytes: hi columns first, then lo); chunk q of row r sits at position q ^ swz(r), which makes the eight row
// addresses of every ldmatrix 8x8 tile fall into eight different 16-byte bank groups.
template <int B>
__device__ __forceinline__ int split_swz(int r) {
    if constexpr (B == 16) return (r >> 1) & 3;
    else return r & 7;
}

// D(16x8) += A(16x16, row) * B(16x8, col), f16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}


// ---- tensor memory (TMEM) as the second accumulation level of the Gram kernel --------------------------------
// mma.sync accumulates in fp32 with truncation, so the error of an accumulator grows linearly with the length of its
// chain (measured, tools/reorth_accuracy.py: 256-row chains 1.4x, 1024-row chains 4.9x the error of an fp32 sgemm).
// Every FLUSH k-steps each warp therefore adds its MMA accumulators, with round-to-nearest FADDs, to running sums
// and restarts the chain from zero.  The register file has no room for a second accumulator set (106 of 128 registers).
// B = 16: the sums live in shared memory (64 KB, one pipeline stage fewer).  B = 32: 128 KB of sums do not fit next to
// the stages, they live in tensor memory (256 KB per SM, otherwise unused by this kernel): a warp reaches lanes
// [32 (warp % 4), +32) of TMEM, the four warps that share a lane quarter use different columns.  (TMEM sums were
// measured for B = 16 too: the LDTM / STTM traffic shares the tensor datapath with the MMAs and cost 10% of the
// kernel's bandwidth; the shared-memory sums cost 3%.)
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, unsigned ncols) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(d), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&r)[8]) {   // issue only: tmem_wait_ld() before use
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

}  // namespace

// ---- targets: row-pair interleaved f16x2 words  Wp[(r/2) * 2B + t] = (W[r][t], W[r+1][t]) ----------------
template <int B>
__global__ void split_targets_h_kernel(int64_t n, const double* __restrict__ w0, const double* __restrict__ w1,
                                       float scale, unsigned* __restrict__ wh, unsigned* __restrict__ wl) {
    const int64_t npairs = (n + 1) / 2;
    const int64_t total = npairs * 2 * B;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const int64_t rp = e / (2 * B);
        const int t = (int)(e % (2 * B));
        const int64_t ra = 2 * rp, rb = 2 * rp + 1;
        const double* src = (t < B) ? w0 : w1;
        const int tc = (t < B) ? t : t - B;
        const float x0 = (float)src[ra * B + tc];
        const float x1 = (rb < n) ? (float)src[rb * B + tc] : 0.f;
        unsigned hi, lo;
        split_h2(x0, x1, scale, hi, lo);
        wh[e] = hi;
        wl[e] = lo;
    }
}

// =================================================================================================
// Gram.  MMA roles: M = 16 Krylov columns of a stored block, N = 8 targets, K = 16 rows.
// B = 16: 16 warps x 2 stored blocks; B = 32: 16 warps x 1 stored block (two m-tiles, n-tiles in two halves).
// =================================================================================================
template <int B>
struct GramH {
    static constexpr int NW = 16;                // warps per CTA (ncu: with 8 the schedulers had 0.6 eligible warps/cycle)
    static constexpr int WB = 32 / B;            // stored blocks per warp
    static constexpr int JT = NW * WB;           // stored blocks per CTA
    static constexpr int MT = B / 16;            // m-tiles per stored block
    static constexpr int NT = (2 * B) / 8;       // n-tiles
    static constexpr int NH = NT / 4;            // n-tiles are processed four at a time (register budget)
    static constexpr int PA = B;                 // unpadded rows, permuted + swizzled (see gram_slot)
    static constexpr int PW = 2 * B + 8;         // words per staged target row-pair
    static constexpr bool SUM_SMEM = (B == 16);  // second-level sums: shared memory (B = 16) or tensor memory (B = 32)
    static constexpr int NST = SUM_SMEM ? 4 : 5;
    static constexpr int RS = 16;                // rows per stage
    static constexpr int RW = 64;                // rows per shared target chunk (32 row pairs)
    static constexpr int STAGE = WB * RS * PA;   // floats
    static constexpr int WBUF = (RW / 2) * PW;   // words per target buffer (hi or lo)
    static constexpr int NCP = (WB * RS * (B / 4)) / 32;  // cp.async per lane per stage
    static constexpr int NACC = WB * MT * NT * 4;   // fp32 accumulators per thread
    static constexpr size_t smem_bytes = (size_t)(NW * NST * STAGE + 4 * WBUF + (SUM_SMEM ? NW * NACC * 32 : 0)) * sizeof(float);
    static constexpr int TCOLS = 4 * NACC;          // TMEM columns: four warps per lane quarter
    static constexpr int FLUSH = 8;                 // k-steps (of RS rows) per tensor-core accumulation chain
};

// Shared-memory slot of element (row r of a 16-row stage, column c) of a stored block, in floats.
// Rows keep their B*4 contiguous bytes (a multiple of 32 B: both 16-byte halves of a global 32-byte sector land
// in one shared-memory sector, so cp.async fetches every sector once - a padded 80-byte pitch was measured to
// fetch 1.47x the sectors from L2).  Bank conflicts of the fragment loads (rows 2t / 2t+1 / 2t+8 / 2t+9,
// columns g + 8i) are removed by storing row r at position p = (r&1)*8 + (r>>1) and XOR-ing the 16-byte chunk
// index with an even number derived from p (even: 32-byte sector pairs stay together).
template <int B>
__device__ __forceinline__ int gram_slot(int r, int c) {
    const int p = ((r & 1) << 3) + (r >> 1);
    if constexpr (B == 16) return p * 16 + ((((c >> 2) ^ (p & 2))) << 2) + (c & 3);
    else return p * 32 + ((((c >> 2) ^ ((p & 3) << 1))) << 2) + (c & 3);
}

template <int B, bool PS>
__global__ void __launch_bounds__(GramH<B>::NW * 32, 1)
    reorth_gram_h_kernel(int64_t n, int64_t m, const float* __restrict__ buf, int64_t bstride,
                         const unsigned* __restrict__ wh, const unsigned* __restrict__ wl, float scale,
                         float inv_scale2, float* __restrict__ partials, int64_t rows_per_range) {
    using C = GramH<B>;
    constexpr int NW = C::NW, WB = C::WB, JT = C::JT, MT = C::MT, NT = C::NT, NH = C::NH, PA = C::PA, PW = C::PW,
                  NST = C::NST, RS = C::RS, RW = C::RW, STAGE = C::STAGE, WBUF = C::WBUF, NCP = C::NCP, NTHR = NW * 32;
    static_assert(B == 16 || B == 32, "instantiated for B = 16, 32");
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    float* sA = smem + (size_t)warp * NST * STAGE;
    unsigned* sW = reinterpret_cast<unsigned*>(smem + (size_t)NW * NST * STAGE);  // [2][hi,lo][RW/2][PW]
    const int64_t jbase = (int64_t)blockIdx.x * JT + (int64_t)warp * WB;
    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_range;  // multiple of RW
    const int64_t rend = min(n, rbeg + rows_per_range);
    const int64_t nrows = rend - rbeg;

    float acc[WB][MT][NT][4];
#pragma unroll
    for (int b = 0; b < WB; ++b)
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int x = 0; x < NT; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[b][a][x][y] = 0.f;

    // second-level sums (see the note above tmem_alloc)
    constexpr int NACC = C::NACC;
    static_assert(NACC % 16 == 0 && (C::TCOLS & (C::TCOLS - 1)) == 0 && C::TCOLS >= 32 && C::TCOLS <= 512, "TMEM slice shape");
    __shared__ unsigned tmem_base_sh;
    unsigned tsum = 0;
    float4* ssum = nullptr;
    if constexpr (C::SUM_SMEM) {
        // [warp][NACC/4][lane] float4: consecutive lanes read consecutive 16-byte words (conflict-free)
        ssum = reinterpret_cast<float4*>(smem + (size_t)NW * NST * STAGE + 4 * WBUF) + (size_t)warp * (NACC / 4) * 32 + lane;
#pragma unroll
        for (int q = 0; q < NACC / 4; ++q) ssum[q * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        if (warp == 0) tmem_alloc(&tmem_base_sh, C::TCOLS);
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
        tsum = tmem_base_sh + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * NACC;
        const unsigned z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int c0 = 0; c0 < NACC; c0 += 8) tmem_st8(tsum + c0, z);
        tmem_wait_st();
    }
    // sums += acc (round to nearest); acc = 0, or (last) acc = sums
    auto flush = [&](bool last) {
        float* af = &acc[0][0][0][0];
        if constexpr (C::SUM_SMEM) {
#pragma unroll
            for (int q = 0; q < NACC / 4; ++q) {
                float4 v = ssum[q * 32];
                v.x += af[4 * q]; v.y += af[4 * q + 1]; v.z += af[4 * q + 2]; v.w += af[4 * q + 3];
                if (last) {
                    af[4 * q] = v.x; af[4 * q + 1] = v.y; af[4 * q + 2] = v.z; af[4 * q + 3] = v.w;
                } else {
                    ssum[q * 32] = v;
                    af[4 * q] = af[4 * q + 1] = af[4 * q + 2] = af[4 * q + 3] = 0.f;
                }
            }
        } else {
            tmem_wait_st();   // the stores of this warp's previous flush (issued FLUSH k-steps ago)
#pragma unroll
            for (int c0 = 0; c0 < NACC; c0 += 16) {
                unsigned sa[8], sb[8];
                tmem_ld8(tsum + c0, sa);
                tmem_ld8(tsum + c0 + 8, sb);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    sa[e] = __float_as_uint(__uint_as_float(sa[e]) + af[c0 + e]);
                    sb[e] = __float_as_uint(__uint_as_float(sb[e]) + af[c0 + 8 + e]);
                }
                if (last) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        af[c0 + e] = __uint_as_float(sa[e]);
                        af[c0 + 8 + e] = __uint_as_float(sb[e]);
                    }
                } else {
                    tmem_st8(tsum + c0, sa);
                    tmem_st8(tsum + c0 + 8, sb);
#pragma unroll
                    for (int e = 0; e < 16; ++e) af[c0 + e] = 0.f;
                }
            }
        }
    };

    if (nrows > 0) {
        const int nks = (int)((nrows + RS - 1) / RS);
        constexpr int KPC = RW / RS;  // k-steps per target chunk
        // per-lane constants of the NCP copies of a stage: source pointer at stage 0, row inside the stage,
        // destination slot; a stage advances every source by RS*B floats
        const float* src0[NCP];
        int dst0[NCP], row0[NCP];
        bool blk_ok[NCP];
#pragma unroll
        for (int u = 0; u < NCP; ++u) {
            const int q = lane + 32 * u;
            const int blk = q / (RS * (B / 4));
            const int rem = q % (RS * (B / 4));
            const int row = rem / (B / 4), c4 = rem % (B / 4);
            blk_ok[u] = (jbase + blk) < m;
            row0[u] = row;
            dst0[u] = PS ? blk * RS * PA + row * B + ((c4 ^ split_swz<B>(row)) << 2) : blk * RS * PA + gram_slot<B>(row, c4 * 4);
            src0[u] = buf + (size_t)(blk_ok[u] ? jbase + blk : 0) * bstride + (size_t)(rbeg + row) * B + c4 * 4;
        }
        auto issue_a = [&](int ks) {
            float* st = sA + (size_t)(ks % NST) * STAGE;
            const int64_t rem_rows = nrows - (int64_t)ks * RS;  // rows of this range not yet consumed
            const size_t adv = (size_t)ks * RS * B;
#pragma unroll
            for (int u = 0; u < NCP; ++u) {
                const bool ok = blk_ok[u] && (row0[u] < rem_rows);
                cp_async16(st + dst0[u], ok ? src0[u] + adv : buf, ok ? 16 : 0);
            }
        };
        auto issue_w = [&](int chunk) {  // whole CTA: RW/2 row pairs x 2B words, hi and lo
            constexpr int PER = (RW / 2) * (2 * B / 4);  // 16-byte copies per array
            const int64_t rp0 = (rbeg + (int64_t)chunk * RW) / 2;
            const int64_t rp_end = (rend + 1) / 2;
#pragma unroll
            for (int u = 0; u < (2 * PER) / NTHR; ++u) {
                const int idx = tid + NTHR * u;
                const int half = idx / PER, q = idx % PER;
                unsigned* dst = sW + (size_t)((chunk & 1) * 2 + half) * WBUF;
                const unsigned* srcb = half ? wl : wh;
                const int rp = q / (2 * B / 4), c4 = q % (2 * B / 4);
                const bool ok = (rp0 + rp < rp_end);
                cp_async16(dst + rp * PW + c4 * 4, ok ? srcb + (size_t)(rp0 + rp) * 2 * B + c4 * 4 : srcb, ok ? 16 : 0);
            }
        };
        static_assert((2 * (RW / 2) * (2 * B / 4)) % NTHR == 0, "target chunk copy must tile the CTA");
        issue_w(0);
        issue_a(0);
        cp_async_commit();
#pragma unroll
        for (int s = 1; s < NST - 1; ++s) {
            issue_a(s);
            cp_async_commit();
        }
        for (int ks = 0; ks < nks; ++ks) {
            cp_async_wait<NST - 2>();
            if ((ks % KPC) == 0) __syncthreads();
            else __syncwarp();
            if ((ks % KPC) == 0 && (int64_t)(ks / KPC + 1) * RW < nrows) issue_w(ks / KPC + 1);
            issue_a(ks + NST - 1);
            cp_async_commit();

            const float* st = sA + (size_t)(ks % NST) * STAGE;
            const int chunk = ks / KPC;
            const unsigned* ph = sW + (size_t)((chunk & 1) * 2 + 0) * WBUF + (size_t)(ks % KPC) * 8 * PW;
            const unsigned* pl = sW + (size_t)((chunk & 1) * 2 + 1) * WBUF + (size_t)(ks % KPC) * 8 * PW;
            unsigned bh[4][2], bl[4][2];
            auto load_b = [&](int nh) {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int col = (nh * 4 + x) * 8 + g;
                    bh[x][0] = ph[t * PW + col];
                    bh[x][1] = ph[(t + 4) * PW + col];
                    bl[x][0] = pl[t * PW + col];
                    bl[x][1] = pl[(t + 4) * PW + col];
                }
            };
            if constexpr (NH == 1) load_b(0);
#pragma unroll
            for (int b = 0; b < WB; ++b) {
                const float* a = st + b * RS * PA;
                unsigned ah[MT][4], al[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    if constexpr (PS) {
                        // the slab already holds (hi | lo) f16 rows: four transposed 8x8 tiles per m-tile and part.
                        // lanes 8i..8i+7 address tile i: rows (i>>1)*8.., column chunk 2mt + (i&1)
                        const int row = (lane & 7) + ((lane >> 4) << 3);
                        const int q = 2 * mt + ((lane >> 3) & 1);
                        const float* rowp = a + row * B;
                        ldmatrix_x4_trans(ah[mt], rowp + ((q ^ split_swz<B>(row)) << 2));
                        ldmatrix_x4_trans(al[mt], rowp + (((q + B / 8) ^ split_swz<B>(row)) << 2));
                    } else {
                        const int c0 = mt * 16 + g;
                        // A[m = column][k = row]: register halves are rows 2t, 2t+1 (and +8) of columns c0 / c0+8
                        split_h2(a[gram_slot<B>(2 * t, c0)], a[gram_slot<B>(2 * t + 1, c0)], scale, ah[mt][0], al[mt][0]);
                        split_h2(a[gram_slot<B>(2 * t, c0 + 8)], a[gram_slot<B>(2 * t + 1, c0 + 8)], scale, ah[mt][1], al[mt][1]);
                        split_h2(a[gram_slot<B>(2 * t + 8, c0)], a[gram_slot<B>(2 * t + 9, c0)], scale, ah[mt][2], al[mt][2]);
                        split_h2(a[gram_slot<B>(2 * t + 8, c0 + 8)], a[gram_slot<B>(2 * t + 9, c0 + 8)], scale, ah[mt][3], al[mt][3]);
                    }
                }
#pragma unroll
                for (int nh = 0; nh < NH; ++nh) {
                    if constexpr (NH > 1) load_b(nh);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) mma_f16(acc[b][mt][nh * 4 + x], al[mt], bh[x]);
#pragma unroll
                        for (int x = 0; x < 4; ++x) mma_f16(acc[b][mt][nh * 4 + x], ah[mt], bl[x]);
#pragma unroll
                        for (int x = 0; x < 4; ++x) mma_f16(acc[b][mt][nh * 4 + x], ah[mt], bh[x]);
                    }
                }
            }
            // the warps flush in turn (two per k-step): the TMEM round trips of one warp hide behind the MMAs of the others
            if ((ks + 1 + warp) % C::FLUSH == 0 && ks + 1 < nks) flush(false);
        }
        cp_async_wait<0>();
        flush(true);
    }
    if constexpr (!C::SUM_SMEM) {
        tcgen05_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem_base_sh, C::TCOLS);
    }
#pragma unroll
    for (int b = 0; b < WB; ++b) {
        const int64_t j = jbase + b;
        if (j >= m) continue;
        float* out = partials + ((size_t)blockIdx.y * m * B + (size_t)j * B) * (2 * B);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                *reinterpret_cast<float2*>(out + (size_t)(mt * 16 + g) * 2 * B + x * 8 + 2 * t) =
                    make_float2(acc[b][mt][x][0] * inv_scale2, acc[b][mt][x][1] * inv_scale2);
                *reinterpret_cast<float2*>(out + (size_t)(mt * 16 + g + 8) * 2 * B + x * 8 + 2 * t) =
                    make_float2(acc[b][mt][x][2] * inv_scale2, acc[b][mt][x][3] * inv_scale2);
            }
    }
}

// sum over row ranges (double); also the running max |C| (bit pattern of a non-negative float orders like uint)
__global__ void reorth_reduce_max_kernel(const float* __restrict__ partials, int ranges, size_t count,
                                         float* __restrict__ Cout, unsigned* __restrict__ cmax_bits) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float c = 0.f;
    if (e < count) {
        double s = 0.0;
        for (int p = 0; p < ranges; ++p) s += (double)partials[(size_t)p * count + e];
        c = (float)s;
        Cout[e] = c;
    }
    float mx = fabsf(c);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(cmax_bits, __float_as_uint(mx));
}

// coefficients: column-pair interleaved f16x2 words  Cp[(j*B/2 + c/2) * 2B + t] = (C_j[c][t], C_j[c+1][t]),
// scaled by the power of two that brings max|C| into [1024, 2048)
__global__ void split_coeff_h_kernel(size_t nwords, int B, const float* __restrict__ Cin,
                                     const unsigned* __restrict__ cmax_bits, unsigned* __restrict__ ch,
                                     unsigned* __restrict__ cl, float* __restrict__ scale_out) {
    const float cmax = __uint_as_float(*cmax_bits);
    float scale = 1.f;
    if (cmax > 0.f && isfinite(cmax)) {
        int ex;
        frexpf(cmax, &ex);           // cmax = f * 2^ex, f in [0.5, 1)
        scale = ldexpf(1.f, 11 - ex);  // cmax * scale in [1024, 2048)
    }
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) *scale_out = scale;
    if (e >= nwords) return;
    const size_t cp = e / (2 * B);
    const int t = (int)(e % (2 * B));
    const float x0 = Cin[(cp * 2) * (2 * B) + t];
    const float x1 = Cin[(cp * 2 + 1) * (2 * B) + t];
    unsigned hi, lo;
    split_h2(x0, x1, scale, hi, lo);
    ch[e] = hi;
    cl[e] = lo;
}

// =================================================================================================
// Update.  MMA roles: M = 16 rows, N = 8 targets, K = 16 Krylov columns (B/16 k-steps per stored block).
// Every warp owns 32 rows; B = 16: 16 warps, B = 32: 8 warps (shared memory).
// =================================================================================================
template <int B, bool PS>
struct UpdH {
    static constexpr int NW = (B == 16) ? 16 : 8;  // warps per CTA
    static constexpr int MT = 2;
    static constexpr int NT = (2 * B) / 8;
    static constexpr int NH = NT / 4;
    static constexpr int KS = B / 16;
    // words per staged row.  fp32 slab: B + 8 (conflict-free float2 A fragments, 32-byte multiple);
    // pre-split slab: B, swizzled for ldmatrix (split_swz)
    static constexpr int PA = PS ? B : B + 8;
    static constexpr int PC = 2 * B + 8;         // words per staged coefficient row pair
    static constexpr int JC = (B == 16) ? 8 : 4;
    static constexpr int NST = PS ? 4 : 3;
    static constexpr int STAGE = 32 * PA;
    static constexpr int CBUF = JC * (B / 2) * PC;
    static constexpr int ROWS_CTA = NW * 32;
    static constexpr int NCP = (32 * (B / 4)) / 32;
    static constexpr size_t smem_bytes = (size_t)(NW * NST * STAGE + 4 * CBUF) * sizeof(float);
};

template <int B, bool PS>
__global__ void __launch_bounds__(UpdH<B, PS>::NW * 32, 1)
    reorth_update_h_kernel(int64_t n, int64_t m, const float* __restrict__ buf, int64_t bstride,
                           const unsigned* __restrict__ Ch, const unsigned* __restrict__ Cl, float scale_a,
                           const float* __restrict__ scale_c_ptr, double* __restrict__ w0, double* __restrict__ w1,
                           float* __restrict__ store_w1, float* __restrict__ store_w0) {
    using C = UpdH<B, PS>;
    constexpr int NW = C::NW, MT = C::MT, NT = C::NT, NH = C::NH, KS = C::KS, PA = C::PA, PC = C::PC, JC = C::JC,
                  NST = C::NST, STAGE = C::STAGE, CBUF = C::CBUF, NCP = C::NCP, NTHR = NW * 32;
    static_assert(B == 16 || B == 32, "instantiated for B = 16, 32");
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    float* sA = smem + (size_t)warp * NST * STAGE;
    unsigned* sC = reinterpret_cast<unsigned*>(smem + (size_t)NW * NST * STAGE);  // [2][hi,lo][JC][B/2][PC]
    const int64_t r0 = (int64_t)blockIdx.x * C::ROWS_CTA + (int64_t)warp * 32;
    const int mi = (int)m;

    float acc[MT][NT][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int x = 0; x < NT; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[a][x][y] = 0.f;

    // per-lane constants of the copies of a stage (32 rows x B floats of one stored block)
    const float* src0[NCP];
    int dst0[NCP];
    bool row_ok[NCP];
#pragma unroll
    for (int u = 0; u < NCP; ++u) {
        const int q = lane + 32 * u;
        const int row = q / (B / 4), c4 = q % (B / 4);
        row_ok[u] = (r0 + row) < n;
        dst0[u] = PS ? row * PA + ((c4 ^ split_swz<B>(row)) << 2) : row * PA + c4 * 4;
        src0[u] = buf + (size_t)(row_ok[u] ? r0 + row : 0) * B + c4 * 4;
    }
    auto issue_a = [&](int j) {
        float* st = sA + (size_t)(j % NST) * STAGE;
        const size_t adv = (size_t)j * bstride;
#pragma unroll
        for (int u = 0; u < NCP; ++u) {
            const bool ok = row_ok[u] && (j < mi);
            cp_async16(st + dst0[u], ok ? src0[u] + adv : buf, ok ? 16 : 0);
        }
    };
    auto issue_c = [&](int chunk) {  // JC blocks x B/2 column pairs x 2B words, hi and lo
        constexpr int PER = JC * (B / 2) * (2 * B / 4);  // 16-byte copies per array
        const int j0 = chunk * JC;
#pragma unroll
        for (int u = 0; u < (2 * PER) / NTHR; ++u) {
            const int idx = tid + NTHR * u;
            const int half = idx / PER, q = idx % PER;
            unsigned* dst = sC + (size_t)((chunk & 1) * 2 + half) * CBUF;
            const unsigned* srcb = half ? Cl : Ch;
            const int rowc = q / (2 * B / 4), c4 = q % (2 * B / 4);  // rowc = jb*(B/2) + c/2
            const bool ok = (j0 * (B / 2) + rowc) < mi * (B / 2);
            const size_t off = ((size_t)j0 * (B / 2) + rowc) * (2 * B) + c4 * 4;
            cp_async16(dst + rowc * PC + c4 * 4, ok ? srcb + off : srcb, ok ? 16 : 0);
        }
    };
    static_assert((2 * JC * (B / 2) * (2 * B / 4)) % NTHR == 0, "coefficient chunk copy must tile the CTA");
    issue_c(0);
    issue_a(0);
    cp_async_commit();
#pragma unroll
    for (int s = 1; s < NST - 1; ++s) {
        issue_a(s);
        cp_async_commit();
    }
    for (int j = 0; j < mi; ++j) {
        cp_async_wait<NST - 2>();
        if ((j % JC) == 0) __syncthreads();
        else __syncwarp();
        if ((j % JC) == 0 && (j / JC + 1) * JC < mi) issue_c(j / JC + 1);
        issue_a(j + NST - 1);
        cp_async_commit();

        const float* st = sA + (size_t)(j % NST) * STAGE;
        const int chunk = j / JC;
        const unsigned* ph = sC + (size_t)((chunk & 1) * 2 + 0) * CBUF + (size_t)(j % JC) * (B / 2) * PC;
        const unsigned* pl = sC + (size_t)((chunk & 1) * 2 + 1) * CBUF + (size_t)(j % JC) * (B / 2) * PC;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            unsigned ah[MT][4], al[MT][4];
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                if constexpr (PS) {
                    // lanes 8i..8i+7 address tile i: rows 16a + (i&1)*8.., column chunk 2ks + (i>>1)
                    const int row = a * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    const int q = 2 * ks + (lane >> 4);
                    const float* rowp = st + row * PA;
                    ldmatrix_x4(ah[a], rowp + ((q ^ split_swz<B>(row)) << 2));
                    ldmatrix_x4(al[a], rowp + (((q + B / 8) ^ split_swz<B>(row)) << 2));
                } else {
                    const float* ap = st + (a * 16) * PA + ks * 16;
                    // A[m = row][k = column]: register halves are columns 2t, 2t+1 (and +8) of rows g / g+8
                    const float2 v0 = *reinterpret_cast<const float2*>(ap + g * PA + 2 * t);
                    const float2 v1 = *reinterpret_cast<const float2*>(ap + (g + 8) * PA + 2 * t);
                    const float2 v2 = *reinterpret_cast<const float2*>(ap + g * PA + 2 * t + 8);
                    const float2 v3 = *reinterpret_cast<const float2*>(ap + (g + 8) * PA + 2 * t + 8);
                    split_h2(v0.x, v0.y, scale_a, ah[a][0], al[a][0]);
                    split_h2(v1.x, v1.y, scale_a, ah[a][1], al[a][1]);
                    split_h2(v2.x, v2.y, scale_a, ah[a][2], al[a][2]);
                    split_h2(v3.x, v3.y, scale_a, ah[a][3], al[a][3]);
                }
            }
#pragma unroll
            for (int nh = 0; nh < NH; ++nh) {
                unsigned bh[4][2], bl[4][2];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int col = (nh * 4 + x) * 8 + g;
                    bh[x][0] = ph[(ks * 8 + t) * PC + col];
                    bh[x][1] = ph[(ks * 8 + t + 4) * PC + col];
                    bl[x][0] = pl[(ks * 8 + t) * PC + col];
                    bl[x][1] = pl[(ks * 8 + t + 4) * PC + col];
                }
#pragma unroll
                for (int a = 0; a < MT; ++a) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) mma_f16(acc[a][nh * 4 + x], al[a], bh[x]);
#pragma unroll
                    for (int x = 0; x < 4; ++x) mma_f16(acc[a][nh * 4 + x], ah[a], bl[x]);
#pragma unroll
                    for (int x = 0; x < 4; ++x) mma_f16(acc[a][nh * 4 + x], ah[a], bh[x]);
                }
            }
        }
    }
    cp_async_wait<0>();
    const float inv = 1.0f / (scale_a * __ldg(scale_c_ptr));
#pragma unroll
    for (int a = 0; a < MT; ++a) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t row = r0 + a * 16 + g + 8 * h;
            if (row >= n) continue;
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                const int tgt = x * 8 + 2 * t;
                const float d0 = acc[a][x][2 * h] * inv, d1 = acc[a][x][2 * h + 1] * inv;
                if (tgt < B) {
                    double2* p = reinterpret_cast<double2*>(w0 + (size_t)row * B + tgt);
                    double2 v = *p;
                    v.x -= (double)d0;
                    v.y -= (double)d1;
                    *p = v;
                    if (store_w0 != nullptr) {
                        if constexpr (PS) {
                            unsigned hi, lo;
                            split_h2((float)v.x, (float)v.y, scale_a, hi, lo);
                            unsigned* srow = reinterpret_cast<unsigned*>(store_w0) + (size_t)row * B;
                            srow[tgt >> 1] = hi;
                            srow[B / 2 + (tgt >> 1)] = lo;
                        } else {
                            *reinterpret_cast<float2*>(store_w0 + (size_t)row * B + tgt) = make_float2((float)v.x, (float)v.y);
                        }
                    }
                } else {
                    double2* p = reinterpret_cast<double2*>(w1 + (size_t)row * B + (tgt - B));
                    double2 v = *p;
                    v.x -= (double)d0;
                    v.y -= (double)d1;
                    *p = v;
                    if (store_w1 != nullptr) {
                        if constexpr (PS) {
                            unsigned hi, lo;
                            split_h2((float)v.x, (float)v.y, scale_a, hi, lo);
                            unsigned* srow = reinterpret_cast<unsigned*>(store_w1) + (size_t)row * B;
                            srow[(tgt - B) >> 1] = hi;
                            srow[B / 2 + ((tgt - B) >> 1)] = lo;
                        } else {
                            *reinterpret_cast<float2*>(store_w1 + (size_t)row * B + (tgt - B)) = make_float2((float)v.x, (float)v.y);
                        }
                    }
                }
            }
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------------
// scratch layout (32-bit words): Wh, Wl: ((n+1)/2) * 2B each | Ch, Cl: m_cap*(B/2) * 2B each | cmax bits | scale_c
static float pick_scale(int64_t n_global) {
    // orthonormal columns: |x| <= 1, rms 1/sqrt(n); bring the rms to ~8 but never above 2^15 (fp16 max 65504)
    double s = 8.0 * std::sqrt((double)std::max<int64_t>(n_global, 1));
    int e = (int)std::floor(std::log2(s));
    if (e > 15) e = 15;
    if (e < 0) e = 0;
    return std::ldexp(1.0f, e);
}

size_t reorth_h_scratch_words(int B, int64_t n, int64_t m_cap) {
    return 2 * (size_t)((n + 1) / 2) * 2 * B + 2 * (size_t)m_cap * (B / 2) * 2 * B + 64;
}

struct HScratch {
    unsigned *wh, *wl, *ch, *cl, *cmax;
    float* scale_c;
};
static HScratch h_layout(float* scratch, int B, int64_t n, int64_t m_cap) {
    HScratch s;
    unsigned* p = reinterpret_cast<unsigned*>(scratch);
    const size_t wn = (size_t)((n + 1) / 2) * 2 * B, cn = (size_t)m_cap * (B / 2) * 2 * B;
    s.wh = p;
    s.wl = s.wh + wn;
    s.ch = s.wl + wn;
    s.cl = s.ch + cn;
    s.cmax = s.cl + cn;
    s.scale_c = reinterpret_cast<float*>(s.cmax + 16);
    return s;
}

bool reorth_h_supported(int B, int fp32) { return fp32 && (B == 16 || B == 32); }

template <int B, bool PS>
static void gram_h_launch(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t bstride, const double* w0,
                          const double* w1, void* partials, void* Cmat, float* scratch, int64_t m_cap, cudaStream_t st) {
    using G = GramH<B>;
    HScratch s = h_layout(scratch, B, p.n, m_cap);
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_gram_h_kernel<B, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const float scale = pick_scale(n_global);
    const int64_t total = ((p.n + 1) / 2) * 2 * B;
    const int sgrid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
    split_targets_h_kernel<B><<<sgrid, 256, 0, st>>>(p.n, w0, w1, scale, s.wh, s.wl);
    cudaMemsetAsync(s.cmax, 0, 4, st);
    int64_t rpr = (p.n + p.ranges - 1) / p.ranges;
    rpr = (rpr + G::RW - 1) / G::RW * G::RW;
    dim3 grid((unsigned)((p.m + G::JT - 1) / G::JT), p.ranges);
    reorth_gram_h_kernel<B, PS><<<grid, G::NW * 32, G::smem_bytes, st>>>(p.n, p.m, (const float*)buf, bstride, s.wh, s.wl, scale,
                                                               1.0f / (scale * scale), (float*)partials, rpr);
    const size_t count = (size_t)p.m * B * 2 * B;
    reorth_reduce_max_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const float*)partials, p.ranges, count,
                                                                               (float*)Cmat, s.cmax);
}

void launch_reorth_gram_h(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t bstride, const double* w0,
                          const double* w1, void* partials, void* Cmat, float* scratch, int64_t m_cap, int presplit,
                          cudaStream_t st) {
    if (p.B == 16) {
        if (presplit) gram_h_launch<16, true>(p, n_global, buf, bstride, w0, w1, partials, Cmat, scratch, m_cap, st);
        else gram_h_launch<16, false>(p, n_global, buf, bstride, w0, w1, partials, Cmat, scratch, m_cap, st);
    } else {
        if (presplit) gram_h_launch<32, true>(p, n_global, buf, bstride, w0, w1, partials, Cmat, scratch, m_cap, st);
        else gram_h_launch<32, false>(p, n_global, buf, bstride, w0, w1, partials, Cmat, scratch, m_cap, st);
    }
}

float reorth_h_scale(int64_t n_global) { return pick_scale(n_global); }

// (re)build the packed coefficient words from C - after the all-reduce of C in a row-sharded run the local
// maxima differ, so `recompute_max` rescans C first
__global__ void coeff_max_kernel(size_t count, const float* __restrict__ Cin, unsigned* __restrict__ cmax_bits) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mx = (e < count) ? fabsf(Cin[e]) : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(cmax_bits, __float_as_uint(mx));
}

void launch_reorth_coeff_h(const ReorthPlan& p, const void* Cmat, float* scratch, int64_t m_cap, int recompute_max,
                           cudaStream_t st) {
    const int B = p.B;
    HScratch s = h_layout(scratch, B, p.n, m_cap);
    const size_t count = (size_t)p.m * B * 2 * B;
    if (recompute_max) {
        cudaMemsetAsync(s.cmax, 0, 4, st);
        coeff_max_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, (const float*)Cmat, s.cmax);
    }
    const size_t nwords = count / 2;
    split_coeff_h_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(nwords, B, (const float*)Cmat, s.cmax, s.ch, s.cl,
                                                                            s.scale_c);
}

template <int B, bool PS>
static void update_h_launch(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t bstride, double* w0,
                            double* w1, void* store_w1, void* store_w0, float* scratch, int64_t m_cap, cudaStream_t st,
                            int64_t coef_block0) {
    using U = UpdH<B, PS>;
    HScratch s = h_layout(scratch, B, p.n, m_cap);
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_update_h_kernel<B, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::smem_bytes);
    const unsigned grid = (unsigned)((p.n + U::ROWS_CTA - 1) / U::ROWS_CTA);
    const size_t coff = (size_t)coef_block0 * (B / 2) * 2 * B;
    reorth_update_h_kernel<B, PS><<<grid, U::NW * 32, U::smem_bytes, st>>>(p.n, p.m, (const float*)buf, bstride, s.ch + coff, s.cl + coff,
                                                                 pick_scale(n_global), s.scale_c, w0, w1, (float*)store_w1, (float*)store_w0);
}

void launch_reorth_update_h(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t bstride, double* w0,
                            double* w1, void* store_w1, float* scratch, int64_t m_cap, int presplit, cudaStream_t st,
                            void* store_w0, int64_t coef_block0) {
    if (p.B == 16) {
        if (presplit) update_h_launch<16, true>(p, n_global, buf, bstride, w0, w1, store_w1, store_w0, scratch, m_cap, st, coef_block0);
        else update_h_launch<16, false>(p, n_global, buf, bstride, w0, w1, store_w1, store_w0, scratch, m_cap, st, coef_block0);
    } else {
        if (presplit) update_h_launch<32, true>(p, n_global, buf, bstride, w0, w1, store_w1, store_w0, scratch, m_cap, st, coef_block0);
        else update_h_launch<32, false>(p, n_global, buf, bstride, w0, w1, store_w1, store_w0, scratch, m_cap, st, coef_block0);
    }
}

// =================================================================================================
// K6 on the split16 slab: V = Q * S with the same scaled two-term FP16 products.  MMA roles as in the update
// (M = 16 rows, N = 8 Ritz columns, K = 16 Krylov columns).  Every warp owns 16 rows and NT*8 columns; grid.y
// walks column groups.  V is an O(1) quantity (unlike the reorth corrections), so accumulation is two-level:
// every k-step's three products start from a zero accumulator and are added to an fp32 sum with round-to-nearest
// (the tensor core's own accumulation truncates), and the fp32 sums are flushed into fp64 every JC stored blocks.
// =================================================================================================
template <int B, int NT>
struct RitzH {
    static constexpr int NW = 8;
    static constexpr int KS = B / 16;
    static constexpr int PA = B;
    static constexpr int NCOL = NT * 8;
    static constexpr int PC = NCOL + ((NCOL % 16 == 8) ? 16 : 8);  // == 8 or 24 mod 32: conflict-free B-fragment loads
    static constexpr int JC = (B == 16) ? 8 : 4;
    static constexpr int NST = 4;
    static constexpr int STAGE = 16 * PA;
    static constexpr int CBUF = JC * (B / 2) * PC;
    static constexpr int ROWS_CTA = NW * 16;
    static constexpr int NCP = (16 * (B / 4)) / 32;
    static constexpr size_t smem_bytes = (size_t)(NW * NST * STAGE + 4 * CBUF) * sizeof(float);
};

// S (row-major (m*B) x kpad floats) -> column-pair interleaved f16x2 words  Sp[(r/2) * ncols + t] = (S[r][t], S[r+1][t])
__global__ void split_ritz_coeff_kernel(size_t nwords, int kpad, int ncols, const float* __restrict__ S, float scale,
                                        unsigned* __restrict__ sh, unsigned* __restrict__ sl) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nwords) return;
    const size_t pr = e / ncols;
    const int t = (int)(e % ncols);
    const float x0 = (t < kpad) ? S[(pr * 2) * kpad + t] : 0.f;
    const float x1 = (t < kpad) ? S[(pr * 2 + 1) * kpad + t] : 0.f;
    unsigned hi, lo;
    split_h2(x0, x1, scale, hi, lo);
    sh[e] = hi;
    sl[e] = lo;
}

template <int B, int NT, typename VT>
__global__ void __launch_bounds__(RitzH<B, NT>::NW * 32)
    ritz_h_kernel(int64_t n, int64_t m, const float* __restrict__ buf, int64_t bstride, const unsigned* __restrict__ Sh,
                  const unsigned* __restrict__ Sl, int ncols, float inv_scale, VT* __restrict__ V, int64_t ldv, int k,
                  int accumulate) {
    using C = RitzH<B, NT>;
    constexpr int NW = C::NW, KS = C::KS, PA = C::PA, NCOL = C::NCOL, PC = C::PC, JC = C::JC, NST = C::NST,
                  STAGE = C::STAGE, CBUF = C::CBUF, NCP = C::NCP, NTHR = NW * 32;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    float* sA = smem + (size_t)warp * NST * STAGE;
    unsigned* sC = reinterpret_cast<unsigned*>(smem + (size_t)NW * NST * STAGE);  // [2][hi,lo][JC][B/2][PC]
    const int64_t r0 = (int64_t)blockIdx.x * C::ROWS_CTA + (int64_t)warp * 16;
    const int col0 = blockIdx.y * NCOL;
    const int mi = (int)m;

    float acc[NT][4];
    double dacc[NT][4];
#pragma unroll
    for (int x = 0; x < NT; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            acc[x][y] = 0.f;
            dacc[x][y] = 0.0;
        }

    const float* src0[NCP];
    int dst0[NCP];
    bool row_ok[NCP];
#pragma unroll
    for (int u = 0; u < NCP; ++u) {
        const int q = lane + 32 * u;
        const int row = q / (B / 4), c4 = q % (B / 4);
        row_ok[u] = (r0 + row) < n;
        dst0[u] = row * PA + ((c4 ^ split_swz<B>(row)) << 2);
        src0[u] = buf + (size_t)(row_ok[u] ? r0 + row : 0) * B + c4 * 4;
    }
    auto issue_a = [&](int j) {
        float* st = sA + (size_t)(j % NST) * STAGE;
        const size_t adv = (size_t)j * bstride;
#pragma unroll
        for (int u = 0; u < NCP; ++u) {
            const bool ok = row_ok[u] && (j < mi);
            cp_async16(st + dst0[u], ok ? src0[u] + adv : buf, ok ? 16 : 0);
        }
    };
    auto issue_c = [&](int chunk) {  // JC blocks x B/2 column pairs x NCOL words, hi and lo
        constexpr int PER = JC * (B / 2) * (NCOL / 4);
        static_assert((2 * PER) % NTHR == 0, "coefficient chunk copy must tile the CTA");
        const int j0 = chunk * JC;
#pragma unroll
        for (int u = 0; u < (2 * PER) / NTHR; ++u) {
            const int idx = tid + NTHR * u;
            const int half = idx / PER, q = idx % PER;
            unsigned* dst = sC + (size_t)((chunk & 1) * 2 + half) * CBUF;
            const unsigned* srcb = half ? Sl : Sh;
            const int rowc = q / (NCOL / 4), c4 = q % (NCOL / 4);
            const bool ok = (j0 * (B / 2) + rowc) < mi * (B / 2);
            const size_t off = ((size_t)j0 * (B / 2) + rowc) * ncols + col0 + c4 * 4;
            cp_async16(dst + rowc * PC + c4 * 4, ok ? srcb + off : srcb, ok ? 16 : 0);
        }
    };
    issue_c(0);
    issue_a(0);
    cp_async_commit();
#pragma unroll
    for (int s = 1; s < NST - 1; ++s) {
        issue_a(s);
        cp_async_commit();
    }
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < mi; ++j) {
        cp_async_wait<NST - 2>();
        if ((j % JC) == 0) {
            __syncthreads();
#pragma unroll
            for (int x = 0; x < NT; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    dacc[x][y] += (double)acc[x][y];
                    acc[x][y] = 0.f;
                }
        } else {
            __syncwarp();
        }
        if ((j % JC) == 0 && (j / JC + 1) * JC < mi) issue_c(j / JC + 1);
        issue_a(j + NST - 1);
        cp_async_commit();

        const float* st = sA + (size_t)(j % NST) * STAGE;
        const int chunk = j / JC;
        const unsigned* ph = sC + (size_t)((chunk & 1) * 2 + 0) * CBUF + (size_t)(j % JC) * (B / 2) * PC;
        const unsigned* pl = sC + (size_t)((chunk & 1) * 2 + 1) * CBUF + (size_t)(j % JC) * (B / 2) * PC;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            unsigned ah[4], al[4];
            {
                const int row = (lane & 7) + (((lane >> 3) & 1) << 3);
                const int q = 2 * ks + (lane >> 4);
                const float* rowp = st + row * PA;
                ldmatrix_x4(ah, rowp + ((q ^ split_swz<B>(row)) << 2));
                ldmatrix_x4(al, rowp + (((q + B / 8) ^ split_swz<B>(row)) << 2));
            }
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                unsigned bh[2], bl[2];
                const int col = x * 8 + g;
                bh[0] = ph[(ks * 8 + t) * PC + col];
                bh[1] = ph[(ks * 8 + t + 4) * PC + col];
                bl[0] = pl[(ks * 8 + t) * PC + col];
                bl[1] = pl[(ks * 8 + t + 4) * PC + col];
                float d[4] = {zero4[0], zero4[1], zero4[2], zero4[3]};
                mma_f16(d, al, bh);
                mma_f16(d, ah, bl);
                mma_f16(d, ah, bh);
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += d[y];
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t row = r0 + g + 8 * h;
        if (row >= n) continue;
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            const int col = col0 + x * 8 + 2 * t;
            const double v0 = (dacc[x][2 * h] + (double)acc[x][2 * h]) * (double)inv_scale;
            const double v1 = (dacc[x][2 * h + 1] + (double)acc[x][2 * h + 1]) * (double)inv_scale;
            if (col < k) {
                VT* d0 = V + (size_t)col * ldv + row;
                *d0 = accumulate ? (VT)((double)*d0 + v0) : (VT)v0;
            }
            if (col + 1 < k) {
                VT* d1 = V + (size_t)(col + 1) * ldv + row;
                *d1 = accumulate ? (VT)((double)*d1 + v1) : (VT)v1;
            }
        }
    }
}

template <int B, int NT, typename VT>
static void ritz_h_launch_t(int64_t n, int64_t m, int k, const void* buf, int64_t bstride, const unsigned* sh,
                            const unsigned* sl, int ncols, int groups, float inv_scale, void* V, int64_t ldv,
                            cudaStream_t st, int accumulate) {
    using R = RitzH<B, NT>;
    cudaFuncSetAttribute(ritz_h_kernel<B, NT, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::smem_bytes);
    dim3 grid((unsigned)((n + R::ROWS_CTA - 1) / R::ROWS_CTA), (unsigned)groups);
    ritz_h_kernel<B, NT, VT><<<grid, R::NW * 32, R::smem_bytes, st>>>(n, m, (const float*)buf, bstride, sh, sl, ncols, inv_scale,
                                                                     (VT*)V, ldv, k, accumulate);
}

// column groups of at most 64 Ritz columns; NT n-tiles per group from {2, 4, 6, 7, 8}
static void ritz_h_shape(int kpad, int& groups, int& nt) {
    const int tiles = kpad / 8;
    groups = (tiles + 7) / 8;
    const int need = (tiles + groups - 1) / groups;
    nt = need <= 2 ? 2 : need <= 4 ? 4 : need <= 6 ? 6 : need <= 7 ? 7 : 8;
}

size_t ritz_h_scratch_words(int B, int64_t m, int kpad) {
    int groups, nt;
    ritz_h_shape(kpad, groups, nt);
    return 2 * (size_t)m * (B / 2) * (size_t)(groups * nt * 8);
}

void launch_ritz_h(int B, int64_t n, int64_t m, int k, int kpad, const void* buf, int64_t bstride, const void* Smat,
                   void* V, int64_t ldv, int v_fp32, float split_scale, unsigned* scratch, cudaStream_t st, int accumulate) {
    int groups, nt;
    ritz_h_shape(kpad, groups, nt);
    const int ncols = groups * nt * 8;
    const size_t nwords = (size_t)m * (B / 2) * ncols;
    unsigned* sh = scratch;
    unsigned* sl = scratch + nwords;
    const float scale_s = 2048.f;  // |S| <= 1 (unit eigenvectors of T)
    split_ritz_coeff_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(nwords, kpad, ncols, (const float*)Smat, scale_s, sh, sl);
    const float inv = 1.0f / (split_scale * scale_s);
    auto go = [&](auto bc, auto ntc) {
        constexpr int BB = decltype(bc)::value;
        constexpr int NN = decltype(ntc)::value;
        if (v_fp32) ritz_h_launch_t<BB, NN, float>(n, m, k, buf, bstride, sh, sl, ncols, groups, inv, V, ldv, st, accumulate);
        else ritz_h_launch_t<BB, NN, double>(n, m, k, buf, bstride, sh, sl, ncols, groups, inv, V, ldv, st, accumulate);
    };
    auto by_nt = [&](auto bc) {
        switch (nt) {
            case 2: go(bc, std::integral_constant<int, 2>()); break;
            case 4: go(bc, std::integral_constant<int, 4>()); break;
            case 6: go(bc, std::integral_constant<int, 6>()); break;
            case 7: go(bc, std::integral_constant<int, 7>()); break;
            default: go(bc, std::integral_constant<int, 8>()); break;
        }
    };
    if (B == 16) by_nt(std::integral_constant<int, 16>());
    else by_nt(std::integral_constant<int, 32>());
}

}  // namespace rbl

// K5 for the all-fp64 mode (the reference as shipped: FLOAT = DOUBLE = Float64, Julia/common.jl:5-6):
// full re-orthogonalisation against an fp64 Krylov buffer with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64)
// fed by per-warp cp.async rings.  With 2B = 32 targets an fp64 buffer element (8 B) feeds 32 FMAs = 8 flop/B,
// i.e. 52 TFLOP/s at 6.5 TB/s against a measured 37 TFLOP/s fp64 rate (DFMA and DMMA alike): this kernel pair is
// bound by the fp64 pipe at ~70% of the HBM roofline by construction; the SIMT version reached 14%.
//
// Replaces hybrid_part_reorth! / part_reorth_gpu_async!  (Julia/RBL_gpu.jl:59-81, 29-47) with FLOAT = Float64.
#include <cstdio>

#include "kernels.h"

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

// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane (g = lane/4, t = lane%4): a = A[g][t], b = B[t][g],
// c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

}  // namespace

// =================================================================================================
// Gram.  MMA roles: M = 8 Krylov columns of a stored block, N = 8 targets, K = 4 rows.
// 16 warps x 2 stored blocks; 8-row stages; targets [w0 | w1] staged per 64 rows for the whole CTA.
// =================================================================================================
template <int B>
struct GramD {
    static constexpr int NW = 16;
    static constexpr int WB = 2;
    static constexpr int JT = NW * WB;
    static constexpr int MT = B / 8;
    static constexpr int NT = (2 * B) / 8;
    static constexpr int PA = B + 4;             // doubles per staged buffer row (160 B: 32-byte multiple, conflict-free)
    static constexpr int PW = 2 * B + 4;         // doubles per staged target row (288 B)
    static constexpr int NST = 4;
    static constexpr int RS = 8;
    static constexpr int RW = 64;
    static constexpr int STAGE = WB * RS * PA;
    static constexpr int WBUF = RW * PW;
    static constexpr int NCP = (WB * RS * (B / 2)) / 32;   // 16-byte copies per lane per stage
    static constexpr size_t smem_bytes = (size_t)(NW * NST * STAGE + 2 * WBUF) * sizeof(double);
};

template <int B>
__global__ void __launch_bounds__(GramD<B>::NW * 32, 1)
    reorth_gram_d_kernel(int64_t n, int64_t m, const double* __restrict__ buf, int64_t bstride,
                         const double* __restrict__ w0, const double* __restrict__ w1, double* __restrict__ partials,
                         int64_t rows_per_range) {
    using C = GramD<B>;
    constexpr int NW = C::NW, WB = C::WB, JT = C::JT, MT = C::MT, NT = C::NT, PA = C::PA, PW = C::PW, NST = C::NST,
                  RS = C::RS, RW = C::RW, STAGE = C::STAGE, WBUF = C::WBUF, NCP = C::NCP, NTHR = NW * 32;
    static_assert(B == 16, "instantiated for B = 16");
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* sA = smem_d + (size_t)warp * NST * STAGE;
    double* sW = smem_d + (size_t)NW * NST * STAGE;  // [2 buffers][RW][PW]
    const int64_t jbase = (int64_t)blockIdx.x * JT + (int64_t)warp * WB;
    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_range;
    const int64_t rend = min(n, rbeg + rows_per_range);
    const int64_t nrows = rend - rbeg;

    double acc[WB][MT][NT][2];
#pragma unroll
    for (int b = 0; b < WB; ++b)
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int x = 0; x < NT; ++x) acc[b][a][x][0] = acc[b][a][x][1] = 0.0;

    if (nrows > 0) {
        const int nks = (int)((nrows + RS - 1) / RS);
        constexpr int KPC = RW / RS;
        const double* src0[NCP];
        int dst0[NCP], row0[NCP];
        bool blk_ok[NCP];
#pragma unroll
        for (int u = 0; u < NCP; ++u) {
            const int q = lane + 32 * u;
            const int blk = q / (RS * (B / 2));
            const int rem = q % (RS * (B / 2));
            const int row = rem / (B / 2), c2 = rem % (B / 2);
            blk_ok[u] = (jbase + blk) < m;
            row0[u] = row;
            dst0[u] = blk * RS * PA + row * PA + c2 * 2;
            src0[u] = buf + (size_t)(blk_ok[u] ? jbase + blk : 0) * bstride + (size_t)(rbeg + row) * B + c2 * 2;
        }
        auto issue_a = [&](int ks) {
            double* st = sA + (size_t)(ks % NST) * STAGE;
            const int64_t rem_rows = nrows - (int64_t)ks * RS;
            const size_t adv = (size_t)ks * RS * B;
#pragma unroll
            for (int u = 0; u < NCP; ++u) {
                const bool ok = blk_ok[u] && (row0[u] < rem_rows);
                cp_async16(st + dst0[u], ok ? src0[u] + adv : buf, ok ? 16 : 0);
            }
        };
        auto issue_w = [&](int chunk) {  // RW rows x 2B doubles: 64 x 16 copies of 16 B = 1024 -> 2 per thread
            double* dst = sW + (size_t)(chunk & 1) * WBUF;
            const int64_t r0 = rbeg + (int64_t)chunk * RW;
#pragma unroll
            for (int u = 0; u < (RW * (2 * B / 2)) / NTHR; ++u) {
                const int q = tid + NTHR * u;
                const int row = q / B, c2 = q % B;          // c2 in [0, 2B/2)
                const bool ok = (r0 + row < rend);
                const double* src = (c2 < B / 2) ? w0 + (size_t)(r0 + row) * B + c2 * 2
                                                 : w1 + (size_t)(r0 + row) * B + (c2 - B / 2) * 2;
                cp_async16(dst + row * PW + c2 * 2, ok ? src : w0, ok ? 16 : 0);
            }
        };
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

            const double* st = sA + (size_t)(ks % NST) * STAGE;
            const double* wt = sW + (size_t)((ks / KPC) & 1) * WBUF + (size_t)(ks % KPC) * RS * PW;
#pragma unroll
            for (int kk = 0; kk < RS / 4; ++kk) {
                double bf[NT];
#pragma unroll
                for (int x = 0; x < NT; ++x) bf[x] = wt[(kk * 4 + t) * PW + x * 8 + g];
#pragma unroll
                for (int b = 0; b < WB; ++b) {
                    const double* a = st + b * RS * PA + (kk * 4 + t) * PA;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const double af = a[mt * 8 + g];   // A[m = column][k = row]
#pragma unroll
                        for (int x = 0; x < NT; ++x) dmma(acc[b][mt][x], af, bf[x]);
                    }
                }
            }
        }
        cp_async_wait<0>();
    }
#pragma unroll
    for (int b = 0; b < WB; ++b) {
        const int64_t j = jbase + b;
        if (j >= m) continue;
        double* out = partials + ((size_t)blockIdx.y * m * B + (size_t)j * B) * (2 * B);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int x = 0; x < NT; ++x)
                *reinterpret_cast<double2*>(out + (size_t)(mt * 8 + g) * 2 * B + x * 8 + 2 * t) =
                    make_double2(acc[b][mt][x][0], acc[b][mt][x][1]);
    }
}

// =================================================================================================
// Update.  MMA roles: M = 8 rows, N = 8 targets, K = 4 Krylov columns.  16 warps x 16 rows.
// =================================================================================================
template <int B>
struct UpdD {
    static constexpr int NW = 16;
    static constexpr int RWP = 16;               // rows per warp
    static constexpr int MT = RWP / 8;
    static constexpr int NT = (2 * B) / 8;
    static constexpr int PA = B + 4;
    static constexpr int PC = 2 * B + 4;
    static constexpr int JC = 4;
    static constexpr int NST = 4;
    static constexpr int STAGE = RWP * PA;
    static constexpr int CBUF = JC * B * PC;
    static constexpr int ROWS_CTA = NW * RWP;
    static constexpr int NCP = (RWP * (B / 2)) / 32;
    static constexpr size_t smem_bytes = (size_t)(NW * NST * STAGE + 2 * CBUF) * sizeof(double);
};

template <int B>
__global__ void __launch_bounds__(UpdD<B>::NW * 32, 1)
    reorth_update_d_kernel(int64_t n, int64_t m, const double* __restrict__ buf, int64_t bstride,
                           const double* __restrict__ Cmat, double* __restrict__ w0, double* __restrict__ w1,
                           double* __restrict__ store_w1, double* __restrict__ store_w0) {
    using C = UpdD<B>;
    constexpr int NW = C::NW, RWP = C::RWP, MT = C::MT, NT = C::NT, PA = C::PA, PC = C::PC, JC = C::JC, NST = C::NST,
                  STAGE = C::STAGE, CBUF = C::CBUF, NCP = C::NCP, NTHR = NW * 32;
    static_assert(B == 16, "instantiated for B = 16");
    extern __shared__ __align__(16) double smem_d[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* sA = smem_d + (size_t)warp * NST * STAGE;
    double* sC = smem_d + (size_t)NW * NST * STAGE;  // [2 buffers][JC][B][PC]
    const int64_t r0 = (int64_t)blockIdx.x * C::ROWS_CTA + (int64_t)warp * RWP;
    const int mi = (int)m;

    double acc[MT][NT][2];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int x = 0; x < NT; ++x) acc[a][x][0] = acc[a][x][1] = 0.0;

    const double* src0[NCP];
    int dst0[NCP];
    bool row_ok[NCP];
#pragma unroll
    for (int u = 0; u < NCP; ++u) {
        const int q = lane + 32 * u;
        const int row = q / (B / 2), c2 = q % (B / 2);
        row_ok[u] = (r0 + row) < n;
        dst0[u] = row * PA + c2 * 2;
        src0[u] = buf + (size_t)(row_ok[u] ? r0 + row : 0) * B + c2 * 2;
    }
    auto issue_a = [&](int j) {
        double* st = sA + (size_t)(j % NST) * STAGE;
        const size_t adv = (size_t)j * bstride;
#pragma unroll
        for (int u = 0; u < NCP; ++u) {
            const bool ok = row_ok[u] && (j < mi);
            cp_async16(st + dst0[u], ok ? src0[u] + adv : buf, ok ? 16 : 0);
        }
    };
    auto issue_c = [&](int chunk) {  // JC blocks x B rows x 2B doubles = JC*B*B copies of 16 B
        double* dst = sC + (size_t)(chunk & 1) * CBUF;
        const int j0 = chunk * JC;
#pragma unroll
        for (int u = 0; u < (JC * B * B) / NTHR; ++u) {
            const int q = tid + NTHR * u;
            const int rowc = q / B, c2 = q % B;  // rowc = jb*B + c ; c2 indexes the 2B/2 pairs of a row
            const bool ok = (j0 * B + rowc) < mi * B;
            const size_t off = ((size_t)j0 * B + rowc) * (2 * B) + c2 * 2;
            cp_async16(dst + rowc * PC + c2 * 2, ok ? Cmat + off : Cmat, ok ? 16 : 0);
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
    for (int j = 0; j < mi; ++j) {
        cp_async_wait<NST - 2>();
        if ((j % JC) == 0) __syncthreads();
        else __syncwarp();
        if ((j % JC) == 0 && (j / JC + 1) * JC < mi) issue_c(j / JC + 1);
        issue_a(j + NST - 1);
        cp_async_commit();

        const double* st = sA + (size_t)(j % NST) * STAGE;
        const double* cb = sC + (size_t)((j / JC) & 1) * CBUF + (size_t)(j % JC) * B * PC;
#pragma unroll
        for (int kk = 0; kk < B / 4; ++kk) {
            double bf[NT];
#pragma unroll
            for (int x = 0; x < NT; ++x) bf[x] = cb[(kk * 4 + t) * PC + x * 8 + g];
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                const double af = st[(a * 8 + g) * PA + kk * 4 + t];   // A[m = row][k = column]
#pragma unroll
                for (int x = 0; x < NT; ++x) dmma(acc[a][x], af, bf[x]);
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int a = 0; a < MT; ++a) {
        const int64_t row = r0 + a * 8 + g;
        if (row >= n) continue;
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            const int tgt = x * 8 + 2 * t;
            if (tgt < B) {
                double2* p = reinterpret_cast<double2*>(w0 + (size_t)row * B + tgt);
                double2 v = *p;
                v.x -= acc[a][x][0];
                v.y -= acc[a][x][1];
                *p = v;
                if (store_w0 != nullptr) *reinterpret_cast<double2*>(store_w0 + (size_t)row * B + tgt) = v;
            } else {
                double2* p = reinterpret_cast<double2*>(w1 + (size_t)row * B + (tgt - B));
                double2 v = *p;
                v.x -= acc[a][x][0];
                v.y -= acc[a][x][1];
                *p = v;
                if (store_w1 != nullptr) *reinterpret_cast<double2*>(store_w1 + (size_t)row * B + (tgt - B)) = v;
            }
        }
    }
}

__global__ void reorth_reduce_d_kernel(const double* __restrict__ partials, int ranges, size_t count,
                                       double* __restrict__ Cout) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    double s = 0.0;
    for (int p = 0; p < ranges; ++p) s += partials[(size_t)p * count + e];
    Cout[e] = s;
}

// ---- launchers -------------------------------------------------------------------------------------
bool reorth_d_supported(int B, int fp32) { return !fp32 && B == 16; }

void launch_reorth_gram_d(const ReorthPlan& p, const void* buf, int64_t bstride, const double* w0, const double* w1,
                          void* partials, void* Cmat, cudaStream_t st) {
    constexpr int B = 16;
    using G = GramD<B>;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_gram_d_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes);
    int64_t rpr = (p.n + p.ranges - 1) / p.ranges;
    rpr = (rpr + G::RW - 1) / G::RW * G::RW;
    dim3 grid((unsigned)((p.m + G::JT - 1) / G::JT), p.ranges);
    reorth_gram_d_kernel<B><<<grid, G::NW * 32, G::smem_bytes, st>>>(p.n, p.m, (const double*)buf, bstride, w0, w1,
                                                                      (double*)partials, rpr);
    const size_t count = (size_t)p.m * B * 2 * B;
    reorth_reduce_d_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const double*)partials, p.ranges, count,
                                                                            (double*)Cmat);
}

void launch_reorth_update_d(const ReorthPlan& p, const void* buf, int64_t bstride, const void* Cmat, double* w0,
                            double* w1, void* store_w1, cudaStream_t st, void* store_w0) {
    constexpr int B = 16;
    using U = UpdD<B>;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_update_d_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::smem_bytes);
    const unsigned grid = (unsigned)((p.n + U::ROWS_CTA - 1) / U::ROWS_CTA);
    reorth_update_d_kernel<B><<<grid, U::NW * 32, U::smem_bytes, st>>>(p.n, p.m, (const double*)buf, bstride,
                                                                       (const double*)Cmat, w0, w1, (double*)store_w1, (double*)store_w0);
}

}  // namespace rbl

// K5 (tensor-core version): full re-orthogonalisation of the two newest blocks against the fp32 Krylov
// buffer with error-compensated TF32 tensor-core MMAs ("3xTF32": hi*hi + lo*hi + hi*lo, fp32 accumulate,
// i.e. fp32-grade products) and per-warp cp.async pipelines that keep several KB per warp in flight.
//
// Why tensor cores for a bandwidth kernel: with the 2B = 32 target columns of
// [Q_i Q_{i-1}] every fp32 buffer element feeds 32 FMAs = 16 flop/B.  At the measured 6.5 TB/s that is
// 104 TFLOP/s of fp32-grade arithmetic - above the measured 68 TFLOP/s FFMA rate of the SIMT pipes, so a
// CUDA-core kernel is compute-bound at <= 65% of the HBM roofline no matter how it is written.
//
// Replaces hybrid_part_reorth! / part_reorth_gpu_async!  (Julia/RBL_gpu.jl:59-81, 29-47).
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

__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

// D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

}  // namespace

// ---- target split: Whi/Wlo[r][t] (fp32, tf32-representable) from the two fp64 active blocks ----------
template <int B>
__global__ void split_targets_kernel(int64_t n, const double* __restrict__ w0, const double* __restrict__ w1,
                                     float* __restrict__ whi, float* __restrict__ wlo) {
    const int64_t total = n * 2 * B;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const int64_t r = e / (2 * B);
        const int t = (int)(e % (2 * B));
        const float x = (float)((t < B) ? w0[r * B + t] : w1[r * B + (t - B)]);
        unsigned hi, lo;
        split_tf32(x, hi, lo);
        whi[e] = __uint_as_float(hi);
        wlo[e] = __uint_as_float(lo);
    }
}

// =================================================================================================
// Gram:  C_j (B x 2B) = buf_j' * W  for JT stored blocks per CTA, over the CTA's row range.
// MMA roles: M = 16 Krylov columns of a stored block, N = 8 target columns, K = 8 rows.
// Each warp owns WB stored blocks and streams their rows through a private NST-stage cp.async ring
// (8 rows per stage); the target tile (hi and lo parts) is shared by the CTA in 64-row chunks.
// =================================================================================================
template <int B>
struct GramTc {
    static constexpr int WB = 4;                 // stored blocks per warp
    static constexpr int JT = 8 * WB;            // stored blocks per CTA
    static constexpr int NT = (2 * B) / 8;       // n-tiles
    static constexpr int PA = B + 8;             // smem pitch of a buffer row (floats): conflict-free A fragments
    static constexpr int PW = 2 * B + 8;         // smem pitch of a target row
    static constexpr int NST = 5;
    static constexpr int RW = 64;                // rows per shared target chunk
    static constexpr int STAGE = WB * 8 * PA;    // floats per warp stage
    static constexpr int WBUF = RW * PW;         // floats per target buffer (hi or lo)
    static constexpr size_t smem_bytes = (size_t)(8 * NST * STAGE + 4 * WBUF) * sizeof(float);
};

template <int B>
__global__ void __launch_bounds__(256, 1)
    reorth_gram_tc_kernel(int64_t n, int64_t m, const float* __restrict__ buf, int64_t bstride,
                          const float* __restrict__ whi, const float* __restrict__ wlo, float* __restrict__ partials,
                          int64_t rows_per_range) {
    using C = GramTc<B>;
    constexpr int WB = C::WB, JT = C::JT, NT = C::NT, PA = C::PA, PW = C::PW, NST = C::NST, RW = C::RW,
                  STAGE = C::STAGE, WBUF = C::WBUF;
    static_assert(B == 16, "tensor-core Gram is instantiated for B = 16");
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    float* sA = smem + (size_t)warp * NST * STAGE;
    float* sW = smem + (size_t)8 * NST * STAGE;  // [2 buffers][hi, lo][RW][PW]
    const int64_t jbase = (int64_t)blockIdx.x * JT + (int64_t)warp * WB;
    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_range;
    const int64_t rend = min(n, rbeg + rows_per_range);
    const int64_t nrows = rend - rbeg;

    float acc[WB][NT][4];
#pragma unroll
    for (int b = 0; b < WB; ++b)
#pragma unroll
        for (int x = 0; x < NT; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[b][x][y] = 0.f;

    if (nrows > 0) {
        const int nks = (int)((nrows + 7) / 8);
        auto issue_a = [&](int ks) {
            float* st = sA + (size_t)(ks % NST) * STAGE;
            const int64_t r8 = rbeg + (int64_t)ks * 8;
#pragma unroll
            for (int u = 0; u < (WB * 8 * (B / 4)) / 32; ++u) {
                const int q = lane + 32 * u;
                const int blk = q / (8 * (B / 4));
                const int rem = q % (8 * (B / 4));
                const int row = rem / (B / 4), c4 = rem % (B / 4);
                const bool ok = (ks < nks) && (r8 + row < rend) && (jbase + blk < m);
                const float* src = ok ? buf + (size_t)(jbase + blk) * bstride + (size_t)(r8 + row) * B + c4 * 4 : buf;
                cp_async16(st + blk * 8 * PA + row * PA + c4 * 4, src, ok ? 16 : 0);
            }
        };
        auto issue_w = [&](int chunk) {  // whole CTA: RW rows x 2B floats, hi and lo
            float* dsth = sW + (size_t)((chunk & 1) * 2 + 0) * WBUF;
            float* dstl = sW + (size_t)((chunk & 1) * 2 + 1) * WBUF;
            const int64_t r0 = rbeg + (int64_t)chunk * RW;
#pragma unroll
            for (int u = 0; u < (RW * 2 * B / 4) / 256; ++u) {
                const int q = tid + 256 * u;
                const int row = q / (2 * B / 4), c4 = q % (2 * B / 4);
                const bool ok = (r0 + row < rend);
                const size_t off = (size_t)(r0 + row) * 2 * B + c4 * 4;
                cp_async16(dsth + row * PW + c4 * 4, ok ? whi + off : whi, ok ? 16 : 0);
                cp_async16(dstl + row * PW + c4 * 4, ok ? wlo + off : wlo, ok ? 16 : 0);
            }
        };
        // prologue: target chunk 0 and the first NST-1 buffer stages, one commit group per stage
        issue_w(0);
        issue_a(0);
        cp_async_commit();
#pragma unroll
        for (int s = 1; s < NST - 1; ++s) {
            issue_a(s);
            cp_async_commit();
        }
        for (int ks = 0; ks < nks; ++ks) {
            cp_async_wait<NST - 2>();        // stage ks (and any target chunk issued with it or earlier) has landed
            if ((ks & 7) == 0) __syncthreads();  // chunk boundary: everybody's target-chunk copies are visible
            else __syncwarp();
            // refill the slot consumed at ks-1; piggy-back the next target chunk on the first stage of a chunk
            if ((ks & 7) == 0 && (int64_t)(ks / 8 + 1) * RW < nrows) issue_w(ks / 8 + 1);
            issue_a(ks + NST - 1);
            cp_async_commit();

            const float* st = sA + (size_t)(ks % NST) * STAGE;
            const int chunk = ks >> 3;
            const float* wh = sW + (size_t)((chunk & 1) * 2 + 0) * WBUF + (size_t)(ks & 7) * 8 * PW;
            const float* wl = sW + (size_t)((chunk & 1) * 2 + 1) * WBUF + (size_t)(ks & 7) * 8 * PW;
            unsigned bh[NT][2], bl[NT][2];
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                bh[x][0] = __float_as_uint(wh[t * PW + x * 8 + g]);
                bh[x][1] = __float_as_uint(wh[(t + 4) * PW + x * 8 + g]);
                bl[x][0] = __float_as_uint(wl[t * PW + x * 8 + g]);
                bl[x][1] = __float_as_uint(wl[(t + 4) * PW + x * 8 + g]);
            }
#pragma unroll
            for (int b = 0; b < WB; ++b) {
                const float* a = st + b * 8 * PA;
                unsigned ah[4], al[4];
                split_tf32(a[t * PA + g], ah[0], al[0]);
                split_tf32(a[t * PA + g + 8], ah[1], al[1]);
                split_tf32(a[(t + 4) * PA + g], ah[2], al[2]);
                split_tf32(a[(t + 4) * PA + g + 8], ah[3], al[3]);
#pragma unroll
                for (int x = 0; x < NT; ++x) {
                    mma_tf32(acc[b][x], al, bh[x]);
                    mma_tf32(acc[b][x], ah, bl[x]);
                    mma_tf32(acc[b][x], ah, bh[x]);
                }
            }
        }
        cp_async_wait<0>();
    }
    // partials[range][(j*B + c)][2B]: fragment (c0,c1) = (row g, cols 2t,2t+1), (c2,c3) = (row g+8, ...)
#pragma unroll
    for (int b = 0; b < WB; ++b) {
        const int64_t j = jbase + b;
        if (j >= m) continue;
        float* out = partials + ((size_t)blockIdx.y * m * B + (size_t)j * B) * (2 * B);
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            *reinterpret_cast<float2*>(out + (size_t)g * 2 * B + x * 8 + 2 * t) = make_float2(acc[b][x][0], acc[b][x][1]);
            *reinterpret_cast<float2*>(out + (size_t)(g + 8) * 2 * B + x * 8 + 2 * t) = make_float2(acc[b][x][2], acc[b][x][3]);
        }
    }
}

// sum over row ranges (double), emit C and its tf32 hi/lo parts for the update kernel
__global__ void reorth_reduce_split_kernel(const float* __restrict__ partials, int ranges, size_t count,
                                           float* __restrict__ Cout, float* __restrict__ Chi, float* __restrict__ Clo) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    double s = 0.0;
    for (int p = 0; p < ranges; ++p) s += (double)partials[(size_t)p * count + e];
    const float c = (float)s;
    unsigned hi, lo;
    split_tf32(c, hi, lo);
    Cout[e] = c;
    Chi[e] = __uint_as_float(hi);
    Clo[e] = __uint_as_float(lo);
}

// =================================================================================================
// Update:  W(rows x 2B) -= sum_j buf_j(rows x B) * C_j(B x 2B).
// MMA roles: M = 16 rows, N = 8 targets, K = 8 Krylov columns (B/8 k-steps per stored block).
// Each warp owns 32 rows (two m-tiles) for the whole sweep over the m stored blocks and streams its
// 32 x B slice of every block through a private NST-stage cp.async ring; the coefficient blocks
// (hi and lo) are shared by the CTA in chunks of JC blocks (double-buffered, same commit groups).
// =================================================================================================
template <int B>
struct UpdTc {
    static constexpr int MT = 2;
    static constexpr int NT = (2 * B) / 8;
    static constexpr int KS = B / 8;
    static constexpr int PA = B + 4;             // pitch: conflict-free A fragments (rows along g)
    static constexpr int PC = 2 * B + 8;
    static constexpr int JC = 8;
    static constexpr int NST = 4;
    static constexpr int STAGE = 32 * PA;
    static constexpr int CBUF = JC * B * PC;     // floats per coefficient buffer (hi or lo)
    static constexpr int ROWS_CTA = 8 * 32;
    static constexpr size_t smem_bytes = (size_t)(8 * NST * STAGE + 4 * CBUF) * sizeof(float);
};

template <int B>
__global__ void __launch_bounds__(256, 1)
    reorth_update_tc_kernel(int64_t n, int64_t m, const float* __restrict__ buf, int64_t bstride,
                            const float* __restrict__ Chi, const float* __restrict__ Clo, double* __restrict__ w0,
                            double* __restrict__ w1, float* __restrict__ store_w1) {
    using C = UpdTc<B>;
    constexpr int MT = C::MT, NT = C::NT, KS = C::KS, PA = C::PA, PC = C::PC, JC = C::JC, NST = C::NST,
                  STAGE = C::STAGE, CBUF = C::CBUF;
    static_assert(B == 16, "tensor-core update is instantiated for B = 16");
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    float* sA = smem + (size_t)warp * NST * STAGE;
    float* sC = smem + (size_t)8 * NST * STAGE;  // [2 buffers][hi, lo][JC][B][PC]
    const int64_t r0 = (int64_t)blockIdx.x * C::ROWS_CTA + (int64_t)warp * 32;
    const int mi = (int)m;

    float acc[MT][NT][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int x = 0; x < NT; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[a][x][y] = 0.f;

    auto issue_a = [&](int j) {
        float* st = sA + (size_t)(j % NST) * STAGE;
#pragma unroll
        for (int u = 0; u < (32 * (B / 4)) / 32; ++u) {
            const int q = lane + 32 * u;
            const int row = q / (B / 4), c4 = q % (B / 4);
            const bool ok = (j < mi) && (r0 + row < n);
            const float* src = ok ? buf + (size_t)j * bstride + (size_t)(r0 + row) * B + c4 * 4 : buf;
            cp_async16(st + row * PA + c4 * 4, src, ok ? 16 : 0);
        }
    };
    auto issue_c = [&](int chunk) {
        float* dsth = sC + (size_t)((chunk & 1) * 2 + 0) * CBUF;
        float* dstl = sC + (size_t)((chunk & 1) * 2 + 1) * CBUF;
        const int j0 = chunk * JC;
#pragma unroll
        for (int u = 0; u < (JC * B * (2 * B / 4)) / 256; ++u) {
            const int q = tid + 256 * u;
            const int rowc = q / (2 * B / 4), c4 = q % (2 * B / 4);   // rowc = jb*B + c
            const bool ok = (j0 * B + rowc) < mi * B;
            const size_t off = ((size_t)j0 * B + rowc) * (2 * B) + c4 * 4;
            cp_async16(dsth + rowc * PC + c4 * 4, ok ? Chi + off : Chi, ok ? 16 : 0);
            cp_async16(dstl + rowc * PC + c4 * 4, ok ? Clo + off : Clo, ok ? 16 : 0);
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

        const float* st = sA + (size_t)(j % NST) * STAGE;
        const int chunk = j / JC;
        const float* ch = sC + (size_t)((chunk & 1) * 2 + 0) * CBUF + (size_t)(j % JC) * B * PC;
        const float* cl = sC + (size_t)((chunk & 1) * 2 + 1) * CBUF + (size_t)(j % JC) * B * PC;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            unsigned bh[NT][2], bl[NT][2];
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                bh[x][0] = __float_as_uint(ch[(ks * 8 + t) * PC + x * 8 + g]);
                bh[x][1] = __float_as_uint(ch[(ks * 8 + t + 4) * PC + x * 8 + g]);
                bl[x][0] = __float_as_uint(cl[(ks * 8 + t) * PC + x * 8 + g]);
                bl[x][1] = __float_as_uint(cl[(ks * 8 + t + 4) * PC + x * 8 + g]);
            }
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                const float* ap = st + (a * 16) * PA + ks * 8;
                unsigned ah[4], al[4];
                split_tf32(ap[g * PA + t], ah[0], al[0]);
                split_tf32(ap[(g + 8) * PA + t], ah[1], al[1]);
                split_tf32(ap[g * PA + t + 4], ah[2], al[2]);
                split_tf32(ap[(g + 8) * PA + t + 4], ah[3], al[3]);
#pragma unroll
                for (int x = 0; x < NT; ++x) {
                    mma_tf32(acc[a][x], al, bh[x]);
                    mma_tf32(acc[a][x], ah, bl[x]);
                    mma_tf32(acc[a][x], ah, bh[x]);
                }
            }
        }
    }
    cp_async_wait<0>();
    // W -= acc.  Fragment: (c0,c1) = (row g, targets 2t,2t+1), (c2,c3) = (row g+8, ...)
#pragma unroll
    for (int a = 0; a < MT; ++a) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t row = r0 + a * 16 + g + 8 * h;
            if (row >= n) continue;
#pragma unroll
            for (int x = 0; x < NT; ++x) {
                const int tgt = x * 8 + 2 * t;
                const float d0 = acc[a][x][2 * h], d1 = acc[a][x][2 * h + 1];
                if (tgt < B) {
                    double2* p = reinterpret_cast<double2*>(w0 + (size_t)row * B + tgt);
                    double2 v = *p;
                    v.x -= (double)d0;
                    v.y -= (double)d1;
                    *p = v;
                } else {
                    double2* p = reinterpret_cast<double2*>(w1 + (size_t)row * B + (tgt - B));
                    double2 v = *p;
                    v.x -= (double)d0;
                    v.y -= (double)d1;
                    *p = v;
                    if (store_w1 != nullptr)
                        *reinterpret_cast<float2*>(store_w1 + (size_t)row * B + (tgt - B)) = make_float2((float)v.x, (float)v.y);
                }
            }
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------------
bool reorth_tc_supported(int B, int fp32) { return fp32 && B == 16; }

size_t reorth_tc_scratch_floats(int B, int64_t n, int64_t m_cap) {
    // Whi, Wlo (n x 2B each) + Chi, Clo (m_cap*B x 2B each)
    return 2 * (size_t)n * 2 * B + 2 * (size_t)m_cap * B * 2 * B;
}

void launch_reorth_gram_tc(const ReorthPlan& p, const void* buf, int64_t bstride, const double* w0, const double* w1,
                           void* partials, void* Cmat, float* scratch, int64_t m_cap, cudaStream_t st) {
    constexpr int B = 16;
    using G = GramTc<B>;
    float* whi = scratch;
    float* wlo = whi + (size_t)p.n * 2 * B;
    float* chi = wlo + (size_t)p.n * 2 * B;
    float* clo = chi + (size_t)m_cap * B * 2 * B;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_gram_tc_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes);
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t total = p.n * 2 * B;
    const int sgrid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
    split_targets_kernel<B><<<sgrid, 256, 0, st>>>(p.n, w0, w1, whi, wlo);
    int64_t rpr = (p.n + p.ranges - 1) / p.ranges;
    rpr = (rpr + 63) / 64 * 64;
    dim3 grid(p.chunks, p.ranges);
    reorth_gram_tc_kernel<B><<<grid, 256, G::smem_bytes, st>>>(p.n, p.m, (const float*)buf, bstride, whi, wlo,
                                                                (float*)partials, rpr);
    const size_t count = (size_t)p.m * B * 2 * B;
    reorth_reduce_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const float*)partials, p.ranges, count,
                                                                                 (float*)Cmat, chi, clo);
}

__global__ void resplit_kernel(size_t count, const float* __restrict__ Cin, float* __restrict__ Chi, float* __restrict__ Clo) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    unsigned hi, lo;
    split_tf32(Cin[e], hi, lo);
    Chi[e] = __uint_as_float(hi);
    Clo[e] = __uint_as_float(lo);
}

void launch_reorth_gram_tc_resplit(const ReorthPlan& p, void* Cmat, float* scratch, int64_t m_cap, cudaStream_t st) {
    constexpr int B = 16;
    float* chi = scratch + 2 * (size_t)p.n * 2 * B;
    float* clo = chi + (size_t)m_cap * B * 2 * B;
    const size_t count = (size_t)p.m * B * 2 * B;
    resplit_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, (const float*)Cmat, chi, clo);
}

void launch_reorth_update_tc(const ReorthPlan& p, const void* buf, int64_t bstride, double* w0, double* w1,
                             void* store_w1, float* scratch, int64_t m_cap, cudaStream_t st) {
    constexpr int B = 16;
    using U = UpdTc<B>;
    float* chi = scratch + 2 * (size_t)p.n * 2 * B;
    float* clo = chi + (size_t)m_cap * B * 2 * B;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_update_tc_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::smem_bytes);
    const unsigned grid = (unsigned)((p.n + U::ROWS_CTA - 1) / U::ROWS_CTA);
    reorth_update_tc_kernel<B><<<grid, 256, U::smem_bytes, st>>>(p.n, p.m, (const float*)buf, bstride, chi, clo, w0, w1,
                                                                  (float*)store_w1);
}

}  // namespace rbl

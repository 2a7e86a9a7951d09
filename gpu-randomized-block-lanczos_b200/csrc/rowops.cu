// K2/K3 of the per-step work, second generation: ONE fused, software-pipelined kernel for every n x B block pass of a
// Lanczos step (3-term recurrence, CholQR passes, local re-orthogonalisation, slab store), fp64 tensor-core MMAs
// (mma.sync.m8n8k4.f64) for all B x B products and Gram matrices.
//
// Reference call sites replaced (Julia/RBL_gpu.jl): transpose(Qg_d)*U, mul!(U,Qg1_d,transpose(Big)), mul!(U,Qg_d,Ai)
// (:153-154,:177-179, cuBLAS dgemm x3), qr(U) / CuArray(fact.Q) (:155-157,:180-182, cuSOLVER geqrf/orgqr), the
// effective projection of loc_reorth_gpu! (:83-93) and the copy of the new block into the buffer (:168-172).
//
// For every row tile (64 rows) of Y (n x B, fp64, in place):
//     y <- y * Rinv                  (upper triangular, optional)                      CholQR apply
//     y <- y - x1 * M1               (optional; M1 B x B, optionally transposed)       3-term / local reorth
//     Y <- y ; store <- y            (optional slab copy: fp64 / fp32 / split16)
//     G0 += y' y                     (optional)                                        Gram for the next Cholesky
//     G1 += z' y                     (optional)                                        A_i, or the local-reorth overlap
// Tiles are staged with cp.async into padded shared memory (2 stages: the next tile streams in while this one is
// computed); each warp owns 8 rows of the tile for the products and one or more 8 x 8 Gram tiles for the reductions.
// The first-generation kernel (kernels.cu rowop_kernel: thread-per-row FMAs out of shared memory, tile load /
// compute / store serialised inside a CTA) reached 27-52% of the HBM peak and needed 18.5 block passes per step;
// the step now takes 15.5 passes (solver.cu) through this kernel.
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
// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane (g = lane/4, t = lane%4): a = A[g][t], b = B[t][g],
// c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

}  // namespace

template <int B>
struct FusedCfg {
    static constexpr int NW = 8;                 // warps per CTA
    static constexpr int TR = 64;                // rows per tile: 8 per warp
    static constexpr int P = B + 4;              // doubles per staged row: 32-byte multiple, conflict-free 64-bit fragment loads
    static constexpr int NT = B / 8;             // 8-column tiles
    static constexpr int KS = B / 4;             // k-steps of a B x B product
    static constexpr int TILE = TR * P;          // doubles per staged array
    static constexpr int NSTAGE = 2;
    static constexpr int JOBS = NT * NT;         // 8 x 8 tiles of one Gram matrix
    static constexpr int JPW = (2 * JOBS + NW - 1) / NW;   // Gram tiles per warp when both Grams are on
    static constexpr size_t smem_bytes = (size_t)(NSTAGE * 3 * TILE + 2 * B * P) * sizeof(double);
    static constexpr int CTAS_PER_SM = (B == 16) ? 3 : 1;
};

template <int B>
__global__ void __launch_bounds__(FusedCfg<B>::NW * 32, FusedCfg<B>::CTAS_PER_SM) fused_rowop_kernel(FusedArgs a) {
    using C = FusedCfg<B>;
    constexpr int NW = C::NW, TR = C::TR, P = C::P, NT = C::NT, KS = C::KS, TILE = C::TILE, JOBS = C::JOBS, JPW = C::JPW,
                  NTHR = NW * 32;
    if (a.skip_flag != nullptr && *a.skip_flag == 0) return;
    extern __shared__ __align__(16) double fsm[];
    double* sM = fsm + (size_t)C::NSTAGE * 3 * TILE;   // -M1 (row-major [k][col], pitch P)
    double* sRi = sM + B * P;                          // Rinv
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const bool has_x = a.x1 != nullptr, has_z = a.z != nullptr, has_r = a.rinv != nullptr;
    const bool gram_yy = a.gram_yy && (a.gram_flag == nullptr || *a.gram_flag != 0);
    const int64_t ntiles = (a.n + TR - 1) / TR;

    for (int e = tid; e < B * B; e += NTHR) {
        const int r = e / B, c = e % B;
        if (has_x) sM[r * P + c] = -(a.m1_transposed ? a.m1[c * B + r] : a.m1[e]);
        if (has_r) sRi[r * P + c] = a.rinv[e];
    }

    auto stage_ptr = [&](int s, int arr) { return fsm + (size_t)(s * 3 + arr) * TILE; };
    // 16-byte chunks of a tile: TR rows x (B/2) chunks
    auto issue = [&](int64_t tile, int s) {
        const int64_t r0 = tile * TR;
        constexpr int CH = TR * (B / 2);
#pragma unroll
        for (int u = 0; u < CH / NTHR; ++u) {
            const int q = tid + u * NTHR;
            const int r = q / (B / 2), c2 = q % (B / 2);
            const bool ok = (r0 + r) < a.n;
            const size_t off = (size_t)(ok ? r0 + r : 0) * B + c2 * 2;
            cp_async16(stage_ptr(s, 0) + r * P + c2 * 2, a.y + off, ok ? 16 : 0);
            if (has_x) cp_async16(stage_ptr(s, 1) + r * P + c2 * 2, a.x1 + off, ok ? 16 : 0);
            if (has_z) cp_async16(stage_ptr(s, 2) + r * P + c2 * 2, a.z + off, ok ? 16 : 0);
        }
    };
    static_assert((TR * (B / 2)) % NTHR == 0, "tile copy must tile the CTA");

    // Gram tiles owned by this warp: job id = warp + NW * u over [G0 tiles | G1 tiles]
    double gacc[JPW][2];
#pragma unroll
    for (int u = 0; u < JPW; ++u) gacc[u][0] = gacc[u][1] = 0.0;

    int64_t tile = blockIdx.x;
    if (tile < ntiles) issue(tile, 0);
    cp_async_commit();
    int s = 0;
    for (; tile < ntiles; tile += gridDim.x, s ^= 1) {
        const int64_t next = tile + gridDim.x;
        cp_async_wait<0>();
        __syncthreads();            // this tile landed for everybody; the other stage is free (its write-back is done)
        if (next < ntiles) issue(next, s ^ 1);
        cp_async_commit();
        double* sY = stage_ptr(s, 0);
        const double* sX = stage_ptr(s, 1);
        const double* sZ = stage_ptr(s, 2);
        const int64_t r0 = tile * TR;

        // ---- products on this warp's 8 rows -----------------------------------------------------------------
        if (has_x || has_r) {
            double* yrow = sY + (warp * 8) * P;
            double c[NT][2];
            if (has_r) {
                double ay[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) ay[ks] = yrow[g * P + 4 * ks + t];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    c[nt][0] = c[nt][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        if (4 * ks <= 8 * nt + 7)     // Rinv is upper triangular: rows below the column tile are zero
                            dmma(c[nt], ay[ks], sRi[(4 * ks + t) * P + 8 * nt + g]);
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    c[nt][0] = yrow[g * P + 8 * nt + 2 * t];
                    c[nt][1] = yrow[g * P + 8 * nt + 2 * t + 1];
                }
            }
            if (has_x) {
                const double* xrow = sX + (warp * 8) * P;
                double ax[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) ax[ks] = xrow[g * P + 4 * ks + t];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) dmma(c[nt], ax[ks], sM[(4 * ks + t) * P + 8 * nt + g]);
            }
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(yrow + g * P + 8 * nt + 2 * t) = make_double2(c[nt][0], c[nt][1]);
        }
        __syncthreads();

        // ---- Gram tiles over the 64 rows (rows past n are zero) ------------------------------------------------
        if (gram_yy || has_z) {
#pragma unroll
            for (int u = 0; u < JPW; ++u) {
                const int job = warp + NW * u;
                const int sel = job / JOBS;              // 0: y'y, 1: z'y
                const int jt = job % JOBS;
                const bool on = (sel == 0) ? gram_yy : (sel == 1 && has_z);
                if (job < 2 * JOBS && on) {
                    const int ti = jt / NT, tj = jt % NT;
                    const double* left = (sel == 0) ? sY : sZ;
#pragma unroll 4
                    for (int k4 = 0; k4 < TR / 4; ++k4)
                        dmma(gacc[u], left[(4 * k4 + t) * P + 8 * ti + g], sY[(4 * k4 + t) * P + 8 * tj + g]);
                }
            }
        }

        // ---- write back ------------------------------------------------------------------------------------------
        const int rows = (int)min((int64_t)TR, a.n - r0);
        const int nvec = rows * (B / 2);
        if (a.write_y) {
            double2* gy = reinterpret_cast<double2*>(a.y + (size_t)r0 * B);
            for (int idx = tid; idx < nvec; idx += NTHR) {
                const int r = idx / (B / 2), c2 = idx % (B / 2);
                gy[idx] = *reinterpret_cast<const double2*>(sY + r * P + 2 * c2);
            }
        }
        if (a.store != nullptr) {
            if (a.store_fp32 && a.store_split_scale != 0.f) {
                unsigned* gw = reinterpret_cast<unsigned*>(a.store) + (size_t)r0 * B;
                for (int idx = tid; idx < nvec; idx += NTHR) {
                    const int r = idx / (B / 2), p = idx % (B / 2);
                    unsigned hi, lo;
                    split_h2((float)sY[r * P + 2 * p], (float)sY[r * P + 2 * p + 1], a.store_split_scale, hi, lo);
                    gw[r * B + p] = hi;
                    gw[r * B + B / 2 + p] = lo;
                }
            } else if (a.store_fp32) {
                float2* gs = reinterpret_cast<float2*>(reinterpret_cast<float*>(a.store) + (size_t)r0 * B);
                for (int idx = tid; idx < nvec; idx += NTHR) {
                    const int r = idx / (B / 2), c2 = idx % (B / 2);
                    gs[idx] = make_float2((float)sY[r * P + 2 * c2], (float)sY[r * P + 2 * c2 + 1]);
                }
            } else {
                double2* gs = reinterpret_cast<double2*>(reinterpret_cast<double*>(a.store) + (size_t)r0 * B);
                for (int idx = tid; idx < nvec; idx += NTHR) {
                    const int r = idx / (B / 2), c2 = idx % (B / 2);
                    gs[idx] = *reinterpret_cast<const double2*>(sY + r * P + 2 * c2);
                }
            }
        }
    }
    cp_async_wait<0>();
    // ---- per-CTA Gram partials: [cta][2][B*B] (slot 0: y'y, slot 1: z'y); unused slots are written as zeros -------
    if (a.partials != nullptr) {
        double* out = a.partials + (size_t)blockIdx.x * 2 * B * B;
#pragma unroll
        for (int u = 0; u < JPW; ++u) {
            const int job = warp + NW * u;
            if (job < 2 * JOBS) {
                const int sel = job / JOBS, jt = job % JOBS;
                const int ti = jt / NT, tj = jt % NT;
                *reinterpret_cast<double2*>(out + (size_t)sel * B * B + (8 * ti + g) * B + 8 * tj + 2 * t) =
                    make_double2(gacc[u][0], gacc[u][1]);
            }
        }
    }
}

bool fused_rowop_supported(int B) { return B == 16 || B == 32; }

int fused_rowop_grid(int B, int64_t n) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (n + 63) / 64;
    const int64_t cap = (int64_t)sms * (B == 16 ? 3 : 1);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, cap));
}

void launch_fused_rowop(int B, const FusedArgs& a, int grid, cudaStream_t st) {
    if (B == 16) {
        using C = FusedCfg<16>;
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(fused_rowop_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        fused_rowop_kernel<16><<<grid, C::NW * 32, C::smem_bytes, st>>>(a);
    } else if (B == 32) {
        using C = FusedCfg<32>;
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(fused_rowop_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        fused_rowop_kernel<32><<<grid, C::NW * 32, C::smem_bytes, st>>>(a);
    } else {
        std::fprintf(stderr, "rbl: fused row kernel needs a padded block size of 16 or 32\n");
        std::abort();
    }
}

}  // namespace rbl

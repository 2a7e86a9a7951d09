// Hand-written sm_100a kernels of the RBL hot path.  Interface and layout: kernels.h / DESIGN.md.
//
// Reference call sites replaced (Julia/RBL_gpu.jl):
//   spmm_kernel          mul!(U,Ag,Qg_d)                                   :152,:176  (cuSPARSE SpMM)
//   rowop_kernel         transpose(Qg_d)*U, mul!(U,Qg1_d,transpose(Big)),  :153-154,:177-179 (cuBLAS dgemm x3)
//                        mul!(U,Qg_d,Ai), and the apply steps of qr(U)     :155-157,:180-182 (cuSOLVER geqrf/orgqr)
//                        loc_reorth_gpu! effective projection              :83-93
//   chol_kernel          the b x b triangular factor of qr(U)              :159,:184
//   reorth_gram/update   hybrid_part_reorth! / part_reorth_gpu_async!      :59-81,:29-47 (4 cuBLAS gemm per block)
//   ritz_kernel          recover_eigvec                                    :106-132
#include "kernels.h"
#include "split16.h"

#include <cstdio>
#include <cstdlib>

namespace rbl {

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <typename F>
static void dispatch_B(int B, F&& f) {
    switch (B) {
        case 4: f(std::integral_constant<int, 4>()); break;
        case 8: f(std::integral_constant<int, 8>()); break;
        case 16: f(std::integral_constant<int, 16>()); break;
        case 32: f(std::integral_constant<int, 32>()); break;
        default: std::fprintf(stderr, "rbl: unsupported padded block size %d\n", B); std::abort();
    }
}

// =================================================================================================
// K1: block SpMM  (CSR, int32 indices, fp64 values; dense blocks row-major so that one gathered row of
// Q is one contiguous B*8-byte segment = one coalesced 128 B request at B = 16)
// B/2 lanes share a row, each lane owns two adjacent columns (16-byte loads).
// =================================================================================================
// MODE 0: all rows (the hot path: 32 registers, 8 CTAs per SM - the other modes need 40); 1: the rows in `rowlist`; 2: all rows
// except those flagged in `skip`
template <int B, int MODE>
__global__ void __launch_bounds__(256) spmm_kernel(int64_t nrows, const int* __restrict__ rowptr,
                                                   const int* __restrict__ colidx, const double* __restrict__ vals,
                                                   const double* __restrict__ Q, double* U, SpmmCoef cf,
                                                   const double* Z, const int* __restrict__ rowlist,
                                                   const unsigned char* __restrict__ skip) {
    constexpr int LPR = B / 2;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int rsel = lane / LPR;
    // (A blocked assignment - every CTA walking a contiguous row range so that L1 would serve the +-1 / +-N stencil
    // neighbours - was measured and is slower: 117-222 us against 105 us at 2-16 CTAs per SM on config 2.)
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double2* __restrict__ Q2 = reinterpret_cast<const double2*>(Q);
    // (Two rows per thread and iteration - more loads in flight per thread - was measured too: 120 us against 96 us.)
    // rowlist != nullptr: the nrows rows listed there (row-sharded runs: the rows that reference halo columns, computed once
    // the halo has arrived); skip != nullptr: rows flagged there are left alone (the same rows, while the halo is in flight)
    for (int64_t idx = warp * RPW + rsel; idx < nrows; idx += nwarps * RPW) {
        int64_t row = idx;
        if constexpr (MODE == 1) row = (int64_t)__ldg(rowlist + idx);
        if constexpr (MODE == 2) {
            if (skip[row]) continue;
        }
        int p = __ldg(rowptr + row);
        const int p1 = __ldg(rowptr + row + 1);
        double2 acc = make_double2(0.0, 0.0);
        for (; p + 4 <= p1; p += 4) {
            const int c0 = __ldg(colidx + p), c1 = __ldg(colidx + p + 1), c2 = __ldg(colidx + p + 2),
                      c3 = __ldg(colidx + p + 3);
            const double v0 = __ldg(vals + p), v1 = __ldg(vals + p + 1), v2 = __ldg(vals + p + 2),
                         v3 = __ldg(vals + p + 3);
            const double2 q0 = __ldg(Q2 + (size_t)c0 * LPR + sub);
            const double2 q1 = __ldg(Q2 + (size_t)c1 * LPR + sub);
            const double2 q2 = __ldg(Q2 + (size_t)c2 * LPR + sub);
            const double2 q3 = __ldg(Q2 + (size_t)c3 * LPR + sub);
            acc.x = fma(v0, q0.x, acc.x); acc.y = fma(v0, q0.y, acc.y);
            acc.x = fma(v1, q1.x, acc.x); acc.y = fma(v1, q1.y, acc.y);
            acc.x = fma(v2, q2.x, acc.x); acc.y = fma(v2, q2.y, acc.y);
            acc.x = fma(v3, q3.x, acc.x); acc.y = fma(v3, q3.y, acc.y);
        }
        for (; p < p1; ++p) {
            const int c0 = __ldg(colidx + p);
            const double v0 = __ldg(vals + p);
            const double2 q0 = __ldg(Q2 + (size_t)c0 * LPR + sub);
            acc.x = fma(v0, q0.x, acc.x); acc.y = fma(v0, q0.y, acc.y);
        }
        acc.x *= cf.alpha;
        acc.y *= cf.alpha;
        if (cf.beta != 0.0) {
            const double2 q = __ldg(Q2 + (size_t)row * LPR + sub);
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
}

void launch_spmm(int B, int64_t nrows, const int* rowptr, const int* colidx, const double* vals, const double* Q,
                 double* U, SpmmCoef cf, const double* Z, cudaStream_t st, const int* rowlist, const unsigned char* skip) {
    if (nrows <= 0) return;
    dispatch_B(B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        constexpr int RPW = 32 / (BB / 2);
        int64_t rows_per_cta = (int64_t)RPW * 8;
        int64_t want = (nrows + rows_per_cta - 1) / rows_per_cta;
        int grid = (int)std::min<int64_t>(want, (int64_t)num_sms() * 16);
        if (rowlist) spmm_kernel<BB, 1><<<grid, 256, 0, st>>>(nrows, rowptr, colidx, vals, Q, U, cf, Z, rowlist, skip);
        else if (skip) spmm_kernel<BB, 2><<<grid, 256, 0, st>>>(nrows, rowptr, colidx, vals, Q, U, cf, Z, rowlist, skip);
        else spmm_kernel<BB, 0><<<grid, 256, 0, st>>>(nrows, rowptr, colidx, vals, Q, U, cf, Z, rowlist, skip);
    });
}

// =================================================================================================
// K2/K3/K4: fused row-wise block operation (see RowOpArgs).  One CTA = TR threads works on tiles of TR
// rows: coalesced tile loads into padded shared memory, thread-per-row small matrix products with the
// B x B coefficient matrices broadcast from shared memory, then a register-tiled Gram accumulation
// over the tile and a coalesced write-back.
// =================================================================================================
template <int B>
struct RowOpCfg {
    static constexpr int TR = (B == 32) ? 64 : 128;
    static constexpr int P = B + 1;
    static constexpr int TT = (B >= 8) ? 4 : 2;
    static constexpr int NC = B / TT;
    static constexpr int NS = TR / (NC * NC);
    static constexpr size_t smem_bytes = (size_t)(3 * TR * P + 2 * B * B) * sizeof(double);
    static constexpr int CTAS_PER_SM = (B == 32) ? 3 : 4;   // B <= 16: 56 KB and <= 128 registers per thread
};

template <int B>
__global__ void __launch_bounds__(RowOpCfg<B>::TR, RowOpCfg<B>::CTAS_PER_SM) rowop_kernel(RowOpArgs a) {
    using C = RowOpCfg<B>;
    constexpr int TR = C::TR, P = C::P, TT = C::TT, NC = C::NC, NS = C::NS;
    if (a.skip_flag != nullptr && *a.skip_flag == 0) return;
    extern __shared__ double smem[];
    double* sY = smem;
    double* sX = sY + TR * P;
    double* sZ = sX + TR * P;
    double* sM1 = sZ + TR * P;
    double* sRi = sM1 + B * B;
    const int tid = threadIdx.x;

    for (int e = tid; e < B * B; e += TR) {
        if (a.m1 != nullptr) {
            const int r = e / B, c = e % B;
            sM1[e] = a.m1_transposed ? a.m1[c * B + r] : a.m1[e];
        }
        if (a.rinv != nullptr) sRi[e] = a.rinv[e];
    }

    const int combo = tid % (NC * NC);
    const int slot = tid / (NC * NC);
    const int ti = combo / NC, tj = combo % NC;
    double acc[TT][TT];
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[i][j] = 0.0;

    const int64_t ntiles = (a.n + TR - 1) / TR;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * TR;
        const int rows = (int)min((int64_t)TR, a.n - r0);
        const int nvec = rows * (B / 2);
        __syncthreads();  // previous tile fully consumed (also orders the sM1/sRi fill)
        {
            // full tiles: constant trip count, the VPT 16-byte loads of an array are issued before the first
            // shared-memory store - bytes in flight, not a loop-carried load->store dependency, set the bandwidth
            constexpr int VPT = B / 2;  // 16-byte vectors per thread and array
            const double2* gy = reinterpret_cast<const double2*>(a.y + (size_t)r0 * B);
            const double2* gx = reinterpret_cast<const double2*>(a.x1 + (size_t)r0 * B);
            const double2* gz = reinterpret_cast<const double2*>(a.gram_z + (size_t)r0 * B);
            auto put = [&](double* dst, int idx, const double2& v) {
                const int r = idx / (B / 2), c = (idx % (B / 2)) * 2;
                dst[r * P + c] = v.x;
                dst[r * P + c + 1] = v.y;
            };
            if (rows == TR) {
                auto stage = [&](double* dst, const double2* src) {
                    double2 v[VPT];
#pragma unroll
                    for (int u = 0; u < VPT; ++u) v[u] = __ldg(src + tid + u * TR);
#pragma unroll
                    for (int u = 0; u < VPT; ++u) put(dst, tid + u * TR, v[u]);
                };
                {
                    double2 v[VPT];
#pragma unroll
                    for (int u = 0; u < VPT; ++u) v[u] = gy[tid + u * TR];
#pragma unroll
                    for (int u = 0; u < VPT; ++u) put(sY, tid + u * TR, v[u]);
                }
                if (a.x1 != nullptr) stage(sX, gx);
                if (a.gram_z != nullptr) stage(sZ, gz);
            } else {
                for (int idx = tid; idx < nvec; idx += TR) put(sY, idx, gy[idx]);
                if (a.x1 != nullptr)
                    for (int idx = tid; idx < nvec; idx += TR) put(sX, idx, __ldg(gx + idx));
                if (a.gram_z != nullptr)
                    for (int idx = tid; idx < nvec; idx += TR) put(sZ, idx, __ldg(gz + idx));
            }
        }
        __syncthreads();
        if (tid < rows && (a.x1 != nullptr || a.rinv != nullptr)) {
            double y[B];
#pragma unroll
            for (int j = 0; j < B; ++j) y[j] = sY[tid * P + j];
            if (a.x1 != nullptr) {
#pragma unroll 4
                for (int c = 0; c < B; ++c) {
                    const double xv = sX[tid * P + c];
#pragma unroll
                    for (int j = 0; j < B; ++j) y[j] = fma(-xv, sM1[c * B + j], y[j]);
                }
            }
            if (a.rinv != nullptr) {
#pragma unroll
                for (int j = B - 1; j >= 0; --j) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c <= j; ++c) s = fma(y[c], sRi[c * B + j], s);
                    y[j] = s;
                }
            }
#pragma unroll
            for (int j = 0; j < B; ++j) sY[tid * P + j] = y[j];
        }
        __syncthreads();
        if (a.do_gram) {
            const double* zt = (a.gram_z != nullptr) ? sZ : sY;
            for (int r = slot; r < rows; r += NS) {
                double zr[TT], yr[TT];
#pragma unroll
                for (int i = 0; i < TT; ++i) zr[i] = zt[r * P + ti * TT + i];
#pragma unroll
                for (int j = 0; j < TT; ++j) yr[j] = sY[r * P + tj * TT + j];
#pragma unroll
                for (int i = 0; i < TT; ++i)
#pragma unroll
                    for (int j = 0; j < TT; ++j) acc[i][j] = fma(zr[i], yr[j], acc[i][j]);
            }
        }
        if (a.write_y) {
            double2* gy = reinterpret_cast<double2*>(a.y + (size_t)r0 * B);
            for (int idx = tid; idx < nvec; idx += TR) {
                const int r = idx / (B / 2), c = (idx % (B / 2)) * 2;
                gy[idx] = make_double2(sY[r * P + c], sY[r * P + c + 1]);
            }
        }
        if (a.store != nullptr) {
            if (a.store_fp32 && a.store_split_scale != 0.f) {
                unsigned* gw = reinterpret_cast<unsigned*>(a.store) + (size_t)r0 * B;
                for (int idx = tid; idx < nvec; idx += TR) {
                    const int r = idx / (B / 2), p = idx % (B / 2);
                    unsigned hi, lo;
                    split_h2((float)sY[r * P + 2 * p], (float)sY[r * P + 2 * p + 1], a.store_split_scale, hi, lo);
                    gw[r * B + p] = hi;
                    gw[r * B + B / 2 + p] = lo;
                }
            } else if (a.store_fp32) {
                float2* gs = reinterpret_cast<float2*>(reinterpret_cast<float*>(a.store) + (size_t)r0 * B);
                for (int idx = tid; idx < nvec; idx += TR) {
                    const int r = idx / (B / 2), c = (idx % (B / 2)) * 2;
                    gs[idx] = make_float2((float)sY[r * P + c], (float)sY[r * P + c + 1]);
                }
            } else {
                double2* gs = reinterpret_cast<double2*>(reinterpret_cast<double*>(a.store) + (size_t)r0 * B);
                for (int idx = tid; idx < nvec; idx += TR) {
                    const int r = idx / (B / 2), c = (idx % (B / 2)) * 2;
                    gs[idx] = make_double2(sY[r * P + c], sY[r * P + c + 1]);
                }
            }
        }
    }
    if (a.do_gram && a.partials != nullptr) {
        __syncthreads();
        double* red = smem;  // NS * B*B doubles, aliases the tiles (all reads are done)
#pragma unroll
        for (int i = 0; i < TT; ++i)
#pragma unroll
            for (int j = 0; j < TT; ++j) red[(size_t)slot * B * B + (ti * TT + i) * B + (tj * TT + j)] = acc[i][j];
        __syncthreads();
        for (int e = tid; e < B * B; e += TR) {
            double s = 0.0;
            for (int k = 0; k < NS; ++k) s += red[(size_t)k * B * B + e];
            a.partials[(size_t)blockIdx.x * B * B + e] = s;
        }
    }
}

int rowop_grid(int B, int64_t n) {
    int TR = (B == 32) ? 64 : 128;
    int64_t ntiles = (n + TR - 1) / TR;
    int64_t cap = (int64_t)num_sms() * (B == 32 ? 3 : 4);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, cap));
}

void launch_rowop(int B, const RowOpArgs& a, int grid, cudaStream_t st) {
    dispatch_B(B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        using C = RowOpCfg<BB>;
        static PerDeviceOnce once;
        if (once.first()) cudaFuncSetAttribute(rowop_kernel<BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        rowop_kernel<BB><<<grid, C::TR, C::smem_bytes, st>>>(a);
    });
}

// one warp per output element: lanes stride over the partials, fixed-order shuffle tree (deterministic)
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nparts, int count,
                                       double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= count) return;
    double s = 0.0;
    for (int p = lane; p < nparts; p += 32) s += partials[(size_t)p * count + warp];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out[warp] = s;
}

void launch_reduce_partials(const double* partials, int nparts, int count, double* out, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (count * 32 + threads - 1) / threads;
    reduce_partials_kernel<<<blocks, threads, 0, st>>>(partials, nparts, count, out);
}

// =================================================================================================
// Block-QR small step: Cholesky of the B x B Gram matrix, explicit inverse of the triangular factor,
// accumulation of R across passes, deflation bookkeeping.  One CTA of B*B threads.
// =================================================================================================
template <int B>
__global__ void __launch_bounds__(B* B) chol_kernel(const double* __restrict__ G, QrState* st, int pass,
                                                     double nrows_global, int reset_ref, double defl_rel,
                                                     const double* __restrict__ overlap) {
    __shared__ double A[B][B + 1];
    __shared__ double Rm[B][B + 1];
    __shared__ double Ro[B][B + 1];
    __shared__ double diag0[B];
    __shared__ int defl[B];
    __shared__ double s_shift, s_ref, s_minratio, s_piv;
    __shared__ int s_defl_now, s_bad;
    const int tid = threadIdx.x;
    const int i = tid / B, j = tid % B;
    if (pass >= 3 && st->need_more == 0) return;  // optional third pass not requested

    {
        const double gij = G[i * B + j], gji = G[j * B + i];
        A[i][j] = 0.5 * (gij + gji);
        Rm[i][j] = 0.0;
        Ro[i][j] = (pass == 1) ? 0.0 : st->R[i * B + j];
        if (tid == 0) s_bad = 0;
    }
    __syncthreads();
    if (i == j) {
        diag0[i] = A[i][i];
        defl[i] = (pass == 1) ? 0 : st->deflated[i];
        if (!isfinite(A[i][i])) s_bad = 1;
    }
    __syncthreads();
    if (tid == 0) {
        double tr = 0.0, mx = 0.0;
        for (int c = 0; c < B; ++c) {
            tr += diag0[c];
            mx = fmax(mx, diag0[c]);
        }
        double ref = (pass == 1) ? (reset_ref ? 0.0 : st->ref) : st->ref;
        if (pass == 1) ref = fmax(ref, sqrt(fmax(mx, 0.0)));
        s_ref = ref;
        // Fukaya et al. shifted CholeskyQR: s = 11 (m n + n (n+1)) u ||X||_2^2, ||X||_2^2 <= trace(G)
        s_shift = (pass == 1) ? 11.0 * (nrows_global * B + (double)B * (B + 1)) * 1.1102230246251565e-16 * tr : 0.0;
        s_minratio = 1.0;
    }
    __syncthreads();
    if (pass == 1 && i == j) A[i][i] += s_shift;
    __syncthreads();

    for (int jj = 0; jj < B; ++jj) {
        if (tid == 0) {
            double d = A[jj][jj];
            int dn = defl[jj];
            if (pass == 1) {
                if (!(diag0[jj] > 0.0)) dn = 1;              // exactly zero (or non-finite) column
                else if (!(d > 0.0)) d = fmax(s_shift, 1e-300);
            } else if (!dn) {
                if (!(d > 1e-14 * diag0[jj])) dn = 1;         // dependent even after scaling
                else if (sqrt(d) * fabs(Ro[jj][jj]) <= defl_rel * s_ref) dn = 1;  // below the noise floor of the run
                else s_minratio = fmin(s_minratio, d / diag0[jj]);
            }
            s_defl_now = dn;
            defl[jj] = dn;
            s_piv = dn ? 1.0 : sqrt(d);
        }
        __syncthreads();
        if (i == jj) {
            if (s_defl_now) Rm[jj][j] = (j == jj) ? 1.0 : 0.0;
            else Rm[jj][j] = (j < jj) ? 0.0 : ((j == jj) ? s_piv : A[jj][j] / s_piv);
        }
        __syncthreads();
        if (i > jj && j > jj) A[i][j] -= Rm[jj][i] * Rm[jj][j];
        __syncthreads();
    }

    // Rinv of the upper triangular Rm: one thread per column
    __shared__ double Ri[B][B + 1];
    Ri[i][j] = 0.0;
    __syncthreads();
    if (tid < B) {
        const int c = tid;
        Ri[c][c] = 1.0 / Rm[c][c];
        for (int r = c - 1; r >= 0; --r) {
            double s = 0.0;
            for (int k = r + 1; k <= c; ++k) s += Rm[r][k] * Ri[k][c];
            Ri[r][c] = -s / Rm[r][r];
        }
    }
    __syncthreads();
    // deflated columns produce an exactly zero Q column
    const double rinv_ij = (defl[j] || defl[i]) ? 0.0 : Ri[i][j];
    st->Rinv[i * B + j] = rinv_ij;
    if (overlap != nullptr) {
        // coefficients of the fused local re-orthogonalisation: (Q_i' U) * Rinv, U the block the Gram was taken of
        __syncthreads();
        Ri[i][j] = rinv_ij;
        __syncthreads();
        double s = 0.0;
        for (int k = 0; k <= j; ++k) s += overlap[i * B + k] * Ri[k][j];
        st->Mloc[i * B + j] = s;
    }
    // accumulated R = Rm * Ro  (pass 1: Rm); rows of deflated columns are zero
    double racc;
    if (pass == 1) {
        racc = Rm[i][j];
    } else {
        racc = 0.0;
        for (int k = i; k <= j; ++k) racc += Rm[i][k] * Ro[k][j];
    }
    if (defl[i] || j < i) racc = 0.0;
    st->R[i * B + j] = racc;
    if (i == j) st->deflated[i] = defl[i];
    if (tid == 0) {
        st->ref = s_ref;
        if (pass == 1) st->need_more = 0;
        if (pass == 2) st->need_more = (s_minratio < 1e-2) ? 1 : 0;
        int nd = 0;
        for (int c = 0; c < B; ++c) nd += defl[c];
        st->ndeflated = nd;
        if (s_bad) st->bad = 1;
    }
}

void launch_chol(int B, const double* G, QrState* st, int pass, int64_t nrows_global, int reset_ref, double defl_rel,
                 cudaStream_t stream, const double* overlap) {
    dispatch_B(B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        chol_kernel<BB><<<1, BB * BB, 0, stream>>>(G, st, pass, (double)nrows_global, reset_ref, defl_rel, overlap);
    });
}

// =================================================================================================
// K5a: streaming tall-skinny Gram against the Krylov buffer (SIMT version).
// grid = (chunks of JT stored blocks) x (row ranges).  A thread owns a TC x 8 tile of the
// (B x 2B) coefficient block of ONE stored block and walks all rows of the CTA's range: buffer values
// come straight from global memory (each stored element is read exactly once from HBM), the 2B target
// columns of the two active blocks are staged per 32-row tile in shared memory (converted to the
// buffer's arithmetic type).
// =================================================================================================
template <int B, typename S>
struct GramCfg {
    static constexpr int TC = (B == 4) ? 4 : ((sizeof(S) == 4) ? 8 : 4);
    static constexpr int TTG = 8;
    static constexpr int CG = B / TC;
    static constexpr int TG = (2 * B) / TTG;
    static constexpr int TPB = CG * TG;
    static constexpr int JT = 256 / TPB;
    static constexpr int RT = 32;
};

template <typename S, int N>
__device__ __forceinline__ void load_vec(const S* __restrict__ p, S (&v)[N]) {
    if constexpr (sizeof(S) == 4) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(p) + q);
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    }
}

template <typename S, int N>
__device__ __forceinline__ void load_vec_smem(const S* p, S (&v)[N]) {
    if constexpr (sizeof(S) == 4) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 t = *(reinterpret_cast<const float4*>(p) + q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const double2 t = *(reinterpret_cast<const double2*>(p) + q);
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    }
}

template <int B, typename S>
__global__ void __launch_bounds__(256) reorth_gram_kernel(int64_t n, int64_t m, const S* __restrict__ buf,
                                                          int64_t bstride, const double* __restrict__ w0,
                                                          const double* __restrict__ w1, S* __restrict__ partials,
                                                          int64_t rows_per_range) {
    using C = GramCfg<B, S>;
    constexpr int TC = C::TC, TTG = C::TTG, TG = C::TG, TPB = C::TPB, JT = C::JT, RT = C::RT;
    __shared__ __align__(16) S sW[RT * 2 * B];
    const int tid = threadIdx.x;
    const int jj = tid / TPB;
    const int sub = tid % TPB;
    const int cg = sub / TG, tg = sub % TG;
    const int64_t j = (int64_t)blockIdx.x * JT + jj;
    const bool active = j < m;
    const S* __restrict__ bj = buf + (size_t)(active ? j : 0) * bstride + cg * TC;

    S acc[TC][TTG];
#pragma unroll
    for (int c = 0; c < TC; ++c)
#pragma unroll
        for (int t = 0; t < TTG; ++t) acc[c][t] = S(0);

    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_range;
    const int64_t rend = min(n, rbeg + rows_per_range);
    for (int64_t r0 = rbeg; r0 < rend; r0 += RT) {
        const int rows = (int)min((int64_t)RT, rend - r0);
        __syncthreads();
        for (int e = tid; e < RT * 2 * B; e += 256) {
            const int r = e / (2 * B), t = e % (2 * B);
            double v = 0.0;
            if (r < rows) v = (t < B) ? __ldg(w0 + (size_t)(r0 + r) * B + t) : __ldg(w1 + (size_t)(r0 + r) * B + (t - B));
            sW[e] = (S)v;
        }
        __syncthreads();
        if (active) {
            const S* __restrict__ bp = bj + (size_t)r0 * B;
            int r = 0;
            for (; r + 4 <= rows; r += 4) {
                S b0[TC], b1[TC], b2[TC], b3[TC];
                load_vec<S, TC>(bp + (size_t)(r + 0) * B, b0);
                load_vec<S, TC>(bp + (size_t)(r + 1) * B, b1);
                load_vec<S, TC>(bp + (size_t)(r + 2) * B, b2);
                load_vec<S, TC>(bp + (size_t)(r + 3) * B, b3);
                S wv[TTG];
                load_vec_smem<S, TTG>(sW + (r + 0) * 2 * B + tg * TTG, wv);
#pragma unroll
                for (int c = 0; c < TC; ++c)
#pragma unroll
                    for (int t = 0; t < TTG; ++t) acc[c][t] = fma(b0[c], wv[t], acc[c][t]);
                load_vec_smem<S, TTG>(sW + (r + 1) * 2 * B + tg * TTG, wv);
#pragma unroll
                for (int c = 0; c < TC; ++c)
#pragma unroll
                    for (int t = 0; t < TTG; ++t) acc[c][t] = fma(b1[c], wv[t], acc[c][t]);
                load_vec_smem<S, TTG>(sW + (r + 2) * 2 * B + tg * TTG, wv);
#pragma unroll
                for (int c = 0; c < TC; ++c)
#pragma unroll
                    for (int t = 0; t < TTG; ++t) acc[c][t] = fma(b2[c], wv[t], acc[c][t]);
                load_vec_smem<S, TTG>(sW + (r + 3) * 2 * B + tg * TTG, wv);
#pragma unroll
                for (int c = 0; c < TC; ++c)
#pragma unroll
                    for (int t = 0; t < TTG; ++t) acc[c][t] = fma(b3[c], wv[t], acc[c][t]);
            }
            for (; r < rows; ++r) {
                S b0[TC], wv[TTG];
                load_vec<S, TC>(bp + (size_t)r * B, b0);
                load_vec_smem<S, TTG>(sW + r * 2 * B + tg * TTG, wv);
#pragma unroll
                for (int c = 0; c < TC; ++c)
#pragma unroll
                    for (int t = 0; t < TTG; ++t) acc[c][t] = fma(b0[c], wv[t], acc[c][t]);
            }
        }
    }
    if (active) {
        S* out = partials + ((size_t)blockIdx.y * m * B + (size_t)j * B + cg * TC) * (2 * B) + tg * TTG;
#pragma unroll
        for (int c = 0; c < TC; ++c)
#pragma unroll
            for (int t = 0; t < TTG; ++t) out[(size_t)c * 2 * B + t] = acc[c][t];
    }
}

template <typename S>
__global__ void reorth_reduce_kernel(const S* __restrict__ partials, int ranges, size_t count, S* __restrict__ Cout) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    // accumulate the per-range partial sums in double: the ranges are few (<= ~600) and this is tiny
    double s = 0.0;
    for (int p = 0; p < ranges; ++p) s += (double)partials[(size_t)p * count + e];
    Cout[e] = (S)s;
}

static int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

ReorthPlan reorth_plan(int B, int fp32, int64_t n, int64_t m) {
    ReorthPlan p;
    p.B = B; p.fp32 = fp32; p.n = n; p.m = m;
    int tc = (B == 4) ? 4 : (fp32 ? 8 : 4);
    int tpb = (B / tc) * ((2 * B) / 8);
    p.jt = 256 / tpb;
    p.chunks = (int)((m + p.jt - 1) / p.jt);
    if (p.chunks < 1) p.chunks = 1;
    int64_t target_ctas = (int64_t)num_sms() * 16;  // >= 8 waves of 1-CTA/SM kernels: short tail
    int64_t ranges = (target_ctas + p.chunks - 1) / p.chunks;
    int64_t max_ranges = std::max<int64_t>(1, n / 256);
    static const int64_t rows_override = [] { const char* e = std::getenv("RBL_GRAM_ROWS"); return e ? std::atoll(e) : 0ll; }();
    if (rows_override > 0) {   // experiment knob: rows per range (accumulation chain length of the Gram kernels)
        ranges = (n + rows_override - 1) / rows_override;
        max_ranges = std::max<int64_t>(1, n / 64);
    }
    ranges = std::max<int64_t>(1, std::min<int64_t>(ranges, max_ranges));
    ranges = std::min<int64_t>(ranges, 65535);
    p.ranges = (int)ranges;
    p.partial_elems = (size_t)p.ranges * (size_t)m * B * 2 * B;
    return p;
}

size_t reorth_max_partial_elems(int B, int fp32, int64_t n, int64_t m_cap) {
    size_t mx = 0;
    // partial_elems is piecewise monotone in m; scan the chunk boundaries
    ReorthPlan p0 = reorth_plan(B, fp32, n, 1);
    int jt = p0.jt;
    for (int64_t m = 1; m <= m_cap; m = (m % jt == 0 ? m + 1 : std::min<int64_t>(round_up(m, jt), m_cap))) {
        ReorthPlan p = reorth_plan(B, fp32, n, m);
        mx = std::max(mx, p.partial_elems);
        if (m == m_cap) break;
    }
    return mx;
}

template <int B, typename S>
static void gram_launch_t(const ReorthPlan& p, const void* buf, int64_t bstride, const double* w0, const double* w1,
                          void* partials, cudaStream_t st) {
    int64_t rpr = round_up((p.n + p.ranges - 1) / p.ranges, 32);
    dim3 grid(p.chunks, p.ranges);
    reorth_gram_kernel<B, S><<<grid, 256, 0, st>>>(p.n, p.m, (const S*)buf, bstride, w0, w1, (S*)partials, rpr);
}

void launch_reorth_gram(const ReorthPlan& p, const void* buf, int64_t bstride, const double* w0, const double* w1,
                        void* partials, void* Cmat, cudaStream_t st) {
    if (p.m <= 0) return;
    dispatch_B(p.B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        if (p.fp32) gram_launch_t<BB, float>(p, buf, bstride, w0, w1, partials, st);
        else gram_launch_t<BB, double>(p, buf, bstride, w0, w1, partials, st);
    });
    launch_reorth_gram_reduce(p, partials, Cmat, st);
}

void launch_reorth_gram_reduce(const ReorthPlan& p, const void* partials, void* Cmat, cudaStream_t st) {
    size_t count = (size_t)p.m * p.B * 2 * p.B;
    int threads = 256;
    unsigned blocks = (unsigned)((count + threads - 1) / threads);
    if (p.fp32) reorth_reduce_kernel<float><<<blocks, threads, 0, st>>>((const float*)partials, p.ranges, count, (float*)Cmat);
    else reorth_reduce_kernel<double><<<blocks, threads, 0, st>>>((const double*)partials, p.ranges, count, (double*)Cmat);
}

// =================================================================================================
// K5b: streaming update  W -= Qbuf * C  (SIMT version).  B/4 adjacent lanes share a row, each lane owns
// 4 of the B stored columns and accumulates their contribution to all 2B targets; the coefficient
// blocks are staged through shared memory in chunks (skewed by 16 B per lane group: conflict-free
// 128-bit reads); the lanes of a row are summed with shuffles at the very end.
// =================================================================================================
template <int B, typename S>
struct UpdCfg {
    static constexpr int LPR = B / 4;
    static constexpr int RPT = (sizeof(S) == 4 && B <= 16) ? 2 : 1;
    static constexpr int ROWS_W = 32 / LPR;                 // row lanes per warp
    static constexpr int ROWS_CTA = 8 * ROWS_W * RPT;
    static constexpr int SK = 16 / sizeof(S);                // skew elements
    static constexpr int BLK = B * 2 * B + LPR * SK;         // smem elements per coefficient block
    static constexpr int JC_RAW = (32 * 1024) / (BLK * (int)sizeof(S));
    static constexpr int JC = JC_RAW < 1 ? 1 : (JC_RAW > 64 ? 64 : JC_RAW);
};

template <int B, typename S>
__global__ void __launch_bounds__(256) reorth_update_kernel(int64_t n, int64_t m, const S* __restrict__ buf,
                                                            int64_t bstride, const S* __restrict__ Cmat,
                                                            double* __restrict__ w0, double* __restrict__ w1,
                                                            S* __restrict__ store_w1, S* __restrict__ store_w0) {
    using C = UpdCfg<B, S>;
    constexpr int LPR = C::LPR, RPT = C::RPT, ROWS_W = C::ROWS_W, ROWS_CTA = C::ROWS_CTA, SK = C::SK, BLK = C::BLK,
                  JC = C::JC;
    constexpr int VW = 16 / sizeof(S);  // elements per 16-byte vector
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S* sC = reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int rl = lane / LPR, q = lane % LPR;
    const int64_t row_a = (int64_t)blockIdx.x * ROWS_CTA + warp * (ROWS_W * RPT) + rl;
    int64_t rows[RPT];
    bool valid[RPT];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        rows[u] = row_a + u * ROWS_W;
        valid[u] = rows[u] < n;
    }
    S acc[RPT][2 * B];
#pragma unroll
    for (int u = 0; u < RPT; ++u)
#pragma unroll
        for (int t = 0; t < 2 * B; ++t) acc[u][t] = S(0);

    for (int64_t j0 = 0; j0 < m; j0 += JC) {
        const int jn = (int)min((int64_t)JC, m - j0);
        __syncthreads();
        // stage coefficient blocks j0 .. j0+jn: global row-major (B x 2B) -> skewed smem
        {
            const int vec_per_block = B * 2 * B / VW;
            for (int e = tid; e < jn * vec_per_block; e += 256) {
                const int jb = e / vec_per_block, v = e % vec_per_block;
                const int c = (v * VW) / (2 * B), t = (v * VW) % (2 * B);
                const S* src = Cmat + ((size_t)(j0 + jb) * B + c) * (2 * B) + t;
                S* dst = sC + (size_t)jb * BLK + c * 2 * B + (c / 4) * SK + t;
                if constexpr (sizeof(S) == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
                else *reinterpret_cast<double2*>(dst) = __ldg(reinterpret_cast<const double2*>(src));
            }
        }
        __syncthreads();
        S bv[RPT][4], bn[RPT][4];
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            if (valid[u]) load_vec<S, 4>(buf + (size_t)j0 * bstride + (size_t)rows[u] * B + q * 4, bv[u]);
            else { bv[u][0] = bv[u][1] = bv[u][2] = bv[u][3] = S(0); }
        }
        for (int jb = 0; jb < jn; ++jb) {
            // prefetch next block's buffer values
            if (jb + 1 < jn) {
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    if (valid[u]) load_vec<S, 4>(buf + (size_t)(j0 + jb + 1) * bstride + (size_t)rows[u] * B + q * 4, bn[u]);
                    else { bn[u][0] = bn[u][1] = bn[u][2] = bn[u][3] = S(0); }
                }
            }
            const S* cb = sC + (size_t)jb * BLK + (q * 4) * 2 * B + q * SK;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                for (int t0 = 0; t0 < 2 * B; t0 += VW) {
                    S cv[VW];
                    load_vec_smem<S, VW>(cb + cc * 2 * B + t0, cv);
#pragma unroll
                    for (int u = 0; u < RPT; ++u)
#pragma unroll
                        for (int x = 0; x < VW; ++x) acc[u][t0 + x] = fma(bv[u][cc], cv[x], acc[u][t0 + x]);
                }
            }
            if (jb + 1 < jn) {
#pragma unroll
                for (int u = 0; u < RPT; ++u)
#pragma unroll
                    for (int x = 0; x < 4; ++x) bv[u][x] = bn[u][x];
            }
        }
    }
    // sum the LPR lanes of a row
#pragma unroll
    for (int u = 0; u < RPT; ++u)
#pragma unroll
        for (int t = 0; t < 2 * B; ++t) {
            S v = acc[u][t];
#pragma unroll
            for (int off = 1; off < LPR; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            acc[u][t] = v;
        }
    // lane q writes targets [8q, 8q+8)
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        if (!valid[u]) continue;
        const size_t ro = (size_t)rows[u] * B;
#pragma unroll
        for (int t = 0; t < 2 * B; ++t) {
            if ((t >> 3) == q) {
                if (t < B) {
                    const double nv = w0[ro + t] - (double)acc[u][t];
                    w0[ro + t] = nv;
                    if (store_w0 != nullptr) store_w0[ro + t] = (S)nv;
                } else {
                    const double nv = w1[ro + (t - B)] - (double)acc[u][t];
                    w1[ro + (t - B)] = nv;
                    if (store_w1 != nullptr) store_w1[ro + (t - B)] = (S)nv;
                }
            }
        }
    }
}

template <int B, typename S>
static void update_launch_t(const ReorthPlan& p, const void* buf, int64_t bstride, const void* Cmat, double* w0,
                            double* w1, void* store_w1, void* store_w0, cudaStream_t st) {
    using C = UpdCfg<B, S>;
    const size_t smem = (size_t)C::JC * C::BLK * sizeof(S);
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(reorth_update_kernel<B, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned grid = (unsigned)((p.n + C::ROWS_CTA - 1) / C::ROWS_CTA);
    reorth_update_kernel<B, S><<<grid, 256, smem, st>>>(p.n, p.m, (const S*)buf, bstride, (const S*)Cmat, w0, w1,
                                                         (S*)store_w1, (S*)store_w0);
}

void launch_reorth_update(const ReorthPlan& p, const void* buf, int64_t bstride, const void* Cmat, double* w0,
                          double* w1, void* store_w1, cudaStream_t st, void* store_w0) {
    if (p.m <= 0) return;
    dispatch_B(p.B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        if (p.fp32) update_launch_t<BB, float>(p, buf, bstride, Cmat, w0, w1, store_w1, store_w0, st);
        else update_launch_t<BB, double>(p, buf, bstride, Cmat, w0, w1, store_w1, store_w0, st);
    });
}

// =================================================================================================
// K6: Ritz vectors  V = Qbuf * S  (SIMT version; ~1% of a solve).  A thread owns RPT rows x 16 targets;
// consecutive lanes are consecutive rows so that the column-major V stores coalesce.
// =================================================================================================
template <int B, typename S, typename VT, bool SPLIT>
__global__ void __launch_bounds__(256) ritz_kernel(int64_t n, int64_t m, int k, int kpad, const S* __restrict__ buf,
                                                   int64_t bstride, const S* __restrict__ Smat, VT* __restrict__ V,
                                                   int64_t ldv, int RL, int JC, float split_inv_scale, int accumulate) {
    constexpr int RPT = 2;
    constexpr int TT = 16;
    constexpr int VW = 16 / sizeof(S);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S* sS = reinterpret_cast<S*>(smem_raw);  // [JC][B][kpad]
    const int tid = threadIdx.x;
    const int KG = kpad / TT;
    const int rl = tid % RL, tg = tid / RL;
    const bool worker = tg < KG;
    int64_t rows[RPT];
    bool valid[RPT];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        rows[u] = (int64_t)blockIdx.x * RL * RPT + u * RL + rl;
        valid[u] = worker && rows[u] < n;
    }
    // fp32 buffers: every shared-memory chunk (JC stored blocks) is accumulated in fp32 and the chunk sums in
    // fp64 - a single fp32 accumulator over all m*B terms loses ~sqrt(m*B) ulps and was measured to push the Ritz
    // residual of a b = 32 problem to 1.3e-6 ||A|| (bar: 1e-6)
    double dacc[RPT][TT];
    S acc[RPT][TT];
#pragma unroll
    for (int u = 0; u < RPT; ++u)
#pragma unroll
        for (int t = 0; t < TT; ++t) {
            acc[u][t] = S(0);
            dacc[u][t] = 0.0;
        }

    for (int64_t j0 = 0; j0 < m; j0 += JC) {
        const int jn = (int)min((int64_t)JC, m - j0);
        __syncthreads();
        {
            const int vecs = jn * B * kpad / VW;
            const S* src = Smat + (size_t)j0 * B * kpad;
            for (int e = tid; e < vecs; e += 256) {
                if constexpr (sizeof(S) == 4) reinterpret_cast<float4*>(sS)[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
                else reinterpret_cast<double2*>(sS)[e] = __ldg(reinterpret_cast<const double2*>(src) + e);
            }
        }
        __syncthreads();
        if (!worker) continue;
        for (int jb = 0; jb < jn; ++jb) {
            S bv[RPT][B];
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                if (valid[u]) {
                    load_vec<S, B>(buf + (size_t)(j0 + jb) * bstride + (size_t)rows[u] * B, bv[u]);
                    if constexpr (SPLIT && sizeof(S) == 4) {  // split16 row: words [hi pairs | lo pairs]
                        float dec[B];
#pragma unroll
                        for (int p = 0; p < B / 2; ++p) {
                            const float2 x = join_h2(__float_as_uint(bv[u][p]), __float_as_uint(bv[u][B / 2 + p]), split_inv_scale);
                            dec[2 * p] = x.x;
                            dec[2 * p + 1] = x.y;
                        }
#pragma unroll
                        for (int c = 0; c < B; ++c) bv[u][c] = dec[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < B; ++c) bv[u][c] = S(0);
                }
            }
            const S* sb = sS + (size_t)jb * B * kpad + tg * TT;
#pragma unroll
            for (int c = 0; c < B; ++c) {
                S sv[TT];
                load_vec_smem<S, TT>(sb + c * kpad, sv);
#pragma unroll
                for (int u = 0; u < RPT; ++u)
#pragma unroll
                    for (int t = 0; t < TT; ++t) acc[u][t] = fma(bv[u][c], sv[t], acc[u][t]);
            }
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u)
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                dacc[u][t] += (double)acc[u][t];
                acc[u][t] = S(0);
            }
    }
    if (!worker) return;
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        if (!valid[u]) continue;
#pragma unroll
        for (int t = 0; t < TT; ++t) {
            const int col = tg * TT + t;
            if (col < k) {
                VT* dst = V + (size_t)col * ldv + rows[u];
                *dst = accumulate ? (VT)((double)*dst + dacc[u][t]) : (VT)dacc[u][t];
            }
        }
    }
}

template <int B, typename S, typename VT, bool SPLIT = false>
static void ritz_launch_t(int64_t n, int64_t m, int k, int kpad, const void* buf, int64_t bstride, const void* Smat,
                          void* V, int64_t ldv, cudaStream_t st, int accumulate, float split_inv_scale = 0.f) {
    const int KG = kpad / 16;
    int RL = 256 / KG;
    if (RL > 32) RL = (RL / 32) * 32;   // whole warps share a target group (broadcast smem reads)
    if (RL < 1) RL = 1;
    int JC = (int)((48 * 1024) / ((size_t)B * kpad * sizeof(S)));
    if (JC < 1) JC = 1;
    if (JC > 16) JC = 16;
    const size_t smem = (size_t)JC * B * kpad * sizeof(S);
    cudaFuncSetAttribute(ritz_kernel<B, S, VT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned grid = (unsigned)((n + (int64_t)RL * 2 - 1) / ((int64_t)RL * 2));
    ritz_kernel<B, S, VT, SPLIT><<<grid, 256, smem, st>>>(n, m, k, kpad, (const S*)buf, bstride, (const S*)Smat, (VT*)V, ldv, RL,
                                                         JC, split_inv_scale, accumulate);
}

void launch_ritz(int B, int fp32, int64_t n, int64_t m, int k, int kpad, const void* buf, int64_t bstride,
                 const void* Smat, void* V, int64_t ldv, int v_fp32, float split_scale, cudaStream_t st, int accumulate) {
    if (kpad / 16 > 256) { std::fprintf(stderr, "rbl: k too large for ritz kernel\n"); std::abort(); }
    dispatch_B(B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        if (fp32 && split_scale != 0.f) {
            if constexpr (BB >= 16) {
                const float inv = 1.0f / split_scale;
                if (v_fp32) ritz_launch_t<BB, float, float, true>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate, inv);
                else ritz_launch_t<BB, float, double, true>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate, inv);
            }
        } else if (fp32) {
            if (v_fp32) ritz_launch_t<BB, float, float>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate);
            else ritz_launch_t<BB, float, double>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate);
        } else {
            if (v_fp32) ritz_launch_t<BB, double, float>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate);
            else ritz_launch_t<BB, double, double>(n, m, k, kpad, buf, bstride, Smat, V, ldv, st, accumulate);
        }
    });
}

// =================================================================================================
// layout / conversion helpers
// =================================================================================================
__global__ void colmajor_to_block_kernel(int B, int64_t n, int b, const double* __restrict__ src, int64_t ld,
                                         double* __restrict__ dst) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int c = 0; c < B; ++c) dst[(size_t)r * B + c] = (c < b) ? src[(size_t)c * ld + r] : 0.0;
}
__global__ void block_to_colmajor_kernel(int B, int64_t n, int b, const double* __restrict__ src,
                                         double* __restrict__ dst, int64_t ld) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int c = 0; c < b; ++c) dst[(size_t)c * ld + r] = src[(size_t)r * B + c];
}
void launch_colmajor_to_block(int B, int64_t n, int b, const double* src, int64_t ld, double* dst, cudaStream_t st) {
    if (n <= 0) return;
    colmajor_to_block_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, n, b, src, ld, dst);
}
void launch_block_to_colmajor(int B, int64_t n, int b, const double* src, double* dst, int64_t ld, cudaStream_t st) {
    if (n <= 0) return;
    block_to_colmajor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, n, b, src, dst, ld);
}

template <typename S>
__global__ void convert_kernel(int64_t count, const double* __restrict__ src, S* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = (S)src[i];
}
template <typename S>
__global__ void widen_kernel(int64_t count, const S* __restrict__ src, double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = (double)src[i];
}
void launch_convert_s(int64_t count, const double* src, void* dst, int fp32, cudaStream_t st) {
    if (count <= 0) return;
    int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)num_sms() * 16);
    if (fp32) convert_kernel<float><<<grid, 256, 0, st>>>(count, src, (float*)dst);
    else convert_kernel<double><<<grid, 256, 0, st>>>(count, src, (double*)dst);
}
template <typename T>
__global__ void encode_split_kernel(int B, int64_t rows, const T* __restrict__ src, unsigned* __restrict__ dst, float scale) {
    const int hb = B / 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one column pair per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < rows * hb; i += stride) {
        const int64_t r = i / hb;
        const int p = (int)(i % hb);
        unsigned hi, lo;
        split_h2((float)src[r * B + 2 * p], (float)src[r * B + 2 * p + 1], scale, hi, lo);
        dst[r * B + p] = hi;
        dst[r * B + hb + p] = lo;
    }
}
void launch_encode_split(int B, int64_t rows, const float* src, void* dst, float scale, cudaStream_t st) {
    if (rows <= 0) return;
    int grid = (int)std::min<int64_t>((rows * (B / 2) + 255) / 256, (int64_t)num_sms() * 16);
    encode_split_kernel<float><<<grid, 256, 0, st>>>(B, rows, src, (unsigned*)dst, scale);
}
void launch_store_block(int B, int64_t n, const double* src, void* dst, int fp32, float split_scale, cudaStream_t st) {
    if (fp32 && split_scale != 0.f) {
        if (n <= 0) return;
        int grid = (int)std::min<int64_t>((n * (B / 2) + 255) / 256, (int64_t)num_sms() * 16);
        encode_split_kernel<double><<<grid, 256, 0, st>>>(B, n, src, (unsigned*)dst, split_scale);
        return;
    }
    launch_convert_s(n * B, src, dst, fp32, st);
}
void launch_load_block(int B, int64_t n, const void* src, int fp32, double* dst, cudaStream_t st) {
    int64_t count = n * B;
    if (count <= 0) return;
    int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)num_sms() * 16);
    if (fp32) widen_kernel<float><<<grid, 256, 0, st>>>(count, (const float*)src, dst);
    else widen_kernel<double><<<grid, 256, 0, st>>>(count, (const double*)src, dst);
}

__global__ void decode_split_kernel(int B, int64_t rows, const unsigned* __restrict__ src, double* __restrict__ dst,
                                    float inv_scale) {
    const int hb = B / 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one column pair per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < rows * hb; i += stride) {
        const int64_t r = i / hb;
        const int p = (int)(i % hb);
        const float2 x = join_h2(src[r * B + p], src[r * B + hb + p], inv_scale);
        dst[r * B + 2 * p] = (double)x.x;
        dst[r * B + 2 * p + 1] = (double)x.y;
    }
}
void launch_decode_block(int B, int64_t n, const void* src, int fp32, float split_scale, double* dst, cudaStream_t st) {
    if (n <= 0) return;
    if (fp32 && split_scale != 0.f) {
        int grid = (int)std::min<int64_t>((n * (B / 2) + 255) / 256, (int64_t)num_sms() * 16);
        decode_split_kernel<<<grid, 256, 0, st>>>(B, n, (const unsigned*)src, dst, 1.0f / split_scale);
        return;
    }
    launch_load_block(B, n, src, fp32, dst, st);
}

// orthogonality loss accumulation (see kernels.h); one thread per coefficient
template <typename S>
__global__ void ortho_accumulate_kernel(int B, int64_t m, int64_t j0, int ntargets, const S* __restrict__ C,
                                        double* __restrict__ out) {
    const int64_t total = m * B * 2 * B;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double mx = 0.0, ss = 0.0;
    for (; e < total; e += stride) {
        const int64_t row = e / (2 * B);
        const int t = (int)(e % (2 * B));
        if (t / B >= ntargets) continue;
        double v = (double)C[e];
        if (row == (j0 + t / B) * B + (t % B)) {
            // diagonal: a deflated / padded column is exactly zero in the slab (q'q == 0): not part of the basis
            if (v == 0.0) continue;
            v -= 1.0;
        }
        mx = fmax(mx, fabs(v));
        ss += v * v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 1, ss);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(mx));
    }
}
void launch_ortho_accumulate(int B, int64_t m, int64_t j0, int ntargets, int fp32, const void* C, double* out,
                             cudaStream_t st) {
    if (m <= 0) return;
    const int64_t total = m * B * 2 * B;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 8);
    if (fp32) ortho_accumulate_kernel<float><<<grid, 256, 0, st>>>(B, m, j0, ntargets, (const float*)C, out);
    else ortho_accumulate_kernel<double><<<grid, 256, 0, st>>>(B, m, j0, ntargets, (const double*)C, out);
}

// per-column dots of column-major matrices: grid.y = column, grid.x strides over rows
__global__ void col_dots_kernel(int64_t n, int k, const double* __restrict__ V, const double* __restrict__ W, int64_t ld,
                                double* __restrict__ out) {
    const int c = blockIdx.y;
    const double* v = V + (size_t)c * ld;
    const double* w = W + (size_t)c * ld;
    double vv = 0.0, vw = 0.0, ww = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const double a = v[r], b = w[r];
        vv = fma(a, a, vv);
        vw = fma(a, b, vw);
        ww = fma(b, b, ww);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        vv += __shfl_xor_sync(0xffffffffu, vv, off);
        vw += __shfl_xor_sync(0xffffffffu, vw, off);
        ww += __shfl_xor_sync(0xffffffffu, ww, off);
    }
    __shared__ double red[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = vv; red[1][warp] = vw; red[2][warp] = ww; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
        atomicAdd(out + (size_t)threadIdx.x * k + c, s);
    }
}
void launch_col_dots(int64_t n, int k, const double* V, const double* W, int64_t ld, double* out3k, cudaStream_t st) {
    if (n <= 0 || k <= 0) return;
    dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)k);
    col_dots_kernel<<<grid, 256, 0, st>>>(n, k, V, W, ld, out3k);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// counter-based N(0,1): element i of stream `seed` is a pure function of (seed, offset+i)   [CUDA.randn, RBL_gpu.jl:213]
__global__ void randn_kernel(int64_t count, uint64_t seed, uint64_t offset, double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        const uint64_t ctr = offset + (uint64_t)i;
        const uint64_t a = splitmix64(seed * 0xD1342543DE82EF95ull + 2 * ctr);
        const uint64_t b = splitmix64(seed * 0xD1342543DE82EF95ull + 2 * ctr + 1);
        const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        dst[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}
void launch_randn(int64_t count, uint64_t seed, uint64_t offset, double* dst, cudaStream_t st) {
    if (count <= 0) return;
    int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)num_sms() * 16);
    randn_kernel<<<grid, 256, 0, st>>>(count, seed, offset, dst);
}

template <int B>
__global__ void gather_rows_kernel(int64_t nrows, const int* __restrict__ rows, const double* __restrict__ src,
                                   double* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = idx / (B / 2);
    const int c = (int)(idx % (B / 2));
    if (r >= nrows) return;
    const int sr = rows[r];
    reinterpret_cast<double2*>(dst)[(size_t)r * (B / 2) + c] = __ldg(reinterpret_cast<const double2*>(src) + (size_t)sr * (B / 2) + c);
}
void launch_gather_rows(int B, int64_t nrows, const int* rows, const double* src, double* dst, cudaStream_t st) {
    if (nrows <= 0) return;
    dispatch_B(B, [&](auto bc) {
        constexpr int BB = decltype(bc)::value;
        int64_t total = nrows * (BB / 2);
        gather_rows_kernel<BB><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(nrows, rows, src, dst);
    });
}

}  // namespace rbl

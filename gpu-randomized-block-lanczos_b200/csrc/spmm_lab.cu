// SpMM laboratory: candidate K1 kernels timed side by side on a handle's matrix in ONE process (rbl_spmm_bench).
// Not on the solve path.  Every variant computes U = alpha*A*Q + beta*Q (+ gamma*Z) like kernels.cu spmm_kernel and is
// checked against it before it is timed.
//   0  the product's gather kernel (kernels.cu)
//   1  CSR, software-pipelined: the next row's row pointers are loaded and its CSR lines prefetched (prefetch.global.L2)
//      while the current row's Q rows are in flight
//   2  ELL (row-major, width W = max row length rounded to 4): no row pointers, 16-byte index / value loads, next row prefetched
//  +16 rows visited in patch-schedule order (spmm_sched.cu); ELL rows are stored in that order
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "solver.h"

namespace rbl {

namespace {

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

// ---- variant 1: CSR with the row-pointer / CSR-line fetch of the next row overlapped -----------------------------------
template <int B, int PF>
__global__ void __launch_bounds__(256) spmm_csr_pipe_kernel(int64_t nslots, int64_t nrows, const int* __restrict__ order,
                                                            const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                            const double* __restrict__ vals, const double* __restrict__ Q, double* U,
                                                            SpmmCoef cf, const double* Z) {
    constexpr int LPR = B / 2;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int rsel = lane / LPR;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double2* __restrict__ Q2 = reinterpret_cast<const double2*>(Q);
    auto row_of = [&](int64_t s) -> int64_t {
        if (s >= nslots) return -1;
        return order ? (int64_t)__ldg(order + s) : s;
    };
    int64_t s = warp * RPW + rsel;
    int64_t row = row_of(s);
    int p = 0, p1 = 0;
    if (row >= 0) { p = __ldg(rowptr + row); p1 = __ldg(rowptr + row + 1); }
    while (s < nslots) {     // (slots beyond the last are never reached by a whole warp pass unless all its rows are out)
        const int64_t sn = s + nwarps * RPW;
        const int64_t rown = row_of(sn);
        int pn = 0, pn1 = 0;
        if (rown >= 0) {
            pn = __ldg(rowptr + rown);
            pn1 = __ldg(rowptr + rown + 1);
        }
        if (row >= 0) {
            double2 acc = make_double2(0.0, 0.0);
            for (; p + 4 <= p1; p += 4) {
                const int c0 = __ldg(colidx + p), c1 = __ldg(colidx + p + 1), c2 = __ldg(colidx + p + 2), c3 = __ldg(colidx + p + 3);
                const double v0 = __ldg(vals + p), v1 = __ldg(vals + p + 1), v2 = __ldg(vals + p + 2), v3 = __ldg(vals + p + 3);
                const double2 q0 = __ldg(Q2 + (size_t)c0 * LPR + sub);
                const double2 q1 = __ldg(Q2 + (size_t)c1 * LPR + sub);
                const double2 q2 = __ldg(Q2 + (size_t)c2 * LPR + sub);
                const double2 q3 = __ldg(Q2 + (size_t)c3 * LPR + sub);
                acc.x = fma(v0, q0.x, acc.x); acc.y = fma(v0, q0.y, acc.y);
                acc.x = fma(v1, q1.x, acc.x); acc.y = fma(v1, q1.y, acc.y);
                acc.x = fma(v2, q2.x, acc.x); acc.y = fma(v2, q2.y, acc.y);
                acc.x = fma(v3, q3.x, acc.x); acc.y = fma(v3, q3.y, acc.y);
            }
            if (PF && rown >= 0 && sub == 0) {      // the next row's CSR lines, requested while this row finishes
                if (PF == 1) { prefetch_l2(colidx + pn); prefetch_l2(vals + pn); }
                else { prefetch_l1(colidx + pn); prefetch_l1(vals + pn); }
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
        s = sn; row = rown; p = pn; p1 = pn1;
    }
}

// ---- variant 2: ELL ------------------------------------------------------------------------------------------------------
// ecol / eval: nslots x W, row-major, in VISITING order (slot s holds row erow[s], -1 = padding slot); padding entries have
// value 0 and the row's own index as column (a line that is needed anyway).
__global__ void ell_build_kernel(int64_t nslots, int W, const int* __restrict__ order, const int* __restrict__ rowptr,
                                 const int* __restrict__ colidx, const double* __restrict__ vals, int* __restrict__ ecol,
                                 double* __restrict__ eval) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const int64_t row = order ? (int64_t)order[s] : s;
    int p = 0, p1 = 0;
    if (row >= 0) { p = rowptr[row]; p1 = rowptr[row + 1]; }
    for (int j = 0; j < W; ++j) {
        const bool on = p + j < p1;
        ecol[s * W + j] = on ? colidx[p + j] : (int)(row >= 0 ? row : 0);
        eval[s * W + j] = on ? vals[p + j] : 0.0;
    }
}

template <int B, int W, int PF>
__global__ void __launch_bounds__(256) spmm_ell_kernel(int64_t nslots, const int* __restrict__ order, const int* __restrict__ ecol,
                                                       const double* __restrict__ eval, const double* __restrict__ Q, double* U,
                                                       SpmmCoef cf, const double* Z) {
    constexpr int LPR = B / 2;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int rsel = lane / LPR;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double2* __restrict__ Q2 = reinterpret_cast<const double2*>(Q);
    for (int64_t s = warp * RPW + rsel; s < nslots; s += nwarps * RPW) {
        const int64_t row = order ? (int64_t)__ldg(order + s) : s;
        if (PF) {
            const int64_t sn = s + nwarps * RPW;
            if (sn < nslots && sub == 0) {
                prefetch_l2(ecol + sn * W);
                prefetch_l2(eval + sn * W);
                if (W * 8 > 128) prefetch_l2(eval + sn * W + 16);
            }
        }
        if (row < 0) continue;
        int c[W];
        double v[W];
        const int4* ec = reinterpret_cast<const int4*>(ecol + s * W);
        const double2* ev = reinterpret_cast<const double2*>(eval + s * W);
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
            const int4 t = __ldg(ec + j);
            c[4 * j] = t.x; c[4 * j + 1] = t.y; c[4 * j + 2] = t.z; c[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < W / 2; ++j) {
            const double2 t = __ldg(ev + j);
            v[2 * j] = t.x; v[2 * j + 1] = t.y;
        }
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int j0 = 0; j0 < W; j0 += 4) {
            double2 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = __ldg(Q2 + (size_t)c[j0 + u] * LPR + sub);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x = fma(v[j0 + u], q[u].x, acc.x);
                acc.y = fma(v[j0 + u], q[u].y, acc.y);
            }
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

__global__ void fill_kernel(double* x, int64_t n, unsigned seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        x[i] = ((double)(h & 0xffffff) / 16777216.0) - 0.5;
    }
}
__global__ void maxdiff_kernel(const double* a, const double* b, int64_t n, unsigned long long* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(out, 1ull);
}

}  // namespace

// Runs variant `variant` `iters` times (L2 flushed by a 512 MB memset before every timed launch when flush != 0) and returns
// the mean launch time in microseconds; mismatches against variant 0 (elements that differ in any bit) in *mismatch_out.
double spmm_lab_run(rbl_handle* h, int b, int variant, int grid_mult, int iters, int flush, int with_z, unsigned long long* mismatch_out) {
    const int B = padded_block(b);
    if (B != 16) return -1.0;
    const int64_t n = h->nloc;
    cudaStream_t st = h->stream;
    const int* rowptr = h->wsp->d_rowptr.p;
    const int* colidx = h->wsp->d_colidx.p;
    const double* vals = h->wsp->d_vals.p;
    const bool sched = (variant & 16) != 0 && h->spmm_sched.dims > 0;
    const int kind = variant & 15;
    if (kind == 0 && (variant & 16)) return -4.0;     // the product kernel has no scheduled form
    if ((variant & 16) && !sched) return -5.0;        // no schedule planned for this handle (RBL_SPMM_SCHED=1 at create)
    const int* order = sched ? h->wsp->d_order.p : nullptr;
    const int64_t nslots = sched ? h->spmm_sched.npatch * h->spmm_sched.slots : n;
    DevBuf<double> Q, U, Uref, Zb, flushbuf;
    Q.alloc((size_t)n * B); U.alloc((size_t)n * B); Uref.alloc((size_t)n * B); Zb.alloc((size_t)n * B);
    const size_t flush_elems = (size_t)64 << 20;
    if (flush) flushbuf.alloc(flush_elems);
    const unsigned fg = (unsigned)(((size_t)n * B + 255) / 256);
    fill_kernel<<<fg, 256, 0, st>>>(Q.p, (int64_t)n * B, 1u);
    fill_kernel<<<fg, 256, 0, st>>>(Zb.p, (int64_t)n * B, 2u);
    const SpmmCoef cf = with_z ? SpmmCoef{-0.3, 1.7, -1.0} : SpmmCoef{-1.0, 12.0, 0.0};
    const double* Z = with_z ? Zb.p : nullptr;
    // ELL copy
    DevBuf<int> ecol;
    DevBuf<double> eval;
    int W = 0;
    if (kind == 2) {
        std::vector<int> rp((size_t)n + 1);
        cudaMemcpy(rp.data(), rowptr, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost);
        int mx = 0;
        for (int64_t r = 0; r < n; ++r) mx = std::max(mx, rp[r + 1] - rp[r]);
        W = (mx + 3) & ~3;
        if (W != 8 && W != 12) return -2.0;
        ecol.alloc((size_t)nslots * W);
        eval.alloc((size_t)nslots * W);
        ell_build_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(nslots, W, order, rowptr, colidx, vals, ecol.p, eval.p);
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (nslots + 31) / 32;
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * (grid_mult > 0 ? grid_mult : 16));
    const int pf = (variant >> 8) & 3;
    auto launch = [&](double* out) {
        if (kind == 0) {
            launch_spmm(B, n, rowptr, colidx, vals, Q.p, out, cf, Z, st);
        } else if (kind == 1) {
            if (pf == 0) spmm_csr_pipe_kernel<16, 0><<<grid, 256, 0, st>>>(nslots, n, order, rowptr, colidx, vals, Q.p, out, cf, Z);
            else if (pf == 1) spmm_csr_pipe_kernel<16, 1><<<grid, 256, 0, st>>>(nslots, n, order, rowptr, colidx, vals, Q.p, out, cf, Z);
            else spmm_csr_pipe_kernel<16, 2><<<grid, 256, 0, st>>>(nslots, n, order, rowptr, colidx, vals, Q.p, out, cf, Z);
        } else {
            if (W == 8) {
                if (pf) spmm_ell_kernel<16, 8, 1><<<grid, 256, 0, st>>>(nslots, order, ecol.p, eval.p, Q.p, out, cf, Z);
                else spmm_ell_kernel<16, 8, 0><<<grid, 256, 0, st>>>(nslots, order, ecol.p, eval.p, Q.p, out, cf, Z);
            } else {
                if (pf) spmm_ell_kernel<16, 12, 1><<<grid, 256, 0, st>>>(nslots, order, ecol.p, eval.p, Q.p, out, cf, Z);
                else spmm_ell_kernel<16, 12, 0><<<grid, 256, 0, st>>>(nslots, order, ecol.p, eval.p, Q.p, out, cf, Z);
            }
        }
    };
    // reference result: the product kernel
    launch_spmm(B, n, rowptr, colidx, vals, Q.p, Uref.p, cf, Z, st);
    launch(U.p);
    DevBuf<unsigned long long> bad;
    bad.alloc(1);
    cudaMemsetAsync(bad.p, 0, 8, st);
    maxdiff_kernel<<<fg, 256, 0, st>>>(U.p, Uref.p, (int64_t)n * B, bad.p);
    unsigned long long nbad = 0;
    cudaMemcpyAsync(&nbad, bad.p, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (mismatch_out) *mismatch_out = nbad;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double total = 0.0;
    for (int it = 0; it < iters + 2; ++it) {
        if (flush) cudaMemsetAsync(flushbuf.p, it, flush_elems * 8, st);
        cudaEventRecord(e0, st);
        launch(U.p);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -3.0;
    return total * 1000.0 / iters;
}

}  // namespace rbl

// Device kernels of the RBL hot path (sm_100a) - launch interface.  See DESIGN.md section 4.
//
// HBM layout: every dense block is ROW-major [rows][B] with B in {4,8,16,32}
// (the caller's b is zero-padded up to B; padded columns behave as deflated columns).  The Krylov
// buffer is a slab of `m` such blocks, stored as float (mixed precision) or double.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

namespace rbl {

// Kernel attributes (cudaFuncSetAttribute) are per device: one of these per call site remembers which devices
// of the process have been configured.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        return !(mask.fetch_or(bit) & bit);
    }
};


inline int padded_block(int b) { return b <= 4 ? 4 : (b <= 8 ? 8 : (b <= 16 ? 16 : 32)); }

// ---- K1 block SpMM ------------------------------------------------------------------------------
// U[r,:] = alpha * sum_p vals[p] * Q[colidx[p],:] + beta * Q[r,:] + gamma * Z[r,:]
//   alpha = 1, beta = gamma = 0          replaces mul!(U,Ag,Qg_d)  RBL_gpu.jl:152,176
//   alpha = -1, beta = sigma             the shifted operator sigma*I - A
//   general                              one step of the Chebyshev three-term recurrence of the filtered operator
// Z may alias U (it is read and written row by row by the same thread); Z == nullptr when gamma == 0.
struct SpmmCoef {
    double alpha = 1.0, beta = 0.0, gamma = 0.0;
};
// rowlist: compute only the `nrows` rows listed; skip: leave the rows flagged there untouched (halo overlap, solver.cu)
void launch_spmm(int B, int64_t nrows, const int* rowptr, const int* colidx, const double* vals, const double* Q,
                 double* U, SpmmCoef cf, const double* Z, cudaStream_t st, const int* rowlist = nullptr,
                 const unsigned char* skip = nullptr);

// second generation for banded / stencil matrices (spmm.cu): Q and the CSR stream staged in shared memory by TMA bulk copies
struct SpmmWindows {
    static constexpr int kMax = 8;         // windows
    static constexpr int kTileRows = 64;   // rows per tile
    static constexpr int kMaxStages = 8;   // tiles in flight
    int nwin = 0;            // 0: no window structure (use launch_spmm)
    int lo[kMax], hi[kMax];  // offset range col - row of every window
    int base[kMax];          // first ring row of the window in shared memory (filled by spmm_window_stages)
    int diag_w = -1;         // window that holds offset 0 (-1: none)
    int max_tile_nnz = 0;    // most nonzeros in an aligned 64-row tile
};
// window table from a sample of the (host) CSR rows; columns >= nown (halo) are never part of a window
SpmmWindows spmm_plan_windows(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx);
// number of pipeline stages that fit shared memory for block size B (0: none), ring bases, total bytes
int spmm_window_stages(SpmmWindows& wt, int B, size_t* smem_bytes_out);
// rel[p] = (window << 24) | (col - row - lo[window]) or -1 - col for entries outside every window; *irregular += their count
void launch_spmm_build_rel(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx, const SpmmWindows& wt, int* rel,
                           unsigned long long* irregular, cudaStream_t st);
bool spmm_window_supported(int B);
// rowptr / rel / vals need 8 elements of slack behind their last entry (16-byte rounded bulk copies)
void launch_spmm_window(int B, int64_t nrows, int64_t nown, const int* rowptr, const int* rel, const double* vals, const double* Q,
                        double* U, SpmmCoef cf, const double* Z, const SpmmWindows& wt, cudaStream_t st);

// row schedule for stencil / banded matrices (spmm_sched.cu; used by the SpMM laboratory only): rows grouped into compact
// patches of the grid the stencil offsets imply
struct SpmmSchedule {
    int dims = 0;                 // 0: no schedule (use launch_spmm)
    int64_t stride[3] = {1, 0, 0};   // row = x + stride[1] * y + stride[2] * z
    int64_t ext[3] = {1, 1, 1};      // grid extents
    int halo[3] = {0, 0, 0};      // stencil reach per dimension
    int patch[3] = {1, 1, 1};     // patch shape (points)
    int slots = 0;                // schedule entries per patch (>= patch volume, padded with -1)
    int64_t npatch = 0;
    double fetch_model = 0.0;     // modelled rows of Q crossing L2->SM per row (x padding penalty)
};
}  // namespace rbl
#include <vector>
namespace rbl {
bool spmm_plan_schedule(int64_t nrows, int64_t nown, const int* rowptr, const int* colidx, int slots, std::vector<int>& order,
                        SpmmSchedule* info);
int spmm_sched_default_slots(int B);

// ---- K2/K3/K4 fused row-wise block operations on fp64 blocks --------------------------------------
// For every row r of Y (n x B, fp64):
//     y <- y - x1[r,:] * M1 - x2[r,:] * M2        (either may be absent; M row-major B x B, device)
//     y <- y * Rinv                               (upper triangular right-multiply, if rinv != null)
//     Y[r,:] <- y                                 (if write_y)
//     store[r,:] <- (float|double) y              (Krylov-buffer copy, if store != null)
//     G += z[r,:]' * y                            (z = gram_z if given, else y itself; if partials != null)
// G partials are written per CTA to partials[cta][B*B]; reduce with launch_reduce_partials.
struct RowOpArgs {
    int64_t n = 0;
    double* y = nullptr;
    const double* x1 = nullptr;
    const double* m1 = nullptr;
    int m1_transposed = 0;        // use M1' (U -= Q_{i-1} * B_{i-1}')          RBL_gpu.jl:177
    const double* x2 = nullptr;
    const double* m2 = nullptr;
    const double* rinv = nullptr;
    const int* skip_flag = nullptr;  // device flag: kernel is a no-op when *skip_flag == 0 (optional 3rd QR pass)
    int write_y = 0;
    void* store = nullptr;
    int store_fp32 = 0;
    float store_split_scale = 0.f;  // != 0 (with store_fp32): write the split16 row format (split16.h) with this scale
    const double* gram_z = nullptr;
    int do_gram = 0;
    double* partials = nullptr;   // [grid][B*B]
};
int rowop_grid(int B, int64_t n);
void launch_rowop(int B, const RowOpArgs& a, int grid, cudaStream_t st);
// out[e] = sum_p partials[p][e], e < count   (fixed order: deterministic)
void launch_reduce_partials(const double* partials, int nparts, int count, double* out, cudaStream_t st);

// ---- second-generation fused row kernel (rowops.cu): B = 16, 32 -----------------------------------------------------
// per row tile:  y <- y*Rinv (optional) ; y <- y - x1*M1 (optional) ; Y <- y ; store <- y ; G0 += y'y ; G1 += z'y
struct FusedArgs {
    int64_t n = 0;
    double* y = nullptr;
    const double* rinv = nullptr;     // upper triangular B x B (row-major), applied first
    const double* x1 = nullptr;
    const double* m1 = nullptr;
    int m1_transposed = 0;
    const double* z = nullptr;        // left operand of the second Gram
    int gram_yy = 0;
    const int* gram_flag = nullptr;   // device flag: y'y is accumulated only when *gram_flag != 0 (optional 3rd QR pass)
    const int* skip_flag = nullptr;   // device flag: the kernel is a no-op when *skip_flag == 0
    int write_y = 1;
    void* store = nullptr;
    int store_fp32 = 0;
    float store_split_scale = 0.f;
    double* partials = nullptr;       // [grid][2][B*B]: slot 0 = y'y, slot 1 = z'y
};
bool fused_rowop_supported(int B);
int fused_rowop_grid(int B, int64_t n);
void launch_fused_rowop(int B, const FusedArgs& a, int grid, cudaStream_t st);

// ---- block-QR small step (single CTA) -----------------------------------------------------------
struct QrState {             // lives in device memory
    double R[32 * 32];       // accumulated R (row-major B x B)
    double Rinv[32 * 32];    // inverse of the current pass' triangular factor
    double ref;              // running norm scale for the deflation test
    int deflated[32];
    int need_more;           // 1: a further re-orthogonalisation pass is required
    int ndeflated;
    int bad;                 // non-finite input seen
    double Mloc[32 * 32];    // pass 2 with an overlap matrix P = Q_i' U: P * Rinv (coefficients of the fused local reorth)
};
// pass 1: shifted Cholesky of G (+ exact-zero column deflation); pass >= 2: plain Cholesky with
// deflation of columns whose accumulated R_jj <= defl_rel * ref.  `nrows_global` enters the shift.
// `overlap` (optional, pass 2): B x B matrix P; st->Mloc = P * Rinv of this pass.
void launch_chol(int B, const double* G, QrState* st, int pass, int64_t nrows_global, int reset_ref, double defl_rel,
                 cudaStream_t stream, const double* overlap = nullptr);

// ---- K5 partial (full) re-orthogonalisation against the Krylov buffer ----------------------------
// gram:   C[(j*B+c), t] = sum_r buf_j[r,c] * W[r,t],  W = [w0 | w1] (fp64 active blocks), t < 2B
// update: w0 -= sum_j buf_j * C_j[:, 0:B],  w1 -= sum_j buf_j * C_j[:, B:2B]; optionally refresh the
//         buffer copy of w1 (copyto!(Qgpu[i-1],Qg1), RBL_gpu.jl:76) and, when the newest block is already stored
//         (fused local reorth, solver.cu), that of w0.
// C and the partials are float when the buffer is float, else double.
struct ReorthPlan {
    int B = 0, fp32 = 0;
    int64_t n = 0, m = 0;
    int chunks = 0, ranges = 0, jt = 0;
    size_t partial_elems = 0;   // elements of the partials array needed
};
ReorthPlan reorth_plan(int B, int fp32, int64_t n, int64_t m);
size_t reorth_max_partial_elems(int B, int fp32, int64_t n, int64_t m_cap);
void launch_reorth_gram(const ReorthPlan& p, const void* buf, int64_t block_stride_elems, const double* w0,
                        const double* w1, void* partials, void* C, cudaStream_t st);
void launch_reorth_gram_reduce(const ReorthPlan& p, const void* partials, void* C, cudaStream_t st);
void launch_reorth_update(const ReorthPlan& p, const void* buf, int64_t block_stride_elems, const void* C, double* w0,
                          double* w1, void* store_w1, cudaStream_t st, void* store_w0 = nullptr);

// scaled two-term FP16 split on mma.sync m16n8k16 (reorth_tc16.cu): same interface, half the tensor-pipe time.
// `n_global` fixes the power-of-two operand scale (orthonormal columns of global length n_global).
bool reorth_h_supported(int B, int fp32);   // fp32 buffer, B = 16 or 32
size_t reorth_h_scratch_words(int B, int64_t n, int64_t m_cap);
void launch_reorth_gram_h(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t block_stride_elems,
                          const double* w0, const double* w1, void* partials, void* C, float* scratch, int64_t m_cap,
                          int presplit, cudaStream_t st);
float reorth_h_scale(int64_t n_global);   // the operand scale S; the split16 slab is written with it
void launch_reorth_coeff_h(const ReorthPlan& p, const void* C, float* scratch, int64_t m_cap, int recompute_max,
                           cudaStream_t st);
void launch_reorth_update_h(const ReorthPlan& p, int64_t n_global, const void* buf, int64_t block_stride_elems,
                            double* w0, double* w1, void* store_w1, float* scratch, int64_t m_cap, int presplit,
                            cudaStream_t st, void* store_w0 = nullptr, int64_t coef_block0 = 0);
// coef_block0: `buf` holds stored blocks coef_block0 .. coef_block0+p.m of the coefficient set built by launch_reorth_coeff_h
// (host-spilled blocks are streamed through a staging buffer chunk by chunk)
// `presplit`: buf (and store_w1) hold split16 rows (split16.h) instead of fp32 values

// all-fp64 mode: FP64 tensor-core MMA (mma.sync.m8n8k4.f64) + cp.async rings, B = 16 (reorth_f64.cu)
bool reorth_d_supported(int B, int fp32);
void launch_reorth_gram_d(const ReorthPlan& p, const void* buf, int64_t block_stride_elems, const double* w0,
                          const double* w1, void* partials, void* C, cudaStream_t st);
void launch_reorth_update_d(const ReorthPlan& p, const void* buf, int64_t block_stride_elems, const void* C, double* w0,
                            double* w1, void* store_w1, cudaStream_t st, void* store_w0 = nullptr);

// ---- K6 Ritz vectors -----------------------------------------------------------------------------
// V[:, t] = sum_j buf_j * S[(j*B .. j*B+B), t];  S device, row-major (m*B) x kpad in the buffer's type;
// V column-major n x k (ldv), float or double as the buffer.                  RBL_gpu.jl:106-132
void launch_ritz(int B, int fp32, int64_t n, int64_t m, int k, int kpad, const void* buf, int64_t block_stride_elems,
                 const void* S, void* V, int64_t ldv, int v_fp32, float split_scale, cudaStream_t st, int accumulate = 0);
// split_scale != 0: the fp32 buffer holds split16 rows written with that scale

// tensor-core K6 for a split16 buffer (reorth_tc16.cu): B = 16 or 32, S fp32; scratch: ritz_h_scratch_words words
size_t ritz_h_scratch_words(int B, int64_t m, int kpad);
void launch_ritz_h(int B, int64_t n, int64_t m, int k, int kpad, const void* buf, int64_t block_stride_elems, const void* S,
                   void* V, int64_t ldv, int v_fp32, float split_scale, unsigned* scratch, cudaStream_t st, int accumulate = 0);

// ---- layout / conversion helpers -------------------------------------------------------------------
// column-major n x b (ld) fp64  ->  row-major n x B fp64 (zero padded)
void launch_colmajor_to_block(int B, int64_t n, int b, const double* src, int64_t ld, double* dst, cudaStream_t st);
void launch_block_to_colmajor(int B, int64_t n, int b, const double* src, double* dst, int64_t ld, cudaStream_t st);
void launch_store_block(int B, int64_t n, const double* src, void* dst, int fp32, float split_scale, cudaStream_t st);
// fp32 rows -> split16 rows (out of place)
void launch_encode_split(int B, int64_t rows, const float* src, void* dst, float scale, cudaStream_t st);
void launch_load_block(int B, int64_t n, const void* src, int fp32, double* dst, cudaStream_t st);
// slab block (fp64 / fp32 / split16 when split_scale != 0) -> fp64 row-major block
void launch_decode_block(int B, int64_t n, const void* src, int fp32, float split_scale, double* dst, cudaStream_t st);
// Orthogonality loss of the Krylov basis: C is the (m*B) x 2B coefficient matrix of a Gram pass whose targets are the
// decoded stored blocks j0 and j0+1 (ntargets = 1: only j0), so the exact answer is delta = 1 at
// (row (j0 + t/B)*B + t%B, column t).  Accumulates out[0] = max |C - delta|, out[1] = sum (C - delta)^2 over calls
// (zero `out` first).  Rows / columns of padded or deflated (all-zero) columns are skipped via `skip` (m*B flags).
void launch_ortho_accumulate(int B, int64_t m, int64_t j0, int ntargets, int fp32, const void* C, double* out,
                             cudaStream_t st);
// per-column dot products of two column-major n x k matrices (ld): out[c] = v_c'v_c, out[k+c] = v_c'w_c,
// out[2k+c] = w_c'w_c (accumulated with atomics: zero `out` first).  Rayleigh quotients / residuals of Ritz vectors.
void launch_col_dots(int64_t n, int k, const double* V, const double* W, int64_t ld, double* out3k, cudaStream_t st);
void launch_randn(int64_t count, uint64_t seed, uint64_t offset, double* dst, cudaStream_t st);
void launch_convert_s(int64_t count, const double* src, void* dst, int fp32, cudaStream_t st);
void launch_gather_rows(int B, int64_t nrows, const int* rows, const double* src, double* dst, cudaStream_t st);

// ---- micro-benchmarks ------------------------------------------------------------------------------
double microbench(int which, int64_t size, int iters);

}  // namespace rbl

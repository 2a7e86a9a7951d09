/*
 * rbl_b200.h - C ABI of the B200-native randomized block Lanczos (RBL) eigensolver.
 *
 * This is the drop-in boundary for the reference's two Julia entry points
 *     RBL    (A, k, b) -> (D, V)        Julia/RBL.jl:119-142
 *     RBL_gpu(A, k, b) -> (D, V)        Julia/RBL_gpu.jl:205-221
 * Host code (Julia `ccall`, or the Python `ctypes` mirror used by the tests in this repository)
 * passes the arrays exactly as Julia's SparseMatrixCSC lays them out (Int64 colptr/rowval,
 * 1-based, Float64 nzval, column-major dense blocks) and receives the k eigenvalues of largest
 * magnitude (descending |lambda|, RBL.jl:116) and the n x k Ritz vectors (column-major, host).
 *
 * Everything below `rbl_solve` is a kernel-level export: one entry per device kernel family, used
 * by the parity tests (against oracle/) and by the profiling scripts.  Each cites the reference
 * call site it replaces.  No torch / CUDA types appear in any signature: device buffers are plain
 * `void*` device addresses, streams are not exposed.
 *
 * Threading: one solve per handle at a time; the library may start host worker threads.
 * Errors: every function returns an rbl_status; rbl_last_error() gives the message of the last
 * failure on the calling thread.
 */
#ifndef RBL_B200_H
#define RBL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RBL_OK = 0,
    RBL_NOT_CONVERGED = 1, /* cap reached; best-effort results of the last check are returned   */
    RBL_BREAKDOWN = 2,     /* reserved: rank-deficient residual block (handled by deflation)      */
    RBL_OOM = 3,
    RBL_INVALID = 4,
    RBL_CUDA_ERROR = 5,
    RBL_NCCL_ERROR = 6,
    RBL_NO_DEVICE = 7      /* no CUDA device: the product has no CPU fallback                    */
} rbl_status;

enum { RBL_PRECISION_FP64 = 0, RBL_PRECISION_MIXED = 1 };
enum { RBL_OP_A = 0, RBL_OP_SHIFT_MINUS_A = 1 };

/* Options.  Defaults (rbl_options_default) reproduce the constants hard-coded in the reference. */
typedef struct {
    int64_t max_kryl_sz;   /* Krylov column cap; 1200 = RBL_gpu.jl:211 (RBL.jl:133 uses 1400)     */
    double tol;            /* absolute residual-bound tolerance, 1e-7 = RBL_gpu.jl:189            */
    int32_t reorth_period; /* full reorth of the two newest blocks every 2nd step, RBL_gpu.jl:164 */
    int32_t check_period;  /* convergence check every 4th step, RBL_gpu.jl:186                    */
    int32_t precision;     /* RBL_PRECISION_FP64 = shipped FLOAT=Float64 (common.jl:5);
                              RBL_PRECISION_MIXED = README.md:69 split: fp32 Krylov buffer, fp32
                              reorth + Ritz arithmetic, fp64 active blocks / SpMM / QR / T        */
    int32_t op;            /* RBL_OP_A: operator is A; RBL_OP_SHIFT_MINUS_A: sigma*I - A (lowest
                              eigenpairs of A as largest of the shifted operator)  */
    double sigma;          /* shift for RBL_OP_SHIFT_MINUS_A                                      */
    int32_t device;        /* CUDA device ordinal (-1: current)                                   */
    int32_t async_check;   /* 1: run the host T eigen-check on a worker thread while the device
                              keeps iterating (results identical to the synchronous order); row-
                              sharded solves never wait at a check point (a check that is still
                              running postpones the next one: the accepted step may vary by a few
                              check periods between runs).  2: that non-waiting form on one GPU too.
                              0: synchronous, as the reference (RBL_gpu.jl:186-193)               */
    int32_t host_threads;  /* worker threads for the final host eigensolve (0: hardware)          */
    int32_t v_fp32;        /* 1: V_out is float (reference's FLOAT=Float32 build), else double    */
    int32_t verbose;
    int32_t reorth_impl;   /* 0: auto; 1: SIMT; 3: tensor-core scaled FP16 split of an fp32 buffer; 4: the same
                              with the buffer stored pre-split (auto picks 4 in mixed precision, B = 16/32) */
    int32_t seed;          /* stream of the counter-based device generator used when omega == NULL (the reference
                              never seeds CUDA.randn, RBL_gpu.jl:213); 0 selects the default stream           */
    int32_t ngpus;         /* > 1: rbl_create row-shards A over devices [device, device+ngpus) of THIS process
                              (one host thread + one NCCL rank per device); rbl_solve then takes / returns the
                              full n x b Omega and n x k V like the single-GPU call                          */
    int32_t filter_degree; /* 0: iterate with the operator itself (the reference); d > 0: iterate with the
                              degree-d Chebyshev-filtered operator p(op(A)) (N1, generalises restarted.jl);
                              -1: pick a degree                                                              */
    int32_t restart;       /* 1: when the Krylov cap (max_kryl_sz or device memory) is reached, lock the converged
                              Ritz pairs and restart from the best unconverged ones (restarted.jl:23-146)
                              instead of returning RBL_NOT_CONVERGED                                         */
    int32_t spill;         /* 1: Krylov blocks that do not fit device memory are kept in pinned host memory and
                              streamed back for the re-orthogonalisation (hybrid_part_reorth!, RBL_gpu.jl:59-81) */
    int32_t probe_steps;   /* filtered solves: plain block steps used to locate the wanted end of the spectrum
                              (0: automatic)                                                                 */
    int32_t mem_limit_mb;  /* device-memory budget of a solve in MiB (0: what cudaMemGetInfo reports); lets the
                              capped / spill / restart paths be exercised on small problems                  */
} rbl_options;

/* Per-solve statistics; phase labels are the reference's TimerOutputs labels (RBL_gpu.jl:152-187,219). */
typedef struct {
    int64_t iterations;      /* block steps executed before the accepting check (RBL_gpu.jl:195)  */
    int64_t kryl_sz;         /* iterations * b                                                    */
    int64_t iterations_run;  /* block steps the device actually ran (>= iterations when async)    */
    int32_t converged;
    int32_t checks;          /* number of host eigen-checks                                       */
    int32_t full_checks;     /* how many of them computed all k pairs                             */
    int32_t deflated;        /* columns deflated by the block QR                                  */
    double t_total;          /* wall seconds of rbl_solve                                         */
    double t_spmm;           /* "AQ"           device seconds (CUDA events)                       */
    double t_3term;          /* "3-term"                                                          */
    double t_qr;             /* "qr"                                                              */
    double t_part_reorth;    /* "part reorth"                                                     */
    double t_loc_reorth;     /* "loc reorth"                                                      */
    double t_eig;            /* "eig"          host seconds                                       */
    double t_ritz;           /* "Ritz vectors" device seconds                                     */
    double t_h2d;            /* upload of A and Omega (e2e accounting)                            */
    double t_d2h;            /* download of V                                                     */
    double t_eig_wait;       /* seconds the DEVICE sat idle because the host was still waiting for the
                                result of a convergence check (measured by polling the stream)     */
    double bytes_part_reorth;/* algorithmic HBM bytes streamed by the reorth Gram+update kernels  */
    double bytes_spmm;       /* algorithmic bytes of all SpMM launches                            */
    int64_t kernel_launches; /* kernels of this library launched during the solve                 */
    double t_reorth_gram;    /* device seconds inside the K5a Gram kernels (sum over launches)     */
    double t_reorth_update;  /* device seconds inside the K5b update kernels                       */
    double bytes_reorth_gram;
    double bytes_reorth_update;
    int64_t launches_reorth_gram;
    int64_t launches_reorth_update;
    int64_t launches_spmm;
    double t_ritz_kernel;    /* device seconds inside the K6 kernel                               */
    double bytes_ritz;
    double flops_ritz;
    int64_t host_factorizations; /* band factorisations done by the host eigen-checks              */
    int64_t restarts;        /* restart cycles after the first (opts.restart)                      */
    int64_t locked;          /* Ritz pairs locked by restarts                                      */
    int64_t spilled_blocks;  /* Krylov blocks that lived in pinned host memory (opts.spill)        */
    int64_t buffer_blocks;   /* Krylov blocks the device buffer of this solve could hold           */
    double max_residual;     /* filtered solves: max ||A v - lambda v|| over the returned pairs, measured on
                                the device in fp64 with op(A) itself (0 when not computed)          */
    double t_host_blocked;   /* host seconds the solve thread spent blocked on check results        */
    double filter_cut;       /* |lambda| below which the filter damps (0: no filter)               */
    int32_t filter_degree;   /* degree actually used                                               */
    int32_t filter_two_sided;/* 1: wanted pairs on both ends of the spectrum (odd filter)          */
} rbl_stats;

typedef struct rbl_handle rbl_handle;

const char* rbl_last_error(void);
const char* rbl_version(void);
int rbl_device_count(void);
int rbl_options_default(rbl_options* opts);
/* sizeof(rbl_options), sizeof(rbl_stats) as compiled: lets a binding (ccall / ctypes) verify its struct mirrors. */
int rbl_struct_sizes(int64_t* options_bytes, int64_t* stats_bytes);

/* Replaces `Ag = adapt(CuArray, A)` (RBL_gpu.jl:209) + `matrix_size` (RBL_gpu.jl:8-22).
 * colptr (n+1), rowval (nnz), nzval (nnz): Julia SparseMatrixCSC fields of a SYMMETRIC matrix
 * (CSC == CSR); index_base 1 for Julia, 0 for SciPy.  Indices are narrowed to int32 on the device
 * (as CUSPARSE.CuSparseMatrixCSC{Float64,Int32} does); n or nnz >= 2^31 is RBL_INVALID. */
int rbl_create(int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval,
               int index_base, const rbl_options* opts, rbl_handle** out);

/* Dense-A variant of the same entry (`Matrix{DOUBLE}` in the signature, RBL_gpu.jl:20-22,205):
 * the dense column-major n x n matrix is stored as CSR with n entries per row. */
int rbl_create_dense(int64_t n, const double* a_colmajor, const rbl_options* opts, rbl_handle** out);

/* Row-sharded variant (new; the reference is single-GPU): this process owns rows [row0,row0+nloc)
 * of the global n x n matrix, given as local CSR with GLOBAL column indices.  `nccl_uid` is the
 * 128-byte ncclUniqueId created by rbl_nccl_unique_id on rank 0 and distributed by the host
 * (torch.distributed / MPI). */
int rbl_nccl_unique_id(void* uid128);
int rbl_create_sharded(int64_t n, int64_t row0, int64_t nloc, int64_t nnz_loc, const int64_t* rowptr,
                       const int64_t* colidx_global, const double* vals, int index_base, int rank, int world,
                       const void* nccl_uid, const rbl_options* opts, rbl_handle** out);

int rbl_destroy(rbl_handle* h);
/* The Krylov slab of a destroyed handle is kept for the next handle on the same device (allocation of tens
 * of GB costs hundreds of ms); this returns it to the driver - the analogue of CUDA.reclaim() (RBL_gpu.jl:201). */
int rbl_release_cached_memory(void);

/* Replaces the body of RBL_gpu (RBL_gpu.jl:211-220): random start (213-214), lanczos_iteration
 * (134-203) and recover_eigvec (106-132).
 *   omega   n x b column-major start block (the reference draws CUDA.randn, never seeded; NULL
 *           draws N(0,1) from a counter-based device generator with opts seed 0).  Sharded handles
 *           pass their nloc x b row slice (column-major, leading dimension nloc).
 *   d_out   k eigenvalues, descending |lambda| (of the operator selected by opts.op)
 *   v_out   n x k (nloc x k when sharded) column-major Ritz vectors, double or float (opts.v_fp32)
 */
int rbl_solve(rbl_handle* h, int64_t k, int64_t b, const double* omega, double* d_out, void* v_out,
              rbl_stats* stats);

/* Same solve with inputs and outputs resident in device memory (bench `value` leg):
 * omega_dev is n x b column-major fp64 on the device, v_dev n x k column-major. */
int rbl_solve_device(rbl_handle* h, int64_t k, int64_t b, const void* omega_dev, double* d_out, void* v_dev,
                     rbl_stats* stats);

/* `gpu_buffer_size` (RBL_gpu.jl:95-104): number of Krylov blocks of width b that fit the device now.
 * rbl_plan_blocks is the memory plan rbl_solve(h, k, b) itself uses (it also reserves the n x k Ritz-vector
 * buffer); rbl_buffer_blocks == rbl_plan_blocks with k = 0. */
int rbl_buffer_blocks(rbl_handle* h, int64_t b, int64_t* blocks_out);
int rbl_plan_blocks(rbl_handle* h, int64_t k, int64_t b, int64_t* blocks_out);

/* Debug / parity exports on the Krylov basis of the LAST solve of this handle (the reference keeps it in the
 * `Q` / `Qgpu` vectors, RBL_gpu.jl:136-151).  Single-GPU handles.
 *   rbl_krylov_info   number of stored blocks (locked blocks of restarted solves first) and block width b
 *   rbl_krylov_block  block j decoded to fp64, n x b column-major, on the host
 *   rbl_orthogonality max_abs_out = max |(Q'Q - I)_ij|, fro_out = ||Q'Q - I||_F over all stored columns, computed
 *                     on the device with the K5a Gram kernels (an upper bound of the 2-norm the north star quotes) */
int rbl_krylov_info(rbl_handle* h, int64_t* blocks_out, int64_t* b_out);
int rbl_krylov_block(rbl_handle* h, int64_t j, double* out_colmajor);
int rbl_orthogonality(rbl_handle* h, double* max_abs_out, double* fro_out);
/* `CUDA.available_memory()` (RBL_gpu.jl:25,96). */
int rbl_query_memory(int device, int64_t* free_bytes, int64_t* total_bytes);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level exports (host buffers in, host buffers out; the kernels run on the device).
 * Dense blocks are ROW-major n x B here (the library's internal HBM layout).
 * ---------------------------------------------------------------------------------------------- */

/* K1: U = op(A) * Q   - replaces `mul!(U,Ag,Qg_d)` (RBL_gpu.jl:152,176).  q,u: n x b row-major. */
int rbl_spmm(rbl_handle* h, int64_t b, const double* q, double* u);

/* K2: C = X' * Y (b x b, row-major) - replaces `transpose(Qg_d)*U` (RBL_gpu.jl:153,178) and the
 * Gram of `loc_reorth_gpu!` (RBL_gpu.jl:87). */
int rbl_gram(int64_t n, int64_t b, const double* x, const double* y, double* c);

/* K3: thin block QR  U = Q R  (R upper triangular, b x b row-major) by shifted CholQR with
 * re-orthogonalisation passes - replaces `qr(U)` + `CuArray(fact.Q)` + `Array(fact.R)`
 * (RBL_gpu.jl:155-159,180-184).  deflated_out[j] = 1 where column j was numerically dependent. */
int rbl_block_qr(int64_t n, int64_t b, double* u_inout, double* r_out, int32_t* deflated_out);

/* K4/K5: W -= Qbuf * (Qbuf' * W) for m stored blocks at once (block classical Gram-Schmidt) -
 * replaces hybrid_part_reorth! / part_reorth_gpu_async! (RBL_gpu.jl:59-81,29-47).
 *   qbuf   m blocks, each n x b row-major, contiguous (fp32 when storage_fp32, else fp64)
 *   w      n x (2b) given as two n x b row-major fp64 blocks w0,w1 (Q_i and Q_{i-1})
 *   c_out  optional m*b x 2b row-major coefficients (float or double as storage), may be NULL
 *   impl   0 auto, 1 SIMT, 3 tensor-core scaled FP16 split of an fp32 slab, 4 the same on the pre-split slab
 *          format the solver uses (both need |entries| <= 1) */
int rbl_reorth(int64_t n, int64_t b, int64_t m, int storage_fp32, const void* qbuf, double* w0, double* w1,
               void* c_out, int impl);

/* K6: V = Qbuf * S - replaces recover_eigvec (RBL_gpu.jl:106-132).  s: (m*b) x k row-major fp64,
 * v_out: n x k column-major (double, or float when storage_fp32).
 *   impl   0 auto, 1 SIMT, 4 tensor-core scaled FP16 split on the pre-split buffer format (fp32 storage, padded
 *          block size 16 or 32; what the solver uses in mixed precision; auto picks it when available) */
int rbl_ritz(int64_t n, int64_t b, int64_t m, int64_t k, int storage_fp32, const void* qbuf, const double* s,
             void* v_out, int impl);

/* Host side of the path (no device needed) ---------------------------------------------------- */

/* `dsbev('V','L',T)` + `sort_eig_abs` + `check_convergence` (common.jl:36-65) without the O(N^3)
 * all-eigenvectors solve: band storage ab is (kd+1) x N column-major, LAPACK lower band, as built
 * by insertA!/insertB! (common.jl:9-26).  bi is the b x b upper-triangular B_i (column-major, may
 * be NULL).  Outputs: d_out k eigenvalues by descending |lambda|, s_out N x k column-major
 * eigenvectors, resid_out k residual bounds ||B_i S[end-b+1:end, j]||, converged_out the decision
 * of check_convergence(B_i,S,b,k,tol). */
int rbl_band_eig_topk(int64_t N, int64_t kd, const double* ab, int64_t k, const double* bi, int64_t b,
                      double tol, int threads, double* d_out, double* s_out, double* resid_out,
                      int32_t* converged_out);

/* Stateful variant used by the solver between checks (keeps the witness Ritz pair of the previous check):
 * returns the decision of check_convergence; when all k pairs were computed (*have_all_out = 1) d_out/s_out/
 * resid_out are filled as in rbl_band_eig_topk.  stats_out[0] = band factorisations of this call,
 * stats_out[1] = 1 if the full k-pair path ran. */
typedef struct rbl_checker rbl_checker;
int rbl_checker_create(int threads, rbl_checker** out);
int rbl_checker_check(rbl_checker* c, int64_t N, int64_t kd, const double* ab, int64_t k, const double* bi, int64_t b,
                      double tol, int force_full, double* d_out, double* s_out, double* resid_out,
                      int32_t* converged_out, int32_t* have_all_out, int64_t* stats_out);
/* Hands the checker k Ritz pairs of an EARLIER (smaller) T as starting points of its next full check - what the solver's
 * background tracker thread does between checks (d: k values, s: n_seed x k column-major).  resid (may be NULL): their k
 * residual bounds at that time; the pairs with the largest ones join the checker's witnesses. */
int rbl_checker_set_seeds(rbl_checker* c, int64_t n_seed, int64_t k, const double* d, const double* s, const double* resid);
/* fn(user, N) is called when a full check is about to start without usable seeds (N = size of T): the last moment at which
 * rbl_checker_set_seeds still helps.  The solver uses the same hook to wait for its tracker's pass in flight. */
typedef void (*rbl_need_seeds_fn)(void* user, int64_t N);
int rbl_checker_set_need_seeds(rbl_checker* c, rbl_need_seeds_fn fn, void* user);
int rbl_checker_destroy(rbl_checker* c);

/* Number of eigenvalues of the band matrix strictly below x (Sturm count by row-wise elimination). */
int rbl_band_count_below(int64_t N, int64_t kd, const double* ab, double x, int64_t* count_out);

/* 1-D contiguous row partition + SpMM halo plan of rank `rank` of `world` (new).
 * Inputs are this rank's local CSR rows with global columns (0-based).  Two-call protocol: call with
 * the *_out pointers NULL to get sizes, then with buffers.
 *   row_starts   world+1 global row offsets of the partition (input)
 *   halo_cols    sorted global column indices outside [row0,row0+nloc) referenced by local rows
 *   halo_owner_ptr  world+1 offsets into halo_cols by owning rank
 *   colidx_local nnz remapped indices: own rows -> [0,nloc), halo -> nloc + position in halo_cols */
int rbl_partition_rows(int64_t n, int world, int64_t* row_starts_out);
int rbl_halo_plan(int64_t n, int world, const int64_t* row_starts, int rank, int64_t nloc, int64_t nnz_loc,
                  const int64_t* rowptr, const int64_t* colidx_global, int64_t* n_halo_out, int64_t* halo_cols_out,
                  int64_t* halo_owner_ptr_out, int32_t* colidx_local_out);

/* Row schedule of the patch-scheduled SpMM (host only; new - the reference leaves the SpMM to cuSPARSE, RBL_gpu.jl:152).
 * From 0-based int32 CSR arrays of the local rows (columns >= nown are halo columns) it detects the grid a stencil matrix
 * lives on and returns the rows patch by patch: `slots` entries per patch, -1 = padding.  Two-call protocol: with
 * order_out NULL it returns the number of entries (0: no stencil structure, the plain gather kernel is used).
 *   info_out[12]: dims, stride1, stride2, ext0, ext1, ext2, patch0, patch1, patch2, slots, npatch, halo0 */
int rbl_spmm_schedule(int64_t nrows, int64_t nown, const int32_t* rowptr, const int32_t* colidx, int slots,
                      int64_t* entries_out, int32_t* order_out, int64_t* info_out);

/* Matrix Market loader (host only) - what `mmread` gives the reference's benchmark driver (benchmark.jl:3,21,28): reads a
 * square `coordinate` file (real / integer / pattern; general / symmetric / skew-symmetric) into the CSC arrays rbl_create
 * takes (full matrix, symmetric storage expanded, duplicates summed, row indices sorted, index_base 0 or 1).
 * The arrays returned by rbl_matrix_arrays stay valid until rbl_matrix_free. */
typedef struct rbl_matrix rbl_matrix;
int rbl_matrix_market_read(const char* path, int index_base, rbl_matrix** out, int64_t* n_out, int64_t* nnz_out);
int rbl_matrix_arrays(rbl_matrix* m, const int64_t** colptr, const int64_t** rowval, const double** nzval);
int rbl_matrix_free(rbl_matrix* m);

/* Micro-benchmarks used by bench.py / profiles (device): achieved copy GB/s and pipe rates. */
/* K1 laboratory (spmm_lab.cu): mean launch time in microseconds of SpMM variant `variant` on the handle's matrix, Q resident,
 * L2 flushed before every launch when flush != 0; *mismatch_out = elements that differ from the default kernel's result. */
int rbl_spmm_bench(rbl_handle* h, int64_t b, int variant, int grid_mult, int iters, int flush, int with_z, double* us_out,
                   int64_t* mismatch_out);
int rbl_microbench(int which, int64_t size, int iters, double* result_out);

#ifdef __cplusplus
}
#endif
#endif /* RBL_B200_H */

// Device-resident driver of the RBL iteration (replaces RBL_gpu / lanczos_iteration / recover_eigvec,
// Julia/RBL_gpu.jl:134-221).  Host side of the C ABI in include/rbl_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rbl_b200.h"
#include "band_eig.h"
#include "comm.h"
#include "kernels.h"

namespace rbl {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define RBL_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            throw ::rbl::Error(e__ == cudaErrorMemoryAllocation ? RBL_OOM : RBL_CUDA_ERROR,         \
                               std::string(#call) + ": " + cudaGetErrorString(e__) + " at " +      \
                                   __FILE__ + ":" + std::to_string(__LINE__));                      \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        count = n;
        if (n) RBL_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    // grow-only: keeps the allocation when it is already large enough (workspace reuse across solves)
    void ensure(size_t n) {
        if (n > count || (n && !p)) alloc(n);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        count = 0;
    }
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t count = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void alloc(size_t n) {
        release();
        count = n;
        if (n) RBL_CUDA(cudaMallocHost((void**)&p, n * sizeof(T)));
    }
    void ensure(size_t n) {
        if (n > count || (n && !p)) alloc(n);
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        count = 0;
    }
};

}  // namespace rbl

namespace rbl {
// Per-handle device workspace, kept between solves (cudaMalloc/cudaFree of the tens-of-GB Krylov slab costs
// hundreds of ms; CUDA.jl's pool allocator gives the reference the same reuse).  The slab itself is parked in
// a process-wide cache when a handle is destroyed and picked up by the next handle on the same device;
// rbl_release_cached_memory() returns it to the driver (the analogue of CUDA.reclaim(), RBL_gpu.jl:201).
struct Workspace {
    DevBuf<double> X[4];                 // three active blocks + the Chebyshev scratch block
    DevBuf<unsigned char> buf;
    DevBuf<double> part, small;
    DevBuf<QrState> qr;
    DevBuf<unsigned char> Cmat, rpart;
    DevBuf<float> tc_scratch;
    DevBuf<double> sendbuf;
    DevBuf<unsigned char> ritzS, ritzV;   // Ritz coefficient matrix and (host-output solves) the device copy of V
    DevBuf<unsigned> ritz_words;          // f16 hi/lo words of S for the tensor-core Ritz kernel
    DevBuf<double> omega;
    PinnedBuf<unsigned char> hslab;       // host tier of the Krylov slab (opts.spill)
    DevBuf<unsigned char> stage;          // 2 staging chunks + 3 staging store blocks of the host tier
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t spill_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<double> Vacc;                  // restarted / filtered solves: fp64 locked + final Ritz vectors (nloc x k)
    DevBuf<double> ctrl, share;           // row-sharded runs: decision flag and the shared (D, S) of the accepting check
    PinnedBuf<double> h_ctrl;
    PinnedBuf<double> hA, hB;
    PinnedBuf<unsigned char> d2h_pin;     // two staging chunks of the pageable-destination download of V
    cudaEvent_t d2h_ev[2] = {nullptr, nullptr};
    PinnedBuf<QrState> hqr;
    // what a handle needs besides the solve buffers: parked with the workspace so that a create / destroy cycle does
    // no cudaFree / cudaStreamDestroy / cudaEventDestroy (measured: sporadic 0.4-1.8 s stalls in rbl_destroy)
    DevBuf<int> d_rowptr, d_colidx, d_send_rows, d_rel;   // d_rel: window-relative column encoding of the TMA SpMM
    DevBuf<int> d_order;                  // patch-ordered row schedule of the SpMM (spmm_sched.cu)
    DevBuf<int> d_bnd_rows;               // row-sharded: local rows that reference halo columns
    DevBuf<unsigned char> d_bnd_flag;     // the same as one flag per local row
    cudaStream_t comm_stream = nullptr;   // halo exchange overlapped with the interior rows of the SpMM
    cudaEvent_t ev_q = nullptr, ev_halo = nullptr;
    DevBuf<double> d_vals;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> event_pool;
    ~Workspace() {
        for (auto e : event_pool) cudaEventDestroy(e);
        for (auto e : spill_ev)
            if (e) cudaEventDestroy(e);
        for (auto e : d2h_ev)
            if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_q) cudaEventDestroy(ev_q);
        if (ev_halo) cudaEventDestroy(ev_halo);
        if (comm_stream) cudaStreamDestroy(comm_stream);
        if (stream) cudaStreamDestroy(stream);
    }
};
// process-wide cache of whole workspaces (one per device): a destroyed handle parks its workspace, the next
// handle on that device adopts it
Workspace* workspace_take(int device);
void workspace_park(int device, Workspace* ws);
void slab_cache_release_all();
}  // namespace rbl

namespace rbl {
// what the last solve left in the Krylov slab (debug / parity exports rbl_krylov_block, rbl_orthogonality)
struct KrylovInfo {
    int B = 0, b = 0;
    int64_t blocks = 0, bstride = 0, m_cap = 0, m_dev = 0;
    bool fp32 = false, use_h = false, use_d = false;
    float split_scale = 0.f;
    size_t ssz = 8;
};
}  // namespace rbl

struct rbl_handle {
    rbl_options opt{};
    // opts.ngpus > 1: this is a group handle; the per-device handles do the work (one host thread each)
    std::vector<rbl_handle*> parts;
    std::vector<int64_t> part_rows;       // parts.size()+1 global row offsets
    double gersh_lo = 0.0, gersh_hi = 0.0;  // Gershgorin interval of A (global)
    int64_t n_bnd_rows = 0;                // rows in d_bnd_rows (0: no overlap)
    rbl::SpmmSchedule spmm_sched;          // dims > 0: patch-scheduled gather SpMM (default for stencil matrices)
    rbl::SpmmWindows spmm_wt;              // nwin > 0: the matrix has band structure, SpMM stages Q through shared memory
    rbl::KrylovInfo last;
    rbl::Workspace* wsp = nullptr;   // adopted from / returned to the process-wide cache
    rbl::Workspace& ws_ref() { return *wsp; }
    int device = 0;
    int64_t n = 0;      // global order
    int64_t row0 = 0;   // first owned row
    int64_t nloc = 0;   // owned rows
    int64_t nnz = 0;    // local nonzeros
    int64_t n_halo = 0;
    int rank = 0, world = 1;
    // CSR arrays, send lists, stream and timing events live in the workspace (wsp->d_rowptr, ...)
    // halo exchange plan
    std::vector<int64_t> halo_owner_ptr;  // world+1: halo rows received from each owner
    std::vector<int64_t> send_ptr;        // world+1: rows sent to each peer
    rbl::Comm comm;
    cudaStream_t stream = nullptr;        // == wsp->stream
    double t_h2d_create = 0.0;
    ~rbl_handle();
};

namespace rbl {

rbl_handle* handle_create(int64_t n, int64_t row0, int64_t nloc, int64_t nnz, const int64_t* rowptr,
                          const int64_t* colidx, const double* vals, int index_base, int rank, int world,
                          const void* nccl_uid, const rbl_options* opts);
// where the start block comes from and where the Ritz vectors go: column-major, leading dimensions in elements
// (a row-sharded part of a group solve addresses its row slice of the caller's full arrays)
struct SolveIO {
    const double* omega = nullptr;
    int64_t ld_omega = 0;
    bool omega_on_device = false;
    void* v = nullptr;
    int64_t ldv = 0;
    bool v_on_device = false;
};
int solve(rbl_handle* h, int64_t k, int64_t b, const SolveIO& io, double* d_out, rbl_stats* stats);
struct MemPlan {
    double budget = 0, fixed = 0, per_block = 0;
    int64_t m_fit = 0;
};
MemPlan plan_memory(rbl_handle* h, int64_t k, int b, int64_t m_req);
void krylov_block(rbl_handle* h, int64_t j, double* out_colmajor);
void orthogonality(rbl_handle* h, double* max_abs, double* fro);
void default_options(rbl_options* o);

// spmm_lab.cu
double spmm_lab_run(rbl_handle* h, int b, int variant, int grid_mult, int iters, int flush, int with_z, unsigned long long* mismatch_out);

}  // namespace rbl

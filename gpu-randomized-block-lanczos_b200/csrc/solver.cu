// Device-resident RBL driver.  Reference restated by structure, not by call sequence:
//   handle_create   <- Ag = adapt(CuArray, A)                         RBL_gpu.jl:209
//   solve           <- RBL_gpu body + lanczos_iteration               RBL_gpu.jl:134-203, 211-220
//   Run::ritz       <- recover_eigvec                                 RBL_gpu.jl:106-132
//   Run::restart    <- RBL_gpu_restarted / lanczos_iteration_res      restarted.jl:23-146 (generalised, see below)
// Differences by design: no per-statement synchronisation (the only host waits
// are the convergence checks), Krylov buffer is one slab written by kernel epilogues (no F<->D copy
// kernels, no per-block host mirror), full re-orthogonalisation is two streaming passes (block CGS)
// over the slab instead of 4 small GEMMs + a sync per stored block, block QR is shifted CholQR with
// re-orthogonalisation instead of Householder geqrf/orgqr, and the host eigen-check computes only the
// k wanted Ritz pairs and may overlap the device iteration.
//
// Restart / filter (CPU twin: oracle/rbl_restart_oracle.py):
//   * the slab holds [locked Ritz vectors (whole blocks) | Krylov blocks of the current cycle]; the K5 kernels
//     stream both, so every reorth pass also enforces orthogonality against the locked vectors
//     (restart_reorth_gpu!, restarted.jl:1-21,:40,:58-59);
//   * when the cap is reached: every wanted pair whose residual bound passed is locked (restarted.jl:122-131),
//     the best b others form the restart block (:133-135);
//   * with opts.filter_degree the cycles iterate with p(op(A)) = rho T_d((op(A) - c)/e); eigenvalues are
//     recovered as Rayleigh quotients with op(A) and the true residuals are measured on the device.
#include "solver.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <unistd.h>

#include "partition.h"

using namespace rbl;

namespace rbl {
namespace {
std::mutex g_ws_mu;
std::vector<std::pair<int, Workspace*>> g_parked_ws;  // (device, workspace)
size_t solve_bytes(const Workspace& w) {
    size_t x = 0;
    for (int i = 0; i < 4; ++i) x += w.X[i].count;
    return w.buf.count + w.stage.count + w.ritzV.count + w.ritzS.count + w.Cmat.count + w.rpart.count + w.tc_scratch.count * 4 +
           w.ritz_words.count * 4 + (x + w.omega.count + w.part.count + w.small.count + w.sendbuf.count + w.Vacc.count) * 8;
}
size_t ws_bytes(const Workspace& w) {
    return solve_bytes(w) + (w.d_rowptr.count + w.d_colidx.count + w.d_send_rows.count) * 4 + w.d_vals.count * 8;
}
}  // namespace
Workspace* workspace_take(int device) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (size_t i = 0; i < g_parked_ws.size(); ++i)
        if (g_parked_ws[i].first == device) {
            Workspace* w = g_parked_ws[i].second;
            g_parked_ws.erase(g_parked_ws.begin() + i);
            return w;
        }
    return new Workspace();
}
void workspace_park(int device, Workspace* ws) {
    if (!ws) return;
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (auto& e : g_parked_ws)
        if (e.first == device) {  // keep the larger one
            if (ws_bytes(*e.second) >= ws_bytes(*ws)) {
                delete ws;
            } else {
                delete e.second;
                e.second = ws;
            }
            return;
        }
    g_parked_ws.push_back({device, ws});
}
void slab_cache_release_all() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (auto& e : g_parked_ws) {
        cudaSetDevice(e.first);
        delete e.second;
    }
    g_parked_ws.clear();
}
}  // namespace rbl

rbl_handle::~rbl_handle() {
    for (rbl_handle* p : parts) delete p;
    parts.clear();
    if (!wsp) return;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    rbl::workspace_park(device, wsp);
    wsp = nullptr;
    comm.destroy();
    stream = nullptr;
}

namespace rbl {

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void default_options(rbl_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->max_kryl_sz = 1200;  // RBL_gpu.jl:211
    o->tol = 1e-7;          // RBL_gpu.jl:189
    o->reorth_period = 2;   // RBL_gpu.jl:164
    o->check_period = 4;    // RBL_gpu.jl:186
    o->precision = RBL_PRECISION_FP64;  // common.jl:5-6 as shipped
    o->op = RBL_OP_A;
    o->sigma = 0.0;
    o->device = -1;
    o->async_check = 1;
    o->host_threads = 0;
    o->v_fp32 = 0;
    o->verbose = 0;
    o->reorth_impl = 0;
    o->seed = 0;
    o->ngpus = 1;
    o->filter_degree = 0;   // the reference iterates with the operator itself
    o->restart = 0;         // RBL_gpu gives up at the cap (RBL_gpu.jl:162); restarted.jl is a separate entry
    o->spill = 0;
    o->probe_steps = 0;
    o->mem_limit_mb = 0;
}

// ------------------------------------------------------------------------------------------------ create
rbl_handle* handle_create(int64_t n, int64_t row0, int64_t nloc, int64_t nnz, const int64_t* rowptr,
                          const int64_t* colidx, const double* vals, int index_base, int rank, int world,
                          const void* nccl_uid, const rbl_options* opts) {
    if (n <= 0 || nloc < 0 || nnz < 0 || !rowptr || (nnz > 0 && (!colidx || !vals)))
        throw Error(RBL_INVALID, "rbl_create: bad arguments");
    if (n >= (int64_t)1 << 31 || nnz >= (int64_t)1 << 31)
        throw Error(RBL_INVALID, "rbl_create: n and nnz must be < 2^31 (device indices are int32)");
    if (index_base != 0 && index_base != 1) throw Error(RBL_INVALID, "rbl_create: index_base must be 0 or 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        throw Error(RBL_NO_DEVICE, "no CUDA device visible: rbl_b200 has no CPU fallback");
    std::unique_ptr<rbl_handle> h(new rbl_handle());
    if (opts) h->opt = *opts; else default_options(&h->opt);
    int dev = h->opt.device;
    if (dev < 0) RBL_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) throw Error(RBL_INVALID, "rbl_create: device ordinal out of range");
    h->device = dev;
    RBL_CUDA(cudaSetDevice(dev));
    h->wsp = workspace_take(dev);
    if (!h->wsp->stream) RBL_CUDA(cudaStreamCreateWithFlags(&h->wsp->stream, cudaStreamNonBlocking));
    h->stream = h->wsp->stream;
    h->n = n; h->row0 = row0; h->nloc = nloc; h->nnz = nnz; h->rank = rank; h->world = world;
    const double t0 = now_s();

    std::vector<int> rp((size_t)nloc + 1), ci((size_t)nnz);
    if (rowptr[nloc] - rowptr[0] != nnz) throw Error(RBL_INVALID, "rbl_create: rowptr[n]-rowptr[0] != nnz");
    for (int64_t r = 0; r <= nloc; ++r) {
        int64_t v = rowptr[r] - rowptr[0];
        if (v < 0 || v > nnz || (r > 0 && v < rp[r - 1])) throw Error(RBL_INVALID, "rbl_create: rowptr not monotone");
        rp[r] = (int)v;
    }
    const int64_t* cidx = colidx;  // entries of the first local row start at colidx[0]
    // Gershgorin interval of A (rigorous spectrum bounds; places the damped interval of the Chebyshev filter)
    double glo = 1e300, ghi = -1e300;
    for (int64_t r = 0; r < nloc; ++r) {
        double d = 0.0, off = 0.0;
        const int64_t gr = row0 + r;
        for (int p = rp[r]; p < rp[r + 1]; ++p) {
            const int64_t c = cidx[p] - index_base;
            if (c == gr) d += vals[p]; else off += std::fabs(vals[p]);
        }
        glo = std::min(glo, d - off);
        ghi = std::max(ghi, d + off);
    }
    if (nloc == 0) { glo = 0.0; ghi = 0.0; }
    if (world <= 1) {
        for (int64_t p = 0; p < nnz; ++p) {
            int64_t c = cidx[p] - index_base;
            if (c < 0 || c >= n) throw Error(RBL_INVALID, "rbl_create: column index out of range");
            ci[p] = (int)c;
        }
        h->n_halo = 0;
    } else {
        std::string err;
        if (!h->comm.init(nccl_uid, rank, world, err)) throw Error(RBL_NCCL_ERROR, err);
        std::vector<int64_t> starts((size_t)world + 1);
        // all ranks must use the same partition: gather every rank's row0 (and Gershgorin bounds) via NCCL
        DevBuf<int64_t> d_send, d_recv;
        d_send.alloc(4);
        d_recv.alloc((size_t)4 * world);
        int64_t mine[4] = {row0, nloc, 0, 0};
        std::memcpy(&mine[2], &glo, 8);
        std::memcpy(&mine[3], &ghi, 8);
        RBL_CUDA(cudaMemcpyAsync(d_send.p, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
        if (!h->comm.allgather_i64(d_send.p, d_recv.p, 4, h->stream, err)) throw Error(RBL_NCCL_ERROR, err);
        std::vector<int64_t> all((size_t)4 * world);
        RBL_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, all.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        RBL_CUDA(cudaStreamSynchronize(h->stream));
        for (int p = 0; p < world; ++p) {
            starts[p] = all[4 * p];
            if (p > 0 && starts[p] != all[4 * (p - 1)] + all[4 * (p - 1) + 1])
                throw Error(RBL_INVALID, "rbl_create_sharded: row ranges are not contiguous in rank order");
            double lo, hi;
            std::memcpy(&lo, &all[4 * p + 2], 8);
            std::memcpy(&hi, &all[4 * p + 3], 8);
            if (all[4 * p + 1] > 0) { glo = std::min(glo, lo); ghi = std::max(ghi, hi); }
        }
        starts[world] = all[4 * (world - 1)] + all[4 * (world - 1) + 1];
        if (starts[0] != 0 || starts[world] != n) throw Error(RBL_INVALID, "rbl_create_sharded: ranges do not cover [0,n)");
        HaloPlan plan;
        if (!halo_plan(n, world, starts.data(), rank, nloc, nnz, rowptr, colidx, index_base, plan))
            throw Error(RBL_INVALID, "rbl_create_sharded: bad column index or row range");
        for (int64_t p = 0; p < nnz; ++p) ci[p] = plan.colidx_local[p];
        h->n_halo = (int64_t)plan.halo_cols.size();
        h->halo_owner_ptr = plan.halo_owner_ptr;
        // tell every owner which of its rows we need: counts (all-gather), then the index lists (send/recv)
        DevBuf<int64_t> d_cnt_s, d_cnt_r;
        d_cnt_s.alloc(world);
        d_cnt_r.alloc((size_t)world * world);
        std::vector<int64_t> need(world);
        for (int p = 0; p < world; ++p) need[p] = plan.halo_owner_ptr[p + 1] - plan.halo_owner_ptr[p];
        RBL_CUDA(cudaMemcpyAsync(d_cnt_s.p, need.data(), world * 8, cudaMemcpyHostToDevice, h->stream));
        if (!h->comm.allgather_i64(d_cnt_s.p, d_cnt_r.p, world, h->stream, err)) throw Error(RBL_NCCL_ERROR, err);
        std::vector<int64_t> cnt((size_t)world * world);
        RBL_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt_r.p, cnt.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        RBL_CUDA(cudaStreamSynchronize(h->stream));
        h->send_ptr.assign((size_t)world + 1, 0);
        for (int p = 0; p < world; ++p) h->send_ptr[p + 1] = h->send_ptr[p] + cnt[(size_t)p * world + rank];
        const int64_t nsend = h->send_ptr[world];
        DevBuf<int64_t> d_need, d_give;
        d_need.alloc(std::max<int64_t>(1, h->n_halo));
        d_give.alloc(std::max<int64_t>(1, nsend));
        if (h->n_halo)
            RBL_CUDA(cudaMemcpyAsync(d_need.p, plan.halo_cols.data(), h->n_halo * 8, cudaMemcpyHostToDevice, h->stream));
        if (!h->comm.group_start(err)) throw Error(RBL_NCCL_ERROR, err);
        for (int p = 0; p < world; ++p) {
            if (p == rank) continue;
            if (!h->comm.send_bytes(d_need.p + plan.halo_owner_ptr[p], (size_t)need[p] * 8, p, h->stream, err) ||
                !h->comm.recv_bytes(d_give.p + h->send_ptr[p], (size_t)(h->send_ptr[p + 1] - h->send_ptr[p]) * 8, p,
                                    h->stream, err))
                throw Error(RBL_NCCL_ERROR, err);
        }
        if (!h->comm.group_end(err)) throw Error(RBL_NCCL_ERROR, err);
        std::vector<int64_t> give((size_t)std::max<int64_t>(1, nsend));
        if (nsend) RBL_CUDA(cudaMemcpyAsync(give.data(), d_give.p, nsend * 8, cudaMemcpyDeviceToHost, h->stream));
        RBL_CUDA(cudaStreamSynchronize(h->stream));
        std::vector<int> send_rows((size_t)std::max<int64_t>(1, nsend));
        for (int64_t s = 0; s < nsend; ++s) {
            int64_t lr = give[s] - row0;
            if (lr < 0 || lr >= nloc) throw Error(RBL_INVALID, "halo plan: peer requested a row this rank does not own");
            send_rows[s] = (int)lr;
        }
        h->wsp->d_send_rows.ensure(std::max<int64_t>(1, nsend));
        if (nsend) RBL_CUDA(cudaMemcpy(h->wsp->d_send_rows.p, send_rows.data(), nsend * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->gersh_lo = glo;
    h->gersh_hi = ghi;
    h->wsp->d_rowptr.ensure((size_t)nloc + 1 + 8);     // slack: the TMA SpMM copies 16-byte rounded slices
    h->wsp->d_colidx.ensure(std::max<int64_t>(1, nnz));
    h->wsp->d_vals.ensure(std::max<int64_t>(1, nnz) + 8);
    RBL_CUDA(cudaMemcpy(h->wsp->d_rowptr.p, rp.data(), ((size_t)nloc + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (nnz) {
        RBL_CUDA(cudaMemcpy(h->wsp->d_colidx.p, ci.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
        RBL_CUDA(cudaMemcpy(h->wsp->d_vals.p, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    // row-sharded: the rows that reference halo columns.  The SpMM computes all other rows while the halo exchange is in
    // flight on a second stream and these rows once it has arrived (Run::spmm).
    h->n_bnd_rows = 0;
    if (world > 1 && h->n_halo > 0) {
        std::vector<int> bnd;
        std::vector<unsigned char> flag((size_t)nloc, 0);
        for (int64_t r = 0; r < nloc; ++r)
            for (int p = rp[r]; p < rp[r + 1]; ++p)
                if (ci[p] >= nloc) { flag[r] = 1; bnd.push_back((int)r); break; }
        // default on (RBL_HALO_OVERLAP=0 disables): config 5 on 8 GPUs 1.17 ms per SpMM against 1.32 ms with the exchange in
        // front of the kernel (2 x 20 MB per SpMM and rank); neutral on 2 GPUs
        const char* env = std::getenv("RBL_HALO_OVERLAP");
        if (!(env && env[0] == '0') && !bnd.empty() && (int64_t)bnd.size() * 4 <= nloc) {
            h->wsp->d_bnd_rows.ensure(bnd.size());
            h->wsp->d_bnd_flag.ensure((size_t)nloc);
            RBL_CUDA(cudaMemcpy(h->wsp->d_bnd_rows.p, bnd.data(), bnd.size() * sizeof(int), cudaMemcpyHostToDevice));
            RBL_CUDA(cudaMemcpy(h->wsp->d_bnd_flag.p, flag.data(), flag.size(), cudaMemcpyHostToDevice));
            if (!h->wsp->comm_stream) RBL_CUDA(cudaStreamCreateWithFlags(&h->wsp->comm_stream, cudaStreamNonBlocking));
            if (!h->wsp->ev_q) RBL_CUDA(cudaEventCreateWithFlags(&h->wsp->ev_q, cudaEventDisableTiming));
            if (!h->wsp->ev_halo) RBL_CUDA(cudaEventCreateWithFlags(&h->wsp->ev_halo, cudaEventDisableTiming));
            h->n_bnd_rows = (int64_t)bnd.size();
        }
        if (h->opt.verbose)
            std::fprintf(stderr, "[rbl] rank %d: %zu of %lld rows reference halo columns -> halo exchange %s\n", rank, bnd.size(), (long long)nloc,
                         h->n_bnd_rows ? "overlapped with the interior rows" : "before the SpMM");
    }
    // band structure?  then SpMM stages Q through shared-memory rings filled by the TMA engine (spmm.cu)
    h->spmm_wt = SpmmWindows{};
    {
        const char* env = std::getenv("RBL_SPMM_WINDOW");
        // opt-in (RBL_SPMM_WINDOW=1): correct and at its designed L2 traffic (3Q + A), but instruction-bound - 189 us per
        // config-2 launch against 96 us for the gather kernel (profiles/r02_ncu_spmm_window.txt)
        SpmmWindows wt = (env && env[0] == '1') ? spmm_plan_windows(nloc, nloc, rp.data(), ci.data()) : SpmmWindows{};
        if (wt.nwin > 0) {
            h->wsp->d_rel.ensure(std::max<int64_t>(1, nnz) + 8);
            DevBuf<unsigned long long> d_bad;
            d_bad.alloc(1);
            RBL_CUDA(cudaMemsetAsync(d_bad.p, 0, 8, h->stream));
            launch_spmm_build_rel(nloc, nloc, h->wsp->d_rowptr.p, h->wsp->d_colidx.p, wt, h->wsp->d_rel.p, d_bad.p, h->stream);
            unsigned long long bad = 0;
            RBL_CUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, h->stream));
            RBL_CUDA(cudaStreamSynchronize(h->stream));
            if ((double)bad <= 0.10 * (double)nnz) h->spmm_wt = wt;
            if (h->opt.verbose)
                std::fprintf(stderr, "[rbl] SpMM: %d offset windows, %llu of %lld entries outside them -> %s\n", wt.nwin, bad, (long long)nnz,
                             h->spmm_wt.nwin ? "TMA-staged window kernel" : "gather kernel");
        }
    }
    // stencil structure?  the SpMM laboratory can visit the rows patch by patch of the implied grid (spmm_sched.cu)
    h->spmm_sched = SpmmSchedule{};
    {
        const char* env = std::getenv("RBL_SPMM_SCHED");      // laboratory only (tools/spmm_lab.py sets it)
        if (env && env[0] == '1' && h->spmm_wt.nwin == 0) {
            std::vector<int> order;
            SpmmSchedule sc;
            // the schedule does not depend on the block size except through the patch volume: one for all B
            if (spmm_plan_schedule(nloc, nloc, rp.data(), ci.data(), spmm_sched_default_slots(16), order, &sc)) {
                h->wsp->d_order.ensure(order.size());
                RBL_CUDA(cudaMemcpy(h->wsp->d_order.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
                h->spmm_sched = sc;
            }
            if (h->opt.verbose)
                std::fprintf(stderr, "[rbl] SpMM: %s (grid %lld x %lld x %lld, strides 1/%lld/%lld, patch %d x %d x %d in %d slots, %lld patches, model %.2f Q rows per row)\n",
                             h->spmm_sched.dims ? "row schedule planned (laboratory)" : "no stencil structure",
                             (long long)sc.ext[0], (long long)sc.ext[1], (long long)sc.ext[2], (long long)sc.stride[1], (long long)sc.stride[2],
                             sc.patch[0], sc.patch[1], sc.patch[2], sc.slots, (long long)sc.npatch, sc.fetch_model);
        }
    }
    h->t_h2d_create = now_s() - t0;
    return h.release();
}

// ------------------------------------------------------------------------------------------------ memory plan
// gpu_buffer_size (RBL_gpu.jl:95-104): how many Krylov blocks of width b fit, given everything else a solve of k
// pairs allocates.  One function for rbl_solve and for rbl_plan_blocks / rbl_buffer_blocks.
MemPlan plan_memory(rbl_handle* h, int64_t k, int b, int64_t m_req) {
    MemPlan p;
    const rbl_options& opt = h->opt;
    const int B = padded_block(b);
    const bool fp32 = opt.precision == RBL_PRECISION_MIXED;
    const size_t ssz = fp32 ? 4 : 8;
    const int64_t nloc = h->nloc, next = h->nloc + h->n_halo;
    size_t free_b = 0, total_b = 0;
    RBL_CUDA(cudaMemGetInfo(&free_b, &total_b));
    free_b += solve_bytes(h->ws_ref());  // memory already held by this handle's workspace is available to the solve
    double budget = 0.92 * (double)free_b;
    if (opt.mem_limit_mb > 0) budget = std::min(budget, (double)opt.mem_limit_mb * 1048576.0);
    const bool extra = opt.restart || opt.filter_degree != 0;
    const int nX = 3 + (opt.filter_degree != 0 ? 1 : 0);
    const int rgrid = std::max(rowop_grid(B, nloc), fused_rowop_supported(B) ? 2 * fused_rowop_grid(B, nloc) : 0);
    const size_t vsz = opt.v_fp32 ? 4 : 8;
    size_t fixed = (size_t)nX * next * B * 8                    // active blocks
                   + (size_t)nloc * 4 * B * 4                    // packed target words of the tensor-core reorth
                   + (size_t)rgrid * B * B * 8                   // Gram partials
                   + (size_t)nloc * b * 8                        // Omega
                   + (size_t)nloc * (size_t)std::max<int64_t>(k, 0) * vsz   // device copy of V
                   + ((size_t)16 << 20);
    if (extra) fixed += (size_t)nloc * (size_t)(2 * k + b) * 8;  // fp64 accumulation of locked + final vectors, restart scratch
    const size_t per_block = (size_t)nloc * B * ssz + (size_t)B * 2 * B * (ssz + 8);
    p.budget = budget;
    p.per_block = (double)per_block;
    // the Gram partials grow with the number of stored blocks too (piecewise): largest m <= m_req that fits
    auto need = [&](int64_t m) { return (double)fixed + (double)reorth_max_partial_elems(B, fp32, nloc, std::max<int64_t>(m, 1)) * ssz + (double)m * per_block; };
    int64_t lo = 0, hi = std::max<int64_t>(m_req, 1);
    if (need(hi) <= budget) {
        lo = hi;
    } else {
        while (hi - lo > 1) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (need(mid) <= budget) lo = mid; else hi = mid;
        }
    }
    p.m_fit = lo;
    p.fixed = need(lo) - (double)lo * per_block;
    return p;
}

// ------------------------------------------------------------------------------------------------ solve
namespace {

enum Phase { PH_NONE = 0, PH_SPMM, PH_3TERM, PH_QR, PH_LOC, PH_RGRAM, PH_RUPD, PH_RITZ, PH_COUNT };

struct PhaseTimer {
    rbl_handle* h;
    cudaStream_t st;
    std::vector<std::pair<int, int>> marks;  // (phase that STARTS here, event index)
    size_t used = 0;
    bool enabled = true;
    void mark(int phase) {
        if (!enabled) return;
        if (used == h->wsp->event_pool.size()) {
            cudaEvent_t e;
            RBL_CUDA(cudaEventCreate(&e));
            h->wsp->event_pool.push_back(e);
        }
        RBL_CUDA(cudaEventRecord(h->wsp->event_pool[used], st));
        marks.emplace_back(phase, (int)used);
        ++used;
    }
    void collect(double* sec /* PH_COUNT */) {
        for (int p = 0; p < PH_COUNT; ++p) sec[p] = 0.0;
        for (size_t i = 0; i + 1 < marks.size(); ++i) {
            if (marks[i].first == PH_NONE) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->wsp->event_pool[marks[i].second], h->wsp->event_pool[marks[i + 1].second]) == cudaSuccess)
                sec[marks[i].first] += ms * 1e-3;
        }
    }
};

// p(x) = rho * T_d((x - c)/e): damps [a, b], grows outside (oracle/rbl_restart_oracle.py ChebFilter)
struct FilterPlan {
    int degree = 0;
    double a = 0, b = 0, rho = 1;
    bool two_sided = false;
    double c() const { return 0.5 * (a + b); }
    double e() const { return 0.5 * (b - a); }
    double eval(double lam) const {
        if (degree == 0) return lam;
        const double x = (lam - c()) / e(), ax = std::fabs(x);
        double t = ax <= 1.0 ? std::cos(degree * std::acos(std::max(-1.0, std::min(1.0, x))))
                             : std::cosh(degree * std::acosh(ax)) * ((x < 0 && (degree & 1)) ? -1.0 : 1.0);
        return rho * t;
    }
    // eigenvalue estimate of op(A) from a Ritz value theta of p(op(A)) (the wanted side of the filter); `side`: +1 wanted
    // above the damped interval, -1 below, 0 two-sided (odd degree keeps the sign).  Returns false inside the interval.
    bool invert(double theta, int side, double* lam) const {
        if (degree == 0) { *lam = theta; return true; }
        const double y = theta / rho, ay = std::fabs(y);
        if (!(ay > 1.0)) return false;
        if (side > 0 && y < 0) return false;                            // a Ritz value of the other side (not yet converged)
        if (side < 0 && ((y < 0) != ((degree & 1) == 1))) return false;
        const double x = std::cosh(std::acosh(ay) / degree);
        const double sgn = side != 0 ? (double)side : (y < 0 ? -1.0 : 1.0);
        *lam = c() + e() * sgn * x;
        return true;
    }
    // Largest degree <= `want` for which p(lam1) / p(edge of the damped interval) stays below kMaxDynamicRange (lam1: the
    // eigenvalue of largest magnitude).  A one-sided Chebyshev filter of high degree over a WIDE wanted interval makes
    // ||p(A)|| exceed the last wanted Ritz value by many orders of magnitude; those Ritz values are then computed with an
    // absolute error relative to ||p(A)|| and the absolute tolerance of check_convergence (common.jl:56-65) loses its
    // meaning.  A wide wanted interval does not need a high degree.  (oracle: cap_degree)
    static constexpr double kMaxDynamicRange = 1e3;
    int cap_degree(double lam1, int want) const {
        const double x1 = e() > 0 ? std::fabs((lam1 - c()) / e()) : 1.0;
        int d = want;
        if (x1 > 1.0) d = std::min(d, (int)std::floor(std::acosh(kMaxDynamicRange) / std::acosh(x1)));
        d = std::max(d, 2);
        if (two_sided && !(d & 1)) ++d;
        return d;
    }
    // rho such that p(lam_k) = norm (the Ritz value of the last wanted pair gets the magnitude of ||op(A)||)
    void scale_to(double lam_k, double norm) {
        rho = 1.0;
        const double tk = std::fabs(eval(lam_k));
        const double xk = std::fabs((lam_k - c()) / e());
        rho = (tk > 0 && xk > 1.0) ? norm / tk : 1.0;
    }
};

struct CycleOut {
    bool converged = false;
    int64_t final_i = 0;         // blocks of this cycle that span the accepted / final Ritz pairs
    int64_t iterations_run = 0;  // block steps the device ran in this cycle
    TopKResult res;
};

struct Run {
    rbl_handle* h;
    const rbl_options& opt;
    cudaStream_t st;
    Workspace& w;
    int b = 0, B = 0;
    int64_t k = 0;
    bool fp32 = false;
    size_t ssz = 8;
    int64_t nloc = 0, next = 0, m_cap = 0, bstride = 0, kryl_sz = 0;
    bool use_h = false, use_d = false;
    float split_scale = 0.f;  // != 0: the Krylov slab holds split16 rows (split16.h) written with this scale
    int rgrid = 1;
    int reorth_period = 2, check_period = 4;
    bool async_ok = false, multi = false, is_root = true;
    int64_t launches = 0;
    PhaseTimer tm;
    double bytes_rgram = 0, bytes_rupd = 0, bytes_spmm = 0;
    int64_t n_rgram = 0, n_rupd = 0, n_spmm = 0;
    double *cur = nullptr, *prev = nullptr, *U = nullptr, *P1 = nullptr;
    SpmmCoef base;       // op(A) q = base.alpha * A q + base.beta * q
    FilterPlan flt;
    // host bookkeeping over the whole solve
    int checks = 0, full_checks = 0;
    int64_t host_factorizations = 0;
    double t_eig = 0.0, t_blocked = 0.0, t_idle = 0.0;
    int64_t total_steps = 0, total_steps_run = 0;
    cudaEvent_t tail_event = nullptr;

    Run(rbl_handle* hh) : h(hh), opt(hh->opt), st(hh->stream), w(hh->ws_ref()) {}
    ~Run() {
        if (tail_event) cudaEventDestroy(tail_event);
        for (auto e : post_event)
            if (e) cudaEventDestroy(e);
    }

    // small device matrices (B x B each): [G | P] are adjacent so that one reduction / all-reduce covers both
    double* G() { return w.small.p; }
    double* Pov() { return w.small.p + (size_t)B * B; }       // second Gram of a fused pass (A_i, or Q_i'U)
    double* Ai() { return w.small.p + 2 * (size_t)B * B; }
    double* Bp() { return w.small.p + 3 * (size_t)B * B; }
    double* Gloc() { return w.small.p + 4 * (size_t)B * B; }
    static constexpr int kSmallMats = 6;
    bool fused = false;
    int fgrid = 1;
    bool spmm_use_window = false;
    void* slot(int64_t j) { return w.buf.p + (size_t)j * bstride * ssz; }
    // ---- host tier of the Krylov slab (opts.spill; hybrid_part_reorth!, RBL_gpu.jl:59-81,127-130,168-169) ------------
    // slab slots [0, m_dev) live in HBM, slots [m_dev, m_cap) in pinned host memory.  Spilled blocks are written through
    // a device staging block and streamed back, kChunk blocks at a time (double-buffered on a copy stream), for every
    // Gram / update / Ritz pass.
    static constexpr int kChunk = 4;
    int64_t m_dev = 0;
    int64_t spilled_hi = 0;     // highest spilled slot written + 1
    size_t blk_bytes() const { return (size_t)bstride * ssz; }
    bool is_spilled(int64_t j) const { return j >= m_dev; }
    unsigned char* hslot(int64_t j) { return w.hslab.p + (size_t)(j - m_dev) * blk_bytes(); }
    unsigned char* stage_chunk(int which) { return w.stage.p + (size_t)which * kChunk * blk_bytes(); }
    unsigned char* stage_store(int which) { return w.stage.p + (size_t)(2 * kChunk + which) * blk_bytes(); }
    // where a kernel writes slab block j (j < 0: nowhere); flush_store afterwards moves a staged block to the host tier
    void* store_dst(int64_t j, int which) { return j < 0 ? nullptr : (is_spilled(j) ? (void*)stage_store(which) : slot(j)); }
    void flush_store(int64_t j, int which) {
        if (j < 0 || !is_spilled(j)) return;
        RBL_CUDA(cudaMemcpyAsync(hslot(j), stage_store(which), blk_bytes(), cudaMemcpyDeviceToHost, st));
        spilled_hi = std::max(spilled_hi, j + 1);
    }
    // for every kChunk-block piece [c0, c0+cn) of the spilled range [from, to): f(device pointer, c0, cn); the next
    // piece is copied in on the copy stream while f works on this one
    template <typename F>
    void for_spilled_chunks(int64_t from, int64_t to, F&& f) {
        if (to <= from) return;
        cudaStream_t cs = w.copy_stream;
        RBL_CUDA(cudaEventRecord(w.spill_ev[4], st));          // host tier writes issued so far on the main stream
        RBL_CUDA(cudaStreamWaitEvent(cs, w.spill_ev[4], 0));
        auto prefetch = [&](int64_t c0, int which) {
            const int64_t cn = std::min<int64_t>(kChunk, to - c0);
            RBL_CUDA(cudaStreamWaitEvent(cs, w.spill_ev[2 + which], 0));   // the kernel that last read this staging chunk
            RBL_CUDA(cudaMemcpyAsync(stage_chunk(which), hslot(c0), (size_t)cn * blk_bytes(), cudaMemcpyHostToDevice, cs));
            RBL_CUDA(cudaEventRecord(w.spill_ev[which], cs));
        };
        int which = 0;
        prefetch(from, 0);
        for (int64_t c0 = from; c0 < to; c0 += kChunk, which ^= 1) {
            const int64_t cn = std::min<int64_t>(kChunk, to - c0);
            if (c0 + kChunk < to) prefetch(c0 + kChunk, which ^ 1);
            RBL_CUDA(cudaStreamWaitEvent(st, w.spill_ev[which], 0));
            f((void*)stage_chunk(which), c0, cn);
            RBL_CUDA(cudaEventRecord(w.spill_ev[2 + which], st));
        }
    }

    void nccl(bool ok, const std::string& err) {
        if (!ok) throw Error(RBL_NCCL_ERROR, err);
    }
    void allreduce(double* p, size_t count) {
        if (!h->comm.active()) return;
        std::string err;
        nccl(h->comm.allreduce_f64(p, count, st, err), err);
    }
    // sum the per-CTA Gram partials into `out` (B*B) and across ranks
    void finish_gram(double* out) {
        launch_reduce_partials(w.part.p, rgrid, B * B, out, st);
        ++launches;
        allreduce(out, (size_t)B * B);
    }
    void rowop(const RowOpArgs& a) {
        launch_rowop(B, a, rgrid, st);
        ++launches;
    }
    // fused row kernel (rowops.cu) and the reduction of its [grid][2][B*B] partials into G() / Pov()
    void fpass(const FusedArgs& f) {
        launch_fused_rowop(B, f, fgrid, st);
        ++launches;
    }
    // which = 1: y'y -> G(); 2: z'y -> Pov(); 3: both
    void finish_fused(int which) {
        launch_reduce_partials(w.part.p, fgrid, 2 * B * B, G(), st);
        ++launches;
        if (which == 3) allreduce(G(), (size_t)2 * B * B);
        else if (which == 1) allreduce(G(), (size_t)B * B);
        else allreduce(Pov(), (size_t)B * B);
    }
    // ---- the four block passes of a step.  cur = Q_i, prev = Q_{i-1}, U = op(A) Q_i on entry ----
    // B: U -= Q_{i-1} B_{i-1}' ; A_i = Q_i' U                                      RBL_gpu.jl:177-178
    void pass_B(bool have_prev) {
        if (fused) {
            FusedArgs f;
            f.n = nloc; f.y = U; f.z = cur; f.partials = w.part.p; f.write_y = have_prev ? 1 : 0;
            if (have_prev) { f.x1 = prev; f.m1 = Bp(); f.m1_transposed = 1; }
            fpass(f);
            finish_fused(2);
            RBL_CUDA(cudaMemcpyAsync(Ai(), Pov(), (size_t)B * B * 8, cudaMemcpyDeviceToDevice, st));
        } else {
            RowOpArgs a;
            a.n = nloc; a.y = U; a.gram_z = cur; a.do_gram = 1; a.partials = w.part.p;
            if (have_prev) { a.x1 = prev; a.m1 = Bp(); a.m1_transposed = 1; a.write_y = 1; }
            rowop(a);
            finish_gram(Ai());
        }
    }
    // C: U -= Q_i A_i ; G = U'U                                                     :179 (+ Gram for the QR)
    void pass_C() {
        if (fused) {
            FusedArgs f;
            f.n = nloc; f.y = U; f.x1 = cur; f.m1 = Ai(); f.gram_yy = 1; f.partials = w.part.p;
            fpass(f);
            finish_fused(1);
        } else {
            RowOpArgs a;
            a.n = nloc; a.y = U; a.x1 = cur; a.m1 = Ai(); a.write_y = 1; a.do_gram = 1; a.partials = w.part.p;
            rowop(a);
            finish_gram(G());
        }
    }
    // D: U <- U R1^-1 ; G = U'U ; P = Q_i' U                                        CholQR pass 1 (:180-182)
    void pass_D() {
        if (fused) {
            FusedArgs f;
            f.n = nloc; f.y = U; f.rinv = w.qr.p->Rinv; f.gram_yy = 1; f.z = cur; f.partials = w.part.p;
            fpass(f);
            finish_fused(3);
        } else {
            RowOpArgs a;
            a.n = nloc; a.y = U; a.rinv = w.qr.p->Rinv; a.write_y = 1; a.do_gram = 1; a.partials = w.part.p;
            rowop(a);
            finish_gram(G());
            RowOpArgs a2;
            a2.n = nloc; a2.y = U; a2.gram_z = cur; a2.do_gram = 1; a2.partials = w.part.p;
            rowop(a2);
            finish_gram(Pov());
        }
    }
    // E: Q_{i+1} = U R2^-1 - Q_i (P R2^-1): CholQR pass 2 fused with the local re-orthogonalisation of the NEXT step
    // (loc_reorth_gpu! :83-93,:167) and the copy into the Krylov slab (:168-172); Gram only if a third pass is due
    void pass_E(void* store_slot) {
        if (fused) {
            FusedArgs f;
            f.n = nloc; f.y = U; f.rinv = w.qr.p->Rinv; f.x1 = cur; f.m1 = w.qr.p->Mloc;
            f.gram_yy = 1; f.gram_flag = &w.qr.p->need_more; f.partials = w.part.p;
            f.store = store_slot; f.store_fp32 = fp32; f.store_split_scale = split_scale;
            fpass(f);
            finish_fused(1);
        } else {
            RowOpArgs a;
            a.n = nloc; a.y = U; a.rinv = w.qr.p->Rinv; a.write_y = 1;
            rowop(a);
            RowOpArgs a2;
            a2.n = nloc; a2.y = U; a2.x1 = cur; a2.m1 = w.qr.p->Mloc; a2.write_y = 1; a2.do_gram = 1; a2.partials = w.part.p;
            a2.store = store_slot; a2.store_fp32 = fp32; a2.store_split_scale = split_scale;
            rowop(a2);
            finish_gram(G());
        }
    }
    // optional CholQR pass 3 (device flag need_more; no-ops otherwise).  Right-multiplication keeps Q_i' Q_{i+1} = 0.
    void pass_3(void* store_slot) {
        launch_chol(B, G(), w.qr.p, 3, h->n, 0, 1e-12, st);
        ++launches;
        if (fused) {
            FusedArgs f;
            f.n = nloc; f.y = U; f.rinv = w.qr.p->Rinv; f.skip_flag = &w.qr.p->need_more;
            f.store = store_slot; f.store_fp32 = fp32; f.store_split_scale = split_scale;
            fpass(f);
        } else {
            RowOpArgs a3;
            a3.n = nloc; a3.y = U; a3.rinv = w.qr.p->Rinv; a3.write_y = 1; a3.skip_flag = &w.qr.p->need_more;
            a3.store = store_slot; a3.store_fp32 = fp32; a3.store_split_scale = split_scale;
            rowop(a3);
        }
    }
    // one block step after the operator was applied: U = op(A) Q_i  ->  Q_{i+1} (in U), A_i, B_i
    void step_after_op(bool have_prev, int64_t store_j, int reset_ref) {
        void* store_slot = store_dst(store_j, 0);
        tm.mark(PH_3TERM);
        pass_B(have_prev);
        pass_C();
        tm.mark(PH_QR);
        launch_chol(B, G(), w.qr.p, 1, h->n, reset_ref, 1e-12, st);
        ++launches;
        pass_D();
        launch_chol(B, G(), w.qr.p, 2, h->n, 0, 1e-12, st, Pov());
        ++launches;
        tm.mark(PH_LOC);
        pass_E(store_slot);
        pass_3(store_slot);
        flush_store(store_j, 0);
    }

    void halo(double* Xblk, cudaStream_t hs = nullptr) {
        if (!h->comm.active()) return;
        if (!hs) hs = st;
        std::string err;
        const int64_t nsend = h->send_ptr[h->world];
        launch_gather_rows(B, nsend, w.d_send_rows.p, Xblk, w.sendbuf.p, hs);
        ++launches;
        nccl(h->comm.group_start(err), err);
        for (int p = 0; p < h->world; ++p) {
            if (p == h->rank) continue;
            const size_t sb = (size_t)(h->send_ptr[p + 1] - h->send_ptr[p]) * B * sizeof(double);
            const size_t rb = (size_t)(h->halo_owner_ptr[p + 1] - h->halo_owner_ptr[p]) * B * sizeof(double);
            nccl(h->comm.send_bytes(w.sendbuf.p + (size_t)h->send_ptr[p] * B, sb, p, hs, err), err);
            nccl(h->comm.recv_bytes(Xblk + (size_t)(nloc + h->halo_owner_ptr[p]) * B, rb, p, hs, err), err);
        }
        nccl(h->comm.group_end(err), err);
    }
    // U = cf.alpha * A Q + cf.beta * Q + cf.gamma * Z
    void spmm(double* Q, double* Uo, SpmmCoef cf, const double* Z) {
        if (h->comm.active() && h->n_bnd_rows > 0 && !spmm_use_window) {
            // the halo travels on the second stream while the rows that need none of it are computed; the rows that do
            // (flagged, skipped by the first launch: Z may alias U, so they must not be written twice) follow
            RBL_CUDA(cudaEventRecord(w.ev_q, st));
            RBL_CUDA(cudaStreamWaitEvent(w.comm_stream, w.ev_q, 0));
            halo(Q, w.comm_stream);
            RBL_CUDA(cudaEventRecord(w.ev_halo, w.comm_stream));
            launch_spmm(B, nloc, w.d_rowptr.p, w.d_colidx.p, w.d_vals.p, Q, Uo, cf, Z, st, nullptr, w.d_bnd_flag.p);
            RBL_CUDA(cudaStreamWaitEvent(st, w.ev_halo, 0));
            launch_spmm(B, h->n_bnd_rows, w.d_rowptr.p, w.d_colidx.p, w.d_vals.p, Q, Uo, cf, Z, st, w.d_bnd_rows.p, nullptr);
            launches += 2;
            ++n_spmm;
            bytes_spmm += 12.0 * (double)h->nnz + 4.0 * (double)(nloc + 1) + 16.0 * (double)nloc * B +
                          (cf.gamma != 0.0 ? 8.0 * (double)nloc * B : 0.0);
            return;
        }
        halo(Q);
        if (spmm_use_window)
            launch_spmm_window(B, nloc, nloc, w.d_rowptr.p, w.d_rel.p, w.d_vals.p, Q, Uo, cf, Z, h->spmm_wt, st);
        else
            launch_spmm(B, nloc, w.d_rowptr.p, w.d_colidx.p, w.d_vals.p, Q, Uo, cf, Z, st);
        ++launches;
        ++n_spmm;
        bytes_spmm += 12.0 * (double)h->nnz + 4.0 * (double)(nloc + 1) + 16.0 * (double)nloc * B +
                      (cf.gamma != 0.0 ? 8.0 * (double)nloc * B : 0.0);
    }
    // Uo = op(A) Q (plain) - mul!(U,Ag,Qg_d), RBL_gpu.jl:152,176
    void apply_plain(double* Q, double* Uo) { spmm(Q, Uo, base, nullptr); }
    // Uo = p(op(A)) Q by the Chebyshev three-term recurrence; Q is preserved, P1 is scratch.
    void apply_op(double* Q, double* Uo) {
        const int d = flt.degree;
        if (d == 0) {
            apply_plain(Q, Uo);
            return;
        }
        const double c = flt.c(), e = flt.e();
        SpmmCoef c1{base.alpha / e, (base.beta - c) / e, 0.0};
        SpmmCoef cj{2.0 * base.alpha / e, 2.0 * (base.beta - c) / e, -1.0};
        auto scaled = [&](SpmmCoef x) { x.alpha *= flt.rho; x.beta *= flt.rho; x.gamma *= flt.rho; return x; };
        // t_j lands alternately in A (odd j) and Bf (even j); the last one must land in Uo
        double* bufA = (d & 1) ? Uo : P1;
        double* bufB = (d & 1) ? P1 : Uo;
        spmm(Q, bufA, d == 1 ? scaled(c1) : c1, nullptr);                        // t1
        for (int j = 2; j <= d; ++j) {
            double* tj1 = (j & 1) ? bufB : bufA;      // t_{j-1}
            double* out = (j & 1) ? bufA : bufB;      // t_j overwrites t_{j-2} (aliasing Z is allowed), j = 2: Z = Q
            const double* z = (j == 2) ? Q : out;
            spmm(tj1, out, j == d ? scaled(cj) : cj, z);
        }
    }
    // thin QR of the block in `Ub` (in place).  The Gram Ub'Ub must already be in the rowop partials.
    void block_qr(double* Ub, int reset_ref) {
        const double defl_rel = 1e-12;
        finish_gram(G());
        launch_chol(B, G(), w.qr.p, 1, h->n, reset_ref, defl_rel, st);
        ++launches;
        RowOpArgs a;
        a.n = nloc; a.y = Ub; a.rinv = w.qr.p->Rinv; a.write_y = 1; a.do_gram = 1; a.partials = w.part.p;
        rowop(a);
        finish_gram(G());
        launch_chol(B, G(), w.qr.p, 2, h->n, 0, defl_rel, st);
        ++launches;
        rowop(a);  // apply pass 2, Gram for the optional pass 3
        finish_gram(G());
        launch_chol(B, G(), w.qr.p, 3, h->n, 0, defl_rel, st);
        ++launches;
        RowOpArgs a3;
        a3.n = nloc; a3.y = Ub; a3.rinv = w.qr.p->Rinv; a3.write_y = 1; a3.skip_flag = &w.qr.p->need_more;
        rowop(a3);
    }
    void gram_then_qr(double* Ub, int reset_ref) {
        RowOpArgs a;
        a.n = nloc; a.y = Ub; a.do_gram = 1; a.partials = w.part.p;
        rowop(a);
        block_qr(Ub, reset_ref);
    }

    // K5a: C = Qbuf[0..m)' * [w0 | w1] into w.Cmat (all-reduced over ranks); K5b: w -= Qbuf * C, optional refresh
    // of one slab block with the updated w1.  hybrid_part_reorth! / part_reorth_gpu_async!, RBL_gpu.jl:59-81,29-47
    // Gram of `mc` stored blocks at device address `buf` into rows [row0, row0+mc) of the coefficient matrix
    void gram_part(const void* buf, int64_t mc, int64_t row0, double* w0, double* w1) {
        ReorthPlan p = reorth_plan(B, fp32, nloc, mc);
        unsigned char* Cdst = w.Cmat.p + (size_t)row0 * B * 2 * B * ssz;
        if (use_d) {
            launch_reorth_gram_d(p, buf, bstride, w0, w1, w.rpart.p, Cdst, st);
            launches += 2;
        } else if (use_h) {
            launch_reorth_gram_h(p, h->n, buf, bstride, w0, w1, w.rpart.p, Cdst, w.tc_scratch.p, m_cap, split_scale != 0.f, st);
            launches += 4;
        } else {
            launch_reorth_gram(p, buf, bstride, w0, w1, w.rpart.p, Cdst, st);
            launches += 2;
        }
    }
    void reorth_gram(int64_t m, double* w0, double* w1) {
        const int64_t md = std::min(m, m_dev);
        if (md > 0) gram_part(w.buf.p, md, 0, w0, w1);
        for_spilled_chunks(md, m, [&](void* dev, int64_t c0, int64_t cn) { gram_part(dev, cn, c0, w0, w1); });
        if (multi) {
            std::string err;
            const size_t cnt = (size_t)m * B * 2 * B;
            nccl((fp32 ? h->comm.allreduce_f32((float*)w.Cmat.p, cnt, st, err) : h->comm.allreduce_f64((double*)w.Cmat.p, cnt, st, err)), err);
        }
        ++n_rgram;
        bytes_rgram += (double)ssz * (double)nloc * (double)m * B + 8.0 * (double)nloc * 2 * B;
    }
    void update_part(const void* buf, int64_t mc, int64_t row0, double* w0, double* w1, void* store_w1, void* store_w0) {
        ReorthPlan p = reorth_plan(B, fp32, nloc, mc);
        const unsigned char* Csrc = w.Cmat.p + (size_t)row0 * B * 2 * B * ssz;
        if (use_d) launch_reorth_update_d(p, buf, bstride, Csrc, w0, w1, store_w1, st, store_w0);
        else if (use_h) launch_reorth_update_h(p, h->n, buf, bstride, w0, w1, store_w1, w.tc_scratch.p, m_cap, split_scale != 0.f, st, store_w0, row0);
        else launch_reorth_update(p, buf, bstride, Csrc, w0, w1, store_w1, st, store_w0);
        ++launches;
    }
    // j_w1 / j_w0: slab slots whose stored copies are refreshed with the updated targets (-1: none)
    void reorth_update(int64_t m, double* w0, double* w1, int64_t j_w1, int64_t j_w0 = -1) {
        const int64_t md = std::min(m, m_dev);
        if (use_h) {
            ReorthPlan p = reorth_plan(B, fp32, nloc, m);
            launch_reorth_coeff_h(p, w.Cmat.p, w.tc_scratch.p, m_cap, (multi || md < m) ? 1 : 0, st);
            ++launches;
        }
        // the refresh must see the fully updated targets: it rides on the LAST partial update
        void* s1 = store_dst(j_w1, 1);
        void* s0 = store_dst(j_w0, 2);
        if (md > 0) update_part(w.buf.p, md, 0, w0, w1, md == m ? s1 : nullptr, md == m ? s0 : nullptr);
        for_spilled_chunks(md, m, [&](void* dev, int64_t c0, int64_t cn) {
            const bool last = c0 + cn >= m;
            update_part(dev, cn, c0, w0, w1, last ? s1 : nullptr, last ? s0 : nullptr);
        });
        flush_store(j_w1, 1);
        flush_store(j_w0, 2);
        ++n_rupd;
        bytes_rupd += (double)ssz * (double)nloc * (double)m * B + 2 * 8.0 * (double)nloc * 2 * B +
                      ((j_w1 >= 0 ? 1.0 : 0.0) + (j_w0 >= 0 ? 1.0 : 0.0)) * (double)ssz * (double)nloc * B;
    }

    // decision shared by all ranks: 0 continue, 1 accept, 2 abort (the root's host check failed); `step` rides along (the
    // block step the accepted check belongs to).  Slots [0,4) of the control buffers; the posted form uses [4, 4 + 4*kPostSlots).
    int agree(int local_code, int64_t* step = nullptr) {
        if (!multi) return local_code;
        DevBuf<double>& d_ctrl = w.ctrl;
        PinnedBuf<double>& h_ctrl = w.h_ctrl;
        h_ctrl.p[0] = is_root ? (double)local_code : 0.0;
        h_ctrl.p[1] = (is_root && step) ? (double)*step : 0.0;
        RBL_CUDA(cudaMemcpyAsync(d_ctrl.p, h_ctrl.p, 16, cudaMemcpyHostToDevice, st));
        allreduce(d_ctrl.p, 2);
        RBL_CUDA(cudaMemcpyAsync(h_ctrl.p + 2, d_ctrl.p, 16, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaStreamSynchronize(st));
        if (step) *step = (int64_t)std::llround(h_ctrl.p[3]);
        return (int)std::lround(h_ctrl.p[2]);
    }
    // The same agreement without draining the stream: posted now (the all-reduce is stream-ordered behind the steps issued
    // so far), read at the next check point, when it has long completed.
    static constexpr int kPostSlots = 4;
    cudaEvent_t post_event[kPostSlots] = {nullptr, nullptr, nullptr, nullptr};
    int post_next = 0, post_pending = -1;
    void post_agreement(int local_code, int64_t step) {
        const int slot = post_next;
        post_next = (post_next + 1) % kPostSlots;
        double* hsend = w.h_ctrl.p + 4 + 4 * slot;
        double* dbuf = w.ctrl.p + 4 + 2 * slot;
        hsend[0] = is_root ? (double)local_code : 0.0;
        hsend[1] = is_root ? (double)step : 0.0;
        RBL_CUDA(cudaMemcpyAsync(dbuf, hsend, 16, cudaMemcpyHostToDevice, st));
        allreduce(dbuf, 2);
        RBL_CUDA(cudaMemcpyAsync(hsend + 2, dbuf, 16, cudaMemcpyDeviceToHost, st));
        // (blocking sync: a rank that waits here - for the root, which is busy with a full check - sleeps instead of spinning on a
        // core the check could use)
        if (!post_event[slot]) RBL_CUDA(cudaEventCreateWithFlags(&post_event[slot], cudaEventDisableTiming | cudaEventBlockingSync));
        RBL_CUDA(cudaEventRecord(post_event[slot], st));
        post_pending = slot;
    }
    int read_agreement(int64_t* step) {
        const int slot = post_pending;
        post_pending = -1;
        RBL_CUDA(cudaEventSynchronize(post_event[slot]));
        const double* h = w.h_ctrl.p + 4 + 4 * slot;
        *step = (int64_t)std::llround(h[3]);
        return (int)std::lround(h[2]);
    }
    // the root's (d, s, resid) of `cols` pairs over Nrows rows of T, made identical on every rank
    void share_result(TopKResult& r, int64_t Nrows, int64_t cols) {
        if (!multi) return;
        const size_t cnt = (size_t)2 * cols + (size_t)Nrows * cols;
        std::vector<double> hbuf(cnt, 0.0);
        if (is_root) {
            std::copy(r.d.begin(), r.d.begin() + cols, hbuf.begin());
            std::copy(r.resid.begin(), r.resid.begin() + cols, hbuf.begin() + cols);
            std::copy(r.s.begin(), r.s.begin() + (size_t)Nrows * cols, hbuf.begin() + 2 * cols);
        }
        DevBuf<double>& dbuf = w.share;
        dbuf.ensure(cnt);
        RBL_CUDA(cudaMemcpyAsync(dbuf.p, hbuf.data(), cnt * 8, cudaMemcpyHostToDevice, st));
        allreduce(dbuf.p, cnt);
        RBL_CUDA(cudaMemcpyAsync(hbuf.data(), dbuf.p, cnt * 8, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaStreamSynchronize(st));
        r.N = Nrows;
        r.d.assign(hbuf.begin(), hbuf.begin() + cols);
        r.resid.assign(hbuf.begin() + cols, hbuf.begin() + 2 * cols);
        r.s.assign(hbuf.begin() + 2 * cols, hbuf.end());
        r.have_all = true;
    }

    CycleOut cycle(int64_t nlb, int64_t k_rem, int64_t kk_end, bool probe, int64_t max_steps);
    void ritz(int64_t nlb, int64_t mfin, const TopKResult& res, const std::vector<int64_t>& cols, void* Vdev, int64_t ldv,
              bool out_fp32, rbl_stats& stats);
    void rayleigh(double* V, int64_t ncols, std::vector<double>& lam, std::vector<double>& resn);
};

// Download of `ncols` columns of `width` bytes into PAGEABLE host memory (what a Julia Matrix or a NumPy array is).
// cudaMemcpy2D to pageable memory goes through the driver's staging buffer with one thread doing the host-side copy - and
// taking the page faults of a freshly allocated destination: 4.9 GB/s measured for the 800 MB V of config 2 (0.16 s of a
// 2.1 s call).  Here the copy is pipelined through two pinned chunks on the solve's stream while several host threads
// move the previous chunk into place.  Pinned / registered destinations and small copies take the plain 2-D copy.
static void download_columns(Workspace& w, cudaStream_t st, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                             int64_t ncols) {
    const size_t total = width * (size_t)ncols;
    cudaPointerAttributes attr{};
    const bool known = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered;
    cudaGetLastError();  // (older runtimes report an unregistered pointer as an error)
    static const bool staged_off = [] { const char* e = std::getenv("RBL_D2H_STAGED"); return e && e[0] == '0'; }();
    if (known || staged_off || total < ((size_t)32 << 20)) {
        RBL_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, (size_t)ncols, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaStreamSynchronize(st));
        return;
    }
    const size_t CH = (size_t)8 << 20;
    w.d2h_pin.ensure(2 * CH);
    for (auto& e : w.d2h_ev)
        if (!e) RBL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    struct Piece { size_t soff, doff, bytes; };
    std::vector<Piece> pcs;
    for (int64_t c = 0; c < ncols; ++c)
        for (size_t off = 0; off < width; off += CH) pcs.push_back({(size_t)c * spitch + off, (size_t)c * dpitch + off, std::min(CH, width - off)});
    const int P = (int)pcs.size();
    const int T = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    std::atomic<int> ready{0};                       // pieces [0, ready) are complete in their pinned chunk
    std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[(size_t)P]);
    for (int i = 0; i < P; ++i) done[i].store(0);
    unsigned char* const pin = w.d2h_pin.p;
    std::vector<std::thread> movers;
    for (int t = 0; t < T; ++t)
        movers.emplace_back([&, t]() {
            for (int j = 0; j < P; ++j) {
                while (ready.load(std::memory_order_acquire) <= j) std::this_thread::yield();
                const Piece& pc = pcs[(size_t)j];
                const size_t per = ((pc.bytes + (size_t)T - 1) / (size_t)T + 63) & ~(size_t)63;
                const size_t lo = std::min(pc.bytes, per * (size_t)t), hi = std::min(pc.bytes, per * (size_t)(t + 1));
                if (hi > lo) std::memcpy((unsigned char*)dst + pc.doff + lo, pin + (size_t)(j & 1) * CH + lo, hi - lo);
                done[j].fetch_add(1, std::memory_order_release);
            }
        });
    std::exception_ptr err;
    try {
        for (int i = 0; i < P; ++i) {
            if (i >= 2)
                while (done[i - 2].load(std::memory_order_acquire) < T) std::this_thread::yield();   // chunk free again
            RBL_CUDA(cudaMemcpyAsync(pin + (size_t)(i & 1) * CH, (const unsigned char*)src + pcs[(size_t)i].soff, pcs[(size_t)i].bytes,
                                     cudaMemcpyDeviceToHost, st));
            RBL_CUDA(cudaEventRecord(w.d2h_ev[i & 1], st));
            if (i >= 1) {
                RBL_CUDA(cudaEventSynchronize(w.d2h_ev[(i - 1) & 1]));
                ready.store(i, std::memory_order_release);
            }
        }
        RBL_CUDA(cudaEventSynchronize(w.d2h_ev[(P - 1) & 1]));
    } catch (...) {
        err = std::current_exception();
    }
    ready.store(P, std::memory_order_release);       // (on an error the movers run through and are joined)
    for (auto& th : movers) th.join();
    if (err) std::rethrow_exception(err);
}

// One Lanczos cycle (lanczos_iteration, RBL_gpu.jl:134-203) on the operator apply_op, starting from the orthonormal
// block in `cur`; slab slots [0, nlb) hold locked vectors, the cycle's blocks go to slots nlb, nlb+1, ...
//   probe      no convergence checks; run max_steps steps, then return the kk_end leading Ritz pairs of T
//   otherwise  checks every check_period steps for k_rem pairs; when the cap is reached the result holds the
//              kk_end leading pairs of the last T with their residual bounds
CycleOut Run::cycle(int64_t nlb, int64_t k_rem, int64_t kk_end, bool probe, int64_t max_steps) {
    CycleOut out;
    // ---- host-side T bookkeeping (insertA!/insertB!, common.jl:9-26) ---------------------------------
    BandSym T;
    T.reset(0, b);
    BandTopK checker;
    checker.threads = opt.host_threads > 0 ? opt.host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    checker.verbose = opt.verbose;
    int64_t t_blocks = 0;  // blocks already inserted into T
    std::vector<cudaEvent_t> step_event((size_t)m_cap + 2, nullptr);
    struct EventGuard {
        std::vector<cudaEvent_t>& v;
        ~EventGuard() {
            for (auto e : v)
                if (e) cudaEventDestroy(e);
        }
    } event_guard{step_event};
    auto grow_T = [&](int64_t upto_blocks) {
        // A_j for j < upto, B_j for j < upto-1 (B_i of the newest block is applied after the check, common.jl:113)
        const int W = 2 * b + 1;
        T.F.resize((size_t)upto_blocks * b * W, 0.0);
        T.N = upto_blocks * b;
        for (int64_t j = t_blocks; j < upto_blocks; ++j) {
            const double* A = w.hA.p + (size_t)j * B * B;
            for (int r = 0; r < b; ++r)
                for (int cc = 0; cc <= r; ++cc) T.set_sym(j * b + r, j * b + cc, A[r * B + cc]);
            if (j > 0) {
                const double* Bm = w.hB.p + (size_t)(j - 1) * B * B;  // couples block j-1 and j
                for (int cc = 0; cc < b; ++cc)
                    for (int m = 0; m <= cc; ++m) T.set_sym(j * b + m, (j - 1) * b + cc, Bm[m * B + cc]);
            }
        }
        t_blocks = upto_blocks;
        T.update_norm();
    };
    // Shadow tracker (rank 0): from the first checks on a background thread keeps computing
    // ALL k Ritz pairs of the latest T snapshot - by slicing the first time, by refining its own previous pairs
    // afterwards (BandTopK::refine_seeds) - so that the accepting check only has to refine fresh seeds instead of
    // solving the eigenproblem from scratch while the device sits idle.
    struct Shadow {
        std::mutex mu;
        std::condition_variable cv;
        std::thread th;
        bool active = false, stop = false, have_req = false, have_res = false;
        bool busy = false;                 // a pass is running ...
        int64_t busy_N = 0;                // ... on a snapshot of this size
        std::condition_variable cv_done;   // signalled at the end of every pass
        std::atomic<bool> cancel{false};   // raised with `stop`: the tracker abandons the eigensolve it is in
        std::atomic<bool> pause{false};    // raised while the main checker computes all k pairs: the tracker's threads sleep
        BandSym req;
        std::vector<double> req_bi;        // B_i of that snapshot: the tracker's pairs come back with their residual bounds
        bool have_gift = false;            // all k pairs of a full check of the main checker: fresher starting points than
        TopKResult gift;                   // the tracker's own previous pass
        TopKResult res;
    } shadow;
    checker.full_flag = &shadow.pause;
    // a full check without usable seeds waits for the tracker's pass in flight (if its snapshot is recent enough to help)
    // and takes its pairs, instead of pausing the tracker and solving from scratch
    checker.need_seeds = [&shadow, &checker, k_rem](int64_t Nnow) {
        std::unique_lock<std::mutex> lk(shadow.mu);
        if (!shadow.active) return;
        if (!shadow.have_res) {
            if (!shadow.busy || (double)shadow.busy_N < 0.75 * (double)Nnow) return;
            // (bounded: a pass takes tens to hundreds of milliseconds; the full check works without its pairs)
            shadow.cv_done.wait_for(lk, std::chrono::seconds(2), [&] { return !shadow.busy || shadow.stop; });
        }
        if (shadow.have_res) {
            checker.set_seeds(shadow.res.d, shadow.res.s, shadow.res.N, k_rem,
                              shadow.res.resid.size() == (size_t)k_rem ? &shadow.res.resid : nullptr);
            shadow.have_res = false;
        }
    };
    const int shadow_verbose = opt.verbose;
    const int bb = b;
    auto shadow_loop = [&shadow, k_rem, bb, shadow_verbose](int nthreads) {
        BandTopK tracker;
        tracker.threads = nthreads;
        for (;;) {
            BandSym Tc;
            std::vector<double> bic;
            TopKResult gift;
            bool got_gift = false;
            {
                std::unique_lock<std::mutex> lk(shadow.mu);
                shadow.cv.wait(lk, [&] { return shadow.stop || shadow.have_req; });
                if (shadow.stop) return;
                Tc = std::move(shadow.req);
                bic = shadow.req_bi;
                shadow.have_req = false;
                shadow.busy = true;
                shadow.busy_N = Tc.N;
                if (shadow.have_gift) {
                    gift = std::move(shadow.gift);
                    shadow.have_gift = false;
                    got_gift = true;
                }
            }
            if (got_gift) tracker.set_seeds(gift.d, gift.s, gift.N, k_rem);  // (ignored when older than its own)
            Tc.cancel = &shadow.cancel;
            Tc.pause = &shadow.pause;
            struct BusyGuard {   // whatever way the pass ends, nobody keeps waiting for it
                Shadow& s;
                ~BusyGuard() {
                    {
                        std::lock_guard<std::mutex> lk(s.mu);
                        s.busy = false;
                    }
                    s.cv_done.notify_all();
                }
            } busy_guard{shadow};
            TopKResult r;
            const double ts0 = now_s();
            const int64_t f0 = tracker.total_factorizations;
            try {
                // (tolerance 0: nothing counts as converged, the bounds are only reported)
                r = tracker.check(Tc, bic.size() == (size_t)bb * bb ? bic.data() : nullptr, bb, k_rem, 0.0, true);
            } catch (const Cancelled&) {
                return;
            } catch (...) {
                return;  // the tracker is an accelerator only: the checks work without its seeds
            }
            if (shadow_verbose > 1)
                std::fprintf(stderr, "[rbl] tracker N=%lld: %.1f ms, %lld factorisations, all pairs %d\n", (long long)Tc.N,
                             (now_s() - ts0) * 1e3, (long long)(tracker.total_factorizations - f0), (int)r.have_all);
            std::lock_guard<std::mutex> lk(shadow.mu);
            if (r.have_all) {
                shadow.res = std::move(r);
                shadow.have_res = true;
            }
        }
    };
    struct ShadowJoin {  // stops the tracker on every exit path
        Shadow& s;
        ~ShadowJoin() {
            {
                std::lock_guard<std::mutex> lk(s.mu);
                s.stop = true;
                s.cancel = true;
            }
            s.cv.notify_all();
            s.cv_done.notify_all();
            if (s.th.joinable()) s.th.join();
        }
    } shadow_join{shadow};
    auto run_check = [&](int64_t it, bool force_full, int64_t kwant) -> TopKResult {
        cudaSetDevice(h->device);
        cudaEventSynchronize(step_event[it]);
        const double t0 = now_s();
        grow_T(it);
        if (shadow.active && kwant == k_rem) {
            std::lock_guard<std::mutex> lk(shadow.mu);
            if (shadow.have_res) {
                checker.set_seeds(shadow.res.d, shadow.res.s, shadow.res.N, k_rem,
                                  shadow.res.resid.size() == (size_t)k_rem ? &shadow.res.resid : nullptr);
                shadow.have_res = false;
            }
        }
        std::vector<double> Bi((size_t)b * b);
        const double* Bm = w.hB.p + (size_t)(it - 1) * B * B;
        for (int r = 0; r < b; ++r)
            for (int cc = 0; cc < b; ++cc) Bi[(size_t)r * b + cc] = Bm[r * B + cc];
        TopKResult r = checker.check(T, Bi.data(), b, std::min<int64_t>(kwant, T.N), opt.tol, force_full);
        if (!r.converged && !shadow.active && !force_full && T.N >= 2 * k_rem) {
            shadow.active = true;
            shadow.th = std::thread(shadow_loop, std::max(1, checker.threads - 1));
        }
        if (!r.converged && shadow.active && !force_full) {
            {
                std::lock_guard<std::mutex> lk(shadow.mu);
                shadow.req = T;  // snapshot (a few MB)
                shadow.req_bi = Bi;
                shadow.have_req = true;
                if (r.have_all && kwant == k_rem && (int64_t)r.d.size() >= k_rem) {
                    shadow.gift.d = r.d;
                    shadow.gift.s = r.s;
                    shadow.gift.N = r.N;
                    shadow.have_gift = true;
                }
            }
            shadow.cv.notify_all();
        }
        t_eig += now_s() - t0;
        if (opt.verbose > 1)
            std::fprintf(stderr, "[rbl] check it=%lld N=%lld nfac=%d took %.2f ms conv=%d\n", (long long)it, (long long)T.N,
                         r.factorizations, (now_s() - t0) * 1e3, (int)r.converged);
        return r;
    };

    auto record_step = [&](int64_t it) {
        RBL_CUDA(cudaMemcpyAsync(w.hA.p + (size_t)(it - 1) * B * B, Ai(), (size_t)B * B * 8, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaMemcpyAsync(Bp(), w.qr.p->R, (size_t)B * B * 8, cudaMemcpyDeviceToDevice, st));
        RBL_CUDA(cudaMemcpyAsync(w.hB.p + (size_t)(it - 1) * B * B, w.qr.p->R, (size_t)B * B * 8, cudaMemcpyDeviceToHost, st));
    };
    auto mark_event = [&](int64_t it) {
        if (!step_event[it]) RBL_CUDA(cudaEventCreateWithFlags(&step_event[it], cudaEventDisableTiming));
        RBL_CUDA(cudaEventRecord(step_event[it], st));
    };

    // ---- first step (i = 1)                                                     RBL_gpu.jl:149-161 ----
    // Invariant of the loop below: `cur` = Q_i is already locally re-orthogonalised against Q_{i-1} and stored in
    // slab slot nlb+i-1 (both happen in pass E of the step that produced it).
    tm.mark(PH_LOC);
    launch_store_block(B, nloc, cur, store_dst(nlb, 0), fp32, split_scale, st);
    flush_store(nlb, 0);
    ++launches;
    tm.mark(PH_SPMM);
    apply_op(cur, U);
    step_after_op(false, (nlb + 1 < m_cap) ? nlb + 1 : -1, 1);
    record_step(1);
    tm.mark(PH_NONE);
    { double* t = prev; prev = cur; cur = U; U = t; }

    // ---- main loop                                                               RBL_gpu.jl:162-194 ----
    int64_t i = 1;
    bool converged = false;
    std::future<TopKResult> pending;
    int64_t pending_i = 0;
    bool check_in_flight = false;
    std::string root_error;
    // Adaptive check cadence.  The reference tests convergence every check_period-th step (RBL_gpu.jl:186); most of those
    // checks happen while the slowest wanted pair is orders of magnitude away from the tolerance.  From the residual
    // bound rho of the witness pair at two consecutive checks the root estimates its decay per step and postpones the
    // next check by at most HALF the steps that witness still needs (and at most 8 periods): the witness cannot have
    // converged by then, so no accepting check is skipped unless the convergence rate more than doubles.  The
    // acceptance rule itself is unchanged.  RBL_CHECK_ADAPTIVE=0 restores a check at every period.
    static const bool adaptive_checks = [] { const char* e = std::getenv("RBL_CHECK_ADAPTIVE"); return !(e && e[0] == '0'); }();
    double w_rho = -1.0, w_theta = 0.0;
    int64_t w_it = 0, next_check_i = 0;
    auto extra_periods = [&](const TopKResult& r, int64_t it) -> int {
        int extra = 0;
        if (adaptive_checks && r.witness_rho > 0 && w_rho > 0 && it > w_it && r.witness_rho < w_rho &&
            std::fabs(r.witness_theta - w_theta) <= 1e-3 * std::max(std::fabs(w_theta), 1e-300)) {   // (the same pair: Ritz values still drift early on)
            const double rate = std::log(w_rho / r.witness_rho) / (double)(it - w_it);      // decay of log(rho) per step
            if (rate > 0 && r.witness_rho > opt.tol) {
                const double steps_left = std::log(r.witness_rho / opt.tol) / rate;
                extra = (int)std::floor(0.5 * steps_left / check_period) - 1;
                extra = std::max(0, std::min(extra, 7));
            }
        }
        w_rho = r.witness_rho; w_theta = r.witness_theta; w_it = it;
        return extra;
    };
    // waits for the in-flight check; true when it accepted.  Row-sharded runs: only rank 0 evaluates the host
    // check; the decision code (and, on acceptance, D, S and the bounds) is summed over ranks with every other
    // rank contributing zeros, so all ranks follow the same control flow and use the same Ritz basis; a failure
    // of the root's check is broadcast as code 2 and every rank leaves with the same error instead of hanging.
    auto harvest = [&]() -> bool {
        if (!check_in_flight) return false;
        check_in_flight = false;
        TopKResult r;
        int code = 0;
        if (is_root) {
            const double t0 = now_s();
            double idle_from = -1.0;
            try {
                while (pending.wait_for(std::chrono::microseconds(200)) != std::future_status::ready)
                    if (idle_from < 0 && cudaEventQuery(tail_event) == cudaSuccess) idle_from = now_s();
                r = pending.get();
                code = r.converged ? 1 : 4 * extra_periods(r, pending_i);   // 0 / 1 / 2 in the low bits, postponement above
            } catch (const std::exception& e) {
                root_error = e.what();
                code = 2;
            }
            const double t1 = now_s();
            t_blocked += t1 - t0;
            if (idle_from >= 0) t_idle += t1 - idle_from;
            if (opt.verbose > 1) std::fprintf(stderr, "[rbl] harvest it=%lld blocked %.2f ms\n", (long long)pending_i, (t1 - t0) * 1e3);
        }
        ++checks;
        code = agree(code);
        const int extra = code / 4;
        code %= 4;
        next_check_i = pending_i + (int64_t)(1 + extra) * check_period;
        if (code >= 2) throw Error(RBL_BREAKDOWN, "rbl_solve: host eigen-check failed: " + (root_error.empty() ? std::string("(on rank 0)") : root_error));
        if (code == 1) {
            converged = true;
            out.final_i = pending_i;
            out.res = std::move(r);
            share_result(out.res, out.final_i * b, k_rem);
            return true;
        }
        return false;
    };

    // ---- non-blocking form (row-sharded runs, or async_check = 2) ---------------------------------------------------
    // With several GPUs a block step takes a fraction of a millisecond while a host check takes several: waiting for every
    // check (harvest) made the device idle for a third of an 8-GPU solve.  Here a check point never waits: the root polls
    // its worker; while a check is still running no new one is started; the outcome travels to the other ranks through an
    // agreement POSTED on the stream and read one check period later.  The device runs ahead of the accepting check by
    // however long that check takes; those speculative steps are discarded exactly like the blocking form's.  The step at
    // which a solve is accepted can therefore vary by a few check periods from run to run (never earlier than the
    // blocking form's, at most kMaxAhead + check_period later); every accepted result passed the same test.
    const bool nb_mode = async_ok && !probe && (multi || opt.async_check >= 2);
    bool have_accept = false;
    int64_t accept_i = 0;
    TopKResult accept_res;
    auto nb_accept = [&](int64_t step) {
        converged = true;
        out.final_i = step;
        if (is_root) out.res = std::move(accept_res);
        share_result(out.res, out.final_i * b, k_rem);
    };
    // root: collect a finished check / start a new one; returns the code to publish (0 nothing, 1 accepted, 2 failed)
    auto nb_poll = [&](int64_t it, bool may_start, bool block) -> int {
        if (!is_root) return 0;
        if (have_accept) return 1;
        if (check_in_flight && (block || pending.wait_for(std::chrono::seconds(0)) == std::future_status::ready)) {
            check_in_flight = false;
            try {
                const double t0 = now_s();
                double idle_from = -1.0;   // device idle = from the moment everything enqueued has run, as in harvest()
                if (block)
                    while (pending.wait_for(std::chrono::microseconds(200)) != std::future_status::ready)
                        if (idle_from < 0 && cudaEventQuery(tail_event) == cudaSuccess) idle_from = now_s();
                TopKResult r = pending.get();
                if (block) {
                    const double t1 = now_s();
                    t_blocked += t1 - t0;
                    if (idle_from >= 0) t_idle += t1 - idle_from;
                }
                ++checks;
                if (r.converged) {
                    have_accept = true;
                    accept_i = pending_i;
                    accept_res = std::move(r);
                    return 1;
                }
                next_check_i = pending_i + (int64_t)(1 + extra_periods(r, pending_i)) * check_period;
            } catch (const std::exception& e) {
                root_error = e.what();
                return 2;
            }
        }
        if (may_start && !check_in_flight && it >= next_check_i) {
            mark_event(it);
            pending_i = it;
            check_in_flight = true;
            pending = std::async(std::launch::async, [&, it]() { return run_check(it, false, k_rem); });
        }
        return 0;
    };
    auto nb_fail = [&]() { throw Error(RBL_BREAKDOWN, "rbl_solve: host eigen-check failed: " + (root_error.empty() ? std::string("(on rank 0)") : root_error)); };
    // check point of step `it`: true when the solve was accepted (by a check of an earlier step)
    auto nb_checkpoint = [&](int64_t it) -> bool {
        if (multi && post_pending >= 0) {
            int64_t step = 0;
            const int code = read_agreement(&step);
            if (code >= 2) nb_fail();
            if (code == 1) { nb_accept(step); return true; }
        }
        // bounded speculation: a check that started kMaxAhead steps ago is waited for (an accepting full check takes the
        // time of a hundred 8-GPU steps; running on would only burn slab slots and power on steps that get discarded)
        constexpr int64_t kMaxAhead = 16;     // (32 was measured on 2 GPUs: less idle, but a later accepted step and a longer drain - slower)
        const int code = nb_poll(it, true, check_in_flight && it - pending_i >= kMaxAhead);
        if (!multi) {
            if (code >= 2) nb_fail();
            if (code == 1) { nb_accept(accept_i); return true; }
            return false;
        }
        post_agreement(code, accept_i);
        return false;
    };
    // the loop ended at the cap without an acceptance seen so far: collect what is still in flight, this time waiting
    auto nb_drain = [&]() {
        if (multi && post_pending >= 0) {
            int64_t step = 0;
            const int code = read_agreement(&step);
            if (code >= 2) nb_fail();
            if (code == 1) { nb_accept(step); return; }
        }
        int code = nb_poll(i, false, true);
        // the check in flight belonged to an earlier step: the last check point of the cycle gets its own check (the waiting
        // form would have tested it), otherwise a solve that converged just before the cap would be reported as not converged
        const int64_t last = (i / check_period) * check_period;
        if (code == 0 && is_root && last > pending_i && last * b > k_rem) {
            try {
                mark_event(last);
                pending_i = last;
                const double t0 = now_s();
                TopKResult r = run_check(last, false, k_rem);
                t_idle += now_s() - t0;
                ++checks;
                if (r.converged) {
                    have_accept = true;
                    accept_i = last;
                    accept_res = std::move(r);
                    code = 1;
                }
            } catch (const std::exception& e) {
                root_error = e.what();
                code = 2;
            }
        }
        int64_t step = accept_i;
        code = agree(code, &step);
        if (code >= 2) nb_fail();
        if (code == 1) nb_accept(step);
    };

    while (i < max_steps && nlb + i < m_cap) {
        ++i;
        const int64_t m = nlb + i - 2;  // stored blocks the two newest are re-orthogonalised against
        if (i % reorth_period == 0 && m > 0) {
            // hybrid_part_reorth! (+ restart_reorth_gpu! for the locked blocks): project Q_i and Q_{i-1} against
            // everything stored before them, all at once (block CGS); both are already in the slab, so both stored
            // copies are refreshed (copyto!(Qgpu[i-1],Qg1), RBL_gpu.jl:76)
            tm.mark(PH_RGRAM);
            reorth_gram(m, cur, prev);
            tm.mark(PH_RUPD);
            reorth_update(m, cur, prev, nlb + i - 2, nlb + i - 1);
        }
        tm.mark(PH_SPMM);
        apply_op(cur, U);                                                      // :176
        step_after_op(true, (nlb + i < m_cap) ? nlb + i : -1, 0);   // :177-184 and :167-172 of the next step
        record_step(i);
        tm.mark(PH_NONE);
        { double* t = prev; prev = cur; cur = U; U = t; }

        if (!probe && i * b > k_rem && i % check_period == 0) {                 // :186
            RBL_CUDA(cudaEventRecord(tail_event, st));
            if (nb_mode) {
                if (nb_checkpoint(i)) break;
                continue;
            }
            if (harvest()) break;
            if (i < next_check_i) continue;      // postponed (adaptive cadence): no check at this step
            mark_event(i);
            pending_i = i;
            check_in_flight = true;
            const int64_t it = i;
            if (is_root) {
                if (async_ok) {
                    pending = std::async(std::launch::async, [&, it]() { return run_check(it, false, k_rem); });
                } else {
                    std::promise<TopKResult> pr;
                    const double t0 = now_s();
                    try {
                        pr.set_value(run_check(it, false, k_rem));
                    } catch (...) {
                        pr.set_exception(std::current_exception());
                    }
                    t_idle += now_s() - t0;   // synchronous order: the device waits for the whole check
                    pending = pr.get_future();
                }
            }
            if (!async_ok && harvest()) break;
        }
    }
    RBL_CUDA(cudaEventRecord(tail_event, st));
    if (!converged) {
        if (nb_mode) nb_drain();
        else harvest();
    }
    out.iterations_run = i;
    out.converged = converged;
    RBL_CUDA(cudaStreamSynchronize(st));
    if (!converged) {
        // cap reached (or probe finished): the kk_end leading pairs of the last T.  SURVEY Q4: the reference returns a
        // stale check or throws here; this is what restarts lock from and what NOT_CONVERGED returns as best effort.
        const int64_t it = i;
        if (it * b < std::min<int64_t>(k_rem, kk_end)) throw Error(RBL_INVALID, "rbl_solve: Krylov cap smaller than k, no Ritz pairs available");
        mark_event(it);
        int code = 0;
        if (is_root) {
            try {
                out.res = run_check(it, true, kk_end);
                if (!out.res.have_all) throw Error(RBL_BREAKDOWN, "could not isolate the leading Ritz pairs of T");
            } catch (const std::exception& e) {
                root_error = e.what();
                code = 2;
            }
        }
        if (agree(code) >= 2) throw Error(RBL_BREAKDOWN, "rbl_solve: host eigen-check failed: " + (root_error.empty() ? std::string("(on rank 0)") : root_error));
        share_result(out.res, it * b, std::min<int64_t>(kk_end, it * b));
        ++checks;
        out.final_i = it;
    }
    full_checks += checker.full_checks;
    host_factorizations += checker.total_factorizations;
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] cycle: %lld blocks (+%lld locked), ran %lld, converged %d; host checks: witness %d (%.3f s), bracketed %d (%.3f s), full %d (%.3f s); factorisations %lld (+%lld resumed)\n",
                     (long long)out.final_i, (long long)nlb, (long long)i, (int)converged, checker.stage_hits[0], checker.stage_sec[0],
                     checker.stage_hits[1], checker.stage_sec[1], checker.stage_hits[2], checker.stage_sec[2],
                     (long long)checker.total_factorizations, (long long)checker.resumed_factorizations);
    if (const char* dump = std::getenv("RBL_DUMP_T")) {
        // debugging aid: the A_i / B_i blocks of T (what the host checks saw), for tools/replay_dump.py
        if (is_root && dump[0]) {
            if (FILE* f = std::fopen(dump, "wb")) {
                const int64_t hdr[4] = {i, (int64_t)B, (int64_t)b, out.final_i};
                std::fwrite(hdr, sizeof(int64_t), 4, f);
                std::fwrite(w.hA.p, sizeof(double), (size_t)i * B * B, f);
                std::fwrite(w.hB.p, sizeof(double), (size_t)i * B * B, f);
                std::fclose(f);
            }
        }
    }
    return out;
}

// V[:, j] = Qbuf[nlb .. nlb+mfin) * S[:, cols[j]]   (recover_eigvec, RBL_gpu.jl:106-132; S narrowed like cu(), :119)
void Run::ritz(int64_t nlb, int64_t mfin, const TopKResult& res, const std::vector<int64_t>& cols, void* Vdev, int64_t ldv,
               bool out_fp32, rbl_stats& stats) {
    const int64_t kc = (int64_t)cols.size();
    if (kc == 0) return;
    const int kpad = (int)((kc + 15) / 16 * 16);
    std::vector<unsigned char> Sh((size_t)mfin * B * kpad * ssz, 0);
    for (int64_t j = 0; j < mfin; ++j)
        for (int cc = 0; cc < b; ++cc)
            for (int64_t t = 0; t < kc; ++t) {
                const double v = res.s[(size_t)cols[t] * res.N + (size_t)j * b + cc];
                const size_t idx = ((size_t)j * B + cc) * kpad + t;
                if (fp32) reinterpret_cast<float*>(Sh.data())[idx] = (float)v;
                else reinterpret_cast<double*>(Sh.data())[idx] = v;
            }
    DevBuf<unsigned char>& dS = w.ritzS;
    dS.ensure(Sh.size());
    RBL_CUDA(cudaMemcpyAsync(dS.p, Sh.data(), Sh.size(), cudaMemcpyHostToDevice, st));
    tm.mark(PH_RITZ);
    // blocks nlb .. nlb+mfin: the HBM part in one launch, spilled blocks chunk by chunk (accumulating), RBL_gpu.jl:116-130
    auto ritz_part = [&](const void* bufp, int64_t j0, int64_t mc, int accumulate) {
        const unsigned char* Sp = dS.p + (size_t)j0 * B * kpad * ssz;
        if (split_scale != 0.f) {
            w.ritz_words.ensure(ritz_h_scratch_words(B, mc, kpad));
            launch_ritz_h(B, nloc, mc, (int)kc, kpad, bufp, bstride, Sp, Vdev, ldv, out_fp32 ? 1 : 0, split_scale, w.ritz_words.p, st, accumulate);
            ++launches;
        } else {
            launch_ritz(B, fp32, nloc, mc, (int)kc, kpad, bufp, bstride, Sp, Vdev, ldv, out_fp32 ? 1 : 0, 0.f, st, accumulate);
        }
        ++launches;
    };
    const int64_t dev_blocks = std::max<int64_t>(0, std::min(nlb + mfin, m_dev) - nlb);
    if (dev_blocks > 0) ritz_part(slot(nlb), 0, dev_blocks, 0);
    for_spilled_chunks(nlb + dev_blocks, nlb + mfin, [&](void* dev, int64_t c0, int64_t cn) {
        ritz_part(dev, c0 - nlb, cn, (c0 - nlb) > 0 ? 1 : 0);
        if (split_scale != 0.f) RBL_CUDA(cudaStreamSynchronize(st));   // ritz_words is reused by the next piece
    });
    tm.mark(PH_NONE);
    RBL_CUDA(cudaStreamSynchronize(st));   // Sh goes out of scope
    stats.bytes_ritz += (double)ssz * (double)nloc * (double)mfin * B + (out_fp32 ? 4.0 : 8.0) * (double)nloc * (double)kc;
    stats.flops_ritz += 2.0 * (double)nloc * (double)mfin * B * (double)kc;
}

// Rayleigh quotients lam_c = v'op(A)v / v'v and residual norms ||op(A)v - lam v|| / ||v|| of `ncols` column-major
// fp64 vectors (leading dimension nloc), b columns at a time through the block kernels.
void Run::rayleigh(double* V, int64_t ncols, std::vector<double>& lam, std::vector<double>& resn) {
    lam.assign(ncols, 0.0);
    resn.assign(ncols, 0.0);
    std::vector<double> g0((size_t)B * B), g1((size_t)B * B), g2((size_t)B * B), m1((size_t)B * B);
    double* Xa = cur;
    double* Xb = U;
    for (int64_t c0 = 0; c0 < ncols; c0 += b) {
        const int nb = (int)std::min<int64_t>(b, ncols - c0);
        launch_colmajor_to_block(B, nloc, nb, V + (size_t)c0 * nloc, nloc, Xa, st);
        ++launches;
        apply_plain(Xa, Xb);
        RowOpArgs a0;
        a0.n = nloc; a0.y = Xa; a0.do_gram = 1; a0.partials = w.part.p;
        rowop(a0);
        finish_gram(G());
        RowOpArgs a1;
        a1.n = nloc; a1.y = Xb; a1.gram_z = Xa; a1.do_gram = 1; a1.partials = w.part.p;
        rowop(a1);
        finish_gram(Ai());
        RBL_CUDA(cudaMemcpyAsync(g0.data(), G(), g0.size() * 8, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaMemcpyAsync(g1.data(), Ai(), g1.size() * 8, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaStreamSynchronize(st));
        std::fill(m1.begin(), m1.end(), 0.0);
        for (int c = 0; c < nb; ++c) {
            const double vv = g0[(size_t)c * B + c];
            lam[c0 + c] = vv > 0 ? g1[(size_t)c * B + c] / vv : 0.0;
            m1[(size_t)c * B + c] = lam[c0 + c];
        }
        RBL_CUDA(cudaMemcpyAsync(Gloc(), m1.data(), m1.size() * 8, cudaMemcpyHostToDevice, st));
        RowOpArgs a2;
        a2.n = nloc; a2.y = Xb; a2.x1 = Xa; a2.m1 = Gloc(); a2.write_y = 1; a2.do_gram = 1; a2.partials = w.part.p;
        rowop(a2);
        finish_gram(G());
        RBL_CUDA(cudaMemcpyAsync(g2.data(), G(), g2.size() * 8, cudaMemcpyDeviceToHost, st));
        RBL_CUDA(cudaStreamSynchronize(st));
        for (int c = 0; c < nb; ++c) {
            const double vv = g0[(size_t)c * B + c];
            resn[c0 + c] = vv > 0 ? std::sqrt(std::max(0.0, g2[(size_t)c * B + c]) / vv) : 0.0;
        }
    }
}

void fill_stats(rbl_stats* s, Run& c, double* sec) {
    s->t_spmm = sec[PH_SPMM];
    s->t_3term = sec[PH_3TERM];
    s->t_qr = sec[PH_QR];
    s->t_loc_reorth = sec[PH_LOC];
    s->t_part_reorth = sec[PH_RGRAM] + sec[PH_RUPD];
    s->t_reorth_gram = sec[PH_RGRAM];
    s->t_reorth_update = sec[PH_RUPD];
    s->t_ritz = sec[PH_RITZ];
    s->t_ritz_kernel = sec[PH_RITZ];
    s->bytes_reorth_gram = c.bytes_rgram;
    s->bytes_reorth_update = c.bytes_rupd;
    s->bytes_part_reorth = c.bytes_rgram + c.bytes_rupd;
    s->bytes_spmm = c.bytes_spmm;
    s->launches_reorth_gram = c.n_rgram;
    s->launches_reorth_update = c.n_rupd;
    s->launches_spmm = c.n_spmm;
    s->kernel_launches = c.launches;
}

}  // namespace

int solve(rbl_handle* h, int64_t k, int64_t b_in, const SolveIO& io, double* d_out, rbl_stats* stats_out) {
    const double t_begin = now_s();
    rbl_stats stats;
    std::memset(&stats, 0, sizeof(stats));
    const rbl_options& opt = h->opt;
    if (k <= 0 || b_in <= 0 || b_in > 32) throw Error(RBL_INVALID, "rbl_solve: need k >= 1 and 1 <= b <= 32");
    if (k > h->n) throw Error(RBL_INVALID, "rbl_solve: k > n");
    if (!d_out || !io.v) throw Error(RBL_INVALID, "rbl_solve: null output");
    RBL_CUDA(cudaSetDevice(h->device));
    Run c(h);
    Workspace& w = c.w;
    c.b = (int)b_in; c.B = padded_block(c.b); c.k = k;
    c.fp32 = opt.precision == RBL_PRECISION_MIXED;
    c.ssz = c.fp32 ? 4 : 8;
    c.nloc = h->nloc; c.next = h->nloc + h->n_halo;
    c.bstride = c.nloc * c.B;
    c.tm.h = h; c.tm.st = c.st;
    c.multi = h->comm.active();
    c.is_root = (h->rank == 0);
    const int b = c.b, B = c.B;
    c.kryl_sz = std::max<int64_t>(opt.max_kryl_sz, b);
    int64_t m_req = (c.kryl_sz + b - 1) / b;
    c.reorth_period = std::max(1, opt.reorth_period);
    c.check_period = std::max(1, opt.check_period);
    // the asynchronous order equals the synchronous one only if the speculative steps never touch an accepted block:
    // checks fall on reorth steps and the refresh of slot(i-2) stays behind them (reorth_period >= 2)
    c.async_ok = opt.async_check && c.reorth_period >= 2 && (c.check_period % c.reorth_period == 0);
    int fdeg = opt.filter_degree < 0 ? 8 : opt.filter_degree;
    int filt_side = 0;          // +1 / -1: wanted pairs above / below the damped interval; 0: both ends
    double filt_norm = 0.0;     // |lambda_1| estimate: the filter is scaled so that p(lambda_k) ~ ||op(A)||
    double filt_lam1 = 0.0;     // lambda_1 estimate (signed)
    bool filt_settled = false;  // the cut has stopped moving: cycles get the whole buffer
    int filt_want_degree = 0;   // requested degree (the dynamic-range cap may hold the actual one below it)
    const bool filtering = fdeg > 0;
    const bool extra = filtering || opt.restart;
    c.base = (opt.op == RBL_OP_SHIFT_MINUS_A) ? SpmmCoef{-1.0, opt.sigma, 0.0} : SpmmCoef{1.0, 0.0, 0.0};
    RBL_CUDA(cudaEventCreateWithFlags(&c.tail_event, cudaEventDisableTiming));
    w.ctrl.ensure(4 + 2 * Run::kPostSlots);
    w.h_ctrl.ensure(4 + 4 * Run::kPostSlots);

    // ---- memory plan (gpu_buffer_size, RBL_gpu.jl:95-104: how many Krylov blocks fit) ---------------
    MemPlan plan = plan_memory(h, k, b, m_req);
    int64_t m_fit = plan.m_fit;
    if (c.multi) {
        // every rank must leave the iteration at the same step: agree on the smallest capacity
        DevBuf<int64_t> d_s, d_r;
        d_s.alloc(1);
        d_r.alloc(h->world);
        std::string err;
        RBL_CUDA(cudaMemcpyAsync(d_s.p, &m_fit, 8, cudaMemcpyHostToDevice, c.st));
        c.nccl(h->comm.allgather_i64(d_s.p, d_r.p, 1, c.st, err), err);
        std::vector<int64_t> all(h->world);
        RBL_CUDA(cudaMemcpyAsync(all.data(), d_r.p, all.size() * 8, cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaStreamSynchronize(c.st));
        m_fit = *std::min_element(all.begin(), all.end());
    }
    if (m_fit < 3) throw Error(RBL_OOM, "rbl_solve: problem does not fit device memory (fewer than 3 Krylov blocks)");
    int64_t m_cap = std::min(m_req, m_fit);
    int64_t m_dev = m_cap;
    const bool mem_capped = m_fit < m_req;
    if (mem_capped && opt.spill) {
        // host tier (hybrid_part_reorth!, RBL_gpu.jl:59-81): the staging buffers come out of the device budget, the
        // blocks that do not fit go to pinned host memory (at most half of the physical RAM)
        const int64_t stage_blocks = 2 * Run::kChunk + 3;
        if (m_fit - stage_blocks < 3) throw Error(RBL_OOM, "rbl_solve: device memory too small for the host-spill staging buffers");
        m_dev = m_fit - stage_blocks;
        const double host_budget = 0.5 * (double)sysconf(_SC_PHYS_PAGES) * (double)sysconf(_SC_PAGE_SIZE);
        const int64_t host_blocks = (int64_t)(host_budget / ((double)c.bstride * c.ssz));
        m_cap = std::min(m_req, m_dev + host_blocks);
    }
    if (mem_capped && !opt.restart && m_cap < m_req)
        std::fprintf(stderr, "[rbl] warning: Krylov buffer capped at %lld blocks by device memory (%lld requested); "
                             "set opts.restart to continue past the cap\n", (long long)m_cap, (long long)m_req);
    c.m_cap = m_cap;
    c.m_dev = m_dev;
    stats.buffer_blocks = m_dev;
    if (m_dev < m_cap) {
        w.stage.ensure((size_t)(2 * Run::kChunk + 3) * c.bstride * c.ssz);
        w.hslab.ensure((size_t)(m_cap - m_dev) * c.bstride * c.ssz);
        if (!w.copy_stream) RBL_CUDA(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
        for (auto& e : w.spill_ev)
            if (!e) RBL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int q = 0; q < 5; ++q) RBL_CUDA(cudaEventRecord(w.spill_ev[q], c.st));
        if (opt.verbose) std::fprintf(stderr, "[rbl] host spill tier: %lld blocks in HBM, up to %lld in pinned host memory\n", (long long)m_dev, (long long)(m_cap - m_dev));
    }
    c.rgrid = rowop_grid(B, c.nloc);
    {
        SpmmWindows wtmp = h->spmm_wt;
        size_t wsm = 0;
        c.spmm_use_window = wtmp.nwin > 0 && spmm_window_supported(B) && spmm_window_stages(wtmp, B, &wsm) > 0;
    }
    c.fused = fused_rowop_supported(B);
    c.fgrid = c.fused ? fused_rowop_grid(B, c.nloc) : 1;
    const int nX = 3 + (filtering ? 1 : 0);
    for (int i = 0; i < nX; ++i) w.X[i].ensure((size_t)c.next * B);
    w.buf.ensure((size_t)m_dev * c.bstride * c.ssz);
    w.part.ensure(std::max((size_t)c.rgrid, (size_t)2 * c.fgrid) * B * B);
    w.small.ensure(Run::kSmallMats * (size_t)B * B);
    w.qr.ensure(1);
    w.Cmat.ensure((size_t)m_cap * B * 2 * B * c.ssz);
    w.rpart.ensure(std::max<size_t>(1, reorth_max_partial_elems(B, c.fp32, c.nloc, m_cap)) * c.ssz);
    if (opt.reorth_impl >= 3 && !reorth_h_supported(B, c.fp32))
        throw Error(RBL_INVALID, "rbl_solve: FP16-split tensor-core reorth needs precision=mixed and padded block size 16 or 32");
    if (opt.reorth_impl == 2) throw Error(RBL_INVALID, "rbl_solve: reorth_impl 2 (3xTF32) was removed; use 0, 1, 3 or 4");
    c.use_h = reorth_h_supported(B, c.fp32) && opt.reorth_impl != 1;
    c.split_scale = (c.use_h && opt.reorth_impl != 3) ? reorth_h_scale(h->n) : 0.f;
    c.use_d = reorth_d_supported(B, c.fp32) && opt.reorth_impl != 1;
    if (c.use_h) w.tc_scratch.ensure(reorth_h_scratch_words(B, c.nloc, m_cap));
    if (c.multi) w.sendbuf.ensure(std::max<int64_t>(1, h->send_ptr[h->world]) * (size_t)B);
    w.hA.ensure((size_t)m_cap * B * B);
    w.hB.ensure((size_t)m_cap * B * B);
    w.hqr.ensure(1);
    if (extra) w.Vacc.ensure((size_t)c.nloc * k);
    const double t_alloc_done = now_s();
    RBL_CUDA(cudaMemsetAsync(w.qr.p, 0, sizeof(QrState), c.st));
    RBL_CUDA(cudaMemsetAsync(w.small.p, 0, Run::kSmallMats * (size_t)B * B * 8, c.st));
    for (int i = 0; i < nX; ++i) RBL_CUDA(cudaMemsetAsync(w.X[i].p, 0, (size_t)c.next * B * 8, c.st));
    h->last = KrylovInfo{};

    // ---- start block: Q1 = thin-Q of qr(A * Omega)                              RBL_gpu.jl:213-214 ----
    DevBuf<double>& d_omega = w.omega;
    const double* om_dev = nullptr;
    int64_t om_ld = c.nloc;
    {
        const double t0 = now_s();
        if (io.omega && io.omega_on_device) {
            om_dev = io.omega;
            om_ld = io.ld_omega;
        } else {
            d_omega.ensure((size_t)c.nloc * b);
            if (io.omega) {
                RBL_CUDA(cudaMemcpy2DAsync(d_omega.p, (size_t)c.nloc * 8, io.omega, (size_t)io.ld_omega * 8, (size_t)c.nloc * 8, b,
                                           cudaMemcpyHostToDevice, c.st));
            } else {
                const uint64_t seed = opt.seed ? (uint64_t)(uint32_t)opt.seed : 0x5eedull;
                for (int col = 0; col < b; ++col)
                    launch_randn(c.nloc, seed, (uint64_t)col * (uint64_t)h->n + (uint64_t)h->row0,
                                 d_omega.p + (size_t)col * c.nloc, c.st);
            }
            om_dev = d_omega.p;
            RBL_CUDA(cudaStreamSynchronize(c.st));
        }
        stats.t_h2d = (now_s() - t0) + h->t_h2d_create;
    }
    c.cur = w.X[1].p; c.prev = w.X[0].p; c.U = w.X[2].p; c.P1 = filtering ? w.X[3].p : nullptr;
    launch_colmajor_to_block(B, c.nloc, b, om_dev, om_ld, w.X[0].p, c.st);
    ++c.launches;
    c.tm.mark(PH_SPMM);
    c.apply_plain(w.X[0].p, c.cur);
    c.tm.mark(PH_QR);
    {
        c.gram_then_qr(c.cur, 1);
        // a start block of rank 0 (Omega = 0, or A*Omega = 0) spans no Krylov space: report it instead of
        // iterating on zero columns (the reference's Householder QR would continue with arbitrary unit vectors)
        RBL_CUDA(cudaMemcpyAsync(w.hqr.p, w.qr.p, sizeof(QrState), cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaStreamSynchronize(c.st));
        if (w.hqr.p->bad) throw Error(RBL_INVALID, "rbl_solve: A*Omega contains non-finite values");
        if (w.hqr.p->ndeflated >= B) throw Error(RBL_BREAKDOWN, "rbl_solve: the start block A*Omega has rank 0");
    }
    c.tm.mark(PH_NONE);

    // ---- filter placement: a short plain probe run locates the wanted end of the spectrum ------------------
    if (filtering) {
        const int64_t kk = k + b;
        int64_t s = opt.probe_steps > 0 ? opt.probe_steps : std::max<int64_t>(8, 2 * ((kk + b - 1) / b));
        s = std::min<int64_t>(s, m_cap);
        if (s * b < k) throw Error(RBL_INVALID, "rbl_solve: Krylov cap too small for the filter probe");
        // keep Q1: the probe rotates the active blocks
        RBL_CUDA(cudaMemcpyAsync(c.P1, c.cur, (size_t)c.nloc * B * 8, cudaMemcpyDeviceToDevice, c.st));
        CycleOut pr = c.cycle(0, k, std::min<int64_t>(kk, s * b), true, s);
        const TopKResult& r = pr.res;
        const int64_t have = (int64_t)r.d.size();
        const int64_t kq = std::min<int64_t>(k, have);
        FilterPlan f;
        f.degree = fdeg;
        const double cut = std::fabs(r.d[have - 1]);
        bool allpos = true, allneg = true;
        for (int64_t j = 0; j < have; ++j) { allpos &= r.d[j] > 0; allneg &= r.d[j] < 0; }
        const double s_lo = c.base.alpha > 0 ? c.base.alpha * h->gersh_lo + c.base.beta : c.base.alpha * h->gersh_hi + c.base.beta;
        const double s_hi = c.base.alpha > 0 ? c.base.alpha * h->gersh_hi + c.base.beta : c.base.alpha * h->gersh_lo + c.base.beta;
        if (allpos) { f.a = s_lo < cut ? s_lo : cut - std::fabs(cut); f.b = cut; }
        else if (allneg) { f.a = -cut; f.b = s_hi > -cut ? s_hi : -cut + std::fabs(cut); }
        else { f.two_sided = true; f.a = -cut; f.b = cut; if (!(f.degree & 1)) ++f.degree; }
        if (!(f.e() > 0)) throw Error(RBL_BREAKDOWN, "rbl_solve: filter probe found a degenerate spectrum interval");
        // (no dynamic-range cap here: the probe's cut lies far below the wanted end, the cap would be computed from a range
        // p never sees among the wanted pairs; it applies from the first re-placement on, where the estimates are good)
        filt_want_degree = f.degree;
        f.scale_to(r.d[kq - 1], std::fabs(r.d[0]));
        c.flt = f;
        filt_side = allpos ? 1 : (allneg ? -1 : 0);
        filt_norm = std::fabs(r.d[0]);
        filt_lam1 = r.d[0];
        stats.filter_cut = cut;
        stats.filter_degree = f.degree;
        stats.filter_two_sided = f.two_sided ? 1 : 0;
        if (opt.verbose)
            std::fprintf(stderr, "[rbl] filter: probe %lld steps, degree %d, damped [%.6g, %.6g], rho %.3e, %s\n", (long long)pr.iterations_run,
                         f.degree, f.a, f.b, f.rho, f.two_sided ? "two-sided" : "one-sided");
        c.total_steps_run += pr.iterations_run;
        // restart from Q1
        c.cur = w.X[1].p; c.prev = w.X[0].p; c.U = w.X[2].p;
        RBL_CUDA(cudaMemcpyAsync(c.cur, c.P1, (size_t)c.nloc * B * 8, cudaMemcpyDeviceToDevice, c.st));
    }

    // ---- cycles: iterate; at the cap lock what converged and restart (restarted.jl:98-146) ---------------
    int64_t nlock = 0;                 // locked Ritz vectors (columns of Vacc)
    std::vector<double> d_lock;
    int status = RBL_OK;
    CycleOut last;
    int64_t nlb = 0;
    int64_t cycles = 0;
    const int64_t max_cycles = 50;
    bool converged = false;
    double t_loop_done = now_s();
    for (;;) {
        ++cycles;
        const int64_t k_rem = k - nlock;
        const int64_t kk_end = opt.restart ? k_rem + b : k_rem;
        int64_t max_steps = std::max<int64_t>(1, (c.kryl_sz + b - 1) / b - nlb);
        // filtered restarts: short "settling" cycles while the filter is still being re-placed (below)
        if (filtering && opt.restart && !filt_settled) max_steps = std::min<int64_t>(max_steps, std::max<int64_t>(8, 3 * ((kk_end + b - 1) / b)));
        last = c.cycle(nlb, k_rem, kk_end, false, max_steps);
        c.total_steps += last.final_i;
        c.total_steps_run += last.iterations_run;
        if (last.converged) { converged = true; break; }
        if (!opt.restart || cycles >= max_cycles) { status = RBL_NOT_CONVERGED; break; }
        // ---- lock + restart -------------------------------------------------------------------------------
        const TopKResult& r = last.res;
        const int64_t have = (int64_t)r.d.size();
        std::vector<int64_t> lock, rest;
        for (int64_t j = 0; j < have; ++j) {
            if (j < k_rem && r.resid[j] <= opt.tol) lock.push_back(j);
            else if ((int64_t)rest.size() < b) rest.push_back(j);
        }
        if (filtering) {
            // Filtered restarts never lock.  p amplifies the leading (first converged) eigenvalues far more than the last
            // wanted ones - by 1e4 per application for a well placed degree-32 filter - so any imperfection of locked vectors
            // re-grows inside the cycle and comes back as ghost Ritz pairs (observed on 8 GPUs: BASELINE config 5 returned
            // pairs with residual 2e-5 ||A||; the CPU twin reproduces it on small grids).  Instead the cycle is repeated with
            // a better filter from the leading b unconverged Ritz vectors: the filter is re-placed from this cycle's Ritz
            // values mapped back through p (short settling cycles until the cut stops moving - a full-length cycle behind a
            // badly placed filter is wasted), then, if a full-length cycle still does not converge, the degree doubles
            // (within the dynamic-range cap).  oracle/rbl_restart_oracle.py RBL_restarted is the same procedure.
            lock.clear();
            rest.clear();
            for (int64_t j = 0; j < have && (int64_t)rest.size() < b; ++j)
                if (!(j < k_rem && r.resid[j] <= opt.tol)) rest.push_back(j);
            double lam_last = 0.0, lam_k = 0.0;
            const int64_t kq = std::min<int64_t>(k_rem, have);
            const bool ok_last = have > 0 && c.flt.invert(r.d[have - 1], filt_side, &lam_last);
            bool ok_k = have > 0 && c.flt.invert(r.d[kq - 1], filt_side, &lam_k);
            FilterPlan f = c.flt;
            const double cut_old = f.two_sided ? f.b : (filt_side > 0 ? f.b : -f.a);
            bool moved = false;
            if (ok_last && ok_k) {
                const double cut_new = std::fabs(lam_last);
                // (the cut must stay below the estimate of the last wanted eigenvalue - itself a lower bound of it - or
                // wanted eigenvalues would be damped; inconsistent estimates leave the filter alone)
                if (cut_new > cut_old + 1e-2 * std::max(filt_norm - cut_old, 0.0) && cut_new < std::fabs(lam_k)) {
                    if (f.two_sided) { f.a = -cut_new; f.b = cut_new; }
                    else if (filt_side > 0) f.b = cut_new;
                    else f.a = -cut_new;
                    f.degree = filt_want_degree;
                    moved = true;
                }
            }
            bool changed = moved;
            if (!moved && !filt_settled) {
                filt_settled = true;          // same filter, whole buffer
            } else if (!moved) {
                filt_want_degree = std::min(2 * f.degree, 256);
                f.degree = filt_want_degree;
                changed = true;
            }
            if (changed) {
                f.degree = f.cap_degree(filt_lam1, f.degree);
                if (f.degree == c.flt.degree && f.a == c.flt.a && f.b == c.flt.b) {
                    status = RBL_NOT_CONVERGED;        // neither the cut nor the degree can move any more: best effort
                    break;
                }
                if (!ok_k) lam_k = f.c() + f.e() * (1.0 + 1e-3) * (filt_side >= 0 ? 1.0 : -1.0);
                f.scale_to(lam_k, filt_norm);
                if (opt.verbose)
                    std::fprintf(stderr, "[rbl] filter %s: degree %d, damped [%.8g, %.8g] (cut %.8g -> %.8g), rho %.3e\n", moved ? "re-placed" : "degree raised",
                                 f.degree, f.a, f.b, cut_old, f.two_sided ? f.b : (filt_side > 0 ? f.b : -f.a), f.rho);
                c.flt = f;
                stats.filter_cut = f.two_sided ? f.b : (filt_side > 0 ? f.b : -f.a);
                stats.filter_degree = f.degree;
            } else if (opt.verbose) {
                std::fprintf(stderr, "[rbl] filter settled: degree %d, damped [%.8g, %.8g]; the next cycle gets the whole buffer\n", f.degree, f.a, f.b);
            }
        }
        {   // is there room for another cycle once these are locked?  If not: best effort from this cycle, as at the cap
            const int64_t nlb_next = (nlock + (int64_t)lock.size() + b - 1) / b;
            const bool done = nlock + (int64_t)lock.size() >= k;
            if (!done && (m_cap - nlb_next < 3 || (c.kryl_sz + b - 1) / b - nlb_next < 3)) { status = RBL_NOT_CONVERGED; break; }
        }
        std::vector<int64_t> cols = lock;
        cols.insert(cols.end(), rest.begin(), rest.end());
        w.ritzV.ensure((size_t)c.nloc * (size_t)std::max<int64_t>(((int64_t)lock.size() + b) * 8, k * (opt.v_fp32 ? 4 : 8)));
        double* Vtmp = reinterpret_cast<double*>(w.ritzV.p);
        // with the tensor-core Ritz kernel on a split16 slab or an fp32 slab the vectors are fp32-grade, as in the reference
        c.ritz(nlb, last.final_i, r, cols, Vtmp, c.nloc, false, stats);
        const int64_t nl = (int64_t)lock.size();
        if (nl) RBL_CUDA(cudaMemcpyAsync(w.Vacc.p + (size_t)nlock * c.nloc, Vtmp, (size_t)nl * c.nloc * 8, cudaMemcpyDeviceToDevice, c.st));
        for (int64_t j = 0; j < nl; ++j) d_lock.push_back(r.d[lock[j]]);
        // the slab head holds the locked vectors in whole blocks of b columns (zero padded)
        const int64_t first_blk = nlock / b;
        nlock += nl;
        const int64_t nlb_new = (nlock + b - 1) / b;
        for (int64_t blk = first_blk; blk < nlb_new; ++blk) {
            const int ncol = (int)std::min<int64_t>(b, nlock - blk * b);
            launch_colmajor_to_block(B, c.nloc, ncol, w.Vacc.p + (size_t)blk * b * c.nloc, c.nloc, c.U, c.st);
            launch_store_block(B, c.nloc, c.U, c.store_dst(blk, 0), c.fp32, c.split_scale, c.st);
            c.flush_store(blk, 0);
            c.launches += 2;
        }
        nlb = nlb_new;
        stats.restarts = cycles;
        stats.locked = nlock;
        if (opt.verbose)
            std::fprintf(stderr, "[rbl] restart %lld: locked %lld (+%lld), %lld blocks of the slab hold locked vectors\n", (long long)cycles,
                         (long long)nlock, (long long)nl, (long long)nlb);
        if (nlock >= k) { converged = true; last.final_i = 0; break; }
        // restart block: the best b not-locked Ritz vectors (missing columns: fresh random directions)
        c.cur = w.X[1].p; c.prev = w.X[0].p; c.U = w.X[2].p;
        const int nr = (int)rest.size();
        if (nr < b) {
            for (int col = nr; col < b; ++col)
                launch_randn(c.nloc, 0x5eedull + (uint64_t)cycles, (uint64_t)col * (uint64_t)h->n + (uint64_t)h->row0,
                             Vtmp + (size_t)(nl + col) * c.nloc, c.st);
        }
        launch_colmajor_to_block(B, c.nloc, b, Vtmp + (size_t)nl * c.nloc, c.nloc, c.cur, c.st);
        ++c.launches;
        RBL_CUDA(cudaMemsetAsync(c.prev, 0, (size_t)c.next * B * 8, c.st));
        if (nlb > 0) {   // restart_reorth_gpu! (restarted.jl:1-21): start block orthogonal to the locked vectors
            c.reorth_gram(nlb, c.cur, c.prev);
            c.reorth_update(nlb, c.cur, c.prev, -1);
        }
        c.gram_then_qr(c.cur, 1);
    }
    t_loop_done = now_s();
    stats.restarts = cycles - 1;
    stats.locked = nlock;

    const double t_final_done = now_s();
    // ---- Ritz vectors V = Qbuf * S                                              RBL_gpu.jl:106-132,219 ----
    const size_t vsz = opt.v_fp32 ? 4 : 8;
    const int64_t mfin = last.final_i;
    const int64_t k_fin = k - nlock;
    std::vector<double> d_all(d_lock);
    for (int64_t t = 0; t < k_fin; ++t) d_all.push_back(last.res.d.size() > (size_t)t ? last.res.d[t] : 0.0);
    std::vector<int64_t> fin_cols((size_t)k_fin);
    std::iota(fin_cols.begin(), fin_cols.end(), 0);
    if (!extra) {
        // the reference's single-cycle path: the Ritz kernel writes straight into the caller's V
        void* Vdev = io.v;
        int64_t ldv = io.ldv;
        if (!io.v_on_device) {
            w.ritzV.ensure((size_t)c.nloc * k * vsz);
            Vdev = w.ritzV.p;
            ldv = c.nloc;
        }
        c.ritz(0, mfin, last.res, fin_cols, Vdev, ldv, opt.v_fp32 != 0, stats);
        if (!io.v_on_device) {
            const double t0 = now_s();
            download_columns(w, c.st, io.v, (size_t)io.ldv * vsz, w.ritzV.p, (size_t)c.nloc * vsz, (size_t)c.nloc * vsz, k);
            stats.t_d2h = now_s() - t0;
        }
        for (int64_t t = 0; t < k; ++t) d_out[t] = d_all[t];
    } else {
        if (k_fin > 0 && mfin > 0) c.ritz(nlb, mfin, last.res, fin_cols, w.Vacc.p + (size_t)nlock * c.nloc, c.nloc, false, stats);
        std::vector<double> lam = d_all, resn;
        if (filtering) {
            // eigenvalues of op(A) from the Ritz vectors of p(op(A)); true residuals measured in fp64
            c.cur = w.X[1].p; c.U = w.X[2].p;
            c.rayleigh(w.Vacc.p, k, lam, resn);
            stats.max_residual = *std::max_element(resn.begin(), resn.end());
        }
        std::vector<int64_t> perm((size_t)k);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int64_t x, int64_t y) { return std::fabs(lam[x]) > std::fabs(lam[y]); });
        for (int64_t t = 0; t < k; ++t) d_out[t] = lam[perm[t]];
        // output column t = Vacc column perm[t] (descending |lambda|, RBL.jl:116)
        const double t0 = now_s();
        if (!opt.v_fp32) {
            for (int64_t t = 0; t < k; ++t)
                RBL_CUDA(cudaMemcpyAsync((char*)io.v + (size_t)t * io.ldv * 8, w.Vacc.p + (size_t)perm[t] * c.nloc, (size_t)c.nloc * 8,
                                         io.v_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.st));
        } else {
            w.ritzV.ensure((size_t)c.nloc * k * 4);
            for (int64_t t = 0; t < k; ++t) {
                launch_convert_s(c.nloc, w.Vacc.p + (size_t)perm[t] * c.nloc, w.ritzV.p + (size_t)t * c.nloc * 4, 1, c.st);
                RBL_CUDA(cudaMemcpyAsync((char*)io.v + (size_t)t * io.ldv * 4, w.ritzV.p + (size_t)t * c.nloc * 4, (size_t)c.nloc * 4,
                                         io.v_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.st));
            }
        }
        RBL_CUDA(cudaStreamSynchronize(c.st));
        if (!io.v_on_device) stats.t_d2h = now_s() - t0;
    }
    RBL_CUDA(cudaMemcpy(w.hqr.p, w.qr.p, sizeof(QrState), cudaMemcpyDeviceToHost));
    h->last.B = B; h->last.b = b; h->last.blocks = nlb + mfin; h->last.fp32 = c.fp32; h->last.split_scale = c.split_scale;
    h->last.bstride = c.bstride; h->last.ssz = c.ssz; h->last.use_h = c.use_h; h->last.use_d = c.use_d; h->last.m_cap = m_cap;
    h->last.m_dev = m_dev;
    stats.spilled_blocks = std::max<int64_t>(0, c.spilled_hi - m_dev);

    double sec[PH_COUNT];
    c.tm.collect(sec);
    fill_stats(&stats, c, sec);
    stats.iterations = c.total_steps;
    stats.kryl_sz = c.total_steps * b;
    stats.iterations_run = c.total_steps_run;
    stats.converged = converged ? 1 : 0;
    stats.checks = c.checks;
    stats.full_checks = c.full_checks;
    stats.host_factorizations = c.host_factorizations;
    stats.deflated = w.hqr.p->ndeflated - (B - b);
    stats.t_eig = c.t_eig;
    stats.t_eig_wait = c.t_idle;
    stats.t_host_blocked = c.t_blocked;
    stats.t_total = now_s() - t_begin;
    if (w.hqr.p->bad) status = RBL_BREAKDOWN;
    if (stats_out) *stats_out = stats;
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] timeline: alloc %.3f  start+loop %.3f  ritz+d2h %.3f (d2h %.3f, h2d %.3f)\n",
                     t_alloc_done - t_begin, t_loop_done - t_alloc_done, now_s() - t_final_done, stats.t_d2h, stats.t_h2d);
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] Iterations: %lld and kryl_sz: %lld (ran %lld), cycles %lld, locked %lld, checks %d (full %d), t=%.3fs eig=%.3fs device-idle=%.3fs\n",
                     (long long)c.total_steps, (long long)(c.total_steps * b), (long long)c.total_steps_run, (long long)cycles, (long long)nlock,
                     c.checks, c.full_checks, stats.t_total, c.t_eig, c.t_idle);
    return status;
}

// ------------------------------------------------------------------------------------------------ basis exports
void krylov_block(rbl_handle* h, int64_t j, double* out_colmajor) {
    const KrylovInfo& L = h->last;
    if (L.blocks <= 0) throw Error(RBL_INVALID, "rbl_krylov_block: no solve has run on this handle");
    if (j < 0 || j >= L.blocks) throw Error(RBL_INVALID, "rbl_krylov_block: block index out of range");
    RBL_CUDA(cudaSetDevice(h->device));
    Workspace& w = h->ws_ref();
    const int64_t nloc = h->nloc;
    double* blk = w.X[0].p;
    const unsigned char* src = w.buf.p + (size_t)j * L.bstride * L.ssz;
    if (j >= L.m_dev) {   // spilled block: through the staging buffer
        RBL_CUDA(cudaMemcpyAsync(w.stage.p, w.hslab.p + (size_t)(j - L.m_dev) * L.bstride * L.ssz, (size_t)L.bstride * L.ssz,
                                 cudaMemcpyHostToDevice, h->stream));
        src = w.stage.p;
    }
    launch_decode_block(L.B, nloc, src, L.fp32, L.split_scale, blk, h->stream);
    DevBuf<double> cm;
    cm.alloc((size_t)nloc * L.b);
    launch_block_to_colmajor(L.B, nloc, L.b, blk, cm.p, nloc, h->stream);
    RBL_CUDA(cudaStreamSynchronize(h->stream));
    RBL_CUDA(cudaMemcpy(out_colmajor, cm.p, (size_t)nloc * L.b * 8, cudaMemcpyDeviceToHost));
}

// max |(Q'Q - I)_ij| and ||Q'Q - I||_F over the stored basis, with the Gram kernels of the re-orthogonalisation
void orthogonality(rbl_handle* h, double* max_abs, double* fro) {
    const KrylovInfo& L = h->last;
    if (L.blocks <= 0) throw Error(RBL_INVALID, "rbl_orthogonality: no solve has run on this handle");
    if (L.blocks > L.m_dev) throw Error(RBL_INVALID, "rbl_orthogonality: not available when Krylov blocks were spilled to the host (use rbl_krylov_block)");
    RBL_CUDA(cudaSetDevice(h->device));
    Run c(h);
    Workspace& w = c.w;
    c.B = L.B; c.b = L.b; c.fp32 = L.fp32; c.ssz = L.ssz; c.nloc = h->nloc; c.next = h->nloc + h->n_halo;
    c.bstride = L.bstride; c.split_scale = L.split_scale; c.use_h = L.use_h; c.use_d = L.use_d; c.m_cap = L.m_cap;
    c.multi = h->comm.active();
    c.is_root = h->rank == 0;
    c.m_dev = L.m_dev;
    c.tm.enabled = false;
    c.tm.h = h; c.tm.st = c.st;
    DevBuf<double> acc;
    acc.alloc(2);
    RBL_CUDA(cudaMemsetAsync(acc.p, 0, 16, c.st));
    const int64_t m = L.blocks;
    if (!L.fp32) {
        for (int64_t j0 = 0; j0 < m; j0 += 2) {
            const int nt = (int)std::min<int64_t>(2, m - j0);
            launch_decode_block(L.B, c.nloc, c.slot(j0), 0, 0.f, w.X[0].p, c.st);
            if (nt == 2) launch_decode_block(L.B, c.nloc, c.slot(j0 + 1), 0, 0.f, w.X[1].p, c.st);
            else RBL_CUDA(cudaMemsetAsync(w.X[1].p, 0, (size_t)c.nloc * L.B * 8, c.st));
            c.reorth_gram(m, w.X[0].p, w.X[1].p);
            launch_ortho_accumulate(L.B, m, j0, nt, 0, w.Cmat.p, acc.p, c.st);
        }
    } else {
        // A 4-byte slab is measured in fp64: the stored values are decoded exactly (chunk-wise, into a temporary fp64
        // slab) and the fp64 Gram kernels accumulate - the fp32-grade reorth kernels themselves have a noise floor
        // of ~1e-8 per entry, the size of the quantity being measured.
        size_t free_b = 0, total_b = 0;
        RBL_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t blk_bytes = (size_t)c.nloc * L.B * 8;
        int64_t chunk = (int64_t)std::min<double>((double)m, std::max(1.0, (0.7 * (double)free_b) / (double)blk_bytes));
        DevBuf<double> tmp, Cd, part;
        tmp.alloc((size_t)chunk * c.nloc * L.B);
        Cd.alloc((size_t)chunk * L.B * 2 * L.B);
        part.alloc(std::max<size_t>(1, reorth_max_partial_elems(L.B, 0, c.nloc, chunk)));
        const bool dmma = reorth_d_supported(L.B, 0);
        for (int64_t c0 = 0; c0 < m; c0 += chunk) {
            const int64_t cn = std::min<int64_t>(chunk, m - c0);
            for (int64_t j = 0; j < cn; ++j)
                launch_decode_block(L.B, c.nloc, c.slot(c0 + j), 1, L.split_scale, tmp.p + (size_t)j * c.nloc * L.B, c.st);
            ReorthPlan p = reorth_plan(L.B, 0, c.nloc, cn);
            for (int64_t j0 = 0; j0 < m; j0 += 2) {
                const int nt = (int)std::min<int64_t>(2, m - j0);
                launch_decode_block(L.B, c.nloc, c.slot(j0), 1, L.split_scale, w.X[0].p, c.st);
                if (nt == 2) launch_decode_block(L.B, c.nloc, c.slot(j0 + 1), 1, L.split_scale, w.X[1].p, c.st);
                else RBL_CUDA(cudaMemsetAsync(w.X[1].p, 0, (size_t)c.nloc * L.B * 8, c.st));
                if (dmma) launch_reorth_gram_d(p, tmp.p, L.bstride, w.X[0].p, w.X[1].p, part.p, Cd.p, c.st);
                else launch_reorth_gram(p, tmp.p, L.bstride, w.X[0].p, w.X[1].p, part.p, Cd.p, c.st);
                if (c.multi) { std::string err; c.nccl(h->comm.allreduce_f64(Cd.p, (size_t)cn * L.B * 2 * L.B, c.st, err), err); }
                launch_ortho_accumulate(L.B, cn, j0 - c0, nt, 0, Cd.p, acc.p, c.st);
            }
        }
        RBL_CUDA(cudaStreamSynchronize(c.st));
    }
    double hacc[2];
    RBL_CUDA(cudaMemcpyAsync(hacc, acc.p, 16, cudaMemcpyDeviceToHost, c.st));
    RBL_CUDA(cudaStreamSynchronize(c.st));
    if (max_abs) *max_abs = hacc[0];
    if (fro) *fro = std::sqrt(hacc[1]);
}

}  // namespace rbl

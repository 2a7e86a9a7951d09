// Device-resident RBL driver.  Reference restated by structure, not by call sequence:
//   handle_create   <- Ag = adapt(CuArray, A)                         RBL_gpu.jl:209
//   solve           <- RBL_gpu body + lanczos_iteration               RBL_gpu.jl:134-203, 211-220
//   ritz            <- recover_eigvec                                 RBL_gpu.jl:106-132
// Differences by design: no per-statement synchronisation (the only host waits
// are the convergence checks), Krylov buffer is one slab written by kernel epilogues (no F<->D copy
// kernels, no per-block host mirror), full re-orthogonalisation is two streaming passes (block CGS)
// over the slab instead of 4 small GEMMs + a sync per stored block, block QR is shifted CholQR with
// re-orthogonalisation instead of Householder geqrf/orgqr, and the host eigen-check computes only the
// k wanted Ritz pairs and may overlap the device iteration.
#include "solver.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <thread>

#include "partition.h"

using namespace rbl;

namespace rbl {
namespace {
std::mutex g_ws_mu;
std::vector<std::pair<int, Workspace*>> g_parked_ws;  // (device, workspace)
size_t ws_bytes(const Workspace& w) {
    return w.buf.count + w.ritzV.count + w.ritzS.count + w.Cmat.count + w.rpart.count + w.tc_scratch.count * 4 +
           (w.X[0].count + w.X[1].count + w.X[2].count + w.omega.count + w.part.count + w.sendbuf.count) * 8;
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
        // all ranks must use the same partition: gather every rank's row0 via NCCL
        DevBuf<int64_t> d_send, d_recv;
        d_send.alloc(2);
        d_recv.alloc((size_t)2 * world);
        int64_t mine[2] = {row0, nloc};
        RBL_CUDA(cudaMemcpyAsync(d_send.p, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
        if (!h->comm.allgather_i64(d_send.p, d_recv.p, 2, h->stream, err)) throw Error(RBL_NCCL_ERROR, err);
        std::vector<int64_t> all((size_t)2 * world);
        RBL_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, all.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        RBL_CUDA(cudaStreamSynchronize(h->stream));
        for (int p = 0; p < world; ++p) {
            starts[p] = all[2 * p];
            if (p > 0 && starts[p] != all[2 * (p - 1)] + all[2 * (p - 1) + 1])
                throw Error(RBL_INVALID, "rbl_create_sharded: row ranges are not contiguous in rank order");
        }
        starts[world] = all[2 * (world - 1)] + all[2 * (world - 1) + 1];
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
    h->wsp->d_rowptr.ensure((size_t)nloc + 1);
    h->wsp->d_colidx.ensure(std::max<int64_t>(1, nnz));
    h->wsp->d_vals.ensure(std::max<int64_t>(1, nnz));
    RBL_CUDA(cudaMemcpy(h->wsp->d_rowptr.p, rp.data(), ((size_t)nloc + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (nnz) {
        RBL_CUDA(cudaMemcpy(h->wsp->d_colidx.p, ci.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
        RBL_CUDA(cudaMemcpy(h->wsp->d_vals.p, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->t_h2d_create = now_s() - t0;
    return h.release();
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

struct Ctx {
    rbl_handle* h = nullptr;
    cudaStream_t st = nullptr;
    int b = 0, B = 0;
    int64_t k = 0;
    bool fp32 = false;
    size_t ssz = 8;
    int64_t nloc = 0, next = 0, m_cap = 0, bstride = 0;
    Workspace& w;
    DevBuf<double> (&X)[3];
    DevBuf<unsigned char>& buf;
    DevBuf<double>&part, &small;  // rowop partials; small: G, Ai, Bp, Gloc (4 * B*B)
    DevBuf<QrState>& qr;
    DevBuf<unsigned char>&Cmat, &rpart;
    DevBuf<float>& tc_scratch;
    DevBuf<double>& sendbuf;
    PinnedBuf<double>&hA, &hB;
    PinnedBuf<QrState>& hqr;
    explicit Ctx(Workspace& ws)
        : w(ws), X(ws.X), buf(ws.buf), part(ws.part), small(ws.small), qr(ws.qr), Cmat(ws.Cmat), rpart(ws.rpart),
          tc_scratch(ws.tc_scratch), sendbuf(ws.sendbuf), hA(ws.hA), hB(ws.hB), hqr(ws.hqr) {}
    bool use_tc = false;
    bool use_h = false;   // FP16-split tensor-core kernels (default when supported); else TF32x3
    float split_scale = 0.f;  // != 0: the Krylov slab holds split16 rows (split16.h) written with this scale
    bool use_d = false;   // all-fp64 mode: FP64 tensor-core kernels
    int rgrid = 1;
    int64_t launches = 0;
    PhaseTimer tm;
    double bytes_rgram = 0, bytes_rupd = 0, bytes_spmm = 0;
    int64_t n_rgram = 0, n_rupd = 0, n_spmm = 0;

    double* G() { return small.p; }
    double* Ai() { return small.p + (size_t)B * B; }
    double* Bp() { return small.p + 2 * (size_t)B * B; }
    double* Gloc() { return small.p + 3 * (size_t)B * B; }
    void* slot(int64_t j) { return buf.p + (size_t)j * bstride * ssz; }

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
        launch_reduce_partials(part.p, rgrid, B * B, out, st);
        ++launches;
        allreduce(out, (size_t)B * B);
    }
    void rowop(const RowOpArgs& a) {
        launch_rowop(B, a, rgrid, st);
        ++launches;
    }
    void halo(double* Xblk) {
        if (!h->comm.active()) return;
        std::string err;
        const int64_t nsend = h->send_ptr[h->world];
        launch_gather_rows(B, nsend, h->wsp->d_send_rows.p, Xblk, sendbuf.p, st);
        ++launches;
        nccl(h->comm.group_start(err), err);
        for (int p = 0; p < h->world; ++p) {
            if (p == h->rank) continue;
            const size_t sb = (size_t)(h->send_ptr[p + 1] - h->send_ptr[p]) * B * sizeof(double);
            const size_t rb = (size_t)(h->halo_owner_ptr[p + 1] - h->halo_owner_ptr[p]) * B * sizeof(double);
            nccl(h->comm.send_bytes(sendbuf.p + (size_t)h->send_ptr[p] * B, sb, p, st, err), err);
            nccl(h->comm.recv_bytes(Xblk + (size_t)(nloc + h->halo_owner_ptr[p]) * B, rb, p, st, err), err);
        }
        nccl(h->comm.group_end(err), err);
    }
    void spmm(double* Q, double* U) {
        halo(Q);
        launch_spmm(B, nloc, h->wsp->d_rowptr.p, h->wsp->d_colidx.p, h->wsp->d_vals.p, Q, U, h->opt.op, h->opt.sigma, st);
        ++launches;
        ++n_spmm;
        bytes_spmm += 12.0 * (double)h->nnz + 4.0 * (double)(nloc + 1) + 16.0 * (double)nloc * B;
    }
    // thin QR of the block in `U` (in place).  The Gram U'U must already be in the rowop partials.
    void block_qr(double* U, int reset_ref) {
        const double defl_rel = 1e-12;
        finish_gram(G());
        launch_chol(B, G(), qr.p, 1, h->n, reset_ref, defl_rel, st);
        ++launches;
        RowOpArgs a;
        a.n = nloc; a.y = U; a.rinv = qr.p->Rinv; a.write_y = 1; a.do_gram = 1; a.partials = part.p;
        rowop(a);
        finish_gram(G());
        launch_chol(B, G(), qr.p, 2, h->n, 0, defl_rel, st);
        ++launches;
        rowop(a);  // apply pass 2, Gram for the optional pass 3
        finish_gram(G());
        launch_chol(B, G(), qr.p, 3, h->n, 0, defl_rel, st);
        ++launches;
        RowOpArgs a3;
        a3.n = nloc; a3.y = U; a3.rinv = qr.p->Rinv; a3.write_y = 1; a3.skip_flag = &qr.p->need_more;
        rowop(a3);
    }
};

void fill_stats(rbl_stats* s, Ctx& c, double* sec) {
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

int solve(rbl_handle* h, int64_t k, int64_t b_in, const double* omega, bool omega_on_device, double* d_out, void* v_out,
          bool v_on_device, rbl_stats* stats_out) {
    const double t_begin = now_s();
    rbl_stats stats;
    std::memset(&stats, 0, sizeof(stats));
    const rbl_options& opt = h->opt;
    if (k <= 0 || b_in <= 0 || b_in > 32) throw Error(RBL_INVALID, "rbl_solve: need k >= 1 and 1 <= b <= 32");
    if (k > h->n) throw Error(RBL_INVALID, "rbl_solve: k > n");
    if (!d_out || !v_out) throw Error(RBL_INVALID, "rbl_solve: null output");
    RBL_CUDA(cudaSetDevice(h->device));
    Ctx c(h->ws_ref());
    c.h = h; c.st = h->stream; c.b = (int)b_in; c.B = padded_block(c.b); c.k = k;
    c.fp32 = opt.precision == RBL_PRECISION_MIXED;
    c.ssz = c.fp32 ? 4 : 8;
    c.nloc = h->nloc; c.next = h->nloc + h->n_halo;
    c.bstride = c.nloc * c.B;
    c.tm.h = h; c.tm.st = c.st;
    const int b = c.b, B = c.B;
    const int64_t kryl_sz = std::max<int64_t>(opt.max_kryl_sz, b);
    int64_t m_cap = (kryl_sz + b - 1) / b;
    const int reorth_period = std::max(1, opt.reorth_period), check_period = std::max(1, opt.check_period);
    const bool async_ok = opt.async_check && (check_period % reorth_period == 0);

    // ---- memory plan (gpu_buffer_size, RBL_gpu.jl:95-104: how many Krylov blocks fit) ---------------
    size_t free_b = 0, total_b = 0;
    RBL_CUDA(cudaMemGetInfo(&free_b, &total_b));
    {   // memory already held by this handle's workspace is available to this solve
        const Workspace& w = h->ws_ref();
        free_b += w.buf.count + w.ritzS.count + w.ritzV.count + w.omega.count * 8 + (w.X[0].count + w.X[1].count + w.X[2].count + w.part.count + w.small.count + w.sendbuf.count) * 8 +
                  w.Cmat.count + w.rpart.count + w.tc_scratch.count * 4;
    }
    c.rgrid = rowop_grid(B, c.nloc);
    const size_t fixed = 3 * (size_t)c.next * B * 8 + (size_t)c.nloc * 4 * B * 4 + (size_t)c.rgrid * B * B * 8 + (size_t)c.nloc * (size_t)(b + k) * 8 +
                         ((size_t)64 << 20);
    const size_t per_block = (size_t)c.bstride * c.ssz + (size_t)B * 2 * B * (c.ssz + 8);
    if ((double)fixed + 2.0 * per_block > 0.92 * (double)free_b) throw Error(RBL_OOM, "rbl_solve: problem does not fit device memory");
    {
        size_t rp_elems = reorth_max_partial_elems(B, c.fp32, c.nloc, m_cap);
        double avail = 0.92 * (double)free_b - (double)fixed - (double)rp_elems * c.ssz;
        int64_t fit = (int64_t)std::floor(avail / (double)per_block);
        if (fit < m_cap) {
            m_cap = std::max<int64_t>(2, fit);
            if (opt.verbose) std::fprintf(stderr, "[rbl] Krylov buffer capped at %lld blocks by device memory\n", (long long)m_cap);
        }
    }
    c.m_cap = m_cap;
    for (int i = 0; i < 3; ++i) c.X[i].ensure((size_t)c.next * B);
    c.buf.ensure((size_t)m_cap * c.bstride * c.ssz);
    c.part.ensure((size_t)c.rgrid * B * B);
    c.small.ensure(4 * (size_t)B * B);
    c.qr.ensure(1);
    c.Cmat.ensure((size_t)m_cap * B * 2 * B * c.ssz);
    c.rpart.ensure(std::max<size_t>(1, reorth_max_partial_elems(B, c.fp32, c.nloc, m_cap)) * c.ssz);
    c.use_tc = reorth_tc_supported(B, c.fp32) && opt.reorth_impl != 1;
    if (opt.reorth_impl == 2 && !c.use_tc)
        throw Error(RBL_INVALID, "rbl_solve: tensor-core reorth needs precision=mixed and padded block size 16");
    if (opt.reorth_impl >= 3 && !reorth_h_supported(B, c.fp32))
        throw Error(RBL_INVALID, "rbl_solve: FP16-split tensor-core reorth needs precision=mixed and padded block size 16 or 32");
    c.use_h = reorth_h_supported(B, c.fp32) && opt.reorth_impl != 1 && !(opt.reorth_impl == 2 && c.use_tc);
    c.split_scale = (c.use_h && opt.reorth_impl != 3) ? reorth_h_scale(h->n) : 0.f;
    c.use_d = reorth_d_supported(B, c.fp32) && opt.reorth_impl != 1;
    if (c.use_tc || c.use_h) c.tc_scratch.ensure(std::max(reorth_tc_scratch_floats(B, c.nloc, m_cap), reorth_h_scratch_words(B, c.nloc, m_cap)));
    if (h->comm.active()) c.sendbuf.ensure(std::max<int64_t>(1, h->send_ptr[h->world]) * (size_t)B);
    c.hA.ensure((size_t)m_cap * B * B);
    c.hB.ensure((size_t)m_cap * B * B);
    c.hqr.ensure(1);
    const double t_alloc_done = now_s();
    RBL_CUDA(cudaMemsetAsync(c.qr.p, 0, sizeof(QrState), c.st));
    RBL_CUDA(cudaMemsetAsync(c.small.p, 0, 4 * (size_t)B * B * 8, c.st));
    for (int i = 0; i < 3; ++i) RBL_CUDA(cudaMemsetAsync(c.X[i].p, 0, (size_t)c.next * B * 8, c.st));

    // ---- start block: Q1 = thin-Q of qr(A * Omega)                              RBL_gpu.jl:213-214 ----
    DevBuf<double>& d_omega = h->ws_ref().omega;
    const double* om_dev = nullptr;
    {
        const double t0 = now_s();
        if (omega && omega_on_device) {
            om_dev = omega;
        } else {
            d_omega.ensure((size_t)c.nloc * b);
            if (omega) {
                RBL_CUDA(cudaMemcpyAsync(d_omega.p, omega, (size_t)c.nloc * b * 8, cudaMemcpyHostToDevice, c.st));
            } else {
                for (int col = 0; col < b; ++col)
                    launch_randn(c.nloc, 0x5eedull, (uint64_t)col * (uint64_t)h->n + (uint64_t)h->row0,
                                 d_omega.p + (size_t)col * c.nloc, c.st);
            }
            om_dev = d_omega.p;
            RBL_CUDA(cudaStreamSynchronize(c.st));
        }
        stats.t_h2d = (now_s() - t0) + h->t_h2d_create;
    }
    double *cur = c.X[1].p, *prev = c.X[0].p, *U = c.X[2].p;
    launch_colmajor_to_block(B, c.nloc, b, om_dev, c.nloc, c.X[0].p, c.st);
    ++c.launches;
    c.tm.mark(PH_SPMM);
    c.spmm(c.X[0].p, cur);
    c.tm.mark(PH_QR);
    {
        RowOpArgs a;
        a.n = c.nloc; a.y = cur; a.do_gram = 1; a.partials = c.part.p;
        c.rowop(a);
        c.block_qr(cur, 1);
        // a start block of rank 0 (Omega = 0, or A*Omega = 0) spans no Krylov space: report it instead of
        // iterating on zero columns (the reference's Householder QR would continue with arbitrary unit vectors)
        RBL_CUDA(cudaMemcpyAsync(c.hqr.p, c.qr.p, sizeof(QrState), cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaStreamSynchronize(c.st));
        if (c.hqr.p->bad) throw Error(RBL_INVALID, "rbl_solve: A*Omega contains non-finite values");
        if (c.hqr.p->ndeflated >= B) throw Error(RBL_BREAKDOWN, "rbl_solve: the start block A*Omega has rank 0");
    }

    // ---- host-side T bookkeeping (insertA!/insertB!, common.jl:9-26) ---------------------------------
    BandSym T;
    T.reset(0, b);
    BandTopK checker;
    checker.threads = opt.host_threads > 0 ? opt.host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    checker.verbose = opt.verbose;
    int64_t t_blocks = 0;  // blocks already inserted into T
    std::vector<cudaEvent_t> step_event((size_t)m_cap + 2, nullptr);
    auto grow_T = [&](int64_t upto_blocks) {
        // A_j for j < upto, B_j for j < upto-1 (B_i of the newest block is applied after the check, common.jl:113)
        const int W = 2 * b + 1;
        T.F.resize((size_t)upto_blocks * b * W, 0.0);
        T.N = upto_blocks * b;
        for (int64_t j = t_blocks; j < upto_blocks; ++j) {
            const double* A = c.hA.p + (size_t)j * B * B;
            for (int r = 0; r < b; ++r)
                for (int cc = 0; cc <= r; ++cc) T.set_sym(j * b + r, j * b + cc, A[r * B + cc]);
            if (j > 0) {
                const double* Bm = c.hB.p + (size_t)(j - 1) * B * B;  // couples block j-1 and j
                for (int cc = 0; cc < b; ++cc)
                    for (int m = 0; m <= cc; ++m) T.set_sym(j * b + m, (j - 1) * b + cc, Bm[m * B + cc]);
            }
        }
        t_blocks = upto_blocks;
        T.update_norm();
    };
    double t_eig = 0.0;
    // Shadow tracker (rank 0): from the first checks on a background thread keeps computing
    // ALL k Ritz pairs of the latest T snapshot - by slicing the first time, by refining its own previous pairs
    // afterwards (BandTopK::refine_seeds) - so that the accepting check only has to refine fresh seeds instead of
    // solving the eigenproblem from scratch while the device sits idle.
    struct Shadow {
        std::mutex mu;
        std::condition_variable cv;
        std::thread th;
        bool active = false, stop = false, have_req = false, have_res = false;
        std::atomic<bool> cancel{false};   // raised with `stop`: the tracker abandons the eigensolve it is in
        BandSym req;
        TopKResult res;
    } shadow;
    const int shadow_verbose = opt.verbose;
    auto shadow_loop = [&shadow, k, b, shadow_verbose](int nthreads) {
        BandTopK tracker;
        tracker.threads = nthreads;
        for (;;) {
            BandSym Tc;
            {
                std::unique_lock<std::mutex> lk(shadow.mu);
                shadow.cv.wait(lk, [&] { return shadow.stop || shadow.have_req; });
                if (shadow.stop) return;
                Tc = std::move(shadow.req);
                shadow.have_req = false;
            }
            Tc.cancel = &shadow.cancel;
            TopKResult r;
            const double ts0 = now_s();
            const int64_t f0 = tracker.total_factorizations;
            try {
                r = tracker.check(Tc, nullptr, b, k, 0.0, true);
            } catch (const Cancelled&) {
                return;
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
            if (s.th.joinable()) s.th.join();
        }
    } shadow_join{shadow};
    auto run_check = [&](int64_t it, bool force_full) -> TopKResult {
        cudaSetDevice(h->device);
        cudaEventSynchronize(step_event[it]);
        const double t0 = now_s();
        grow_T(it);
        if (shadow.active) {
            std::lock_guard<std::mutex> lk(shadow.mu);
            if (shadow.have_res) {
                checker.set_seeds(shadow.res.d, shadow.res.s, shadow.res.N, k);
                shadow.have_res = false;
            }
        }
        std::vector<double> Bi((size_t)b * b);
        const double* Bm = c.hB.p + (size_t)(it - 1) * B * B;
        for (int r = 0; r < b; ++r)
            for (int cc = 0; cc < b; ++cc) Bi[(size_t)r * b + cc] = Bm[r * B + cc];
        TopKResult r = checker.check(T, Bi.data(), b, k, opt.tol, force_full);
        if (!r.converged && !shadow.active && T.N >= 2 * k) {
            shadow.active = true;
            shadow.th = std::thread(shadow_loop, std::max(1, checker.threads - 1));
        }
        if (!r.converged && shadow.active) {
            {
                std::lock_guard<std::mutex> lk(shadow.mu);
                shadow.req = T;  // snapshot (a few MB)
                shadow.have_req = true;
            }
            shadow.cv.notify_all();
        }
        t_eig += now_s() - t0;
        if (opt.verbose > 1)
            std::fprintf(stderr, "[rbl] check it=%lld N=%lld nfac=%d took %.2f ms conv=%d\n", (long long)it, (long long)T.N,
                         r.factorizations, (now_s() - t0) * 1e3, (int)r.converged);
        return r;
    };

    // ---- first step (i = 1)                                                     RBL_gpu.jl:149-161 ----
    auto record_step = [&](int64_t it) {
        RBL_CUDA(cudaMemcpyAsync(c.hA.p + (size_t)(it - 1) * B * B, c.Ai(), (size_t)B * B * 8, cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaMemcpyAsync(c.Bp(), c.qr.p->R, (size_t)B * B * 8, cudaMemcpyDeviceToDevice, c.st));
        RBL_CUDA(cudaMemcpyAsync(c.hB.p + (size_t)(it - 1) * B * B, c.qr.p->R, (size_t)B * B * 8, cudaMemcpyDeviceToHost, c.st));
    };
    c.tm.mark(PH_LOC);
    launch_store_block(B, c.nloc, cur, c.slot(0), c.fp32, c.split_scale, c.st);
    ++c.launches;
    c.tm.mark(PH_SPMM);
    c.spmm(cur, U);
    c.tm.mark(PH_3TERM);
    {
        RowOpArgs a;
        a.n = c.nloc; a.y = U; a.gram_z = cur; a.do_gram = 1; a.partials = c.part.p;
        c.rowop(a);
        c.finish_gram(c.Ai());
        RowOpArgs a2;
        a2.n = c.nloc; a2.y = U; a2.x1 = cur; a2.m1 = c.Ai(); a2.write_y = 1; a2.do_gram = 1; a2.partials = c.part.p;
        c.rowop(a2);
    }
    c.tm.mark(PH_QR);
    c.block_qr(U, 1);
    record_step(1);
    c.tm.mark(PH_NONE);
    { double* t = prev; prev = cur; cur = U; U = t; }

    // ---- main loop                                                               RBL_gpu.jl:162-194 ----
    int64_t i = 1;
    int64_t final_i = 0;
    bool converged = false;
    TopKResult final_res;
    std::future<TopKResult> pending;
    int64_t pending_i = 0;
    int64_t last_check_i = 0;
    int checks = 0;
    double t_wait = 0.0;
    // Row-sharded runs: only rank 0 evaluates the host check; the decision (and, on acceptance, D and S)
    // is summed over ranks with every other rank contributing zeros, so all ranks follow the same control
    // flow and use the same Ritz basis.
    const bool is_root = (h->rank == 0);
    const bool multi = h->comm.active();
    // kept in the workspace: with peer access enabled by NCCL every cudaMalloc / cudaFree / cudaFreeHost is expensive
    DevBuf<double>& d_ctrl = h->ws_ref().ctrl;
    PinnedBuf<double>& h_ctrl = h->ws_ref().h_ctrl;
    if (multi) {
        d_ctrl.ensure(2);
        h_ctrl.ensure(2);
    }
    auto agree_flag = [&](bool local) -> bool {
        if (!multi) return local;
        h_ctrl.p[0] = (is_root && local) ? 1.0 : 0.0;
        RBL_CUDA(cudaMemcpyAsync(d_ctrl.p, h_ctrl.p, 8, cudaMemcpyHostToDevice, c.st));
        c.allreduce(d_ctrl.p, 1);
        RBL_CUDA(cudaMemcpyAsync(h_ctrl.p + 1, d_ctrl.p, 8, cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaStreamSynchronize(c.st));
        return h_ctrl.p[1] > 0.5;
    };
    auto share_result = [&](TopKResult& r, int64_t Nrows) {
        if (!multi) return;
        const size_t cnt = (size_t)k + (size_t)Nrows * k;
        std::vector<double> hbuf(cnt, 0.0);
        if (is_root) {
            std::copy(r.d.begin(), r.d.begin() + k, hbuf.begin());
            std::copy(r.s.begin(), r.s.begin() + (size_t)Nrows * k, hbuf.begin() + k);
        }
        DevBuf<double>& dbuf = h->ws_ref().share;
        dbuf.ensure(cnt);
        RBL_CUDA(cudaMemcpyAsync(dbuf.p, hbuf.data(), cnt * 8, cudaMemcpyHostToDevice, c.st));
        c.allreduce(dbuf.p, cnt);
        RBL_CUDA(cudaMemcpyAsync(hbuf.data(), dbuf.p, cnt * 8, cudaMemcpyDeviceToHost, c.st));
        RBL_CUDA(cudaStreamSynchronize(c.st));
        r.N = Nrows;
        r.d.assign(hbuf.begin(), hbuf.begin() + k);
        r.s.assign(hbuf.begin() + k, hbuf.end());
    };
    bool check_in_flight = false;
    auto harvest = [&]() -> bool {  // wait for the in-flight check; true when it accepted
        if (!check_in_flight) return false;
        check_in_flight = false;
        TopKResult r;
        bool local = false;
        if (is_root) {
            const double t0 = now_s();
            r = pending.get();
            t_wait += now_s() - t0;
            if (opt.verbose > 1) std::fprintf(stderr, "[rbl] harvest it=%lld waited %.2f ms\n", (long long)pending_i, (now_s() - t0) * 1e3);
            local = r.converged;
        }
        ++checks;
        last_check_i = pending_i;
        if (agree_flag(local)) {
            converged = true;
            final_i = pending_i;
            final_res = std::move(r);
            share_result(final_res, final_i * b);
            return true;
        }
        return false;
    };

    while (i * b < kryl_sz && i < m_cap) {
        ++i;
        if (i % reorth_period == 0 && i > 2) {
            // hybrid_part_reorth!: project Q_i and Q_{i-1} against blocks 1..i-2, all at once (block CGS)
            const int64_t m = i - 2;
            ReorthPlan p = reorth_plan(B, c.fp32, c.nloc, m);
            c.tm.mark(PH_RGRAM);
            if (c.use_d) {
                launch_reorth_gram_d(p, c.buf.p, c.bstride, cur, prev, c.rpart.p, c.Cmat.p, c.st);
                c.launches += 2;
                if (h->comm.active()) {
                    std::string err;
                    c.nccl(h->comm.allreduce_f64((double*)c.Cmat.p, (size_t)m * B * 2 * B, c.st, err), err);
                }
                c.tm.mark(PH_RUPD);
                launch_reorth_update_d(p, c.buf.p, c.bstride, c.Cmat.p, cur, prev, c.slot(i - 2), c.st);
                ++c.launches;
            } else if (c.use_h) {
                launch_reorth_gram_h(p, h->n, c.buf.p, c.bstride, cur, prev, c.rpart.p, c.Cmat.p, c.tc_scratch.p, m_cap,
                                     c.split_scale != 0.f, c.st);
                c.launches += 4;
                if (h->comm.active()) {
                    std::string err;
                    const size_t cnt = (size_t)m * B * 2 * B;
                    c.nccl(h->comm.allreduce_f32((float*)c.Cmat.p, cnt, c.st, err), err);
                }
                launch_reorth_coeff_h(p, c.Cmat.p, c.tc_scratch.p, m_cap, h->comm.active() ? 1 : 0, c.st);
                ++c.launches;
                c.tm.mark(PH_RUPD);
                launch_reorth_update_h(p, h->n, c.buf.p, c.bstride, cur, prev, c.slot(i - 2), c.tc_scratch.p, m_cap,
                                       c.split_scale != 0.f, c.st);
                ++c.launches;
            } else if (c.use_tc) {
                launch_reorth_gram_tc(p, c.buf.p, c.bstride, cur, prev, c.rpart.p, c.Cmat.p, c.tc_scratch.p, m_cap, c.st);
                c.launches += 3;
                if (h->comm.active()) {
                    // the hi/lo parts are not additive across ranks: reduce C, then re-split on every rank
                    std::string err;
                    const size_t cnt = (size_t)m * B * 2 * B;
                    c.nccl(h->comm.allreduce_f32((float*)c.Cmat.p, cnt, c.st, err), err);
                    ReorthPlan one = p;
                    one.ranges = 1;
                    launch_reorth_gram_tc_resplit(one, c.Cmat.p, c.tc_scratch.p, m_cap, c.st);
                    ++c.launches;
                }
                c.tm.mark(PH_RUPD);
                launch_reorth_update_tc(p, c.buf.p, c.bstride, cur, prev, c.slot(i - 2), c.tc_scratch.p, m_cap, c.st);
                ++c.launches;
            } else {
                launch_reorth_gram(p, c.buf.p, c.bstride, cur, prev, c.rpart.p, c.Cmat.p, c.st);
                c.launches += 2;
                if (h->comm.active()) {
                    std::string err;
                    const size_t cnt = (size_t)m * B * 2 * B;
                    c.nccl(c.fp32 ? h->comm.allreduce_f32((float*)c.Cmat.p, cnt, c.st, err)
                                  : h->comm.allreduce_f64((double*)c.Cmat.p, cnt, c.st, err), err);
                }
                c.tm.mark(PH_RUPD);
                launch_reorth_update(p, c.buf.p, c.bstride, c.Cmat.p, cur, prev, c.slot(i - 2), c.st);
                ++c.launches;
            }
            ++c.n_rgram; ++c.n_rupd;
            c.bytes_rgram += (double)c.ssz * (double)c.nloc * (double)m * B + 8.0 * (double)c.nloc * 2 * B;
            c.bytes_rupd += (double)c.ssz * (double)c.nloc * (double)m * B + 2 * 8.0 * (double)c.nloc * 2 * B +
                            (double)c.ssz * (double)c.nloc * B;
        }
        // loc_reorth_gpu! (effective): Q_i -= Q_{i-1} (Q_{i-1}' Q_i); then the block joins the buffer (:167-172)
        c.tm.mark(PH_LOC);
        {
            RowOpArgs a;
            a.n = c.nloc; a.y = cur; a.gram_z = prev; a.do_gram = 1; a.partials = c.part.p;
            c.rowop(a);
            c.finish_gram(c.Gloc());
            RowOpArgs a2;
            a2.n = c.nloc; a2.y = cur; a2.x1 = prev; a2.m1 = c.Gloc(); a2.write_y = 1;
            a2.store = c.slot(i - 1); a2.store_fp32 = c.fp32; a2.store_split_scale = c.split_scale;
            c.rowop(a2);
        }
        c.tm.mark(PH_SPMM);
        c.spmm(cur, U);                                                        // :176
        c.tm.mark(PH_3TERM);
        {
            RowOpArgs a;                                                       // :177-178
            a.n = c.nloc; a.y = U; a.x1 = prev; a.m1 = c.Bp(); a.m1_transposed = 1; a.write_y = 1;
            a.gram_z = cur; a.do_gram = 1; a.partials = c.part.p;
            c.rowop(a);
            c.finish_gram(c.Ai());
            RowOpArgs a2;                                                      // :179 (+ Gram for the QR)
            a2.n = c.nloc; a2.y = U; a2.x1 = cur; a2.m1 = c.Ai(); a2.write_y = 1; a2.do_gram = 1; a2.partials = c.part.p;
            c.rowop(a2);
        }
        c.tm.mark(PH_QR);
        c.block_qr(U, 0);                                                      // :180-184
        record_step(i);
        c.tm.mark(PH_NONE);
        { double* t = prev; prev = cur; cur = U; U = t; }

        if (i * b > k && i % check_period == 0) {                               // :186
            if (harvest()) break;
            RBL_CUDA(cudaEventCreateWithFlags(&step_event[i], cudaEventDisableTiming));
            RBL_CUDA(cudaEventRecord(step_event[i], c.st));
            pending_i = i;
            check_in_flight = true;
            const int64_t it = i;
            if (is_root) {
                if (async_ok) {
                    pending = std::async(std::launch::async, [&, it]() { return run_check(it, false); });
                } else {
                    std::promise<TopKResult> pr;
                    pr.set_value(run_check(it, false));
                    pending = pr.get_future();
                }
            }
            if (!async_ok && harvest()) break;
        }
    }
    if (!converged) harvest();
    const double t_loop_done = now_s();
    const int64_t iterations_run = i;
    RBL_CUDA(cudaStreamSynchronize(c.st));
    int status = RBL_OK;
    if (!converged) {
        // cap reached: the reference returns a stale check or throws; here: best effort + status
        status = RBL_NOT_CONVERGED;
        int64_t it = last_check_i > 0 ? last_check_i : i;
        if (it * b < k) throw Error(RBL_INVALID, "rbl_solve: Krylov cap smaller than k, no Ritz pairs available");
        if (!step_event[it]) {
            RBL_CUDA(cudaEventCreateWithFlags(&step_event[it], cudaEventDisableTiming));
            RBL_CUDA(cudaEventRecord(step_event[it], c.st));
        }
        if (it < t_blocks) {  // T already grew past `it` (cannot happen with one check in flight) - rebuild
            T.reset(0, b);
            t_blocks = 0;
        }
        if (is_root) final_res = run_check(it, true);
        share_result(final_res, it * b);
        ++checks;
        final_i = it;
    }
    for (auto e : step_event)
        if (e) cudaEventDestroy(e);

    if (const char* dump = std::getenv("RBL_DUMP_T")) {
        // debugging aid: the A_i / B_i blocks of T (what the host checks saw), for tools/replay_dump.py
        if (is_root && dump[0]) {
            if (FILE* f = std::fopen(dump, "wb")) {
                const int64_t hdr[4] = {iterations_run, (int64_t)B, (int64_t)b, final_i};
                std::fwrite(hdr, sizeof(int64_t), 4, f);
                std::fwrite(c.hA.p, sizeof(double), (size_t)iterations_run * B * B, f);
                std::fwrite(c.hB.p, sizeof(double), (size_t)iterations_run * B * B, f);
                std::fclose(f);
            }
        }
    }
    const double t_final_done = now_s();
    // ---- Ritz vectors V = Qbuf * S                                              RBL_gpu.jl:106-132,219 ----
    const int64_t mfin = final_i;
    const int kpad = (int)((k + 15) / 16 * 16);
    {
        std::vector<unsigned char> Sh((size_t)mfin * B * kpad * c.ssz, 0);
        for (int64_t j = 0; j < mfin; ++j)
            for (int cc = 0; cc < b; ++cc)
                for (int64_t t = 0; t < k; ++t) {
                    const double v = final_res.s[(size_t)t * final_res.N + (size_t)j * b + cc];
                    const size_t idx = ((size_t)j * B + cc) * kpad + t;
                    if (c.fp32) reinterpret_cast<float*>(Sh.data())[idx] = (float)v;   // cu() narrows S, RBL_gpu.jl:119
                    else reinterpret_cast<double*>(Sh.data())[idx] = v;
                }
        const double tr0 = now_s();
        DevBuf<unsigned char>&dS = h->ws_ref().ritzS, &dV = h->ws_ref().ritzV;
        dS.ensure(Sh.size());
        RBL_CUDA(cudaMemcpyAsync(dS.p, Sh.data(), Sh.size(), cudaMemcpyHostToDevice, c.st));
        const size_t vsz = opt.v_fp32 ? 4 : 8;
        void* Vdev = v_out;
        if (!v_on_device) {
            dV.ensure((size_t)c.nloc * k * vsz);
            Vdev = dV.p;
        }
        const double tr1 = now_s();
        c.tm.mark(PH_RITZ);
        DevBuf<unsigned>& ritz_words = h->ws_ref().ritz_words;
        if (c.split_scale != 0.f) {
            ritz_words.ensure(ritz_h_scratch_words(B, mfin, kpad));
            launch_ritz_h(B, c.nloc, mfin, (int)k, kpad, c.buf.p, c.bstride, dS.p, Vdev, c.nloc, opt.v_fp32, c.split_scale,
                          ritz_words.p, c.st);
            ++c.launches;
        } else {
            launch_ritz(B, c.fp32, c.nloc, mfin, (int)k, kpad, c.buf.p, c.bstride, dS.p, Vdev, c.nloc, opt.v_fp32, 0.f, c.st);
        }
        ++c.launches;
        c.tm.mark(PH_NONE);
        RBL_CUDA(cudaStreamSynchronize(c.st));
        stats.bytes_ritz = (double)c.ssz * (double)c.nloc * (double)mfin * B + (double)vsz * (double)c.nloc * (double)k;
        stats.flops_ritz = 2.0 * (double)c.nloc * (double)mfin * B * (double)k;
        if (!v_on_device) {
            const double t0 = now_s();
            RBL_CUDA(cudaMemcpy(v_out, dV.p, (size_t)c.nloc * k * vsz, cudaMemcpyDeviceToHost));
            stats.t_d2h = now_s() - t0;
        }
        if (opt.verbose) std::fprintf(stderr, "[rbl] ritz section: setup %.3f kernel+sync %.3f\n", tr1 - tr0, now_s() - tr1 - stats.t_d2h);
    }
    for (int64_t t = 0; t < k; ++t) d_out[t] = final_res.d[t];
    RBL_CUDA(cudaMemcpy(c.hqr.p, c.qr.p, sizeof(QrState), cudaMemcpyDeviceToHost));

    double sec[PH_COUNT];
    const double tc0 = now_s();
    c.tm.collect(sec);
    if (opt.verbose) std::fprintf(stderr, "[rbl] event collect %.3f s (%zu events)\n", now_s() - tc0, c.tm.used);
    fill_stats(&stats, c, sec);
    stats.iterations = final_i;
    stats.kryl_sz = final_i * b;
    stats.iterations_run = iterations_run;
    stats.converged = converged ? 1 : 0;
    stats.checks = checks;
    stats.full_checks = checker.full_checks;
    stats.host_factorizations = checker.total_factorizations;
    stats.deflated = c.hqr.p->ndeflated - (B - b);
    stats.t_eig = t_eig;
    stats.t_eig_wait = t_wait;
    stats.t_total = now_s() - t_begin;
    if (c.hqr.p->bad) status = RBL_BREAKDOWN;
    if (stats_out) *stats_out = stats;
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] host checks: witness %d (%.3f s), bracketed %d (%.3f s), full %d (%.3f s); factorisations %lld (+%lld resumed)\n",
                     checker.stage_hits[0], checker.stage_sec[0], checker.stage_hits[1], checker.stage_sec[1], checker.stage_hits[2],
                     checker.stage_sec[2], (long long)checker.total_factorizations, (long long)checker.resumed_factorizations);
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] timeline: alloc %.3f  start+loop %.3f  final-check %.3f  ritz+d2h %.3f (d2h %.3f, h2d %.3f)\n",
                     t_alloc_done - t_begin, t_loop_done - t_alloc_done, t_final_done - t_loop_done,
                     now_s() - t_final_done, stats.t_d2h, stats.t_h2d);
    if (opt.verbose)
        std::fprintf(stderr, "[rbl] Iterations: %lld and kryl_sz: %lld (ran %lld), checks %d (full %d), t=%.3fs eig=%.3fs wait=%.3fs\n",
                     (long long)final_i, (long long)(final_i * b), (long long)iterations_run, checks,
                     checker.full_checks, stats.t_total, t_eig, t_wait);
    return status;
}

}  // namespace rbl

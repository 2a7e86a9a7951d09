// extern "C" entry points declared in include/rbl_b200.h.
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mmio.h"
#include "partition.h"
#include "solver.h"

using namespace rbl;

static thread_local std::string g_err;

template <typename F>
static int guarded(F&& f) {
    try {
        return f();
    } catch (const Error& e) {
        g_err = e.what();
        return e.status;
    } catch (const std::bad_alloc&) {
        g_err = "host out of memory";
        return RBL_OOM;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RBL_INVALID;
    }
}

static void need_device() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        throw Error(RBL_NO_DEVICE, "no CUDA device visible: rbl_b200 has no CPU fallback");
}

// ---- single-process multi-GPU (opts.ngpus > 1): one per-device handle + one host thread per device -------------
// `RBL_gpu(A,k,b; ngpus=8)` from ONE caller (Julia is one process).  The group handle owns N
// row-sharded handles on devices [device, device+N); NCCL ranks are the host threads of this process.
namespace {
std::mutex g_group_mu;
std::map<std::pair<int, int>, std::vector<char>> g_group_uid;   // (first device, N) -> unique id of that group

template <typename F>
void run_parts(int N, std::vector<std::string>& errs, std::vector<int>& codes, F&& f) {
    errs.assign(N, std::string());
    codes.assign(N, RBL_OK);
    std::vector<std::thread> th;
    for (int p = 0; p < N; ++p)
        th.emplace_back([&, p] {
            try {
                codes[p] = f(p);
            } catch (const Error& e) {
                errs[p] = e.what();
                codes[p] = e.status;
            } catch (const std::exception& e) {
                errs[p] = e.what();
                codes[p] = RBL_INVALID;
            }
        });
    for (auto& t : th) t.join();
}

rbl_handle* create_group(int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                         int index_base, const rbl_options* opts) {
    const int N = opts->ngpus;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        throw Error(RBL_NO_DEVICE, "no CUDA device visible: rbl_b200 has no CPU fallback");
    int dev0 = opts->device;
    if (dev0 < 0) dev0 = 0;
    if (dev0 + N > ndev) throw Error(RBL_INVALID, "rbl_create: opts.ngpus exceeds the visible devices");
    if (n <= 0 || nnz < 0 || !colptr || (nnz > 0 && (!rowval || !nzval))) throw Error(RBL_INVALID, "rbl_create: bad arguments");
    if (n < N) throw Error(RBL_INVALID, "rbl_create: fewer rows than GPUs");
    if (colptr[n] - colptr[0] != nnz) throw Error(RBL_INVALID, "rbl_create: rowptr[n]-rowptr[0] != nnz");
    std::unique_ptr<rbl_handle> g(new rbl_handle());
    g->opt = *opts;
    g->n = n; g->nloc = n; g->nnz = nnz; g->world = N; g->device = dev0;
    g->part_rows.resize((size_t)N + 1);
    partition_rows(n, N, g->part_rows.data());
    std::vector<char> uid(128);
    {
        std::lock_guard<std::mutex> lk(g_group_mu);
        auto it = g_group_uid.find({dev0, N});
        if (it == g_group_uid.end()) {
            std::string err;
            if (!Comm::unique_id(uid.data(), err)) throw Error(RBL_NCCL_ERROR, err);
            g_group_uid[{dev0, N}] = uid;
        } else {
            uid = it->second;
        }
    }
    g->parts.assign(N, nullptr);
    std::vector<std::string> errs;
    std::vector<int> codes;
    run_parts(N, errs, codes, [&](int p) {
        rbl_options o = *opts;
        o.device = dev0 + p;
        o.ngpus = 1;
        if (p != 0) o.verbose = 0;
        const int64_t r0 = g->part_rows[p], r1 = g->part_rows[p + 1];
        const int64_t e0 = colptr[r0] - colptr[0], e1 = colptr[r1] - colptr[0];
        g->parts[p] = handle_create(n, r0, r1 - r0, e1 - e0, colptr + r0, rowval + e0, nzval + e0, index_base, p, N,
                                    uid.data(), &o);
        return (int)RBL_OK;
    });
    for (int p = 0; p < N; ++p)
        if (codes[p] != RBL_OK) {
            {   // the group id may be half-used: never reuse it
                std::lock_guard<std::mutex> lk(g_group_mu);
                g_group_uid.erase({dev0, N});
            }
            throw Error(codes[p], "rbl_create (device " + std::to_string(dev0 + p) + "): " + errs[p]);
        }
    g->gersh_lo = g->parts[0]->gersh_lo;
    g->gersh_hi = g->parts[0]->gersh_hi;
    return g.release();
}

int solve_group(rbl_handle* g, int64_t k, int64_t b, const double* omega, double* d_out, void* v_out, rbl_stats* stats) {
    const int N = (int)g->parts.size();
    if (!d_out || !v_out) throw Error(RBL_INVALID, "rbl_solve: null output");
    const size_t vsz = g->opt.v_fp32 ? 4 : 8;
    std::vector<std::vector<double>> dpart(N, std::vector<double>((size_t)std::max<int64_t>(k, 1)));
    std::vector<rbl_stats> st(N);
    std::vector<std::string> errs;
    std::vector<int> codes;
    run_parts(N, errs, codes, [&](int p) {
        SolveIO io;
        const int64_t r0 = g->part_rows[p];
        io.omega = omega ? omega + r0 : nullptr;
        io.ld_omega = g->n;
        io.v = (char*)v_out + (size_t)r0 * vsz;
        io.ldv = g->n;
        return solve(g->parts[p], k, b, io, dpart[p].data(), &st[p]);
    });
    for (int p = 0; p < N; ++p)
        if (codes[p] != RBL_OK && codes[p] != RBL_NOT_CONVERGED)
            throw Error(codes[p], "rbl_solve (device " + std::to_string(g->device + p) + "): " + errs[p]);
    for (int64_t t = 0; t < k; ++t) d_out[t] = dpart[0][t];
    if (stats) {
        *stats = st[0];
        for (int p = 1; p < N; ++p) {   // device phases: the slowest rank; launches: all ranks
            stats->t_spmm = std::max(stats->t_spmm, st[p].t_spmm);
            stats->t_3term = std::max(stats->t_3term, st[p].t_3term);
            stats->t_qr = std::max(stats->t_qr, st[p].t_qr);
            stats->t_part_reorth = std::max(stats->t_part_reorth, st[p].t_part_reorth);
            stats->t_loc_reorth = std::max(stats->t_loc_reorth, st[p].t_loc_reorth);
            stats->t_ritz = std::max(stats->t_ritz, st[p].t_ritz);
            stats->t_total = std::max(stats->t_total, st[p].t_total);
            stats->kernel_launches += st[p].kernel_launches;
            stats->bytes_part_reorth += st[p].bytes_part_reorth;
            stats->bytes_spmm += st[p].bytes_spmm;
            stats->max_residual = std::max(stats->max_residual, st[p].max_residual);
        }
    }
    return codes[0];
}
}  // namespace

extern "C" {

const char* rbl_last_error(void) { return g_err.c_str(); }
const char* rbl_version(void) { return "rbl_b200 0.1 (sm_100a)"; }
int rbl_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 0;
    return ndev;
}
int rbl_struct_sizes(int64_t* options_bytes, int64_t* stats_bytes) {
    if (options_bytes) *options_bytes = (int64_t)sizeof(rbl_options);
    if (stats_bytes) *stats_bytes = (int64_t)sizeof(rbl_stats);
    return RBL_OK;
}
int rbl_options_default(rbl_options* opts) {
    if (!opts) return RBL_INVALID;
    default_options(opts);
    return RBL_OK;
}

int rbl_create(int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval,
               int index_base, const rbl_options* opts, rbl_handle** out) {
    return guarded([&] {
        if (!out) throw Error(RBL_INVALID, "rbl_create: out is null");
        if (opts && opts->ngpus > 1) {
            if (index_base != 0 && index_base != 1) throw Error(RBL_INVALID, "rbl_create: index_base must be 0 or 1");
            *out = create_group(n, nnz, colptr, rowval, nzval, index_base, opts);
        } else {
            *out = handle_create(n, 0, n, nnz, colptr, rowval, nzval, index_base, 0, 1, nullptr, opts);
        }
        return (int)RBL_OK;
    });
}

int rbl_create_dense(int64_t n, const double* a, const rbl_options* opts, rbl_handle** out) {
    return guarded([&] {
        if (!out || !a || n <= 0) throw Error(RBL_INVALID, "rbl_create_dense: bad arguments");
        if (n * n >= ((int64_t)1 << 31)) throw Error(RBL_INVALID, "rbl_create_dense: n*n must be < 2^31");
        std::vector<int64_t> rp((size_t)n + 1), ci((size_t)n * n);
        std::vector<double> v((size_t)n * n);
        for (int64_t r = 0; r <= n; ++r) rp[r] = r * n;
        for (int64_t r = 0; r < n; ++r)
            for (int64_t c = 0; c < n; ++c) {
                ci[(size_t)r * n + c] = c;
                v[(size_t)r * n + c] = a[(size_t)c * n + r];  // row r of the column-major matrix
            }
        *out = handle_create(n, 0, n, n * n, rp.data(), ci.data(), v.data(), 0, 0, 1, nullptr, opts);
        return (int)RBL_OK;
    });
}

int rbl_nccl_unique_id(void* uid128) {
    return guarded([&] {
        std::string err;
        if (!uid128 || !Comm::unique_id(uid128, err)) throw Error(RBL_NCCL_ERROR, err.empty() ? "null uid" : err);
        return (int)RBL_OK;
    });
}

int rbl_create_sharded(int64_t n, int64_t row0, int64_t nloc, int64_t nnz_loc, const int64_t* rowptr,
                       const int64_t* colidx_global, const double* vals, int index_base, int rank, int world,
                       const void* nccl_uid, const rbl_options* opts, rbl_handle** out) {
    return guarded([&] {
        if (!out) throw Error(RBL_INVALID, "rbl_create_sharded: out is null");
        if (world < 1 || rank < 0 || rank >= world) throw Error(RBL_INVALID, "rbl_create_sharded: bad rank/world");
        if (world > 1 && !nccl_uid) throw Error(RBL_INVALID, "rbl_create_sharded: nccl_uid is null");
        *out = handle_create(n, row0, nloc, nnz_loc, rowptr, colidx_global, vals, index_base, rank, world, nccl_uid, opts);
        return (int)RBL_OK;
    });
}

int rbl_release_cached_memory(void) {
    slab_cache_release_all();
    Comm::release_cached();
    std::lock_guard<std::mutex> lk(g_group_mu);
    g_group_uid.clear();   // the parked communicators of the group ids are gone
    return RBL_OK;
}

int rbl_destroy(rbl_handle* h) {
    delete h;
    return RBL_OK;
}

int rbl_solve(rbl_handle* h, int64_t k, int64_t b, const double* omega, double* d_out, void* v_out, rbl_stats* stats) {
    return guarded([&] {
        if (!h) throw Error(RBL_INVALID, "rbl_solve: null handle");
        if (!h->parts.empty()) return solve_group(h, k, b, omega, d_out, v_out, stats);
        SolveIO io;
        io.omega = omega; io.ld_omega = h->nloc; io.v = v_out; io.ldv = h->nloc;
        return solve(h, k, b, io, d_out, stats);
    });
}

int rbl_solve_device(rbl_handle* h, int64_t k, int64_t b, const void* omega_dev, double* d_out, void* v_dev,
                     rbl_stats* stats) {
    return guarded([&] {
        if (!h) throw Error(RBL_INVALID, "rbl_solve_device: null handle");
        if (!h->parts.empty()) throw Error(RBL_INVALID, "rbl_solve_device: not available on multi-GPU group handles (use rbl_solve)");
        SolveIO io;
        io.omega = (const double*)omega_dev; io.ld_omega = h->nloc; io.omega_on_device = true;
        io.v = v_dev; io.ldv = h->nloc; io.v_on_device = true;
        return solve(h, k, b, io, d_out, stats);
    });
}

int rbl_query_memory(int device, int64_t* free_bytes, int64_t* total_bytes) {
    return guarded([&] {
        need_device();
        if (device >= 0) RBL_CUDA(cudaSetDevice(device));
        size_t f = 0, t = 0;
        RBL_CUDA(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = (int64_t)f;
        if (total_bytes) *total_bytes = (int64_t)t;
        return (int)RBL_OK;
    });
}

int rbl_plan_blocks(rbl_handle* h, int64_t k, int64_t b, int64_t* blocks_out) {
    return guarded([&] {
        if (!h || !blocks_out || b < 1 || b > 32 || k < 0) throw Error(RBL_INVALID, "rbl_plan_blocks: bad arguments");
        rbl_handle* hh = h->parts.empty() ? h : h->parts[0];
        RBL_CUDA(cudaSetDevice(hh->device));
        MemPlan p = plan_memory(hh, k, (int)b, hh->n / b + 1);   // a Krylov basis never exceeds n columns
        *blocks_out = p.m_fit > 0 ? p.m_fit : 0;
        return (int)RBL_OK;
    });
}
int rbl_buffer_blocks(rbl_handle* h, int64_t b, int64_t* blocks_out) { return rbl_plan_blocks(h, 0, b, blocks_out); }

int rbl_krylov_info(rbl_handle* h, int64_t* blocks_out, int64_t* b_out) {
    return guarded([&] {
        if (!h) throw Error(RBL_INVALID, "rbl_krylov_info: null handle");
        if (!h->parts.empty()) throw Error(RBL_INVALID, "rbl_krylov_info: single-GPU handles only");
        if (blocks_out) *blocks_out = h->last.blocks;
        if (b_out) *b_out = h->last.b;
        return (int)RBL_OK;
    });
}
int rbl_krylov_block(rbl_handle* h, int64_t j, double* out_colmajor) {
    return guarded([&] {
        if (!h || !out_colmajor) throw Error(RBL_INVALID, "rbl_krylov_block: bad arguments");
        if (!h->parts.empty() || h->world != 1) throw Error(RBL_INVALID, "rbl_krylov_block: single-GPU handles only");
        krylov_block(h, j, out_colmajor);
        return (int)RBL_OK;
    });
}
int rbl_orthogonality(rbl_handle* h, double* max_abs_out, double* fro_out) {
    return guarded([&] {
        if (!h) throw Error(RBL_INVALID, "rbl_orthogonality: null handle");
        if (!h->parts.empty()) throw Error(RBL_INVALID, "rbl_orthogonality: per-rank handles only");
        orthogonality(h, max_abs_out, fro_out);
        return (int)RBL_OK;
    });
}

// ---- kernel-level exports (host in, host out) ------------------------------------------------------
int rbl_spmm(rbl_handle* h, int64_t b, const double* q, double* u) {
    return guarded([&] {
        if (!h || !q || !u || b < 1 || b > 32) throw Error(RBL_INVALID, "rbl_spmm: bad arguments");
        if (h->world != 1 || !h->parts.empty()) throw Error(RBL_INVALID, "rbl_spmm: single-GPU handles only");
        RBL_CUDA(cudaSetDevice(h->device));
        const int B = padded_block((int)b);
        const int64_t n = h->nloc;
        std::vector<double> qp((size_t)n * B, 0.0), up((size_t)n * B);
        for (int64_t r = 0; r < n; ++r)
            for (int c = 0; c < b; ++c) qp[(size_t)r * B + c] = q[(size_t)r * b + c];
        DevBuf<double> dq, du;
        dq.alloc(qp.size());
        du.alloc(up.size());
        RBL_CUDA(cudaMemcpy(dq.p, qp.data(), qp.size() * 8, cudaMemcpyHostToDevice));
        const SpmmCoef cf = (h->opt.op == RBL_OP_SHIFT_MINUS_A) ? SpmmCoef{-1.0, h->opt.sigma, 0.0} : SpmmCoef{1.0, 0.0, 0.0};
        SpmmWindows wtmp = h->spmm_wt;
        size_t wsm = 0;
        if (wtmp.nwin > 0 && spmm_window_supported(B) && spmm_window_stages(wtmp, B, &wsm) > 0)
            launch_spmm_window(B, n, n, h->wsp->d_rowptr.p, h->wsp->d_rel.p, h->wsp->d_vals.p, dq.p, du.p, cf, nullptr, h->spmm_wt, h->stream);
        else
            launch_spmm(B, n, h->wsp->d_rowptr.p, h->wsp->d_colidx.p, h->wsp->d_vals.p, dq.p, du.p, cf, nullptr, h->stream);
        RBL_CUDA(cudaStreamSynchronize(h->stream));
        RBL_CUDA(cudaMemcpy(up.data(), du.p, up.size() * 8, cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r < n; ++r)
            for (int c = 0; c < b; ++c) u[(size_t)r * b + c] = up[(size_t)r * B + c];
        return (int)RBL_OK;
    });
}

namespace {
struct PadBlock {
    std::vector<double> host;
    DevBuf<double> dev;
    void upload(int64_t n, int b, int B, const double* src) {
        host.assign((size_t)n * B, 0.0);
        for (int64_t r = 0; r < n; ++r)
            for (int c = 0; c < b; ++c) host[(size_t)r * B + c] = src[(size_t)r * b + c];
        dev.alloc(host.size());
        RBL_CUDA(cudaMemcpy(dev.p, host.data(), host.size() * 8, cudaMemcpyHostToDevice));
    }
    void download(int64_t n, int b, int B, double* dst) {
        RBL_CUDA(cudaMemcpy(host.data(), dev.p, host.size() * 8, cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r < n; ++r)
            for (int c = 0; c < b; ++c) dst[(size_t)r * b + c] = host[(size_t)r * B + c];
    }
};
}  // namespace

int rbl_gram(int64_t n, int64_t b, const double* x, const double* y, double* cout) {
    return guarded([&] {
        need_device();
        if (!x || !y || !cout || b < 1 || b > 32 || n < 1) throw Error(RBL_INVALID, "rbl_gram: bad arguments");
        const int B = padded_block((int)b);
        PadBlock X, Y;
        X.upload(n, (int)b, B, x);
        Y.upload(n, (int)b, B, y);
        const int grid = rowop_grid(B, n);
        DevBuf<double> part, G;
        part.alloc((size_t)grid * B * B);
        G.alloc((size_t)B * B);
        RowOpArgs a;
        a.n = n; a.y = Y.dev.p; a.gram_z = X.dev.p; a.do_gram = 1; a.partials = part.p;
        launch_rowop(B, a, grid, 0);
        launch_reduce_partials(part.p, grid, B * B, G.p, 0);
        RBL_CUDA(cudaDeviceSynchronize());
        std::vector<double> g((size_t)B * B);
        RBL_CUDA(cudaMemcpy(g.data(), G.p, g.size() * 8, cudaMemcpyDeviceToHost));
        for (int r = 0; r < b; ++r)
            for (int c = 0; c < b; ++c) cout[(size_t)r * b + c] = g[(size_t)r * B + c];
        return (int)RBL_OK;
    });
}

int rbl_block_qr(int64_t n, int64_t b, double* u, double* r_out, int32_t* deflated_out) {
    return guarded([&] {
        need_device();
        if (!u || !r_out || b < 1 || b > 32 || n < 1) throw Error(RBL_INVALID, "rbl_block_qr: bad arguments");
        const int B = padded_block((int)b);
        PadBlock U;
        U.upload(n, (int)b, B, u);
        const int grid = rowop_grid(B, n);
        DevBuf<double> part, G;
        DevBuf<QrState> qr;
        part.alloc((size_t)grid * B * B);
        G.alloc((size_t)B * B);
        qr.alloc(1);
        RBL_CUDA(cudaMemset(qr.p, 0, sizeof(QrState)));
        RowOpArgs g0;
        g0.n = n; g0.y = U.dev.p; g0.do_gram = 1; g0.partials = part.p;
        launch_rowop(B, g0, grid, 0);
        RowOpArgs a;
        a.n = n; a.y = U.dev.p; a.rinv = qr.p->Rinv; a.write_y = 1; a.do_gram = 1; a.partials = part.p;
        for (int pass = 1; pass <= 3; ++pass) {
            launch_reduce_partials(part.p, grid, B * B, G.p, 0);
            launch_chol(B, G.p, qr.p, pass, n, 1, 1e-12, 0);
            RowOpArgs ap = a;
            if (pass == 3) { ap.skip_flag = &qr.p->need_more; ap.do_gram = 0; ap.partials = nullptr; }
            launch_rowop(B, ap, grid, 0);
        }
        RBL_CUDA(cudaDeviceSynchronize());
        U.download(n, (int)b, B, u);
        std::unique_ptr<QrState> hq(new QrState);
        RBL_CUDA(cudaMemcpy(hq.get(), qr.p, sizeof(QrState), cudaMemcpyDeviceToHost));
        for (int r = 0; r < b; ++r)
            for (int c = 0; c < b; ++c) r_out[(size_t)r * b + c] = hq->R[(size_t)r * B + c];
        if (deflated_out)
            for (int c = 0; c < b; ++c) deflated_out[c] = hq->deflated[c];
        return (int)RBL_OK;
    });
}

int rbl_reorth(int64_t n, int64_t b, int64_t m, int storage_fp32, const void* qbuf, double* w0, double* w1,
               void* c_out, int impl) {
    return guarded([&] {
        need_device();
        if (!qbuf || !w0 || !w1 || b < 1 || b > 32 || n < 1 || m < 1) throw Error(RBL_INVALID, "rbl_reorth: bad arguments");
        const int B = padded_block((int)b);
        const bool hs = impl != 1 && reorth_h_supported(B, storage_fp32);
        if (impl == 2) throw Error(RBL_INVALID, "rbl_reorth: impl 2 (3xTF32) was removed");
        if (impl >= 3 && !hs) throw Error(RBL_INVALID, "rbl_reorth: tensor-core path needs fp32 storage and padded block size 16 or 32");
        const size_t ssz = storage_fp32 ? 4 : 8;
        // pad the stored blocks
        std::vector<unsigned char> hb((size_t)m * n * B * ssz, 0);
        for (int64_t j = 0; j < m; ++j)
            for (int64_t r = 0; r < n; ++r)
                for (int c = 0; c < b; ++c) {
                    const size_t s = ((size_t)j * n + r) * b + c, d = ((size_t)j * n + r) * B + c;
                    if (storage_fp32) ((float*)hb.data())[d] = ((const float*)qbuf)[s];
                    else ((double*)hb.data())[d] = ((const double*)qbuf)[s];
                }
        DevBuf<unsigned char> dbuf, dC, dpart;
        dbuf.alloc(hb.size());
        RBL_CUDA(cudaMemcpy(dbuf.p, hb.data(), hb.size(), cudaMemcpyHostToDevice));
        PadBlock W0, W1;
        W0.upload(n, (int)b, B, w0);
        W1.upload(n, (int)b, B, w1);
        ReorthPlan p = reorth_plan(B, storage_fp32, n, m);
        dC.alloc((size_t)m * B * 2 * B * ssz);
        dpart.alloc(p.partial_elems * ssz);
        if (impl != 1 && reorth_d_supported(B, storage_fp32)) {
            launch_reorth_gram_d(p, dbuf.p, n * B, W0.dev.p, W1.dev.p, dpart.p, dC.p, 0);
            launch_reorth_update_d(p, dbuf.p, n * B, dC.p, W0.dev.p, W1.dev.p, nullptr, 0);
            RBL_CUDA(cudaDeviceSynchronize());
        } else if (hs) {
            DevBuf<float> scratch;
            scratch.alloc(reorth_h_scratch_words(B, n, m));
            const int presplit = impl != 3;   // default: the split16 slab format the solver uses (split16.h)
            DevBuf<unsigned char> dsplit;
            const void* slab = dbuf.p;
            if (presplit) {
                dsplit.alloc(hb.size());
                launch_encode_split(B, n * m, (const float*)dbuf.p, dsplit.p, reorth_h_scale(n), 0);
                slab = dsplit.p;
            }
            launch_reorth_gram_h(p, n, slab, n * B, W0.dev.p, W1.dev.p, dpart.p, dC.p, scratch.p, m, presplit, 0);
            launch_reorth_coeff_h(p, dC.p, scratch.p, m, 0, 0);
            launch_reorth_update_h(p, n, slab, n * B, W0.dev.p, W1.dev.p, nullptr, scratch.p, m, presplit, 0);
            RBL_CUDA(cudaDeviceSynchronize());
        } else {
            launch_reorth_gram(p, dbuf.p, n * B, W0.dev.p, W1.dev.p, dpart.p, dC.p, 0);
            launch_reorth_update(p, dbuf.p, n * B, dC.p, W0.dev.p, W1.dev.p, nullptr, 0);
        }
        RBL_CUDA(cudaDeviceSynchronize());
        W0.download(n, (int)b, B, w0);
        W1.download(n, (int)b, B, w1);
        if (c_out) {
            std::vector<unsigned char> hc((size_t)m * B * 2 * B * ssz);
            RBL_CUDA(cudaMemcpy(hc.data(), dC.p, hc.size(), cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < m; ++j)
                for (int c = 0; c < b; ++c)
                    for (int t = 0; t < 2 * b; ++t) {
                        const int tt = t < b ? t : (B + (t - (int)b));
                        const size_t s = ((size_t)j * B + c) * 2 * B + tt, d = ((size_t)j * b + c) * 2 * b + t;
                        if (storage_fp32) ((float*)c_out)[d] = ((float*)hc.data())[s];
                        else ((double*)c_out)[d] = ((double*)hc.data())[s];
                    }
        }
        return (int)RBL_OK;
    });
}

int rbl_ritz(int64_t n, int64_t b, int64_t m, int64_t k, int storage_fp32, const void* qbuf, const double* s,
             void* v_out, int impl) {
    return guarded([&] {
        need_device();
        if (!qbuf || !s || !v_out || b < 1 || b > 32 || n < 1 || m < 1 || k < 1) throw Error(RBL_INVALID, "rbl_ritz: bad arguments");
        const int B = padded_block((int)b);
        const size_t ssz = storage_fp32 ? 4 : 8;
        const int kpad = (int)((k + 15) / 16 * 16);
        std::vector<unsigned char> hb((size_t)m * n * B * ssz, 0), hs((size_t)m * B * kpad * ssz, 0);
        for (int64_t j = 0; j < m; ++j)
            for (int64_t r = 0; r < n; ++r)
                for (int c = 0; c < b; ++c) {
                    const size_t so = ((size_t)j * n + r) * b + c, d = ((size_t)j * n + r) * B + c;
                    if (storage_fp32) ((float*)hb.data())[d] = ((const float*)qbuf)[so];
                    else ((double*)hb.data())[d] = ((const double*)qbuf)[so];
                }
        for (int64_t j = 0; j < m; ++j)
            for (int c = 0; c < b; ++c)
                for (int64_t t = 0; t < k; ++t) {
                    const double v = s[((size_t)j * b + c) * k + t];
                    const size_t d = ((size_t)j * B + c) * kpad + t;
                    if (storage_fp32) ((float*)hs.data())[d] = (float)v;
                    else ((double*)hs.data())[d] = v;
                }
        DevBuf<unsigned char> dbuf, dS, dV;
        dbuf.alloc(hb.size());
        dS.alloc(hs.size());
        dV.alloc((size_t)n * k * ssz);
        RBL_CUDA(cudaMemcpy(dbuf.p, hb.data(), hb.size(), cudaMemcpyHostToDevice));
        RBL_CUDA(cudaMemcpy(dS.p, hs.data(), hs.size(), cudaMemcpyHostToDevice));
        const bool use_tc = impl != 1 && reorth_h_supported(B, storage_fp32);
        if (impl == 4 && !use_tc) throw Error(RBL_INVALID, "rbl_ritz: tensor-core path needs fp32 storage and padded block size 16 or 32");
        if (use_tc) {
            DevBuf<unsigned char> dsplit;
            DevBuf<unsigned> words;
            dsplit.alloc(hb.size());
            words.alloc(ritz_h_scratch_words(B, m, kpad));
            const float scale = reorth_h_scale(n);
            launch_encode_split(B, n * m, (const float*)dbuf.p, dsplit.p, scale, 0);
            launch_ritz_h(B, n, m, (int)k, kpad, dsplit.p, n * B, dS.p, dV.p, n, storage_fp32, scale, words.p, 0);
            RBL_CUDA(cudaDeviceSynchronize());
        } else {
            launch_ritz(B, storage_fp32, n, m, (int)k, kpad, dbuf.p, n * B, dS.p, dV.p, n, storage_fp32, 0.f, 0);
        }
        RBL_CUDA(cudaDeviceSynchronize());
        RBL_CUDA(cudaMemcpy(v_out, dV.p, (size_t)n * k * ssz, cudaMemcpyDeviceToHost));
        return (int)RBL_OK;
    });
}

// ---- host-only exports -----------------------------------------------------------------------------
int rbl_band_eig_topk(int64_t N, int64_t kd, const double* ab, int64_t k, const double* bi, int64_t b, double tol,
                      int threads, double* d_out, double* s_out, double* resid_out, int32_t* converged_out) {
    return guarded([&] {
        if (N < 1 || kd < 0 || !ab || k < 1 || k > N) throw Error(RBL_INVALID, "rbl_band_eig_topk: bad arguments");
        BandSym T;
        T.from_lapack_lower(N, (int)kd, ab);
        BandTopK chk;
        chk.threads = threads > 0 ? threads : 1;
        std::vector<double> bir;
        if (bi) {  // column-major in, row-major inside
            bir.resize((size_t)b * b);
            for (int r = 0; r < b; ++r)
                for (int c = 0; c < b; ++c) bir[(size_t)r * b + c] = bi[(size_t)c * b + r];
        }
        TopKResult r = chk.check(T, bi ? bir.data() : nullptr, (int)b, k, tol, true);
        if (!r.have_all) throw Error(RBL_BREAKDOWN, "rbl_band_eig_topk: could not isolate k eigenpairs");
        for (int64_t j = 0; j < k; ++j) {
            if (d_out) d_out[j] = r.d[j];
            if (resid_out) resid_out[j] = r.resid[j];
        }
        if (s_out) std::memcpy(s_out, r.s.data(), (size_t)N * k * sizeof(double));
        if (converged_out) *converged_out = r.converged ? 1 : 0;
        return (int)RBL_OK;
    });
}

struct rbl_checker {
    BandTopK chk;
};

int rbl_checker_create(int threads, rbl_checker** out) {
    return guarded([&] {
        if (!out) throw Error(RBL_INVALID, "rbl_checker_create: null out");
        *out = new rbl_checker();
        (*out)->chk.threads = threads > 0 ? threads : 1;
        if (const char* v = getenv("RBL_CHECK_VERBOSE")) (*out)->chk.verbose = atoi(v);
        return (int)RBL_OK;
    });
}

int rbl_checker_check(rbl_checker* c, int64_t N, int64_t kd, const double* ab, int64_t k, const double* bi, int64_t b,
                      double tol, int force_full, double* d_out, double* s_out, double* resid_out,
                      int32_t* converged_out, int32_t* have_all_out, int64_t* stats_out) {
    return guarded([&] {
        if (!c || N < 1 || kd < 0 || !ab || k < 1 || k > N || !bi) throw Error(RBL_INVALID, "rbl_checker_check: bad arguments");
        BandSym T;
        T.from_lapack_lower(N, (int)kd, ab);
        std::vector<double> bir((size_t)b * b);
        for (int r = 0; r < b; ++r)
            for (int cc = 0; cc < b; ++cc) bir[(size_t)r * b + cc] = bi[(size_t)cc * b + r];
        const int full_before = c->chk.full_checks;
        const auto tc0 = std::chrono::steady_clock::now();
        TopKResult r = c->chk.check(T, bir.data(), (int)b, k, tol, force_full != 0);
        if (c->chk.verbose > 1)
            std::fprintf(stderr, "[rbl]   rbl_checker_check: check() took %.1f ms\n",
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count() * 1e3);
        if (converged_out) *converged_out = r.converged ? 1 : 0;
        if (have_all_out) *have_all_out = r.have_all ? 1 : 0;
        if (stats_out) {
            stats_out[0] = r.factorizations;
            stats_out[1] = c->chk.full_checks - full_before;
        }
        if (r.have_all) {
            for (int64_t j = 0; j < k; ++j) {
                if (d_out) d_out[j] = r.d[j];
                if (resid_out) resid_out[j] = r.resid[j];
            }
            if (s_out) std::memcpy(s_out, r.s.data(), (size_t)N * k * sizeof(double));
        }
        return (int)RBL_OK;
    });
}

int rbl_checker_set_seeds(rbl_checker* c, int64_t n_seed, int64_t k, const double* d, const double* s, const double* resid) {
    return guarded([&] {
        if (!c || n_seed < 1 || k < 1 || !d || !s) throw Error(RBL_INVALID, "rbl_checker_set_seeds: bad arguments");
        std::vector<double> rv;
        if (resid) rv.assign(resid, resid + k);
        c->chk.set_seeds(std::vector<double>(d, d + k), std::vector<double>(s, s + (size_t)n_seed * k), n_seed, k, resid ? &rv : nullptr);
        return (int)RBL_OK;
    });
}

int rbl_checker_set_need_seeds(rbl_checker* c, rbl_need_seeds_fn fn, void* user) {
    return guarded([&] {
        if (!c) throw Error(RBL_INVALID, "rbl_checker_set_need_seeds: null checker");
        if (fn) c->chk.need_seeds = [fn, user](int64_t N) { fn(user, N); };
        else c->chk.need_seeds = nullptr;
        return (int)RBL_OK;
    });
}

int rbl_checker_destroy(rbl_checker* c) {
    delete c;
    return RBL_OK;
}

int rbl_band_count_below(int64_t N, int64_t kd, const double* ab, double x, int64_t* count_out) {
    return guarded([&] {
        if (N < 1 || kd < 0 || !ab || !count_out) throw Error(RBL_INVALID, "rbl_band_count_below: bad arguments");
        BandSym T;
        T.from_lapack_lower(N, (int)kd, ab);
        *count_out = band_count_below(T, x);
        return (int)RBL_OK;
    });
}

int rbl_partition_rows(int64_t n, int world, int64_t* row_starts_out) {
    return guarded([&] {
        if (n < 0 || world < 1 || !row_starts_out) throw Error(RBL_INVALID, "rbl_partition_rows: bad arguments");
        partition_rows(n, world, row_starts_out);
        return (int)RBL_OK;
    });
}

int rbl_halo_plan(int64_t n, int world, const int64_t* row_starts, int rank, int64_t nloc, int64_t nnz_loc,
                  const int64_t* rowptr, const int64_t* colidx_global, int64_t* n_halo_out, int64_t* halo_cols_out,
                  int64_t* halo_owner_ptr_out, int32_t* colidx_local_out) {
    return guarded([&] {
        if (!row_starts || !rowptr || (nnz_loc > 0 && !colidx_global) || !n_halo_out || rank < 0 || rank >= world)
            throw Error(RBL_INVALID, "rbl_halo_plan: bad arguments");
        HaloPlan p;
        if (!halo_plan(n, world, row_starts, rank, nloc, nnz_loc, rowptr, colidx_global, 0, p))
            throw Error(RBL_INVALID, "rbl_halo_plan: bad column index or row range");
        *n_halo_out = (int64_t)p.halo_cols.size();
        if (halo_cols_out) std::memcpy(halo_cols_out, p.halo_cols.data(), p.halo_cols.size() * 8);
        if (halo_owner_ptr_out) std::memcpy(halo_owner_ptr_out, p.halo_owner_ptr.data(), p.halo_owner_ptr.size() * 8);
        if (colidx_local_out) std::memcpy(colidx_local_out, p.colidx_local.data(), p.colidx_local.size() * 4);
        return (int)RBL_OK;
    });
}

int rbl_spmm_schedule(int64_t nrows, int64_t nown, const int32_t* rowptr, const int32_t* colidx, int slots,
                      int64_t* entries_out, int32_t* order_out, int64_t* info_out) {
    return guarded([&] {
        if (!rowptr || !colidx || !entries_out || nrows < 0) throw Error(RBL_INVALID, "rbl_spmm_schedule: bad arguments");
        std::vector<int> order;
        SpmmSchedule sc;
        const bool ok = spmm_plan_schedule(nrows, nown, rowptr, colidx, slots > 0 ? slots : spmm_sched_default_slots(16), order, &sc);
        *entries_out = ok ? (int64_t)order.size() : 0;
        if (ok && order_out) std::memcpy(order_out, order.data(), order.size() * sizeof(int));
        if (ok && info_out) {
            const int64_t v[12] = {sc.dims, sc.stride[1], sc.stride[2], sc.ext[0], sc.ext[1], sc.ext[2],
                                   sc.patch[0], sc.patch[1], sc.patch[2], sc.slots, sc.npatch, sc.halo[0]};
            std::memcpy(info_out, v, sizeof(v));
        }
        return (int)RBL_OK;
    });
}

// ---- Matrix Market loader (benchmark.jl:21,28 mmread) ---------------------------------------------------------------
struct rbl_matrix {
    MmMatrix m;
};
int rbl_matrix_market_read(const char* path, int index_base, rbl_matrix** out, int64_t* n_out, int64_t* nnz_out) {
    return guarded([&] {
        if (!path || !out || (index_base != 0 && index_base != 1)) throw Error(RBL_INVALID, "rbl_matrix_market_read: bad arguments");
        std::unique_ptr<rbl_matrix> M(new rbl_matrix());
        std::string err;
        if (!read_matrix_market(path, index_base, M->m, err)) throw Error(RBL_INVALID, "rbl_matrix_market_read: " + err);
        if (n_out) *n_out = M->m.n;
        if (nnz_out) *nnz_out = (int64_t)M->m.nzval.size();
        *out = M.release();
        return (int)RBL_OK;
    });
}
int rbl_matrix_arrays(rbl_matrix* m, const int64_t** colptr, const int64_t** rowval, const double** nzval) {
    if (!m) return RBL_INVALID;
    if (colptr) *colptr = m->m.colptr.data();
    if (rowval) *rowval = m->m.rowval.data();
    if (nzval) *nzval = m->m.nzval.data();
    return RBL_OK;
}
int rbl_matrix_free(rbl_matrix* m) {
    delete m;
    return RBL_OK;
}

int rbl_spmm_bench(rbl_handle* h, int64_t b, int variant, int grid_mult, int iters, int flush, int with_z, double* us_out,
                   int64_t* mismatch_out) {
    return guarded([&] {
        if (!h || !us_out || iters < 1) throw Error(RBL_INVALID, "rbl_spmm_bench: bad arguments");
        if (h->world != 1 || !h->parts.empty()) throw Error(RBL_INVALID, "rbl_spmm_bench: single-GPU handles only");
        RBL_CUDA(cudaSetDevice(h->device));
        unsigned long long bad = 0;
        *us_out = spmm_lab_run(h, (int)b, variant, grid_mult, iters, flush, with_z, &bad);
        if (mismatch_out) *mismatch_out = (int64_t)bad;
        if (*us_out < 0) throw Error(RBL_INVALID, "rbl_spmm_bench: variant not available for this matrix / block size");
        return (int)RBL_OK;
    });
}

int rbl_microbench(int which, int64_t size, int iters, double* result_out) {
    return guarded([&] {
        need_device();
        if (!result_out) throw Error(RBL_INVALID, "rbl_microbench: null result");
        *result_out = microbench(which, size, iters);
        if (*result_out < 0) throw Error(RBL_CUDA_ERROR, "rbl_microbench failed");
        return (int)RBL_OK;
    });
}

}  // extern "C"

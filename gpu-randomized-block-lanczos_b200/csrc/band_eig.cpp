// Host-side eigen-check of the Lanczos band matrix T.  See band_eig.h.
//
// Reference call sites replaced (same host-side position, same inputs and outputs):
//   common.jl:36-48  dsbev('V','L',T)      -> BandTopK::check (spectrum slicing, k pairs only)
//   common.jl:50-54  sort_eig_abs          -> selection of the k largest |lambda|
//   common.jl:56-65  check_convergence     -> residual bounds ||B_i s_last|| <= tol for all k
#include "band_eig.h"

#include <algorithm>
#include <array>
#if defined(__AVX2__)
#include <immintrin.h>
#endif
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <random>
#include <thread>

namespace rbl {

// ------------------------------------------------------------------------------------------------ BandSym
void BandSym::reset(int64_t n, int kd_) {
    N = n;
    kd = kd_;
    F.assign((size_t)N * (2 * kd + 1), 0.0);
    norm_inf = 0.0;
}

void BandSym::from_lapack_lower(int64_t n, int kd_, const double* ab) {
    reset(n, kd_);
    for (int64_t c = 0; c < N; ++c)
        for (int d = 0; d <= kd; ++d) {
            int64_t r = c + d;
            if (r >= N) break;
            double v = ab[(size_t)c * (kd + 1) + d];
            at(r, c) = v;
            at(c, r) = v;
        }
    update_norm();
}

void BandSym::update_norm() {
    const int W = 2 * kd + 1;
    double m = 0.0, lo = 1e300, hi = -1e300;
    for (int64_t r = 0; r < N; ++r) {
        double s = 0.0;
        const double* row = &F[(size_t)r * W];
        for (int t = 0; t < W; ++t) s += std::fabs(row[t]);
        m = std::max(m, s);
        const double off = s - std::fabs(row[kd]);
        lo = std::min(lo, row[kd] - off);
        hi = std::max(hi, row[kd] + off);
    }
    norm_inf = m;
    gersh_lo = N ? lo : 0.0;
    gersh_hi = N ? hi : 0.0;
}

void BandSym::matvec(const double* x, double* y) const {
    const int W = 2 * kd + 1;
    for (int64_t r = 0; r < N; ++r) {
        const double* row = &F[(size_t)r * W];
        int64_t c0 = r - kd;
        int t0 = c0 < 0 ? (int)(-c0) : 0;
        int t1 = (c0 + W > N) ? (int)(N - c0) : W;
        double s = 0.0;
        for (int t = t0; t < t1; ++t) s += row[t] * x[c0 + t];
        y[r] = s;
    }
}

// ------------------------------------------------------------------------------------------------ BandLU
// Row r is eliminated against the already triangularised rows c = r-kd .. r-1, interchanging the
// working row with the pivot row when its entry is larger (pairwise pivoting).  Only rows <= r take
// part, so after row r the leading (r+1) x (r+1) principal submatrix is upper triangular and
//     sign det T_r = (-1)^{#swaps} prod_{i<=r} sign U_ii ;
// the number of sign changes along r is the number of eigenvalues of T below the shift (Sturm).
void BandLU::factor(const BandSym& T, double shift) {
    if (T.cancel && T.cancel->load(std::memory_order_relaxed)) throw Cancelled{};
    while (T.pause && T.pause->load(std::memory_order_relaxed)) {
        std::this_thread::sleep_for(std::chrono::microseconds(200));
        if (T.cancel && T.cancel->load(std::memory_order_relaxed)) throw Cancelled{};
    }
    N = T.N;
    kd = T.kd;
    shift_ = shift;
    P_ = 1;
    nneg = 0;
    ck_row = -1;
    run(T, 0);
}

// T grew at its end only (Lanczos appends block rows; rows < N_old - kd are unchanged) and the shift is the
// same: restore the state saved just before row N_old - kd and eliminate only the new / changed rows.
bool BandLU::resume(const BandSym& T) {
    if (ck_row < 0 || T.kd != kd || T.N < N || ck_row > T.N) return false;
    const int W = 2 * kd + 1;
    const int64_t r0 = ck_row;
    const int64_t c0 = std::max<int64_t>(0, r0 - kd);
    std::copy(ck_U.begin(), ck_U.begin() + (size_t)(r0 - c0) * W, U.begin() + (size_t)c0 * W);
    P_ = ck_P;
    nneg = ck_nneg;
    N = T.N;
    run(T, r0);
    return true;
}

void BandLU::run(const BandSym& T, int64_t r_start) {
    switch (kd) {  // compile-time band widths for the common block sizes: fixed-length vector loops
        case 4: run_t<4>(T, r_start); break;
        case 8: run_t<8>(T, r_start); break;
        case 16: run_t<16>(T, r_start); break;
        case 32: run_t<32>(T, r_start); break;
        default: run_t<0>(T, r_start); break;
    }
}

template <int KD>
void BandLU::run_t(const BandSym& T, int64_t r_start) {
    const int kd = KD > 0 ? KD : this->kd;
    const int W = 2 * kd + 1;
    U.resize((size_t)N * W);
    L.resize((size_t)N * std::max(kd, 1));
    sw.resize((size_t)N * std::max(kd, 1));
    w.assign((size_t)3 * kd + 2, 0.0);
    const double shift = shift_;
    const double pivmin = std::max(T.norm_inf, 1e-290) * 1e-20;
    int P = P_;
    double* __restrict__ wp = w.data();
    double* __restrict__ Ubase = U.data();
    double* __restrict__ Lbase = L.data();
    uint8_t* __restrict__ swbase = sw.data();
    const double* __restrict__ Fbase = T.F.data();
    const int64_t ck_at = N - kd;  // rows >= this change when T grows
    for (int64_t r = r_start; r < N; ++r) {
        if (r == ck_at && ck_at >= 0) {
            const int64_t c0 = std::max<int64_t>(0, r - kd);
            ck_U.assign(Ubase + (size_t)c0 * W, Ubase + (size_t)r * W);
            ck_P = P;
            ck_nneg = nneg;
            ck_row = r;
        }
        const int64_t base = r - kd;
        const double* __restrict__ row = Fbase + (size_t)r * W;
        for (int t = 0; t < W; ++t) wp[t] = row[t];
        for (int t = W; t < 3 * kd + 1; ++t) wp[t] = 0.0;
        wp[kd] -= shift;
        const int prev = P;
        const int64_t cstart = base < 0 ? 0 : base;
        double* __restrict__ Lr = Lbase + (size_t)r * kd;
        uint8_t* __restrict__ sr = swbase + (size_t)r * kd;
        for (int64_t c = cstart; c < r; ++c) {
            const int wi = (int)(c - base);
            double* __restrict__ Uc = Ubase + (size_t)c * W;
            double* __restrict__ ws = wp + wi;
            const double wc = ws[0];
            const double piv = Uc[0];
            if (wc == 0.0) {
                Lr[wi] = 0.0;
                sr[wi] = 0;
                continue;
            }
            if (std::fabs(wc) > std::fabs(piv)) {
                // interchange: the working row becomes the pivot row; fused swap + elimination
                const double m = piv / wc;
                P = ((piv < 0) != (wc < 0)) ? P : -P;
#pragma GCC ivdep
                for (int t = 0; t < W; ++t) {
                    const double a = ws[t], b = Uc[t];
                    Uc[t] = a;
                    ws[t] = b - m * a;
                }
                ws[0] = 0.0;
                Lr[wi] = m;
                sr[wi] = 1;
            } else {
                const double m = wc / piv;
#pragma GCC ivdep
                for (int t = 1; t < W; ++t) ws[t] -= m * Uc[t];
                ws[0] = 0.0;
                Lr[wi] = m;
                sr[wi] = 0;
            }
        }
        double* __restrict__ Ur = Ubase + (size_t)r * W;
        for (int t = 0; t < W; ++t) Ur[t] = wp[kd + t];
        if (std::fabs(Ur[0]) < pivmin) Ur[0] = -pivmin;
        const int cur = P * (Ur[0] < 0 ? -1 : 1);
        if (cur != prev) ++nneg;
        P = cur;
    }
    P_ = P;
}

void BandLU::solve(double* v) const {
    switch (kd) {
        case 4: solve_t<4>(v); break;
        case 8: solve_t<8>(v); break;
        case 16: solve_t<16>(v); break;
        case 32: solve_t<32>(v); break;
        default: solve_t<0>(v); break;
    }
}

// Forward: the recorded row operations applied to v (a chain of 2x2 transforms through the running entry).  Backward:
// v[r] = (v[r] - sum_t U[r][t] v[r+t]) / U[r][0]; only the t = 1 term depends on the entry computed one row earlier, so
// everything else is summed first, off the serial chain (a plain `s -= U[t]*v[r+t]` loop is a chain of 2kd dependent
// fused multiply-adds per row and took longer than the whole forward sweep).
template <int KD>
void BandLU::solve_t(double* __restrict__ v) const {
    const int kd = KD > 0 ? KD : this->kd;
    const int W = 2 * kd + 1;
    for (int64_t r = 0; r < N; ++r) {
        const int64_t base = r - kd;
        const int64_t cstart = base < 0 ? 0 : base;
        const double* __restrict__ Lr = &L[(size_t)r * kd];
        const uint8_t* __restrict__ sr = &sw[(size_t)r * kd];
        double vr = v[r];
        for (int64_t c = cstart; c < r; ++c) {
            const int wi = (int)(c - base);
            const double vc = v[c];
            const bool s = sr[wi] != 0;
            const double pc = s ? vr : vc, qr = s ? vc : vr;
            v[c] = pc;
            vr = qr - Lr[wi] * pc;
        }
        v[r] = vr;
    }
    const double* __restrict__ Ub = U.data();
    int64_t r = N - 1;
    // the last 2kd rows have short sums
    for (; r >= 0 && r > N - 1 - (W - 1); --r) {
        const double* Ur = Ub + (size_t)r * W;
        double s = v[r];
        const int tmax = (int)(N - 1 - r);
        for (int t = 1; t <= tmax; ++t) s -= Ur[t] * v[r + t];
        v[r] = s / Ur[0];
    }
    if (KD >= 4) {
        for (; r >= 0; --r) {
            const double* __restrict__ Ur = Ub + (size_t)r * W;
            const double* __restrict__ vv = v + r;
            __m256d acc0 = _mm256_setzero_pd(), acc1 = _mm256_setzero_pd();
            int t = 8;
            for (; t + 8 <= 2 * kd; t += 8) {
                acc0 = _mm256_fmadd_pd(_mm256_loadu_pd(Ur + t), _mm256_loadu_pd(vv + t), acc0);
                acc1 = _mm256_fmadd_pd(_mm256_loadu_pd(Ur + t + 4), _mm256_loadu_pd(vv + t + 4), acc1);
            }
            for (; t + 4 <= 2 * kd; t += 4) acc0 = _mm256_fmadd_pd(_mm256_loadu_pd(Ur + t), _mm256_loadu_pd(vv + t), acc0);
            alignas(32) double h[4];
            _mm256_store_pd(h, _mm256_add_pd(acc0, acc1));
            const int tl = kd >= 4 ? 2 * kd : 0;
            double s0 = Ur[2] * vv[2] + Ur[4] * vv[4] + Ur[6] * vv[6] + (h[0] + h[1]);
            double s1 = Ur[3] * vv[3] + Ur[5] * vv[5] + Ur[7] * vv[7] + (h[2] + h[3]);
            s0 += Ur[tl] * vv[tl];
            const double p = vv[0] - (s0 + s1);
            v[r] = (p - Ur[1] * vv[1]) / Ur[0];
        }
    } else {
        for (; r >= 0; --r) {
            const double* Ur = Ub + (size_t)r * W;
            double s = v[r];
            for (int t = 1; t < W; ++t) s -= Ur[t] * v[r + t];
            v[r] = s / Ur[0];
        }
    }
}

int64_t band_count_below(const BandSym& T, double x) {
    BandLU lu;
    lu.factor(T, x);
    return lu.nneg;
}

// ------------------------------------------------------------------------------------------------ helpers
namespace {

// (four independent AVX2 accumulators: a plain `s += a[i]*b[i]` loop cannot be vectorised without reassociation and ran
// at one fma per 4 cycles - the k^2/2 dot products of finalize_pairs alone took 40 ms of the accepting check)
inline double dot(const double* a, const double* b, int64_t n) {
    __m256d s0 = _mm256_setzero_pd(), s1 = s0, s2 = s0, s3 = s0;
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        s0 = _mm256_fmadd_pd(_mm256_loadu_pd(a + i), _mm256_loadu_pd(b + i), s0);
        s1 = _mm256_fmadd_pd(_mm256_loadu_pd(a + i + 4), _mm256_loadu_pd(b + i + 4), s1);
        s2 = _mm256_fmadd_pd(_mm256_loadu_pd(a + i + 8), _mm256_loadu_pd(b + i + 8), s2);
        s3 = _mm256_fmadd_pd(_mm256_loadu_pd(a + i + 12), _mm256_loadu_pd(b + i + 12), s3);
    }
    for (; i + 4 <= n; i += 4) s0 = _mm256_fmadd_pd(_mm256_loadu_pd(a + i), _mm256_loadu_pd(b + i), s0);
    alignas(32) double t[4];
    _mm256_store_pd(t, _mm256_add_pd(_mm256_add_pd(s0, s1), _mm256_add_pd(s2, s3)));
    double s = (t[0] + t[1]) + (t[2] + t[3]);
    for (; i < n; ++i) s += a[i] * b[i];
    return s;
}
inline double nrm2(const double* a, int64_t n) { return std::sqrt(dot(a, a, n)); }
inline void scal(double* a, double s, int64_t n) {
    for (int64_t i = 0; i < n; ++i) a[i] *= s;
}
inline void axpy(double* y, double a, const double* x, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

struct Work {
    BandLU lu;
    std::vector<double> y, t;
    std::mt19937_64 rng{12345};
    int nfac = 0;
    void random_unit(std::vector<double>& x, int64_t N) {
        std::normal_distribution<double> g(0.0, 1.0);
        x.resize(N);
        for (auto& e : x) e = g(rng);
        scal(x.data(), 1.0 / nrm2(x.data(), N), N);
    }
};

// Rayleigh quotient and residual of a unit vector.
void rayleigh(const BandSym& T, const std::vector<double>& x, std::vector<double>& t, double& theta, double& res) {
    t.resize(T.N);
    T.matvec(x.data(), t.data());
    theta = dot(x.data(), t.data(), T.N);
    double s = 0;
    for (int64_t i = 0; i < T.N; ++i) {
        double d = t[i] - theta * x[i];
        s += d * d;
    }
    res = std::sqrt(s);
}

// Rayleigh-quotient iteration from (theta, x).  Stops when the residual is below rtol*||T|| or stagnates.
// If lo < hi the iterate must stay inside (lo,hi); returns false when it leaves.
bool rqi(const BandSym& T, Work& wk, double& theta, std::vector<double>& x, double& res, double lo, double hi,
         int maxit, double rtol) {
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    double r0, th0;
    rayleigh(T, x, wk.t, th0, r0);
    res = r0;
    for (int it = 0; it < maxit; ++it) {
        if (res <= rtol * tn) return true;
        wk.lu.factor(T, theta);
        ++wk.nfac;
        wk.y = x;
        wk.lu.solve(wk.y.data());
        double nn = nrm2(wk.y.data(), N);
        if (!(nn > 0) || !std::isfinite(nn)) return false;
        scal(wk.y.data(), 1.0 / nn, N);
        double th, rs;
        rayleigh(T, wk.y, wk.t, th, rs);
        if (lo < hi && !(th > lo && th < hi)) return false;
        x.swap(wk.y);
        theta = th;
        if (rs > 0.5 * res && rs <= 2e-13 * tn && it >= 1) {  // stagnated at rounding level
            res = rs;
            return true;
        }
        res = rs;
    }
    return res <= 2e-13 * tn;
}

// Cyclic Jacobi for a small dense symmetric matrix H (m x m, row-major); eigenvectors in Y (columns).
void jacobi_eig(std::vector<double>& H, int m, std::vector<double>& evals, std::vector<double>& Y) {
    Y.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) Y[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) (i == j ? diag : off) += H[(size_t)i * m + j] * H[(size_t)i * m + j];
        if (off <= 1e-32 * std::max(diag, 1e-300)) break;
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                double apq = H[(size_t)p * m + q];
                if (apq == 0.0) continue;
                double app = H[(size_t)p * m + p], aqq = H[(size_t)q * m + q];
                double tau = (aqq - app) / (2.0 * apq);
                double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int i = 0; i < m; ++i) {
                    double hip = H[(size_t)i * m + p], hiq = H[(size_t)i * m + q];
                    H[(size_t)i * m + p] = c * hip - s * hiq;
                    H[(size_t)i * m + q] = s * hip + c * hiq;
                }
                for (int i = 0; i < m; ++i) {
                    double hpi = H[(size_t)p * m + i], hqi = H[(size_t)q * m + i];
                    H[(size_t)p * m + i] = c * hpi - s * hqi;
                    H[(size_t)q * m + i] = s * hpi + c * hqi;
                }
                for (int i = 0; i < m; ++i) {
                    double yip = Y[(size_t)i * m + p], yiq = Y[(size_t)i * m + q];
                    Y[(size_t)i * m + p] = c * yip - s * yiq;
                    Y[(size_t)i * m + q] = s * yip + c * yiq;
                }
            }
    }
    evals.resize(m);
    for (int i = 0; i < m; ++i) evals[i] = H[(size_t)i * m + i];
}

// Orthonormalise X[from..] against `against` and among themselves (two MGS passes).  Returns the
// smallest norm seen before normalisation of a column (duplicate detector).
double mgs(std::vector<std::vector<double>>& X, size_t from, const std::vector<const std::vector<double>*>& against,
           int64_t N) {
    double minn = 1e300;
    for (size_t j = from; j < X.size(); ++j) {
        double n0 = nrm2(X[j].data(), N);
        for (int pass = 0; pass < 2; ++pass) {
            for (auto* a : against) axpy(X[j].data(), -dot(a->data(), X[j].data(), N), a->data(), N);
            for (size_t i = 0; i < j; ++i) axpy(X[j].data(), -dot(X[i].data(), X[j].data(), N), X[i].data(), N);
        }
        double n1 = nrm2(X[j].data(), N);
        minn = std::min(minn, n0 > 0 ? n1 / n0 : 0.0);
        if (n1 > 0) scal(X[j].data(), 1.0 / n1, N);
    }
    return minn;
}

// Rayleigh-Ritz inside span(X): X <- X*Y, returns Ritz values and residual norms.
void rayleigh_ritz(const BandSym& T, std::vector<std::vector<double>>& X, std::vector<double>& theta,
                   std::vector<double>& res) {
    const int m = (int)X.size();
    const int64_t N = T.N;
    std::vector<std::vector<double>> TX(m, std::vector<double>(N));
    for (int j = 0; j < m; ++j) T.matvec(X[j].data(), TX[j].data());
    std::vector<double> H((size_t)m * m), Y;
    for (int i = 0; i < m; ++i)
        for (int j = i; j < m; ++j) {
            double h = dot(X[i].data(), TX[j].data(), N);
            H[(size_t)i * m + j] = h;
            H[(size_t)j * m + i] = h;
        }
    jacobi_eig(H, m, theta, Y);
    std::vector<std::vector<double>> Xn(m, std::vector<double>(N, 0.0)), TXn(m, std::vector<double>(N, 0.0));
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
            double y = Y[(size_t)i * m + j];
            if (y == 0.0) continue;
            axpy(Xn[j].data(), y, X[i].data(), N);
            axpy(TXn[j].data(), y, TX[i].data(), N);
        }
    res.resize(m);
    for (int j = 0; j < m; ++j) {
        double s = 0;
        for (int64_t i = 0; i < N; ++i) {
            double d = TXn[j][i] - theta[j] * Xn[j][i];
            s += d * d;
        }
        res[j] = std::sqrt(s);
    }
    X.swap(Xn);
}

// m eigenpairs of a tight cluster around `mu` by block inverse iteration + Rayleigh-Ritz, kept
// orthogonal to `against` (already accepted vectors of neighbouring intervals).
void extract_cluster(const BandSym& T, Work& wk, double mu, int m, const std::vector<const std::vector<double>*>& against,
                     std::vector<Pair>& out) {
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    std::vector<std::vector<double>> X(m);
    for (auto& x : X) wk.random_unit(x, N);
    mgs(X, 0, against, N);
    // shift slightly off the cluster centre so that the factorisation is not exactly singular
    wk.lu.factor(T, mu + 3e-15 * tn);
    ++wk.nfac;
    std::vector<double> theta, res;
    for (int it = 0; it < 8; ++it) {
        for (auto& x : X) {
            wk.lu.solve(x.data());
            double nn = nrm2(x.data(), N);
            if (nn > 0 && std::isfinite(nn)) scal(x.data(), 1.0 / nn, N);
            else wk.random_unit(x, N);
        }
        mgs(X, 0, against, N);
        rayleigh_ritz(T, X, theta, res);
        double worst = 0;
        for (double r : res) worst = std::max(worst, r);
        if (it >= 1 && worst <= 1e-13 * tn) break;
    }
    for (int j = 0; j < m; ++j) {
        Pair p;
        p.theta = theta[j];
        p.res = res[j];
        p.v.swap(X[j]);
        out.push_back(std::move(p));
    }
}

// Orthonormal eigenvector bases for (nearly) degenerate eigenvalues: vectors of eigenvalues closer than a few
// ctol are orthonormalised and rotated to Ritz vectors (a duplicate is regenerated inside the cluster), and
// independently converged vectors of eigenvalues closer than 1e-3 ||T|| get their eps/gap cross-components
// removed (like LAPACK dstein's ortol).  *dup is set when a vector of the loose pass vanishes (two inputs were
// the same eigenvector).
void finalize_pairs(const BandSym& T, std::vector<Pair>& out, int64_t& nfac, bool* dup, int threads) {
    const double tn = std::max(T.norm_inf, 1e-300);
    const double ctol = 2e-11 * tn;
    const auto t_fin0 = std::chrono::steady_clock::now();
    // final pass: neighbouring eigenvalues closer than a few ctol must have orthogonal vectors
    std::sort(out.begin(), out.end(), [](const Pair& a, const Pair& b) { return a.theta < b.theta; });
    // clusters are independent of each other: one task per cluster, spread over the host threads (the 100 lowest
    // eigenvalues of a 3-D Laplacian form ~25 degenerate clusters; one after the other they took 27 ms)
    struct Cluster { size_t g0, g1; std::vector<Pair> repl; bool dup = false; int nfac = 0; };
    std::vector<Cluster> cls;
    for (size_t g0 = 0; g0 < out.size();) {
        size_t g1 = g0 + 1;
        while (g1 < out.size() && out[g1].theta - out[g1 - 1].theta <= 4 * ctol) ++g1;
        if (g1 - g0 >= 2) { cls.emplace_back(); cls.back().g0 = g0; cls.back().g1 = g1; }
        g0 = g1;
    }
    auto do_cluster = [&](Cluster& c) {
        const bool erase_dups = dup != nullptr;
        std::vector<Pair> loc(out.begin() + c.g0, out.begin() + c.g1);
        const double mu_c = 0.5 * (loc.front().theta + loc.back().theta);
        std::vector<std::vector<double>> X;
        for (auto& p : loc) X.push_back(p.v);
        std::vector<const std::vector<double>*> none;
        // duplicate detection: a vector that (nearly) vanishes under MGS is regenerated
        for (size_t j = 0; j < X.size(); ++j) {
            std::vector<std::vector<double>> head(X.begin(), X.begin() + j + 1);
            double keep = mgs(head, j, none, T.N);
            if (keep < 0.5 && erase_dups) {  // seeded mode: two seeds collapsed onto one eigenvector - drop the second
                c.dup = true;
                X.erase(X.begin() + j);
                loc.erase(loc.begin() + j);
                --j;
                continue;
            } else if (keep < 0.5) {
                Work wk2;
                wk2.lu.factor(T, mu_c + 5e-15 * tn);
                ++c.nfac;
                std::vector<const std::vector<double>*> ag;
                for (size_t i = 0; i < j; ++i) ag.push_back(&X[i]);
                std::vector<std::vector<double>> one(1);
                wk2.random_unit(one[0], T.N);
                for (int it = 0; it < 4; ++it) {
                    mgs(one, 0, ag, T.N);
                    wk2.lu.solve(one[0].data());
                    scal(one[0].data(), 1.0 / nrm2(one[0].data(), T.N), T.N);
                }
                mgs(one, 0, ag, T.N);
                X[j] = one[0];
            } else {
                X[j] = head[j];
            }
        }
        std::vector<double> theta, res;
        rayleigh_ritz(T, X, theta, res);
        std::vector<size_t> ord(X.size());
        for (size_t j = 0; j < ord.size(); ++j) ord[j] = j;
        std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return theta[a] < theta[b]; });
        c.repl.resize(X.size());
        for (size_t j = 0; j < ord.size(); ++j) {
            c.repl[j].theta = theta[ord[j]];
            c.repl[j].res = res[ord[j]];
            c.repl[j].v.swap(X[ord[j]]);
        }
    };
    {
        std::atomic<size_t> next{0};
        std::exception_ptr err;
        std::mutex emu;
        auto run = [&]() {
            for (;;) {
                const size_t ci = next.fetch_add(1);
                if (ci >= cls.size()) break;
                try {
                    do_cluster(cls[ci]);
                } catch (...) {
                    std::lock_guard<std::mutex> lk(emu);
                    if (!err) err = std::current_exception();
                }
            }
        };
        const int ntc = (int)std::min<size_t>((size_t)std::max(1, threads), cls.size());
        std::vector<std::thread> th;
        for (int t = 1; t < ntc; ++t) th.emplace_back(run);
        run();
        for (auto& t : th) t.join();
        if (err) std::rethrow_exception(err);
    }
    if (!cls.empty()) {
        std::vector<Pair> merged;
        merged.reserve(out.size());
        size_t pos = 0;
        int nf_cl = 0;
        for (auto& c : cls) {
            for (; pos < c.g0; ++pos) merged.push_back(std::move(out[pos]));
            for (auto& p : c.repl) merged.push_back(std::move(p));
            pos = c.g1;
            if (c.dup && dup) *dup = true;
            nf_cl += c.nfac;
        }
        for (; pos < out.size(); ++pos) merged.push_back(std::move(out[pos]));
        out.swap(merged);
        nfac += nf_cl;
    }
    // loose pass (like LAPACK dstein's ortol): independently converged vectors of eigenvalues closer than
    // 1e-3 ||T|| carry a residual/gap component of each other (refined seeds are accepted at residuals up to
    // 1e-12 ||T||); remove it by Gram-Schmidt in eigenvalue order.
    // The wanted eigenvalues of the BASELINE Laplacians all lie within that distance of each other, so this is k^2/2
    // dot + axpy pairs over N-vectors - 60 ms of the accepting check of config 2 when done one after the other.  The
    // overlaps are tiny, so Gram-Schmidt is applied as threaded sweeps over all vectors at once: G_ij = v_i'v_j (i < j in
    // the window, all from the same snapshot), v_j <- v_j - sum_i G_ij v_i, normalise; the sweep differs from sequential
    // Gram-Schmidt at second order in max|G| and is repeated until max|G| <= 1e-14.  Anything unusual (an overlap above
    // 1e-3, which is what two copies of one eigenvector look like) takes the sequential path below.
    const double otol = 1e-3 * tn;
    bool swept = false;
    const auto t_loose = std::chrono::steady_clock::now();
    int sweeps = 0;
    if (threads > 1 && out.size() >= 8) {
        const size_t m = out.size();
        const int64_t N = T.N;
        std::vector<size_t> lo(m, 0);
        for (size_t j = 1; j < m; ++j) {
            size_t i = lo[j - 1];
            while (out[j].theta - out[i].theta > otol) ++i;
            lo[j] = i;
        }
        std::vector<std::vector<double>> G(m), Wn(m);
        const int nt = (int)std::min<size_t>((size_t)threads, m);
        auto par = [&](auto&& body) {
            std::atomic<size_t> next{1};
            auto run = [&]() {
                for (;;) {
                    const size_t j = next.fetch_add(1);
                    if (j >= m) break;
                    body(j);
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) th.emplace_back(run);
            run();
            for (auto& t : th) t.join();
        };
        swept = true;
        for (int pass = 0; pass < 4; ++pass) {
            ++sweeps;
            std::vector<double> mx(m, 0.0);
            par([&](size_t j) {
                G[j].assign(j - lo[j], 0.0);
                double mj = 0.0;
                for (size_t i = lo[j]; i < j; ++i) {
                    const double g = dot(out[i].v.data(), out[j].v.data(), N);
                    G[j][i - lo[j]] = g;
                    mj = std::max(mj, std::fabs(g));
                }
                mx[j] = mj;
            });
            const double maxg = *std::max_element(mx.begin(), mx.end());
            if (!(maxg <= 1e-3)) { swept = false; break; }      // (also NaN)
            if (maxg <= 1e-14) break;
            par([&](size_t j) {
                if (mx[j] == 0.0) { Wn[j].clear(); return; }
                Wn[j] = out[j].v;
                for (size_t i = lo[j]; i < j; ++i) axpy(Wn[j].data(), -G[j][i - lo[j]], out[i].v.data(), N);
                const double nn = nrm2(Wn[j].data(), N);
                if (nn > 0) scal(Wn[j].data(), 1.0 / nn, N);
            });
            for (size_t j = 1; j < m; ++j)
                if (!Wn[j].empty()) out[j].v.swap(Wn[j]);
        }
    }
    for (size_t j = 1; j < out.size() && !swept; ++j) {
        bool touched = false;
        for (size_t i = j; i-- > 0;) {
            if (out[j].theta - out[i].theta > otol) break;
            axpy(out[j].v.data(), -dot(out[i].v.data(), out[j].v.data(), T.N), out[i].v.data(), T.N);
            touched = true;
        }
        if (touched) {
            double nn = nrm2(out[j].v.data(), T.N);
            if (nn < 0.5 && dup) {  // the same eigenvector twice: drop this copy
                *dup = true;
                out.erase(out.begin() + j);
                --j;
                continue;
            }
            if (nn > 0) scal(out[j].v.data(), 1.0 / nn, T.N);
        }
    }
    if (std::getenv("RBL_FINALIZE_TIMING"))
        std::fprintf(stderr, "[rbl]   finalize_pairs: tight pass %.1f ms, loose pass %.1f ms (%d sweeps, swept %d, %zu pairs)\n",
                     std::chrono::duration<double>(t_loose - t_fin0).count() * 1e3,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loose).count() * 1e3, sweeps, (int)swept, out.size());
}

// Sturm counts #{lambda < x} at several shifts at once, one factorisation per shift spread over the threads.
void parallel_below(const BandSym& T, const std::vector<double>& xs, int threads, std::vector<int64_t>& below, int64_t& nfac) {
    below.assign(xs.size(), 0);
    if (xs.empty()) return;
    std::atomic<size_t> next{0};
    std::atomic<bool> cancelled{false};
    auto run = [&]() {
        BandLU lu;
        try {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= xs.size()) break;
                lu.factor(T, xs[i]);
                below[i] = lu.nneg;
            }
        } catch (const Cancelled&) {
            cancelled = true;
            next = xs.size();
        }
    };
    const int nt = (int)std::min<size_t>((size_t)std::max(1, threads), xs.size());
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(run);
    run();
    for (auto& t : th) t.join();
    if (cancelled) throw Cancelled{};
    nfac += (int64_t)xs.size();
}

struct Interval {
    double lo, hi;
    int64_t clo, chi;  // eigenvalues below lo / below hi
    double tried = 0.0;  // width at which block_extract last failed on an ancestor (0: never tried) ...
    int64_t tried_m = 0; // ... and the number of eigenvalues it held
};

// The q eigenpairs of a NARROW interval (lo,hi) that are not among `against` (already known eigenvectors of the interval;
// the interval holds q + against.size() eigenvalues by its Sturm counts), all at once: block inverse iteration deflated
// against the known vectors + Rayleigh-Ritz.
// Bisection needs log2(width / 2e-11||T||) ~ 25-30 factorisations to squeeze an interval down to a degenerate cluster, and
// the wanted Ritz values of the BASELINE Laplacians are ~25 such clusters (multiplicities 3 and 6): 600-1000 factorisations
// per k = 100 eigensolve, and ~20 for every single copy of a multiple eigenvalue that is missing from a set of seeds.  When
// the eigenvalues of the interval are bunched together and the rest of the spectrum is far away compared with the width,
// q vectors converge in a few solves each with two factorisations (the second at the Ritz values' centre: block
// Rayleigh-quotient iteration).  Success = q orthonormal vectors, orthogonal to the known ones, with residuals at rounding
// level and Ritz values strictly inside the interval; anything else - slow decay because a neighbour sits just outside
// or the interval holds several groups, a Ritz value on the edge - returns false and the caller goes on bisecting.
bool block_extract(const BandSym& T, Work& wk, const Interval& iv, int q, const std::vector<const std::vector<double>*>& against,
                   std::vector<Pair>& got) {
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    const double edge = std::max(1e-11 * tn, 1e-3 * (iv.hi - iv.lo));
    std::vector<std::vector<double>> X(q);
    for (auto& x : X) wk.random_unit(x, N);
    mgs(X, 0, against, N);
    std::vector<double> theta, res;
    double prev = 1e300;
    for (int it = 0; it < 7; ++it) {
        if (it == 0) {
            wk.lu.factor(T, 0.5 * (iv.lo + iv.hi));
            ++wk.nfac;
        } else if (it == 2) {
            double c = 0;
            for (double t : theta) c += t;
            wk.lu.factor(T, c / q + 3e-15 * tn);
            ++wk.nfac;
        }
        for (auto& x : X) {
            wk.lu.solve(x.data());
            const double nn = nrm2(x.data(), N);
            if (!(nn > 0) || !std::isfinite(nn)) return false;
            scal(x.data(), 1.0 / nn, N);
        }
        if (mgs(X, 0, against, N) < 1e-8) return false;  // the block lost rank: fewer directions than the count says
        rayleigh_ritz(T, X, theta, res);
        double worst = 0;
        bool inside = true;
        for (int j = 0; j < q; ++j) {
            worst = std::max(worst, res[j]);
            if (!(theta[j] > iv.lo + edge && theta[j] < iv.hi - edge)) inside = false;
        }
        static const bool block_debug = std::getenv("RBL_BLOCK_DEBUG") != nullptr;
        if (block_debug) std::fprintf(stderr, "[blk] q=%d known=%zu width=%.3e it=%d worst=%.3e inside=%d\n", q, against.size(), iv.hi - iv.lo, it, worst / tn, (int)inside);
        // converged: at rounding level, or stagnating just above it (the attainable residual grows with N)
        if (worst <= 2e-13 * tn || (it >= 3 && worst <= 5e-12 * tn && worst > 0.25 * prev)) {
            if (!inside) return false;
            for (int j = 0; j < q; ++j) {
                Pair p;
                p.theta = theta[j];
                p.res = res[j];
                p.v.swap(X[j]);
                got.push_back(std::move(p));
            }
            return true;
        }
        if (it >= 1 && !inside) return false;              // a Ritz value on the edge / an outside eigenvalue pulled in
        if (it >= 1 && worst > 0.03 * prev) return false;  // several groups, or a neighbour just outside: bisect further
        prev = worst;
    }
    return false;
}

// All eigenpairs with eigenvalue in (lo,hi): recursive bisection on Sturm counts until an interval
// holds one eigenvalue (finished by inverse iteration + RQI) or a tight cluster.
// `known` (optional): eigenpairs of T that are already available (refined seeds).  An interval whose Sturm count equals
// the number of known eigenvalues strictly inside it is complete - its pairs are taken from `known` and the
// bisection below it is skipped, so the factorisations go only where eigenvalues are actually missing.
void slice(const BandSym& T, const std::vector<Interval>& roots, int threads, std::vector<Pair>& out, int64_t& nfac,
           const std::vector<Pair>* known = nullptr, int64_t* reused = nullptr) {
    const double tn = std::max(T.norm_inf, 1e-300);
    const double ctol = 2e-11 * tn;
    std::vector<std::pair<double, size_t>> kn;  // (theta, index into *known), ascending
    if (known) {
        for (size_t i = 0; i < known->size(); ++i) kn.push_back({(*known)[i].theta, i});
        std::sort(kn.begin(), kn.end());
    }
    const double kmargin = 1e-9 * tn;
    std::atomic<int64_t> n_reused{0};
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Interval> queue;
    std::vector<Interval> clusters;
    int active = 0;
    bool cancelled = false;  // under mu: a worker saw T.cancel; everybody drains
    std::atomic<int64_t> fac{0};
    static const bool slice_timing = std::getenv("RBL_SLICE_TIMING") != nullptr;
    const auto ts0 = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count() * 1e3; };
    // The bisection tree starts with one interval per root, so its first ~5 levels run on one or two threads, and with
    // `known` pairs nearly all of the range is complete anyway.  One round of Sturm counts at several points at once - in
    // the gaps between groups of known eigenvalues, or evenly spaced - hands every thread its own interval from the start,
    // already narrow enough for block_extract where something is missing.
    for (auto& r : roots) {
        if (!(r.chi > r.clo && r.hi > r.lo)) continue;
        std::vector<double> pts;
        const size_t want = (size_t)std::min(64, 4 * std::max(1, threads));
        if (threads > 1 && T.N >= 400 && r.chi - r.clo >= 8) {
            if (!kn.empty()) {
                std::vector<std::pair<double, double>> gaps;  // (gap width, midpoint) between consecutive known values in the root
                for (size_t i = 1; i < kn.size(); ++i) {
                    const double a = kn[i - 1].first, b2 = kn[i].first;
                    if (a > r.lo && b2 < r.hi && b2 - a > 64 * kmargin) gaps.push_back({b2 - a, 0.5 * (a + b2)});
                }
                std::sort(gaps.begin(), gaps.end(), [](const std::pair<double, double>& x, const std::pair<double, double>& y) { return x.first > y.first; });
                if (gaps.size() > want) gaps.resize(want);
                for (auto& gp : gaps) pts.push_back(gp.second);
            } else {
                for (size_t i = 1; i < want; ++i) pts.push_back(r.lo + (r.hi - r.lo) * (double)i / (double)want);
            }
        }
        if (pts.empty()) { queue.push_back(r); continue; }
        std::sort(pts.begin(), pts.end());
        std::vector<int64_t> below;
        int64_t nf0 = 0;
        parallel_below(T, pts, threads, below, nf0);
        fac += nf0;
        double lo = r.lo;
        int64_t clo = r.clo;
        for (size_t i = 0; i <= pts.size(); ++i) {
            const double hi = i < pts.size() ? pts[i] : r.hi;
            const int64_t chi = i < pts.size() ? std::min(std::max(below[i], clo), r.chi) : r.chi;
            if (chi > clo && hi > lo) queue.push_back(Interval{lo, hi, clo, chi});
            lo = hi;
            clo = chi;
        }
    }

    const double ms_presplit = ms_since();
    const size_t n_initial = queue.size();
    auto worker = [&](int seed) {
        Work wk;
        wk.rng.seed(987654321ull + 7919ull * seed);
        std::vector<Pair> local;
        for (;;) {
            Interval iv;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || active == 0 || cancelled; });
                if (queue.empty() || cancelled) break;
                iv = queue.front();
                queue.pop_front();
                ++active;
            }
            std::vector<Interval> push;
            const int64_t m = iv.chi - iv.clo;
            const double mid = 0.5 * (iv.lo + iv.hi);
            bool complete = false;
            if (!kn.empty() && iv.hi - iv.lo > 2 * kmargin) {
                auto lo_it = std::upper_bound(kn.begin(), kn.end(), std::make_pair(iv.lo + kmargin, (size_t)-1));
                auto hi_it = std::lower_bound(kn.begin(), kn.end(), std::make_pair(iv.hi - kmargin, (size_t)0));
                // nothing known may sit in the margins (it could belong to either side of the count)
                auto lo_edge = std::lower_bound(kn.begin(), kn.end(), std::make_pair(iv.lo - kmargin, (size_t)0));
                auto hi_edge = std::upper_bound(kn.begin(), kn.end(), std::make_pair(iv.hi + kmargin, (size_t)-1));
                if (hi_it >= lo_it && (int64_t)(hi_it - lo_it) == m && lo_edge == lo_it && hi_edge == hi_it) {
                    for (auto it = lo_it; it != hi_it; ++it) local.push_back((*known)[it->second]);
                    n_reused += m;
                    complete = true;
                }
            }
            // narrow interval with few eigenvalues that are not known yet: try to pull them out in one go
            int64_t q_missing = m;
            std::vector<const std::vector<double>*> in_known;
            std::vector<size_t> in_idx;
            bool edge_known = false;
            if (!complete && !kn.empty()) {
                auto lo_it = std::upper_bound(kn.begin(), kn.end(), std::make_pair(iv.lo + kmargin, (size_t)-1));
                auto hi_it = std::lower_bound(kn.begin(), kn.end(), std::make_pair(iv.hi - kmargin, (size_t)0));
                auto lo_edge = std::lower_bound(kn.begin(), kn.end(), std::make_pair(iv.lo - kmargin, (size_t)0));
                auto hi_edge = std::upper_bound(kn.begin(), kn.end(), std::make_pair(iv.hi + kmargin, (size_t)-1));
                edge_known = !(lo_edge == lo_it && hi_edge == hi_it) || hi_it < lo_it;
                if (!edge_known) {
                    for (auto it = lo_it; it != hi_it; ++it) {
                        in_known.push_back(&(*known)[it->second].v);
                        in_idx.push_back(it->second);
                    }
                    q_missing = m - (int64_t)in_known.size();
                }
            }
            const bool try_block = !complete && !edge_known && q_missing >= 1 && q_missing <= 12 && m >= 2 && (iv.hi - iv.lo) > ctol &&
                                   (iv.hi - iv.lo) <= 1e-4 * tn &&
                                   (iv.tried == 0.0 || (iv.hi - iv.lo) <= iv.tried / 16 || m < iv.tried_m);
            bool extracted = false;
            if (try_block) {
                std::vector<Pair> got;
                if (block_extract(T, wk, iv, (int)q_missing, in_known, got)) {
                    extracted = true;
                    for (auto& p : got) local.push_back(std::move(p));
                    for (size_t ki : in_idx) local.push_back((*known)[ki]);
                    n_reused += (int64_t)in_known.size();
                } else {
                    iv.tried = iv.hi - iv.lo;  // the halves try again only when narrower or split
                    iv.tried_m = m;
                }
            }
            if (complete) {
                // nothing to do below this interval
            } else if (m >= 2 && (iv.hi - iv.lo) <= ctol) {
                std::lock_guard<std::mutex> lk(mu);
                clusters.push_back(iv);
            } else if (extracted) {
                // all missing pairs of the interval came out of one deflated block inverse iteration
            } else {
                wk.lu.factor(T, mid);
                ++wk.nfac;
                const int64_t cmid = std::min(std::max(wk.lu.nneg, iv.clo), iv.chi);
                bool done = false;
                if (m == 1) {
                    // one eigenvalue in (lo,hi): two inverse-iteration steps with this factorisation, then RQI
                    double lo = iv.lo, hi = iv.hi;
                    if (cmid == iv.clo) lo = mid; else hi = mid;
                    std::vector<double> x;
                    wk.random_unit(x, T.N);
                    for (int s = 0; s < 2; ++s) {
                        wk.lu.solve(x.data());
                        double nn = nrm2(x.data(), T.N);
                        if (!(nn > 0) || !std::isfinite(nn)) { wk.random_unit(x, T.N); continue; }
                        scal(x.data(), 1.0 / nn, T.N);
                    }
                    double th, rs;
                    rayleigh(T, x, wk.t, th, rs);
                    if (th > lo && th < hi) {
                        if (rqi(T, wk, th, x, rs, lo, hi, 8, 2e-15)) {
                            Pair p;
                            p.theta = th;
                            p.res = rs;
                            p.v.swap(x);
                            local.push_back(std::move(p));
                            done = true;
                        }
                    }
                    if (!done) {
                        if ((hi - lo) <= 1e-15 * tn) {  // cannot separate further: take the inverse-iteration vector
                            std::lock_guard<std::mutex> lk(mu);
                            clusters.push_back(Interval{lo, hi, iv.clo, iv.chi});
                        } else {
                            push.push_back(Interval{lo, hi, iv.clo, iv.chi, iv.tried, iv.tried_m});
                        }
                    }
                } else {
                    if (cmid > iv.clo) push.push_back(Interval{iv.lo, mid, iv.clo, cmid, iv.tried, iv.tried_m});
                    if (iv.chi > cmid) push.push_back(Interval{mid, iv.hi, cmid, iv.chi, iv.tried, iv.tried_m});
                }
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                for (auto& p : push) queue.push_back(p);
                --active;
            }
            cv.notify_all();
        }
        fac += wk.nfac;
        std::lock_guard<std::mutex> lk(mu);
        for (auto& p : local) out.push_back(std::move(p));
    };

    auto guarded_worker = [&](int seed) {
        try {
            worker(seed);
        } catch (const Cancelled&) {
            std::lock_guard<std::mutex> lk(mu);
            cancelled = true;
            cv.notify_all();
        }
    };
    int nt = std::max(1, threads);
    if (T.N < 400) nt = 1;
    if (nt == 1) {
        guarded_worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(guarded_worker, t);
        for (auto& t : th) t.join();
    }
    if (cancelled) throw Cancelled{};
    const double ms_tree = ms_since();

    // tight clusters (degenerate Ritz values), in parallel: each orthogonal to the already accepted single
    // vectors within a few ctol; clusters that a bisection point split in two are repaired by the final pass
    std::sort(clusters.begin(), clusters.end(), [](const Interval& a, const Interval& b) { return a.lo < b.lo; });
    {
        const size_t nsingle = out.size();
        std::vector<std::vector<Pair>> found(clusters.size());
        std::atomic<size_t> next{0};
        auto cworker = [&](int seed) {
            Work wk;
            wk.rng.seed(424242ull + 31ull * seed);
            for (;;) {
                const size_t ci = next.fetch_add(1);
                if (ci >= clusters.size()) break;
                const Interval& c = clusters[ci];
                std::vector<const std::vector<double>*> against;
                for (size_t j = 0; j < nsingle; ++j)
                    if (out[j].theta > c.lo - 8 * ctol && out[j].theta < c.hi + 8 * ctol) against.push_back(&out[j].v);
                wk.rng.seed(1000003ull * (ci + 1));  // deterministic per cluster, whatever thread runs it
                extract_cluster(T, wk, 0.5 * (c.lo + c.hi), (int)(c.chi - c.clo), against, found[ci]);
            }
            fac += wk.nfac;
        };
        std::atomic<bool> ccancel{false};
        auto guarded_cworker = [&](int seed) {
            try {
                cworker(seed);
            } catch (const Cancelled&) {
                ccancel = true;
                next = clusters.size();
            }
        };
        const int ct = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, clusters.size()));
        if (ct <= 1) {
            guarded_cworker(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < ct; ++t) th.emplace_back(guarded_cworker, t);
            for (auto& t : th) t.join();
        }
        if (ccancel) throw Cancelled{};
        for (auto& f : found)
            for (auto& p : f) out.push_back(std::move(p));
    }

    const double ms_clusters = ms_since();
    {
        int64_t nf2 = 0;
        finalize_pairs(T, out, nf2, nullptr, threads);
        fac += nf2;
    }
    nfac += fac.load();
    if (reused) *reused = n_reused.load();
    if (slice_timing)
        std::fprintf(stderr, "[rbl]   slice: pre-split %.1f ms (%zu intervals), tree done at %.1f, %zu tight clusters done at %.1f, finalize done at %.1f ms; %lld factorisations\n",
                     ms_presplit, n_initial, ms_tree, clusters.size(), ms_clusters, ms_since(), (long long)fac.load());
}

// residual bound ||B_i s[N-b..N)||, B_i row-major upper triangular b x b
double resid_bound(const double* bi, int b, const std::vector<double>& s) {
    if (!bi) return 0.0;
    const int64_t N = (int64_t)s.size();
    const double* last = s.data() + (N - b);
    double acc = 0;
    for (int r = 0; r < b; ++r) {
        double y = 0;
        for (int c = 0; c < b; ++c) y += bi[(size_t)r * b + c] * last[c];
        acc += y * y;
    }
    return std::sqrt(acc);
}

// One witness followed on its own (thread-safe: no state shared with the checker): the zero-padded old Ritz vector is
// refined by inverse iteration with the shift kept on the inner side of the Ritz value, so that the Sturm count of the
// last factorisation bounds from above the number of eigenvalues of larger magnitude (the pair's rank).
struct Followed {
    bool ok = false;
    double theta = 0, res = 0, rho = 0;
    int64_t larger = 0;
    std::vector<double> x;
    int nfac = 0;
};

Followed follow_witness(const BandSym& T, const std::vector<double>& w, const double* bi, int b) {
    Followed f;
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    if ((int64_t)w.size() > N) return f;
    f.x.assign(N, 0.0);
    std::copy(w.begin(), w.end(), f.x.begin());
    const double nn = nrm2(f.x.data(), N);
    if (!(nn > 0)) return f;
    scal(f.x.data(), 1.0 / nn, N);
    Work wk;
    double th, rs;
    rayleigh(T, f.x, wk.t, th, rs);
    double sigma = 0;
    int64_t cnt = 0;
    bool ok = false;
    for (int round = 0; round < 7 && !ok; ++round) {
        const double sg = th < 0 ? -1.0 : 1.0;
        const double off = (rs <= 1e-6 * tn) ? 2.0 * rs : 0.0;
        sigma = th - sg * off;
        wk.lu.factor(T, sigma);
        ++f.nfac;
        cnt = wk.lu.nneg;
        for (int it = 0; it < 4; ++it) {
            wk.y = f.x;
            wk.lu.solve(wk.y.data());
            const double n2 = nrm2(wk.y.data(), N);
            if (!(n2 > 0) || !std::isfinite(n2)) break;
            scal(wk.y.data(), 1.0 / n2, N);
            f.x.swap(wk.y);
            const double prev = rs;
            rayleigh(T, f.x, wk.t, th, rs);
            if (rs <= 2e-13 * tn) { ok = true; break; }
            if (rs > 0.25 * prev) break;
        }
        if (!ok && rs <= 1e-11 * tn && round >= 1) ok = true;
    }
    if (!ok) return f;
    const double margin = 1e-13 * tn;
    const double delta = std::max(1e-10 * std::fabs(th), 1e-12 * tn);
    if (th >= 0) {
        if (!(sigma < th - margin)) {
            wk.lu.factor(T, th - delta);
            ++f.nfac;
            cnt = wk.lu.nneg;
        }
        int64_t neg = 0;
        if (!(-std::fabs(th) < T.gersh_lo)) {
            wk.lu.factor(T, -std::fabs(th));
            ++f.nfac;
            neg = wk.lu.nneg;
        }
        f.larger = (N - cnt - 1) + neg;
    } else {
        if (!(sigma > th + margin)) {
            wk.lu.factor(T, th + delta);
            ++f.nfac;
            cnt = wk.lu.nneg;
        }
        int64_t pos = 0;
        if (std::fabs(th) <= T.gersh_hi) {
            wk.lu.factor(T, std::fabs(th));
            ++f.nfac;
            pos = N - wk.lu.nneg;
        }
        f.larger = (cnt - 1) + pos;
    }
    f.theta = th;
    f.res = rs;
    f.rho = resid_bound(bi, b, f.x);
    f.ok = true;
    return f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ BandTopK
void BandTopK::set_seeds(const std::vector<double>& d, const std::vector<double>& svec, int64_t Ns, int64_t k,
                         const std::vector<double>* resid) {
    if ((int64_t)d.size() < k || (int64_t)svec.size() < Ns * k) return;
    // (pairs of an older T than the ones in hand - a tracker pass that started before this checker's own last full check -
    // would only make the next full check slower)
    if ((int64_t)seeds_.size() == k && !seeds_.empty() && (int64_t)seeds_[0].v.size() >= Ns) return;
    seeds_.clear();
    seeds_.resize(k);
    for (int64_t j = 0; j < k; ++j) {
        seeds_[j].theta = d[j];
        seeds_[j].v.assign(svec.begin() + (size_t)j * Ns, svec.begin() + (size_t)(j + 1) * Ns);
    }
    // The pairs that were furthest from convergence when the seeds were computed are the ones that will converge last
    // (the bounds of the unconverged pairs decay at similar rates): they join the witness list, so that "all witnesses have
    // converged" - the trigger of the full k-pair check - almost always means that everything has.  The witness this
    // checker is following stays first (its factorisation is extended from check to check).
    if (resid && (int64_t)resid->size() >= k) {
        std::vector<int64_t> ord(k);
        for (int64_t j = 0; j < k; ++j) ord[j] = j;
        std::sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b2) { return (*resid)[a] > (*resid)[b2]; });
        if (wit_.size() > 1) { wit_.resize(1); wit_theta_.resize(1); }
        for (int64_t t = 0; t < k && (int)wit_.size() < 1 + kExtraWitnesses; ++t) {
            const int64_t j = ord[t];
            if (!((*resid)[j] > 0.0)) break;
            if (!wit_.empty()) {  // the same pair as the witness in hand?
                const int64_t n = std::min<int64_t>((int64_t)wit_[0].size(), Ns);
                const double ov = dot(wit_[0].data(), seeds_[j].v.data(), n);
                const double na = nrm2(wit_[0].data(), (int64_t)wit_[0].size()), nb = nrm2(seeds_[j].v.data(), Ns);
                if (na > 0 && nb > 0 && std::fabs(ov) >= 0.9 * na * nb) continue;
            }
            wit_.push_back(seeds_[j].v);
            wit_theta_.push_back(seeds_[j].theta);
        }
    }
}

// All k seed pairs refined for the current T by inverse iteration (zero-padded start vectors), in parallel.
static double seed_min_frac() {
    static const double f = [] {
        const char* e = std::getenv("RBL_SEED_MIN_FRAC");
        const double v = e ? std::atof(e) : 0.0;
        return (v > 0.0 && v <= 1.0) ? v : 0.70;
    }();
    return f;
}

// Seeds at least this fresh (size of the T they come from / size of T now) send a check whose witnesses have converged
// straight to the seeded full check, skipping the serial bracket search of stage 2.
static double seed_fresh_frac(int threads) {
    static const double f = [] {
        const char* e = std::getenv("RBL_SEED_FRESH_FRAC");
        const double v = e ? std::atof(e) : 0.0;
        return (v > 0.0 && v <= 1.0) ? v : 0.0;
    }();
    if (f > 0.0) return f;
    return threads >= 4 ? 0.85 : 0.95;
}

bool BandTopK::refine_seeds(const BandSym& T, int64_t k, std::vector<Pair>& pairs, int64_t& nfac) {
    const auto t_ref0 = std::chrono::steady_clock::now();
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    pairs.assign(k, Pair());
    // Work units: seeds whose Ritz values (of the T they come from) coincide to 1e-7 ||T|| - the copies of a multiple
    // eigenvalue, 3 or 6 at a time for the BASELINE Laplacians - share ONE factorisation just inside their common value;
    // each copy then costs a solve or two instead of a factorisation of its own (the factorisation is 3-4 solves' worth).
    // Units are capped so that there are enough of them for all threads.
    std::vector<std::vector<int64_t>> units;
    {
        std::vector<int64_t> ord(k);
        for (int64_t j = 0; j < k; ++j) ord[j] = j;
        std::sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return seeds_[x].theta < seeds_[y].theta; });
        const size_t cap = (size_t)std::max<int64_t>(1, (k + std::max(1, threads) - 1) / std::max(1, threads));
        for (int64_t t = 0; t < k; ++t) {
            const int64_t j = ord[t];
            const bool join = !units.empty() && units.back().size() < std::max<size_t>(cap, 2) &&
                              std::fabs(seeds_[j].theta - seeds_[units.back().back()].theta) <= 1e-7 * tn &&
                              (seeds_[j].theta < 0) == (seeds_[units.back().back()].theta < 0);
            if (join) units.back().push_back(j);
            else units.push_back({j});
        }
    }
    std::atomic<size_t> next{0};
    std::atomic<int64_t> fac{0};
    auto worker = [&]() {
        Work wk;
        struct Member { int64_t j; std::vector<double> x; double th = 0, rs = 0; bool ok = false, shared = false; };
        for (;;) {
            const size_t u = next.fetch_add(1);
            if (u >= units.size()) break;
            std::vector<Member> mem;
            for (int64_t j : units[u]) {
                Member mb;
                mb.j = j;
                mb.x.assign(N, 0.0);
                std::copy(seeds_[j].v.begin(), seeds_[j].v.end(), mb.x.begin());
                const double nn = nrm2(mb.x.data(), N);
                if (!(nn > 0)) continue;  // dropped (pairs[j].v stays empty)
                scal(mb.x.data(), 1.0 / nn, N);
                rayleigh(T, mb.x, wk.t, mb.th, mb.rs);
                mb.ok = mb.rs <= 2e-13 * tn;
                mem.push_back(std::move(mb));
            }
            // inverse iteration of x with the factorisation in wk.lu: up to 4 solves while the residual keeps dropping
            auto iterate = [&](Member& mb) {
                for (int it = 0; it < 4; ++it) {
                    wk.y = mb.x;
                    wk.lu.solve(wk.y.data());
                    const double n2 = nrm2(wk.y.data(), N);
                    if (!(n2 > 0) || !std::isfinite(n2)) break;
                    scal(wk.y.data(), 1.0 / n2, N);
                    mb.x.swap(wk.y);
                    const double prev = mb.rs;
                    rayleigh(T, mb.x, wk.t, mb.th, mb.rs);
                    if (mb.rs <= 2e-13 * tn) { mb.ok = true; break; }
                    if (mb.rs > 0.25 * prev) break;
                }
            };
            // stale seeds (residual up to 1e-4 ||T||) of one multiple value: a first shared round at their mean Rayleigh quotient
            // brings them close enough for the shared round below; without it every copy pays two factorisations of its own
            {
                int nfar = 0;
                double sum = 0, spread_lo = 1e300, spread_hi = -1e300, rmax = 0;
                for (auto& mb : mem)
                    if (!mb.ok && mb.rs > 1e-8 * tn && mb.rs <= 1e-4 * tn) {
                        ++nfar;
                        sum += mb.th;
                        spread_lo = std::min(spread_lo, mb.th);
                        spread_hi = std::max(spread_hi, mb.th);
                        rmax = std::max(rmax, mb.rs);
                    }
                if (nfar >= 2 && spread_hi - spread_lo <= 10.0 * rmax) {
                    wk.lu.factor(T, sum / nfar);
                    ++wk.nfac;
                    for (auto& mb : mem)
                        if (!mb.ok && mb.rs > 1e-8 * tn && mb.rs <= 1e-4 * tn) iterate(mb);
                }
            }
            // shared factorisation: only for copies that are already close (residual <= 1e-8 ||T||), so that the common
            // shift - the innermost of their own "Ritz value minus twice the residual" - stays within ~1e-7 ||T|| of all of them
            int nclose = 0;
            double shift = 0;
            for (auto& mb : mem)
                if (!mb.ok && mb.rs <= 1e-8 * tn) {
                    const double sg = mb.th < 0 ? -1.0 : 1.0;
                    const double mine = mb.th - sg * 2.0 * mb.rs;
                    if (nclose == 0 || std::fabs(mine) < std::fabs(shift)) shift = mine;
                    ++nclose;
                }
            if (nclose >= 2) {
                wk.lu.factor(T, shift);
                ++wk.nfac;
                for (auto& mb : mem)
                    if (!mb.ok && mb.rs <= 1e-8 * tn) {
                        iterate(mb);
                        mb.shared = true;
                        if (!mb.ok && mb.rs <= 1e-11 * tn) mb.ok = true;
                    }
            }
            for (auto& mb : mem) {
                for (int round = mb.shared ? 1 : 0; round < 6 && !mb.ok; ++round) {
                    const double sg = mb.th < 0 ? -1.0 : 1.0;
                    const double off = (mb.rs <= 1e-6 * tn) ? 2.0 * mb.rs : 0.0;
                    wk.lu.factor(T, mb.th - sg * off);
                    ++wk.nfac;
                    iterate(mb);
                    if (!mb.ok && mb.rs <= 1e-11 * tn && round >= 1) mb.ok = true;
                }
                if (!mb.ok) continue;        // dropped: the caller's count-based repair looks for what is missing
                pairs[mb.j].theta = mb.th;
                pairs[mb.j].res = mb.rs;
                pairs[mb.j].v.swap(mb.x);
            }
        }
        fac += wk.nfac;
    };
    std::atomic<bool> rcancel{false};
    auto guarded_worker = [&]() {
        try {
            worker();
        } catch (const Cancelled&) {
            rcancel = true;
        }
    };
    const int nt = (int)std::min<int64_t>(std::max(1, threads), (int64_t)units.size());
    if (nt <= 1) {
        guarded_worker();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(guarded_worker);
        for (auto& t : th) t.join();
    }
    if (rcancel) throw Cancelled{};
    nfac += fac.load();
    const auto t_ref = std::chrono::steady_clock::now();
    pairs.erase(std::remove_if(pairs.begin(), pairs.end(), [](const Pair& p) { return p.v.empty(); }), pairs.end());
    if ((int64_t)pairs.size() * 4 < k * 3) return false;  // too little survived: slicing from scratch is cheaper
    bool dup = false;
    int64_t nf2 = 0;
    finalize_pairs(T, pairs, nf2, &dup, threads);  // seeded mode: duplicates are erased, not regenerated
    nfac += nf2;
    if (verbose > 1)
        std::fprintf(stderr, "[rbl]   refine_seeds: inverse iterations %.1f ms (%lld factorisations, %d threads), finalize %.1f ms\n",
                     std::chrono::duration<double>(t_ref - t_ref0).count() * 1e3, (long long)fac.load(), nt,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ref).count() * 1e3);
    return (int64_t)pairs.size() * 4 >= k * 3;
}

// Decision procedure (identical outcome to dsbev + sort_eig_abs + check_convergence, common.jl:36-65):
//   "not converged" needs ONE Ritz pair among the k largest |lambda| whose bound ||B_i s_last|| exceeds tol;
//   "converged" needs all k of them.
//   stage 1  follow the witness pairs of the previous check by Rayleigh-quotient iteration (a few factorisations)
//   stage 2  isolate the pair(s) at rank k (the slowest to converge) from a bracket that starts at the previous
//            k-th value (the k-th largest |lambda| never decreases when T grows: Cauchy interlacing)
//   stage 3  all k pairs by spectrum slicing (only when stages 1-2 found nothing unconverged)
TopKResult BandTopK::check(const BandSym& T, const double* bi, int b, int64_t k, double tol, bool force_full) {
    TopKResult R;
    R.N = T.N;
    const int64_t N = T.N;
    const double tn = std::max(T.norm_inf, 1e-300);
    Work wk;
    if (k > N) k = N;
    const double g = tn * (1.0 + 1e-12) + 1e-300;

    // #{ |lambda| > x } for x >= 0, remembering the two one-sided Sturm counts.  A side that lies outside the
    // Gershgorin interval of T is empty and costs no factorisation.
    struct Cnt { int64_t below_pos, below_neg, above; };
    double neg_free_from = (T.gersh_lo >= 0.0) ? 0.0 : -1.0;  // for x >= this the negative side is empty
    auto neg_side = [&](double x) -> int64_t {  // #{ lambda < -x }
        if (-x < T.gersh_lo) return 0;
        if (neg_free_from >= 0.0 && x >= neg_free_from) return 0;
        wk.lu.factor(T, -x);
        ++wk.nfac;
        if (wk.lu.nneg == 0) neg_free_from = x;
        return wk.lu.nneg;
    };
    auto count_abs_above = [&](double x) -> Cnt {
        Cnt c;
        if (x > T.gersh_hi) {
            c.below_pos = N;
        } else {
            wk.lu.factor(T, x);
            ++wk.nfac;
            c.below_pos = wk.lu.nneg;
        }
        c.below_neg = neg_side(x);
        c.above = (N - c.below_pos) + c.below_neg;
        return c;
    };
    const auto t_start = std::chrono::steady_clock::now();
    int stage_now = 0;
    auto finish = [&](bool conv) {
        stage_hits[stage_now]++;
        stage_sec[stage_now] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        R.converged = conv;
        R.factorizations = wk.nfac;
        total_factorizations += wk.nfac;
        return R;
    };

    // Refine an approximate eigenvector x (unit norm) by inverse iteration: the shift is kept on the inner
    // side of the Ritz value, so the Sturm count of the last factorisation bounds the rank of the pair from
    // above without any further factorisation.  Returns false if the iteration does not settle.
    struct Refined { double theta, res; int64_t larger; };
    // `lu` is the factorisation object to use; with try_resume it may still hold the factorisation of the
    // previous (smaller) T at a shift next to this Ritz value, which is then extended instead of recomputed
    auto refine = [&](BandLU& lu, bool try_resume, std::vector<double>& x, bool have_factor, double sigma0, int64_t cnt0,
                      Refined& out) -> bool {
        double th, rs;
        rayleigh(T, x, wk.t, th, rs);
        double sigma = sigma0;
        int64_t cnt_sigma = cnt0;
        bool ok = false;
        if (try_resume && !have_factor && lu.ck_row >= 0 && std::fabs(lu.shift_ - th) <= 1e-6 * tn &&
            ((th >= 0) ? (lu.shift_ < th) : (lu.shift_ > th)) && lu.resume(T)) {
            have_factor = true;
            sigma = lu.shift_;
            cnt_sigma = lu.nneg;
            ++resumed_factorizations;
        }
        for (int round = 0; round < 7 && !ok; ++round) {
            if (round > 0 || !have_factor) {
                // Rayleigh-quotient shifts while far from convergence (cubic); once close, step to the inner
                // side of the Ritz value by twice the residual (the eigenvalue is within one residual of it)
                const double sg = th < 0 ? -1.0 : 1.0;
                const double off = (rs <= 1e-6 * tn) ? 2.0 * rs : 0.0;
                sigma = th - sg * off;
                lu.factor(T, sigma);
                ++wk.nfac;
                cnt_sigma = lu.nneg;
            }
            for (int it = 0; it < 4; ++it) {
                wk.y = x;
                lu.solve(wk.y.data());
                const double n2 = nrm2(wk.y.data(), N);
                if (!(n2 > 0) || !std::isfinite(n2)) break;
                scal(wk.y.data(), 1.0 / n2, N);
                x.swap(wk.y);
                const double prev = rs;
                rayleigh(T, x, wk.t, th, rs);
                if (rs <= 2e-13 * tn) { ok = true; break; }
                if (rs > 0.25 * prev) break;  // slow: a closer shift is needed
            }
            if (!ok && rs <= 1e-11 * tn && round >= 1) ok = true;
        }
        if (!ok) return false;
        out.theta = th;
        out.res = rs;
        const double margin = 1e-13 * tn;
        if (th >= 0 && sigma < th - margin) {
            out.larger = (N - cnt_sigma - 1) + neg_side(std::fabs(th));
        } else if (th < 0 && sigma > th + margin) {
            int64_t pos = 0;
            if (std::fabs(th) <= T.gersh_hi) {
                wk.lu.factor(T, std::fabs(th));
                ++wk.nfac;
                pos = N - wk.lu.nneg;
            }
            out.larger = (cnt_sigma - 1) + pos;
        } else if (try_resume) {
            // the shift in hand is on the outer side: one more factorisation just inside the Ritz value gives the
            // rigorous rank bound and is what the next check extends (BandLU::resume)
            const double delta = std::max(1e-10 * std::fabs(th), 1e-12 * tn);
            const double sg = th < 0 ? -1.0 : 1.0;
            lu.factor(T, th - sg * delta);
            ++wk.nfac;
            int64_t other = 0;
            if (sg > 0) {
                other = neg_side(std::fabs(th));
                out.larger = (N - lu.nneg - 1) + other;
            } else {
                if (std::fabs(th) <= T.gersh_hi) {
                    wk.lu.factor(T, std::fabs(th));
                    ++wk.nfac;
                    other = N - wk.lu.nneg;
                }
                out.larger = (lu.nneg - 1) + other;
            }
        } else {
            const double delta = std::max(1e-12 * std::fabs(th), 1e-14 * tn);
            out.larger = count_abs_above(std::fabs(th) + delta).above;
        }
        return true;
    };
    // the rejecting pair becomes the first witness of the next check; `keep`: other witnesses still worth following
    auto reject_with = [&](std::vector<double>& x, double th, double rho, const char* how,
                           std::vector<std::vector<double>>* keep = nullptr, std::vector<double>* keep_theta = nullptr) {
        if (std::strcmp(how, "witness") != 0) wlu_.ck_row = -1;
        std::vector<std::vector<double>> nw;
        std::vector<double> nt;
        nw.push_back(x);
        nt.push_back(th);
        if (keep)
            for (size_t i = 0; i < keep->size(); ++i) {
                nw.push_back(std::move((*keep)[i]));
                nt.push_back((*keep_theta)[i]);
            }
        wit_.swap(nw);
        wit_theta_.swap(nt);
        R.witness_rho = rho;
        R.witness_theta = th;
        if (verbose > 1)
            std::fprintf(stderr, "[rbl] check N=%lld %s theta=%.12g rho=%.3e (nfac=%d, %.2f ms)\n", (long long)N, how, th, rho, wk.nfac,
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() * 1e3);
        return finish(false);
    };

    // the witnesses after the first, each followed on its own thread; true (and R filled by reject_with) when one of them
    // is a wanted pair that has not converged
    auto follow_extras = [&]() -> bool {
        if (wit_.size() < 2) return false;
        const size_t nw = wit_.size() - 1;
        std::vector<Followed> fw(nw);
        std::atomic<size_t> next{0};
        std::atomic<bool> cancelled{false};
        auto run = [&]() {
            try {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= nw) break;
                    fw[i] = follow_witness(T, wit_[i + 1], bi, b);
                }
            } catch (const Cancelled&) {
                cancelled = true;
                next = nw;
            }
        };
        const int nt = (int)std::min<size_t>((size_t)std::max(1, threads), nw);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(run);
        run();
        for (auto& t : th) t.join();
        if (cancelled) throw Cancelled{};
        int best = -1;
        for (size_t i = 0; i < nw; ++i) {
            wk.nfac += fw[i].nfac;
            if (verbose > 2) std::fprintf(stderr, "[rbl]   stage1 w%zu ok=%d theta=%.10g res=%.2e larger=%lld rho=%.3e\n", i + 1, (int)fw[i].ok, fw[i].theta, fw[i].res, (long long)fw[i].larger, fw[i].ok ? fw[i].rho : -1.0);
            if (fw[i].ok && fw[i].rho > tol && fw[i].larger < k && (best < 0 || fw[i].rho > fw[best].rho)) best = (int)i;
        }
        if (best < 0) return false;
        // the slowest unconverged one is followed from now on; the other unconverged ones stay on the list
        std::vector<std::vector<double>> keep;
        std::vector<double> keep_theta;
        for (size_t i = 0; i < nw; ++i)
            if ((int)i != best && fw[i].ok && fw[i].rho > tol && fw[i].larger < k) {
                keep.push_back(std::move(fw[i].x));
                keep_theta.push_back(fw[i].theta);
            }
        wlu_.ck_row = -1;  // the kept factorisation belongs to the converged first witness
        reject_with(fw[best].x, fw[best].theta, fw[best].rho, "witness", &keep, &keep_theta);
        return true;
    };

    // ---- stage 1: witnesses of the previous check (zero-padded old Ritz vectors) -------------------------
    // The first one - the pair this checker has been following - is refined here with its factorisation extended from the
    // previous check.  Only when it has converged are the others looked at, all at once on their own threads.
    if (!force_full && bi && !wit_.empty()) {
        if ((int64_t)wit_[0].size() <= N) {
            std::vector<double> x(N, 0.0);
            std::copy(wit_[0].begin(), wit_[0].end(), x.begin());
            const double nn = nrm2(x.data(), N);
            if (nn > 0) {
                scal(x.data(), 1.0 / nn, N);
                Refined rf;
                const bool rok = refine(wlu_, true, x, false, 0.0, 0, rf);
                if (verbose > 2) std::fprintf(stderr, "[rbl]   stage1 w0 ok=%d theta=%.10g res=%.2e larger=%lld rho=%.3e\n", (int)rok, rf.theta, rf.res, (long long)rf.larger, rok ? resid_bound(bi, b, x) : -1.0);
                if (rok) {
                    const double rho = resid_bound(bi, b, x);
                    if (rho > tol && rf.larger < k) {
                        std::vector<std::vector<double>> keep(wit_.begin() + 1, wit_.end());
                        std::vector<double> keep_theta(wit_theta_.begin() + 1, wit_theta_.end());
                        return reject_with(x, rf.theta, rho, "witness", &keep, &keep_theta);
                    }
                }
            }
        }
        if (follow_extras()) return R;
    }

    stage_now = 1;
    // ---- stage 2: a pair found from a shift x with a <= #{|lambda| > x} <= hi (hi <= k-1) ----------------
    // The eigenvalue nearest to such an x has rank <= hi+1 <= k on either side of x, so whatever inverse
    // iteration at x converges to is one of the k wanted pairs.  First a target a little inside the wanted
    // set (robust while Ritz values still enter it), then the boundary itself.
    // With FRESH seeds (all k pairs of a T at least 95% of this size, from the tracker or from the last full check) the
    // refinement of stage 3 costs about what a failed stage 2 costs (and runs on all threads): when the witnesses of stage 1
    // have converged - which is the situation of the accepting check - go there directly.
    const bool seeds_fresh = (int64_t)seeds_.size() == k && !seeds_.empty() && (int64_t)seeds_[0].v.size() <= N &&
                             (double)seeds_[0].v.size() >= seed_fresh_frac(threads) * (double)N;
    if (!force_full && bi && k >= 1 && !seeds_fresh) {
        const int64_t margin = std::max<int64_t>(1, std::min<int64_t>(k / 6, k - 1));
        struct Target { int64_t lo, hi; double* x; double* step; };
        Target targets[2] = {{std::max<int64_t>(0, k - 1 - 2 * margin), k - 1 - margin, &xA_, &stepA_},
                             {std::max<int64_t>(0, k - margin), k - 1, &xB_, &stepB_}};
        for (int ti = 0; ti < 2; ++ti) {
            const Target& tg = targets[ti];
            if (tg.hi < tg.lo || tg.hi < 0) continue;
            // locate x with lo <= above(x) <= hi; above(.) never decreases from one check to the next, so the
            // previous x is a lower bound
            double xl = 0.0, xh = g, x = -1.0;
            double xs = (*tg.x > 0.0) ? *tg.x : -1.0;
            double step = std::max(*tg.step, 1e-6 * tn);
            bool found = false;
            if (xs >= 0.0) {
                Cnt c = count_abs_above(xs);
                if (c.above >= tg.lo && c.above <= tg.hi) { x = xs; found = true; }
                else if (c.above > tg.hi) {
                    xl = xs;
                    for (int it = 0; it < 40 && !found; ++it) {
                        const double xn = std::min(g, xl + step);
                        c = count_abs_above(xn);
                        if (c.above >= tg.lo && c.above <= tg.hi) { x = xn; found = true; }
                        else if (c.above > tg.hi) { xl = xn; step *= 3.0; if (xn >= g) break; }
                        else { xh = xn; break; }
                    }
                } else {
                    xh = xs;
                }
            }
            for (int it = 0; it < 60 && !found; ++it) {
                if (xh - xl <= 1e-13 * tn) break;
                const double xm = 0.5 * (xl + xh);
                const Cnt c = count_abs_above(xm);
                if (c.above >= tg.lo && c.above <= tg.hi) { x = xm; found = true; }
                else if (c.above > tg.hi) xl = xm;
                else xh = xm;
            }
            if (verbose > 2) std::fprintf(stderr, "[rbl]   stage2 target %d [%lld,%lld] found=%d x=%.10g (nfac=%d)\n", ti, (long long)tg.lo, (long long)tg.hi, (int)found, x, wk.nfac);
            if (!found) continue;
            *tg.step = std::max(2.0 * (x - std::max(*tg.x, 0.0)), 1e-6 * tn);
            *tg.x = x;
            // inverse iteration at +x (or -x if the spectrum there is the nearer one): start random
            wk.lu.factor(T, x);
            ++wk.nfac;
            std::vector<double> v;
            wk.random_unit(v, N);
            for (int it = 0; it < 5; ++it) {
                wk.lu.solve(v.data());
                const double n2 = nrm2(v.data(), N);
                if (!(n2 > 0) || !std::isfinite(n2)) { wk.random_unit(v, N); continue; }
                scal(v.data(), 1.0 / n2, N);
            }
            Refined rf;
            const bool rok = refine(wk.lu, false, v, true, x, wk.lu.nneg, rf);
            if (verbose > 2) std::fprintf(stderr, "[rbl]   stage2 refine ok=%d theta=%.10g res=%.2e larger=%lld rho=%.3e\n", (int)rok, rf.theta, rf.res, (long long)rf.larger, rok ? resid_bound(bi, b, v) : -1.0);
            if (!rok) continue;
            const double rho = resid_bound(bi, b, v);
            if (rho > tol && rf.larger < k) return reject_with(v, rf.theta, rho, ti == 0 ? "inner pair" : "boundary pair");
        }
    }

    // ---- stage 3: all k pairs of largest |lambda| --------------------------------------------------------
    ++full_checks;
    stage_now = 2;
    auto since_start = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() * 1e3; };
    const double ms_stage3_begin = since_start();
    auto seeds_usable = [&]() {
        return (int64_t)seeds_.size() == k && !seeds_.empty() && (int64_t)seeds_[0].v.size() <= N &&
               (double)seeds_[0].v.size() >= seed_min_frac() * (double)N;  // staler seeds rarely survive the refinement
    };
    // No usable seeds: before solving from scratch, the owner gets a chance to hand some over (the solver waits here for
    // the tracker's pass in flight - finishing that pass and refining its pairs is never slower than starting over with
    // the tracker paused, and usually several times faster).
    if (!seeds_usable() && need_seeds && !force_full) {
        const size_t had = seeds_.empty() ? 0 : seeds_[0].v.size();
        need_seeds(N);
        // fresh seeds bring fresh witnesses (the pairs that were slowest when the seeds were computed): they get their turn
        // before all k pairs are refined
        if (bi && !seeds_.empty() && seeds_[0].v.size() != had && follow_extras()) return R;
    }
    // all k pairs from here on: every thread of the box is wanted (the accepting check is the one the device waits for)
    struct FullFlag {
        std::atomic<bool>* f;
        explicit FullFlag(std::atomic<bool>* p) : f(p) { if (f) f->store(true); }
        ~FullFlag() { if (f) f->store(false); }
    } full_guard(full_flag);
    double ms_refined = 0, ms_validated = 0;
    std::vector<Pair> pairs;
    std::vector<Pair> known_pairs;
    bool from_seeds = false;
    if (seeds_usable()) {
        // Fast path: the k pairs of an earlier full solve are refined in parallel (seeds that do not converge or that
        // collapse onto the same eigenvector are dropped).  Sturm counts then say exactly how many eigenvalues are
        // missing and where - above the smallest value found (Ritz values that entered the wanted set), inside its
        // cluster, or below it - and those are extracted by deflated RQI.  A final count validates the set.
        int64_t nf = 0;
        const int64_t nseed = (int64_t)seeds_[0].v.size();
        const bool refined = refine_seeds(T, k, pairs, nf);
        wk.nfac += (int)nf;
        ms_refined = since_start();
        const char* why = refined ? "" : "too few seeds survived";
        bool good = refined;
        auto by_mag = [](const Pair& a, const Pair& b2) { return std::fabs(a.theta) > std::fabs(b2.theta); };
        if (good) {
            std::stable_sort(pairs.begin(), pairs.end(), by_mag);
            const int64_t kf = (int64_t)pairs.size();
            const double tk = std::fabs(pairs[kf - 1].theta);
            const double delta = std::max(1e-10 * tk, 1e-12 * tn);
            if (!(pairs[kf - 1].theta > 0) || neg_side(std::max(0.0, tk - delta)) != 0) {
                // two-sided spectra: only the plain (nothing missing) case is handled here
                int64_t fa = 0;
                for (int64_t j = 0; j < kf; ++j)
                    if (std::fabs(pairs[j].theta) > tk + delta) ++fa;
                good = (kf == k) && (count_abs_above(tk + delta).above == fa) && (count_abs_above(std::max(0.0, tk - delta)).above >= k);
                if (!good) why = "two-sided spectrum with missing pairs";
            } else {
                std::vector<double> f(kf);
                for (int64_t j = 0; j < kf; ++j) f[j] = pairs[j].theta;
                // the two counts of the validation (just above / just below the k-th value) are taken at once on two
                // threads and remembered: when nothing is missing the "final validation" below asks for the same points
                std::vector<std::pair<double, int64_t>> memo;
                {
                    std::vector<double> xs;
                    for (double x : {tk + delta, std::max(0.0, tk - delta)})
                        if (x <= T.gersh_hi) xs.push_back(x);
                    std::vector<int64_t> below;
                    int64_t nf0 = 0;
                    parallel_below(T, xs, threads, below, nf0);
                    wk.nfac += (int)nf0;
                    for (size_t i = 0; i < xs.size(); ++i) memo.push_back({xs[i], N - below[i]});
                }
                auto above = [&](double x) -> int64_t {  // #{lambda > x}
                    if (x > T.gersh_hi) return 0;
                    for (auto& mm : memo)
                        if (mm.first == x) return mm.second;
                    wk.lu.factor(T, x);
                    ++wk.nfac;
                    return N - wk.lu.nneg;
                };
                auto found_above = [&](double x) -> int64_t {
                    int64_t c = 0;
                    for (int64_t i = 0; i < kf; ++i)
                        if (f[i] > x) ++c;
                    return c;
                };
                // Sturm counts against the refined set: anything missing (Ritz values that entered the wanted set since
                // the seeds were computed) sends the check to the seed-assisted slicing below, which finds the missing
                // pairs in parallel and keeps the refined ones
                const int64_t c_hi = above(tk + delta);
                const int64_t fa_hi = found_above(tk + delta);
                if (c_hi != fa_hi) {
                    good = false;
                    why = "eigenvalues missing above the smallest refined one";
                } else if (kf < k) {
                    good = false;
                    why = "fewer than k refined pairs";
                }
                if (good) {
                    std::stable_sort(pairs.begin(), pairs.end(), by_mag);
                    if ((int64_t)pairs.size() < k) { good = false; why = "fewer than k refined pairs"; }
                }
                if (good) {
                    pairs.resize(k);
                    // final validation: every eigenvalue strictly above the (possibly degenerate) k-th one is in the set
                    const double tk2 = std::fabs(pairs[k - 1].theta);
                    int64_t fa2 = 0;
                    for (int64_t j = 0; j < k; ++j)
                        if (std::fabs(pairs[j].theta) > tk2 + delta) ++fa2;
                    good = (above(tk2 + delta) == fa2) && (above(std::max(0.0, tk2 - delta)) >= k);
                    if (!good) why = "final count validation failed";
                }
            }
        }
        from_seeds = good;
        ms_validated = since_start();
        if (!from_seeds) {
            // the refined seeds are still eigenpairs of T: the slicing below only has to find what is missing
            known_pairs.clear();
            if (refined)  // (deduplicated by refine_seeds)
                for (auto& p : pairs)
                    if ((int64_t)p.v.size() == N && p.res <= 1.01e-11 * tn) known_pairs.push_back(std::move(p));
            pairs.clear();
        }
        if (verbose > 0)
            std::fprintf(stderr, "[rbl] full check N=%lld from seeds (N_seed=%lld): %s%s%s\n", (long long)N,
                         (long long)nseed, from_seeds ? "ok" : "rejected", from_seeds ? "" : ": ", why);
    }
    // bracket the k-th largest |lambda|: largest x_lo with #{|lambda| > x_lo} >= k (within a modest surplus)
    double x_lo = 0.0, x_hi = g;
    Cnt c_lo{0, 0, N};
    bool have_clo = false;
    std::vector<Interval> roots;
    const bool hinted = false;
    if (!hinted) {
        // one-sided spectra (nothing below -x for the x in question: the shifted BASELINE operators) with several threads:
        // multi-section - `threads` Sturm counts per round, all at once - instead of one bisection step after the other
        const double x_onesided = std::max(0.0, -T.gersh_lo) * (1.0 + 1e-12) + 1e-300;  // above this nothing lies below -x
        while (!from_seeds && threads > 1 && N >= 400 && !(c_lo.above >= k && c_lo.above <= k + std::max<int64_t>(2, k / 8)) &&
               x_hi - std::max(x_lo, x_onesided) > 1e-10 * tn) {
            const double lo_eff = std::max(x_lo, x_onesided);
            const int np = std::min(threads, 16);
            std::vector<double> xs;
            for (int i = (lo_eff > x_lo ? 0 : 1); i <= np; ++i) xs.push_back(lo_eff + (x_hi - lo_eff) * (double)i / (double)(np + 1));
            std::vector<int64_t> below;
            int64_t nf0 = 0;
            parallel_below(T, xs, threads, below, nf0);
            wk.nfac += (int)nf0;
            bool moved = false;
            for (size_t i = 0; i < xs.size(); ++i) {
                const int64_t above = N - below[i];
                if (above >= k) { x_lo = xs[i]; c_lo = Cnt{below[i], 0, above}; have_clo = true; moved = true; }
                else { x_hi = xs[i]; break; }
            }
            if (!moved && lo_eff > x_lo) break;  // the k-th value lies where the spectrum is two-sided: bisection below
        }
        for (int it = 0; it < 60 && !from_seeds; ++it) {
            if (c_lo.above >= k && c_lo.above <= k + std::max<int64_t>(2, k / 8)) break;
            if (x_hi - x_lo <= 1e-13 * tn) break;
            const double xm = 0.5 * (x_lo + x_hi);
            const Cnt cm = count_abs_above(xm);
            if (cm.above >= k) { x_lo = xm; c_lo = cm; have_clo = true; } else { x_hi = xm; }
        }
        if (x_lo > 0 && have_clo) {
            roots.push_back(Interval{x_lo, g, c_lo.below_pos, N});
            if (c_lo.below_neg > 0) roots.push_back(Interval{-g, -x_lo, 0, c_lo.below_neg});
        } else {
            roots.push_back(Interval{-g, g, 0, N});
        }
    }
    if (!from_seeds) {
        int64_t nf = 0, reused = 0;
        slice(T, roots, threads, pairs, nf, known_pairs.empty() ? nullptr : &known_pairs, &reused);
        wk.nfac += (int)nf;
        if (verbose > 0 && !known_pairs.empty())
            std::fprintf(stderr, "[rbl] full check N=%lld: slicing reused %lld of %lld refined seeds (%lld factorisations)\n",
                         (long long)N, (long long)reused, (long long)known_pairs.size(), (long long)nf);
    }
    // sort_eig_abs: k largest |lambda|, returned by descending |lambda|
    std::stable_sort(pairs.begin(), pairs.end(),
                     [](const Pair& a, const Pair& b2) { return std::fabs(a.theta) > std::fabs(b2.theta); });
    if ((int64_t)pairs.size() > k) pairs.resize(k);
    const int64_t kk = (int64_t)pairs.size();
    R.d.assign(k, 0.0);
    R.resid.assign(k, 0.0);
    R.s.assign((size_t)N * k, 0.0);
    bool all_ok = (kk == k);
    std::vector<std::pair<double, int64_t>> order;
    for (int64_t j = 0; j < kk; ++j) {
        R.d[j] = pairs[j].theta;
        std::copy(pairs[j].v.begin(), pairs[j].v.end(), R.s.begin() + (size_t)j * N);
        const double rho = resid_bound(bi, b, pairs[j].v);
        R.resid[j] = rho;
        if (rho > tol) all_ok = false;
        order.emplace_back(rho, j);
    }
    R.have_all = (kk == k);
    // witnesses for the next check: the worst few pairs
    std::sort(order.begin(), order.end(), [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b2) { return a.first > b2.first; });
    wit_.clear();
    wit_theta_.clear();
    for (size_t j = 0; j < order.size() && j < (size_t)(1 + kExtraWitnesses); ++j) {
        wit_.push_back(pairs[order[j].second].v);
        wit_theta_.push_back(pairs[order[j].second].theta);
    }
    if (kk == k) seeds_ = pairs;  // the next full check starts from these
    if (kk == k) {  // seeds for the stage-2 brackets of the next check
        const int64_t margin = std::max<int64_t>(1, std::min<int64_t>(k / 6, k - 1));
        const int64_t ia = std::max<int64_t>(0, k - 1 - margin - margin / 2);
        xB_ = std::fabs(pairs[k - 1].theta) * (1.0 - 1e-12);
        xA_ = std::fabs(pairs[ia].theta) * (1.0 - 1e-12);
        if (stepA_ <= 0) stepA_ = 1e-5 * tn;
        if (stepB_ <= 0) stepB_ = 1e-5 * tn;
    }
    if (verbose > 1)
        std::fprintf(stderr, "[rbl]   timeline of this check [ms]: stages 1-2 %.1f, seeds refined at %.1f, validated at %.1f, done at %.1f\n",
                     ms_stage3_begin, ms_refined, ms_validated, since_start());
    if (verbose > 0)
        std::fprintf(stderr, "[rbl] full check N=%lld (%s) found=%lld worst rho=%.3e conv=%d (nfac=%d)\n", (long long)N,
                     from_seeds ? "seeds refined" : (hinted ? "hinted slicing" : "slicing"),
                     (long long)kk, order.empty() ? 0.0 : order[0].first, (int)(all_ok && bi != nullptr), wk.nfac);
    return finish(all_ok && bi != nullptr);
}

}  // namespace rbl

// Host-side eigen-check of the block-tridiagonal Lanczos matrix T (no CUDA in this file).
//
// Replaces, on the host and with the same inputs/outputs, the reference's per-check sequence
//     D,V = dsbev('V','L',T); D,V = sort_eig_abs(D,V,k); check_convergence(Bi,V,b,k,tol)
// (Julia/common.jl:36-65, called from RBL_gpu.jl:186-192).  dsbev computes ALL N eigenvectors in
// O(N^3); here only what the decision needs is computed, by spectrum slicing on the band matrix:
// Sturm counts from a row-wise elimination with pairwise pivoting (leading-principal-minor signs),
// inverse / Rayleigh-quotient iteration with the same factorisation for the vectors.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <vector>

namespace rbl {

// Symmetric band matrix, half-bandwidth kd, stored as full band rows: F[r*(2kd+1) + (c - r + kd)].
// thrown out of BandTopK::check when BandSym::cancel was raised (background tracker being stopped)
struct Cancelled {};

struct BandSym {
    int64_t N = 0;
    int kd = 0;
    std::vector<double> F;
    double norm_inf = 0.0;
    double gersh_lo = 0.0, gersh_hi = 0.0;  // Gershgorin interval containing the spectrum
    const std::atomic<bool>* cancel = nullptr;  // polled at every factorisation; set -> Cancelled is thrown
    const std::atomic<bool>* pause = nullptr;   // polled at every factorisation; while set the caller sleeps (the background
                                                // tracker yields the cores to a full check of the main checker)

    void reset(int64_t n, int kd_);
    inline double& at(int64_t r, int64_t c) { return F[(size_t)r * (2 * kd + 1) + (size_t)(c - r + kd)]; }
    inline double at(int64_t r, int64_t c) const { return F[(size_t)r * (2 * kd + 1) + (size_t)(c - r + kd)]; }
    void set_sym(int64_t r, int64_t c, double v) { at(r, c) = v; at(c, r) = v; }
    void from_lapack_lower(int64_t n, int kd_, const double* ab);  // (kd+1) x N column-major
    void update_norm();
    void matvec(const double* x, double* y) const;
};

// LU of (T - x I) by row-wise elimination with pairwise pivoting; also yields the Sturm count.
struct BandLU {
    int64_t N = 0;
    int kd = 0;
    std::vector<double> U;    // N x (2kd+1): row i holds columns i .. i+2kd
    std::vector<double> L;    // N x kd multipliers
    std::vector<uint8_t> sw;  // N x kd swap flags
    std::vector<double> w;
    int64_t nneg = 0;         // number of eigenvalues of T below the shift
    double shift_ = 0.0;
    void factor(const BandSym& T, double shift);
    bool resume(const BandSym& T);  // same shift, T extended at its end: re-eliminates only the last rows
    void solve(double* v) const;
    template <int KD>
    void solve_t(double* v) const;
    // state saved just before row N - kd (the first row that changes when T grows)
    int64_t ck_row = -1;
    std::vector<double> ck_U;
    int ck_P = 1;
    int64_t ck_nneg = 0;

private:
    void run(const BandSym& T, int64_t r_start);
    template <int KD>
    void run_t(const BandSym& T, int64_t r_start);
    int P_ = 1;

public:
};

struct Pair {
    double theta = 0;
    double res = 0;  // ||T v - theta v||
    std::vector<double> v;
};

struct TopKResult {
    bool converged = false;
    int64_t N = 0;
    std::vector<double> d;      // k eigenvalues, descending |lambda|
    std::vector<double> s;      // N x k column-major eigenvectors (same order)
    std::vector<double> resid;  // k residual bounds ||B_i s_last||
    bool have_all = false;      // d/s/resid hold all k pairs (full check ran)
    double witness_rho = -1.0;  // residual bound of the pair that proved non-convergence (stages 1-2), else -1
    double witness_theta = 0.0; // its Ritz value
    int factorizations = 0;
};

// Stateful checker: keeps a "witness" Ritz pair between checks so that the common (not yet
// converged) case costs a handful of band factorisations instead of a full eigensolve.
class BandTopK {
public:
    int threads = 1;
    int verbose = 0;
    std::atomic<bool>* full_flag = nullptr;     // raised while this checker computes all k pairs (stage 3)
    // called (with the size of T) when a full check is about to start without usable seeds: last chance for set_seeds
    std::function<void(int64_t)> need_seeds;
    int64_t total_factorizations = 0;
    int64_t resumed_factorizations = 0;
    // where the decisions came from and what they cost (seconds): [0] witness, [1] bracketed pair, [2] full
    int stage_hits[3] = {0, 0, 0};
    double stage_sec[3] = {0, 0, 0};  // witness factorisations extended from the previous check instead of recomputed
    int full_checks = 0;

    // T: current N x N band matrix; bi: b x b upper-triangular B_i row-major (bi[r*b+c]); k wanted.
    // force_full: compute all k pairs even when a witness already proves non-convergence.
    TopKResult check(const BandSym& T, const double* bi, int b, int64_t k, double tol, bool force_full);
    void reset() { wit_.clear(); wit_theta_.clear(); xA_ = xB_ = stepA_ = stepB_ = 0; seeds_.clear(); }
    // k Ritz pairs of an EARLIER T (d: k values, svec: Ns x k column-major): starting points of the next full check
    // resid (optional): the k residual bounds of those pairs at the time they were computed; the worst few join the witnesses
    void set_seeds(const std::vector<double>& d, const std::vector<double>& svec, int64_t Ns, int64_t k,
                   const std::vector<double>* resid = nullptr);
    static constexpr int kExtraWitnesses = 8;
    bool has_seeds() const { return !seeds_.empty(); }

private:
    bool refine_seeds(const BandSym& T, int64_t k, std::vector<Pair>& pairs, int64_t& nfac);
    std::vector<Pair> seeds_;
    BandLU wlu_;                            // factorisation kept for the first witness between checks
    std::vector<std::vector<double>> wit_;  // Ritz vectors of T that failed the bound at the last check
    std::vector<double> wit_theta_;
    double xA_ = 0, xB_ = 0;                // stage-2 bracket points of the last check (lower bounds for the next)
    double stepA_ = 0, stepB_ = 0;          // how far they moved
};

int64_t band_count_below(const BandSym& T, double x);

}  // namespace rbl

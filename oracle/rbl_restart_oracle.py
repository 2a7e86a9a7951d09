"""CPU twin of the restarted / filtered RBL solve.

TEST INFRASTRUCTURE ONLY (see rbl_oracle.py header): imported by tests/ and bench.py's CPU arm, never by the product.

The reference's restart is ``Julia/restarted.jl``: ``RBL_restarted`` (:196-246, CPU) / ``RBL_gpu_restarted``
(:98-146, GPU) run a short Lanczos cycle (``new_lanczos_iteration`` :148-194 / ``lanczos_iteration_res`` :23-96),
take the Ritz pairs in descending order, LOCK the leading ones whose residual bound is below 1e-7
(:122-131 / :222-227: ``push!(Qlock, qv)``), restart from the first unconverged Ritz vector (:133-135 / :229-230)
and keep every new Lanczos block orthogonal to the locked vectors (``restart_reorth_gpu!`` :1-21, called at :40,:58-59;
``part_reorth!(length(Qlock),Qlock,...)`` :172,:187).  It is hard-wired to b = 1, never fills ``V`` (:100,:145) and
has no tests.  What is restated here - and built on the device in csrc/solver.cu - is its generalisation:

* block size b: the restart block is the b best not-yet-locked Ritz vectors (restarted.jl restarts from one);
* every pair among the wanted ones whose bound is below tol is locked (restarted.jl stops at the first
  unconverged one); the answer is the k largest |lambda| of locked + final pairs, and ``V`` is returned;
* optionally the cycle iterates with a Chebyshev-filtered operator p(A) = rho * T_d((A - c)/e) that damps the
  unwanted interval [a, b] = [c - e, c + e]: Ritz VECTORS of p(A) are Ritz vectors of A, the eigenvalues are
  recovered as Rayleigh quotients with A, and ||A v - lambda v|| is measured explicitly.  The interval comes from a
  short plain probe run (its Ritz values) and the Gershgorin bounds of A.

Functions cite the reference lines they generalise; the arithmetic inside a cycle is rbl_oracle.lanczos_iteration
(RBL.jl:74-117) with the locked vectors added to the periodic re-orthogonalisation.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import rbl_oracle as ro

DOUBLE = np.float64


# --------------------------------------------------------------------------- filter
@dataclass
class ChebFilter:
    degree: int = 0          # 0: identity (plain operator)
    a: float = 0.0           # damped interval [a, b]
    b: float = 0.0
    rho: float = 1.0         # p = rho * T_d((x - c)/e)
    two_sided: bool = False

    @property
    def c(self):
        return 0.5 * (self.a + self.b)

    @property
    def e(self):
        return 0.5 * (self.b - self.a)

    def scalar(self, lam):
        """p(lam) for scalars / arrays (host-side bookkeeping only)."""
        if self.degree == 0:
            return np.asarray(lam, dtype=DOUBLE)
        x = (np.asarray(lam, dtype=DOUBLE) - self.c) / self.e
        ax = np.abs(x)
        out = np.where(ax <= 1.0, np.cos(self.degree * np.arccos(np.clip(x, -1, 1))),
                       np.cosh(self.degree * np.arccosh(np.maximum(ax, 1.0))) * np.where((x < 0) & (self.degree % 2 == 1), -1.0, 1.0))
        return self.rho * out

    def invert(self, theta, side):
        """Eigenvalue estimate of A from a Ritz value theta of p(A) on the wanted side (side: +1 above the damped
        interval, -1 below, 0 two-sided); None inside the interval or when the sign of theta belongs to the other side."""
        if self.degree == 0:
            return float(theta)
        y = theta / self.rho
        if not abs(y) > 1.0:
            return None
        if side > 0 and y < 0:
            return None
        if side < 0 and (y < 0) != (self.degree % 2 == 1):
            return None
        x = np.cosh(np.arccosh(abs(y)) / self.degree)
        sgn = float(side) if side != 0 else (-1.0 if y < 0 else 1.0)
        return float(self.c + self.e * sgn * x)

    def apply(self, A, Q):
        """p(A) Q by the three-term recurrence t_{j+1} = 2 (A - c)/e t_j - t_{j-1}; rho folded into the last step."""
        if self.degree == 0:
            return np.asarray(A @ Q, dtype=DOUBLE)
        c, e = self.c, self.e
        t_prev = Q
        t = (np.asarray(A @ Q, dtype=DOUBLE) - c * Q) / e
        if self.degree == 1:
            return self.rho * t
        for j in range(2, self.degree + 1):
            s = self.rho if j == self.degree else 1.0
            t_next = s * ((2.0 / e) * (np.asarray(A @ t, dtype=DOUBLE) - c * t) - t_prev)
            t_prev, t = t, t_next
        return t


def gershgorin(A):
    """Rigorous bounds [lo, hi] of the spectrum of a symmetric sparse / dense matrix."""
    if sp.issparse(A):
        M = sp.csr_matrix(A)
        d = M.diagonal()
        r = np.asarray(abs(M).sum(axis=1)).ravel() - np.abs(d)
    else:
        M = np.asarray(A)
        d = np.diag(M)
        r = np.abs(M).sum(axis=1) - np.abs(d)
    return float(np.min(d - r)), float(np.max(d + r))


def place_filter(theta: np.ndarray, k: int, kk: int, degree: int, glo: float, ghi: float) -> ChebFilter:
    """Filter from the probe's Ritz values (sorted by descending |theta|, at least kk of them).

    one-sided (all k leading Ritz values of one sign): damp [gershgorin end, theta_kk]
    two-sided: damp [-|theta_kk|, |theta_kk|] with an odd degree (p(-x) = -p(x) keeps +lambda and -lambda apart)
    rho makes p(theta_k) = |theta_1| so that the absolute tolerance of check_convergence (common.jl:56-65) keeps its
    meaning relative to ||A||.
    """
    th = np.asarray(theta, dtype=DOUBLE)
    assert len(th) >= kk >= k >= 1
    cut = abs(th[kk - 1])
    lead = th[:kk]
    f = ChebFilter(degree=degree)
    if np.all(lead > 0):
        f.a, f.b = (glo if glo < cut else cut - abs(cut)), cut
    elif np.all(lead < 0):
        f.a, f.b = -cut, (ghi if ghi > -cut else -cut + abs(cut))
    else:
        f.two_sided = True
        f.a, f.b = -cut, cut
        if degree % 2 == 0:
            f.degree = degree + 1
    xk = abs((th[k - 1] - f.c) / f.e)
    f.rho = 1.0
    tk = abs(float(f.scalar(th[k - 1])))
    f.rho = abs(th[0]) / tk if tk > 0 and xk > 1.0 else 1.0
    return f


# --------------------------------------------------------------------------- one Lanczos cycle with locking
@dataclass
class CycleResult:
    converged: bool
    iterations: int
    D: np.ndarray            # kk Ritz values of the cycle operator, descending |theta|
    S: np.ndarray            # (iterations*b) x kk eigenvectors of T
    bounds: np.ndarray       # kk residual bounds ||B_i S[end-b+1:end, j]||
    Q: list = field(default_factory=list)


def _orth_against(Y, W):
    """restart_reorth_gpu! (restarted.jl:1-21): W -= Y (Y' W), in place."""
    if Y is not None and Y.shape[1] > 0:
        W -= Y @ (Y.T @ W)


def lanczos_cycle(A, flt: ChebFilter, k_rem: int, kk: int, b: int, max_blocks: int, Q1: np.ndarray, Ylock, *,
                  tol: float, reorth_period: int = 2, check_period: int = 4, run_to_cap: bool = False) -> CycleResult:
    """rbl_oracle.lanczos_iteration (RBL.jl:74-117) on the operator flt(A), blocks kept orthogonal to `Ylock`
    (restarted.jl:40,58-59,172).  Stops at the first check whose k_rem leading bounds are all <= tol, or when
    max_blocks blocks are stored; the result then carries the kk leading Ritz pairs of the last T."""
    Q = [Q1]
    Qi = Q1
    U = flt.apply(A, Qi)
    Ai = Qi.T @ U
    U -= Qi @ Ai
    Qn, R = np.linalg.qr(U)
    Qi, Bi = Qn, R
    T = ro.insert_a(Ai, b)
    ro.insert_b(Bi, T, b, 1)
    i = 1
    out = None
    while i < max_blocks:
        i += 1
        Q.append(Qi)
        if i % reorth_period == 0:
            if Ylock is not None and Ylock.shape[1] > 0:      # locked vectors take part in the periodic reorth
                _orth_against(Ylock, Q[i - 1])
                _orth_against(Ylock, Q[i - 2])
            ro.part_reorth(Q)
        ro.loc_reorth(Q[i - 1], Q[i - 2])
        U = flt.apply(A, Q[i - 1])
        U -= Q[i - 2] @ Bi.T
        Ai = Q[i - 1].T @ U
        U -= Q[i - 1] @ Ai
        Qn, R = np.linalg.qr(U)
        Qi, Bi = Qn, R
        T = np.hstack([T, ro.insert_a(Ai, b)])
        last = i >= max_blocks
        if (i * b > k_rem and i % check_period == 0 and not run_to_cap) or last:
            Dall, Vall = ro.dsbev(T)
            want = min(kk, T.shape[1])
            D, S = ro.sort_eig_abs(Dall, Vall, want)
            D, S = D[::-1].copy(), S[:, ::-1].copy()
            bounds = ro.residual_bounds(Bi, S, b)
            ok = bool(np.all(bounds[:k_rem] <= tol)) and not run_to_cap
            out = CycleResult(ok, i, D, S, bounds, Q)
            if ok:
                return out
        ro.insert_b(Bi, T, b, i)
    return out


MAX_FILTER_DEGREE = 256


MAX_DYNAMIC_RANGE = 1e3


def cap_degree(f: "ChebFilter", lam1: float, degree: int) -> int:
    """Largest degree <= `degree` for which p(lam1) / p(edge of the damped interval) stays below MAX_DYNAMIC_RANGE (lam1:
    estimate of the eigenvalue of largest magnitude).  A one-sided Chebyshev filter of high degree over a WIDE wanted
    interval makes ||p(A)|| exceed the last wanted Ritz value by many orders of magnitude; those Ritz values are then
    computed with an absolute error relative to ||p(A)|| and the absolute tolerance of the convergence test loses its
    meaning.  A wide wanted interval does not need a high degree."""
    x1 = abs((lam1 - f.c) / f.e) if f.e > 0 else 1.0
    d = degree
    if x1 > 1.0:
        d = min(d, int(np.floor(np.arccosh(MAX_DYNAMIC_RANGE) / np.arccosh(x1))))
    d = max(d, 2)
    if f.two_sided and d % 2 == 0:
        d += 1
    return d


def settle_steps(kk: int, b: int) -> int:
    """Length of a settling cycle of the filtered restart: enough blocks for rough Ritz values of all kk pairs."""
    return max(8, 3 * -(-kk // b))


# --------------------------------------------------------------------------- driver
@dataclass
class RestartStats:
    cycles: int = 0
    locked: int = 0
    block_steps: int = 0          # filtered block steps over all cycles
    probe_steps: int = 0
    operator_applications: int = 0
    converged: bool = False
    filter: ChebFilter | None = None
    max_residual: float = 0.0


def RBL_restarted(A, k: int, b: int, Omega: np.ndarray, *, max_blocks: int, tol: float = 1e-7, filter_degree: int = 0,
                  probe_steps: int = 0, restart: bool = True, max_cycles: int = 50, reorth_period: int = 2,
                  check_period: int = 4, return_details: bool = False, replace_filter: bool = True):
    """Generalised RBL_restarted / RBL_gpu_restarted (restarted.jl:98-146,196-246).

    Returns (D, V): the k eigenvalues of A of largest magnitude (descending |lambda|) and their vectors.
    `max_blocks` is the number of Krylov blocks the buffer holds (locked vectors included)."""
    n = A.shape[0]
    st = RestartStats()
    Q1 = np.linalg.qr(np.asarray(A @ np.asarray(Omega, dtype=DOUBLE), dtype=DOUBLE))[0]      # RBL.jl:137
    flt = ChebFilter(degree=0)
    side, norm_a, lam1_est, want_degree = 0, 0.0, 0.0, 0
    if filter_degree != 0:
        d = filter_degree if filter_degree > 0 else 8
        kk = k + b
        s = probe_steps if probe_steps > 0 else max(8, 2 * -(-kk // b))
        s = min(s, max_blocks)
        pr = lanczos_cycle(A, flt, k, min(kk, s * b), b, s, Q1.copy(), None, tol=tol, reorth_period=reorth_period,
                           check_period=check_period, run_to_cap=True)
        st.probe_steps = pr.iterations
        glo, ghi = gershgorin(A)
        flt = place_filter(pr.D, min(k, len(pr.D)), len(pr.D), d, glo, ghi)
        side = 0 if flt.two_sided else (1 if pr.D[0] > 0 else -1)
        norm_a = abs(pr.D[0])
        lam1_est = float(pr.D[0])
        want_degree = flt.degree
        # (no dynamic-range cap at the probe: its cut lies far below the wanted end; the cap applies from the first
        # re-placement on, where the estimates are good)
        st.operator_applications += pr.iterations
    st.filter = flt
    Y = np.zeros((n, 0))
    Dlock = np.zeros(0)
    start = Q1
    settled = not replace_filter
    final = None
    while True:
        st.cycles += 1
        k_rem = k - Y.shape[1]
        nlb = -(-Y.shape[1] // b)                       # locked vectors occupy whole buffer blocks
        room = max_blocks - nlb
        if room < 3:
            break
        kk = k_rem + b
        steps = room
        if flt.degree > 0 and restart and not settled:
            steps = min(room, settle_steps(kk, b))
        res = lanczos_cycle(A, flt, k_rem, kk, b, steps, start, Y, tol=tol, reorth_period=reorth_period,
                            check_period=check_period)
        st.block_steps += res.iterations
        st.operator_applications += res.iterations * max(1, flt.degree)
        if res.converged or not restart or st.cycles >= max_cycles:
            final = res
            st.converged = res.converged
            break
        if flt.degree > 0:
            # Filtered restarts never lock.  p amplifies the leading (first converged) eigenvalues far more than the last
            # wanted ones - by 1e4 per application for a well placed degree-32 filter - so any imperfection of locked vectors
            # re-grows inside the cycle and comes back as ghost Ritz pairs (observed: BASELINE config 5 returned pairs with
            # residual 2e-5 ||A||, smaller cases garbage).  Instead the cycle is repeated with a better filter from the
            # leading b unconverged Ritz vectors: the filter is re-placed from this cycle's Ritz values mapped back through p
            # (short "settling" cycles until the cut stops moving - a full-length cycle behind a badly placed filter is
            # wasted), then, if a full-length cycle still does not converge, the degree doubles.
            nb = res.iterations
            Qm = np.hstack(res.Q[:nb])
            have = len(res.D)
            kq = min(k_rem, have)
            lam_last = flt.invert(res.D[-1], side) if have else None
            lam_k = flt.invert(res.D[kq - 1], side) if have else None
            cut_old = flt.b if (flt.two_sided or side > 0) else -flt.a
            f2 = None
            if replace_filter and lam_last is not None and lam_k is not None:
                cut_new = abs(lam_last)
                # (the cut must stay below the estimate of the last wanted eigenvalue - itself a lower bound of it)
                if cut_new > cut_old + 1e-2 * max(norm_a - cut_old, 0.0) and cut_new < abs(lam_k):
                    f2 = ChebFilter(degree=want_degree, a=flt.a, b=flt.b, rho=1.0, two_sided=flt.two_sided)
                    if f2.two_sided:
                        f2.a, f2.b = -cut_new, cut_new
                    elif side > 0:
                        f2.b = cut_new
                    else:
                        f2.a = -cut_new
            if f2 is None and not settled:
                settled = True                      # the cut has stopped moving: the next cycle gets the whole buffer
                f2 = flt
            elif f2 is None:
                want_degree = min(2 * flt.degree, MAX_FILTER_DEGREE)
                f2 = ChebFilter(degree=want_degree, a=flt.a, b=flt.b, rho=1.0, two_sided=flt.two_sided)
            if f2 is not flt:
                f2.degree = cap_degree(f2, lam1_est, f2.degree)      # (the requested degree, as far as the dynamic range allows)
                if f2.degree == flt.degree and f2.a == flt.a and f2.b == flt.b:
                    final = res                   # neither the cut nor the degree can move any more: best effort
                    break
                if lam_k is None:
                    lam_k = f2.c + f2.e * (1.0 + 1e-3) * (1.0 if side >= 0 else -1.0)
                f2.rho = 1.0
                tk = abs(float(f2.scalar(lam_k)))
                xk = abs((lam_k - f2.c) / f2.e)
                f2.rho = norm_a / tk if tk > 0 and xk > 1.0 else 1.0
                flt = f2
                st.filter = flt
            rest = [j for j in range(have) if not (j < k_rem and res.bounds[j] <= tol)][:b]
            start = Qm @ res.S[:, rest]
            if start.shape[1] < b:
                start = np.hstack([start, np.random.default_rng(st.cycles).standard_normal((n, b - start.shape[1]))])
            start = np.linalg.qr(start)[0]
            continue
        # lock every wanted pair whose bound passed (restarted.jl:122-131), restart from the best b others (:133-135)
        nb = res.iterations
        Qm = np.hstack(res.Q[:nb])
        lock = [j for j in range(min(k_rem, len(res.D))) if res.bounds[j] <= tol]
        rest = [j for j in range(len(res.D)) if j not in lock][:b]
        if lock:
            Y = np.hstack([Y, Qm @ res.S[:, lock]])
            Dlock = np.concatenate([Dlock, res.D[lock]])
        start = Qm @ res.S[:, rest]
        if start.shape[1] < b:
            start = np.hstack([start, np.random.default_rng(st.cycles).standard_normal((n, b - start.shape[1]))])
        _orth_against(Y, start)
        start = np.linalg.qr(start)[0]
        st.locked = Y.shape[1]
        if Y.shape[1] >= k:
            break
    # assemble: locked + the leading pairs of the last cycle
    if final is not None:
        k_rem = k - Y.shape[1]
        Qm = np.hstack(final.Q[:final.iterations])
        Vf = Qm @ final.S[:, :k_rem]
        V = np.hstack([Y, Vf])
    else:
        V = Y[:, :k]
    # eigenvalues of A: Rayleigh quotients (identity for the plain operator up to roundoff), then order by |lambda|
    W = np.asarray(A @ V, dtype=DOUBLE)
    lam = np.einsum("ij,ij->j", V, W) / np.einsum("ij,ij->j", V, V)
    order = np.argsort(-np.abs(lam), kind="stable")
    lam, V, W = lam[order], V[:, order], W[:, order]
    st.max_residual = float(np.max(np.linalg.norm(W - V * lam[None, :], axis=0)))
    if return_details:
        return lam, V, st
    return lam, V

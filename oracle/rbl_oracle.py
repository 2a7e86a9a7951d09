"""CPU restatement of the reference's randomized block Lanczos (RBL) path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / reported CPU baseline.  The product path is the C-ABI library
``librbl_b200.so`` (hand-written sm_100a kernels) and fails loudly when it is missing.

What is restated (all citations relative to the reference tree, ``Julia/``):

* ``common.jl:9-17``   ``insertA!``          -> :func:`insert_a`
* ``common.jl:20-26``  ``insertB!``          -> :func:`insert_b`
* ``common.jl:36-48``  ``dsbev``             -> :func:`dsbev` (LAPACK ``dsbev`` through SciPy, 'V','L')
* ``common.jl:50-54``  ``sort_eig_abs``      -> :func:`sort_eig_abs`
* ``common.jl:56-65``  ``check_convergence`` -> :func:`check_convergence`
* ``RBL.jl:4-13``      ``loc_reorth!``       -> :func:`loc_reorth` (EFFECTIVE semantics, see below)
* ``RBL.jl:30-48``     ``part_reorth!(Q)``   -> :func:`part_reorth`
* ``RBL.jl:61-71``     ``recover_eigvec``    -> :func:`recover_eigvec`
* ``RBL.jl:74-117``    ``lanczos_iteration`` -> :func:`lanczos_iteration`
* ``RBL.jl:119-142``   ``RBL``               -> :func:`RBL`

The reference cannot be executed here (no Julia toolchain in the image, none on the GPU boxes), so
this restatement is pinned against the ONLY results the reference's own tests hold for the path:
the 15 known-answer eigenvalue cases of ``Julia/Unit Testing/{slow,mod,step}_dec.jl`` through
``test.jl:10-50`` (bar: ``norm((d - eig) ./ eig) < 1e-13``); see ``tests/test_oracle_kat.py``.
The reference's GPU path (``RBL_gpu.jl``) has no tests at all ("parity unpinned" at that boundary);
its arithmetic is the same algorithm on CUDA library calls, restated by the same functions.

Effective semantics that differ from a naive reading of the source:

* ``loc_reorth!`` (``RBL.jl:4-13``) loops 2b times over (project, QR) but rebinds the local name
  ``U1`` to a fresh matrix after the first projection (``U1 = Matrix(qr(U1).Q)`` - ``qr`` copies), so
  only the first projection ``U1 -= U2 (U2' U1)`` reaches the caller's array; the closing
  ``U1[:,:] = U1`` assigns the local to itself.  ``loc_reorth(..., literal_dead_work=True)`` also
  performs the discarded work so that the CPU baseline can be timed like the reference.
* ``push!(Q,Qi)`` stores a reference, so ``Q[i] === Qi`` inside the loop (``RBL.jl:92,101``).
* ``U::Matrix{DOUBLE}`` is a typed local: ``U = Matrix{FLOAT}(U)`` rounds through FLOAT but is stored
  as DOUBLE (``RBL.jl:80,83,100``).  With the shipped ``FLOAT = Float64`` (``common.jl:5``) it is a no-op.
* ``insertB!`` runs AFTER the convergence check (``RBL.jl:106-113``), so the trailing band of the last
  block column is zero when ``dsbev`` runs.
* The random start block is never seeded in the reference (``RBL.jl:136``); here the caller passes
  ``Omega`` explicitly so the device path and the oracle start from the same block.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
from scipy.linalg import lapack as _lapack

DOUBLE = np.float64

# Phase labels of the reference's TimerOutputs calls (RBL.jl:80-107,140).
PHASES = ("A*Q", "3-term", "QR", "Part reorth", "Loc reorth", "eig", "Ritz vectors")


class NotConverged(RuntimeError):
    """Raised where the reference would return a stale check or throw a BoundsError."""


@dataclass
class OracleStats:
    iterations: int = 0
    kryl_sz: int = 0
    converged: bool = False
    seconds: dict = field(default_factory=lambda: {p: 0.0 for p in PHASES})
    checks: list = field(default_factory=list)  # (iteration, converged?) for every dsbev call


class _Timer:
    def __init__(self, stats, label):
        self.stats, self.label = stats, label

    def __enter__(self):
        self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        self.stats.seconds[self.label] += time.perf_counter() - self.t0


# --------------------------------------------------------------------------- common.jl
def insert_a(Ai: np.ndarray, b: int) -> np.ndarray:
    """common.jl:9-17 - lower triangle of the b x b block into (b+1) x b LAPACK lower-band columns."""
    T = np.zeros((b + 1, b), dtype=DOUBLE)
    n = Ai.shape[1]
    for j in range(n):
        size = n - j
        T[0:size, j] = Ai[j:n, j]
    return T


def insert_b(Bi: np.ndarray, T: np.ndarray, b: int, it: int) -> None:
    """common.jl:20-26 - upper-triangular B into the sub-diagonal band rows of block column `it` (1-based)."""
    n = Bi.shape[1]
    start = (it - 1) * b
    rows = T.shape[0]
    for j in range(1, n + 1):
        T[rows - j:rows, start + j - 1] = Bi[0:j, j - 1]


def dsbev(T: np.ndarray):
    """common.jl:36-48 - LAPACK dsbev('V','L') on a copy of the band matrix; `info` is ignored there."""
    w, z, _info = _lapack.dsbev(np.asfortranarray(T), compute_v=1, lower=1)
    return w, z


def sort_eig_abs(D: np.ndarray, V: np.ndarray, k: int):
    """common.jl:50-54 - the k eigenvalues of largest magnitude, ascending |lambda| (stable sort)."""
    perm = np.argsort(np.abs(D), kind="stable")
    perm_k = perm[len(perm) - k:]
    return D[perm_k], V[:, perm_k]


def check_convergence(Bi: np.ndarray, V: np.ndarray, b: int, k: int, tol: float) -> bool:
    """common.jl:56-65 - every column of B * V[end-b+1:end, :] must have 2-norm <= tol (absolute)."""
    Y = Bi @ V[V.shape[0] - b:, :]
    for i in range(k):
        if np.linalg.norm(Y[:, i]) > tol:
            return False
    return True


def residual_bounds(Bi: np.ndarray, V: np.ndarray, b: int) -> np.ndarray:
    Y = Bi @ V[V.shape[0] - b:, :]
    return np.linalg.norm(Y, axis=0)


# --------------------------------------------------------------------------- RBL.jl
def _thin_qr(U: np.ndarray):
    """Julia's qr(U) is LAPACK Householder QR; Matrix(F.Q) is the thin factor (RBL.jl:84-86,102-104)."""
    return np.linalg.qr(U, mode="reduced")


def loc_reorth(U1: np.ndarray, U2: np.ndarray, literal_dead_work: bool = False) -> None:
    """RBL.jl:4-13 - effective semantics: U1 -= U2 (U2' U1), once, in place, no renormalisation."""
    temp = U2.T @ U1
    U1 -= U2 @ temp
    if literal_dead_work:
        p = U1.shape[1]
        W = _thin_qr(U1)[0]
        for _ in range(2 * p - 1):
            temp = U2.T @ W
            W = W - U2 @ temp
            W = _thin_qr(W)[0]


def part_reorth(Q: list) -> None:
    """RBL.jl:30-48 - project the two newest blocks against blocks 1..i-2, block by block, in place."""
    i = len(Q)
    for j in range(i - 2):
        Uj = Q[j]
        Q[i - 1] -= Uj @ (Uj.T @ Q[i - 1])
        Q[i - 2] -= Uj @ (Uj.T @ Q[i - 2])


def recover_eigvec(Q: list, V_trunc: np.ndarray, k: int) -> np.ndarray:
    """RBL.jl:61-71 - V = sum_i Q[i] * V_trunc[(i-1)b+1:ib, :]."""
    n, b = Q[0].shape
    V = np.zeros((n, k), dtype=Q[0].dtype)
    for i, Qi in enumerate(Q):
        V += Qi @ V_trunc[i * b:(i + 1) * b, :].astype(Qi.dtype, copy=False)
    return V


def lanczos_iteration(A, k: int, b: int, kryl_sz: int, Qi: np.ndarray, Q: list, *, tol: float = 1e-7,
                      FLOAT=np.float64, reorth_period: int = 2, check_period: int = 4,
                      literal_dead_work: bool = False, stats: OracleStats | None = None,
                      max_iterations: int | None = None):
    """RBL.jl:74-117.  `max_iterations` truncates the loop (bounded CPU-baseline samples only)."""
    st = stats if stats is not None else OracleStats()
    D = None
    V = None
    Q.append(Qi)                                               # :79
    with _Timer(st, "A*Q"):
        U = np.asarray(A @ Qi, dtype=DOUBLE)                   # :80
    with _Timer(st, "3-term"):
        Ai = (Qi.T @ U).astype(DOUBLE)                         # :81
        U -= Qi @ Ai                                           # :82
    U = U.astype(FLOAT).astype(DOUBLE)                         # :83 (typed local, Q14)
    with _Timer(st, "QR"):
        Qn, R = _thin_qr(U)                                    # :84
    Qi = Qn.astype(FLOAT)                                      # :85
    Bi = R.astype(DOUBLE)                                      # :86
    T = insert_a(Ai, b)                                        # :87
    insert_b(Bi, T, b, 1)                                      # :88
    i = 1
    last_check_i = 0
    while i * b < kryl_sz:                                     # :90
        if max_iterations is not None and i >= max_iterations:
            break
        i += 1
        Q.append(Qi)                                           # :92 (by reference)
        if i % reorth_period == 0:                             # :93
            with _Timer(st, "Part reorth"):
                part_reorth(Q)
        with _Timer(st, "Loc reorth"):
            loc_reorth(Q[i - 1], Q[i - 2], literal_dead_work)  # :96
        with _Timer(st, "A*Q"):
            U = np.asarray(A @ Q[i - 1], dtype=DOUBLE)         # :97
        with _Timer(st, "3-term"):
            U -= Q[i - 2] @ Bi.T                               # :98
            Ai = (Q[i - 1].T @ U).astype(DOUBLE)               # :99
        U = U.astype(FLOAT).astype(DOUBLE)                     # :100
        with _Timer(st, "3-term"):
            U -= Qi @ Ai                                       # :101 (Qi === Q[i])
        with _Timer(st, "QR"):
            Qn, R = _thin_qr(U)                                # :102
        Qi = Qn.astype(FLOAT)                                  # :103
        Bi = R.astype(DOUBLE)                                  # :104
        T = np.hstack([T, insert_a(Ai, b)])                    # :105
        if (i * b > k) and (i % check_period == 0):            # :106
            with _Timer(st, "eig"):
                Dall, Vall = dsbev(T)                          # :107
            D, V = sort_eig_abs(Dall, Vall, k)                 # :108
            last_check_i = i
            ok = check_convergence(Bi, V, b, k, tol)           # :109
            st.checks.append((i, ok))
            if ok:
                st.converged = True
                break                                          # :110
        insert_b(Bi, T, b, i)                                  # :113
    st.iterations = i
    st.kryl_sz = len(Q) * b                                    # :115
    if max_iterations is not None and not st.converged:
        # truncated CPU-baseline sample: no result is claimed, only the phase timings are used
        return np.zeros(k), np.zeros((i * b, k)), T
    if D is None:
        raise NotConverged("no convergence check ran (k >= cap?) - reference indexes a 0-dim D (Q4)")
    if last_check_i != i:
        raise NotConverged("cap reached between checks - reference throws BoundsError in recover_eigvec (Q4)")
    return D[::-1].copy(), V[:, ::-1].copy(), T            # :116


def RBL(A, k: int, b: int, Omega: np.ndarray | None = None, *, max_kryl_sz: int = 1400, tol: float = 1e-7,
        FLOAT=np.float64, seed: int | None = None, reorth_period: int = 2, check_period: int = 4,
        literal_dead_work: bool = False, return_details: bool = False, max_iterations: int | None = None):
    """RBL.jl:119-142.  Returns (D, V): k eigenvalues by descending |lambda| and the n x k Ritz vectors.

    `Omega` replaces the unseeded ``randn(DOUBLE,n,b)`` of RBL.jl:136 (same block is handed to the device
    path).  ``max_kryl_sz`` defaults to the CPU cap of RBL.jl:133 (the GPU path uses 1200, RBL_gpu.jl:211).
    """
    n = A.shape[1]
    if Omega is None:
        Omega = np.random.default_rng(seed).standard_normal((n, b))
    Omega = np.asarray(Omega, dtype=DOUBLE)
    assert Omega.shape == (n, b)
    st = OracleStats()
    Q: list = []
    Qi = _thin_qr(np.asarray(A @ Omega, dtype=DOUBLE))[0].astype(FLOAT)   # :137
    D, S, T = lanczos_iteration(A, k, b, max_kryl_sz, Qi, Q, tol=tol, FLOAT=FLOAT, reorth_period=reorth_period,
                                check_period=check_period, literal_dead_work=literal_dead_work, stats=st,
                                max_iterations=max_iterations)
    with _Timer(st, "Ritz vectors"):
        nb = S.shape[0] // b
        V = recover_eigvec(Q[:nb], S.astype(FLOAT), k)                      # :140
    if return_details:
        return D, V, {"stats": st, "Q": Q, "S": S, "T": T}
    return D, V


# --------------------------------------------------------------------------- metrics used by the parity tests
def ritz_residuals(A, D: np.ndarray, V: np.ndarray, norm_a: float | None = None) -> np.ndarray:
    """||A v - lambda v||_2 / ||A||_2 per Ritz pair (||A||_2 estimated by the largest |lambda| if not given)."""
    V = np.asarray(V, dtype=DOUBLE)
    R = A @ V - V * D[None, :]
    na = norm_a if norm_a is not None else float(np.max(np.abs(D)))
    return np.linalg.norm(R, axis=0) / na


def orthogonality_loss(blocks) -> float:
    """||Q'Q - I||_2 of the concatenated Krylov basis."""
    Qall = np.hstack([np.asarray(q, dtype=DOUBLE) for q in blocks])
    G = Qall.T @ Qall
    G[np.diag_indices_from(G)] -= 1.0
    return float(np.linalg.norm(G, 2))


def dense_band_from_T(T: np.ndarray) -> np.ndarray:
    """Expand LAPACK lower-band storage ((b+1) x N) to a dense symmetric N x N matrix (tests only)."""
    kd = T.shape[0] - 1
    N = T.shape[1]
    M = np.zeros((N, N))
    for c in range(N):
        for r in range(kd + 1):
            if c + r < N:
                M[c + r, c] = T[r, c]
                M[c, c + r] = T[r, c]
    return M


def as_csr(A) -> sp.csr_matrix:
    return sp.csr_matrix(A)

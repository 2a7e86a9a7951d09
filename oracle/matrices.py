"""Synthetic matrices for the parity tests and the bench (TEST INFRASTRUCTURE, see rbl_oracle.py header).

* the three known-answer generators of the reference's own tests, ``Julia/Unit Testing/test.jl:17-50``
  (diagonal matrices with slow / moderate / step eigenvalue decay);
* the BASELINE.json configs (2-D 5-point and 3-D 7-point Dirichlet Laplacians, symmetric Erdos-Renyi,
  8-neighbour image-grid graph Laplacian), none of which exist in the reference.

The reference returns the k eigenvalues of LARGEST magnitude only (``common.jl:50-54``); "lowest"
eigenpairs of a Laplacian L are therefore computed as the largest of ``sigma*I - L`` on both sides of
every parity check (sigma = 8 in 2-D, 12 in 3-D >= lambda_max).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# ---- reference fixtures (Unit Testing/test.jl) -------------------------------------------------
def slow_decay(n: int, k: int):
    """test.jl:31-37 - diag = 1..n; wanted eigenvalues n, n-1, ..., n-k+1."""
    a = np.arange(1.0, n + 1.0)
    return sp.diags(a, format="csc"), a[::-1][:k].copy()


def moderate_decay(n: int, k: int):
    """test.jl:17-28 - diag = triangular numbers i(i+1)/2."""
    a = np.cumsum(np.arange(1.0, n + 1.0))
    return sp.diags(a, format="csc"), a[::-1][:k].copy()


def step_decay(n: int, k: int):
    """test.jl:40-50 - first 2k entries (2k)n, (2k-1)n, ..., n, the rest 1; wanted: the first k."""
    a = np.ones(n)
    sz = 2 * k
    for i in range(1, sz + 1):
        a[sz - i] = float(i) * n
    return sp.diags(a, format="csc"), a[:k].copy()


# ---- BASELINE configs ---------------------------------------------------------------------------
def _tridiag(N: int):
    return sp.diags([-np.ones(N - 1), 2.0 * np.ones(N), -np.ones(N - 1)], [-1, 0, 1], format="csr")


def laplacian_2d(N: int):
    """5-point Dirichlet Laplacian on an N x N grid: kron(I,T) + kron(T,I); n = N^2."""
    T = _tridiag(N)
    I = sp.identity(N, format="csr")
    return (sp.kron(I, T) + sp.kron(T, I)).tocsr()


def laplacian_3d(N: int):
    """7-point Dirichlet Laplacian on an N^3 grid; n = N^3, nnz = 7N^3 - 6N^2."""
    T = _tridiag(N)
    I = sp.identity(N, format="csr")
    return (sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I)).tocsr()


def laplacian_eigs(N: int, dim: int, k: int):
    """The k smallest exact eigenvalues of the Dirichlet Laplacian: sum_d 4 sin^2(i_d pi / (2(N+1)))."""
    lam1 = 4.0 * np.sin(np.arange(1, N + 1) * np.pi / (2.0 * (N + 1))) ** 2
    m = min(N, max(8, int(np.ceil(k ** (1.0 / dim))) * 3 + 4))
    l = lam1[:m]
    if dim == 2:
        allv = (l[:, None] + l[None, :]).ravel()
    else:
        allv = (l[:, None, None] + l[None, :, None] + l[None, None, :]).ravel()
    return np.sort(allv)[:k]


def shifted(A, sigma: float):
    """sigma*I - A (largest-|lambda| of this = lowest of A when sigma >= lambda_max)."""
    n = A.shape[0]
    return (sigma * sp.identity(n, format="csr") - A).tocsr()


def erdos_renyi_sym(n: int, nnz_per_row: int = 32, seed: int = 0):
    """Symmetric ER: n*nnz_per_row/2 (i,j) pairs, N(0,1) weights, U = triu(.,1), A = U + U'."""
    rng = np.random.default_rng(seed)
    m = n * nnz_per_row // 2
    i = rng.integers(0, n, m)
    j = rng.integers(0, n, m)
    w = rng.standard_normal(m)
    U = sp.triu(sp.coo_matrix((w, (i, j)), shape=(n, n)).tocsr(), k=1)
    return (U + U.T).tocsr()


def image_graph_laplacian(H: int, W: int, seed: int = 0, sigma2: float = 0.05):
    """8-neighbour graph Laplacian L = D - W of a synthetic H x W image, w = exp(-(dI)^2/sigma2)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = 0.5 + 0.25 * np.sin(2 * np.pi * xx / max(W, 1) * 3) * np.cos(2 * np.pi * yy / max(H, 1) * 2)
    img += 0.25 * ((xx > W // 2) ^ (yy > H // 3))
    img += 0.02 * rng.standard_normal((H, W))
    idx = (yy * W + xx)
    rows, cols, vals = [], [], []
    for dy, dx in ((0, 1), (1, 0), (1, 1), (1, -1)):
        ys = slice(0, H - dy)
        if dx >= 0:
            a = idx[ys, 0:W - dx]; bq = idx[dy:H, dx:W]
            ia = img[ys, 0:W - dx]; ib = img[dy:H, dx:W]
        else:
            a = idx[ys, -dx:W]; bq = idx[dy:H, 0:W + dx]
            ia = img[ys, -dx:W]; ib = img[dy:H, 0:W + dx]
        w = np.exp(-((ia - ib) ** 2) / sigma2).ravel()
        rows += [a.ravel(), bq.ravel()]
        cols += [bq.ravel(), a.ravel()]
        vals += [w, w]
    Wm = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                       shape=(H * W, H * W)).tocsr()
    d = np.asarray(Wm.sum(axis=1)).ravel()
    return (sp.diags(d) - Wm).tocsr()

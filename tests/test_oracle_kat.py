"""Pins the oracle against the ONLY results the reference's own tests hold for this path:
Julia/Unit Testing/{slow,mod,step}_dec.jl:3-6 through test.jl:10-50 - 15 known-answer cases, k = b = 5,
bar  norm((d - eig) ./ eig) < 1e-13  (CPU `RBL`, unseeded start block)."""
import numpy as np
import pytest

from oracle import matrices, rbl_oracle

BAR = 1e-13  # slow_dec.jl:5, mod_dec.jl:5, step_dec.jl:5


@pytest.mark.parametrize("n", [100, 300, 500, 700, 900])       # slow_dec.jl:4  `for i in 100:200:1000`
def test_slow_decay(n):
    A, eig = matrices.slow_decay(n, 5)
    d, _ = rbl_oracle.RBL(A, 5, 5, seed=n)
    assert np.linalg.norm((d - eig) / eig) < BAR


@pytest.mark.parametrize("n", [100, 300, 500, 700, 900])       # mod_dec.jl:4
def test_moderate_decay(n):
    A, eig = matrices.moderate_decay(n, 5)
    d, _ = rbl_oracle.RBL(A, 5, 5, seed=n + 1)
    assert np.linalg.norm((d - eig) / eig) < BAR


@pytest.mark.parametrize("n", [100000, 300000, 500000, 700000, 900000])   # step_dec.jl:4
def test_step_decay(n):
    A, eig = matrices.step_decay(n, 5)
    d, _ = rbl_oracle.RBL(A, 5, 5, seed=n + 2)
    assert np.linalg.norm((d - eig) / eig) < BAR


def test_insert_a_b_band_layout():
    """common.jl:9-26: T[(i)b+m, (i-1)b+j] = B[m,j] (m <= j) and the lower triangle of A on the diagonal block."""
    b = 3
    rng = np.random.default_rng(0)
    A1 = rng.standard_normal((b, b)); A1 = A1 + A1.T
    A2 = rng.standard_normal((b, b)); A2 = A2 + A2.T
    B1 = np.triu(rng.standard_normal((b, b)))
    T = rbl_oracle.insert_a(A1, b)
    rbl_oracle.insert_b(B1, T, b, 1)
    T = np.hstack([T, rbl_oracle.insert_a(A2, b)])
    M = rbl_oracle.dense_band_from_T(T)
    ref = np.block([[A1, B1.T], [B1, A2]])
    assert np.allclose(M, ref)


def test_oracle_residuals_config1_small():
    """The oracle's Ritz pairs satisfy the north-star bars on a reduced config-1 (2-D Laplacian, lowest via 8I-A)."""
    N, k, b = 30, 6, 4
    L = matrices.laplacian_2d(N)
    A = matrices.shifted(L, 8.0)
    D, V, det = rbl_oracle.RBL(A, k, b, seed=3, return_details=True)
    exact = 8.0 - matrices.laplacian_eigs(N, 2, k)
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=8.0)) < 1e-6
    nb = det["S"].shape[0] // b
    assert rbl_oracle.orthogonality_loss(det["Q"][:nb]) < 1e-10

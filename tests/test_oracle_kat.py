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


# ---- the restart / filter twin (oracle/rbl_restart_oracle.py) pinned against the analytic spectrum -----------------------
def test_restart_oracle_filtered_restarts_do_not_return_ghost_pairs():
    """Regression for the filtered restart: with locking behind a re-placed high-degree filter the twin (and the device
    solver, on BASELINE config 5) returned ghost copies of locked eigenvalues - 'converged' pairs with residuals of order
    one.  Filtered restarts now re-place the filter / raise its degree and never lock; the dynamic range of p is capped."""
    from oracle import rbl_restart_oracle as rr
    N, k, b = 24, 40, 16
    A = matrices.shifted(matrices.laplacian_3d(N), 12.0)
    Om = np.random.default_rng(1).standard_normal((N ** 3, b))
    lam, V, st = rr.RBL_restarted(A, k, b, Om, max_blocks=16, filter_degree=16, return_details=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and st.locked == 0
    assert np.max(np.abs(lam - exact) / exact) < 1e-12
    assert st.max_residual < 1e-7
    assert np.linalg.norm(V.T @ V - np.eye(k), 2) < 1e-10
    f = st.filter
    assert f.b < exact[-1]                                   # every wanted eigenvalue stays outside the damped interval
    x1 = (exact[0] - f.c) / f.e
    assert np.cosh(f.degree * np.arccosh(x1)) <= 1.01 * rr.MAX_DYNAMIC_RANGE


def test_chebyshev_filter_inverse_map_rejects_the_wrong_side():
    from oracle import rbl_restart_oracle as rr
    f = rr.ChebFilter(degree=8, a=0.0, b=10.0, rho=2.0)
    lam = 11.3
    th = float(f.scalar(lam))
    assert abs(f.invert(th, 1) - lam) < 1e-12
    assert f.invert(-th, 1) is None and f.invert(1.5, 1) is None      # other side / inside the damped interval
    g = rr.ChebFilter(degree=7, a=-10.0, b=0.0, rho=1.0)
    thn = float(g.scalar(-11.0))
    assert thn < 0 and abs(g.invert(thn, -1) + 11.0) < 1e-12 and g.invert(-thn, -1) is None

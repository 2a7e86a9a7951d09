"""N1: restart with locking (restarted.jl:23-146 generalised) and the Chebyshev-filtered operator,
device path vs the CPU twin oracle/rbl_restart_oracle.py on the same seeded inputs.

Bars: eigenvalues rel <= 1e-8 vs the twin and vs a dense / analytic reference, true residuals <= 1e-6 ||A||,
orthonormal V, and comparable work (block steps / restart cycles) since both sides run the same algorithm."""
import numpy as np
import pytest

from oracle import matrices, rbl_restart_oracle as rr

pytestmark = pytest.mark.gpu


def _resid(A, D, V, norm_a):
    return float(np.max(np.linalg.norm(A @ V - V * D[None, :], axis=0)) / norm_a)


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
@pytest.mark.parametrize("degree", [4, 8])
def test_filtered_operator_lap3d(gpu, precision, degree):
    N, k, b = 24, 30, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    n = N ** 3
    Om = np.random.default_rng(2).standard_normal((n, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=3200, filter_degree=degree,
                           return_stats=True)
    Dt, Vt, tw = rr.RBL_restarted(A, k, b, Om, max_blocks=200, filter_degree=degree, restart=False, return_details=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and tw.converged
    assert st.filter_degree == tw.filter.degree and st.filter_two_sided == 0
    assert abs(st.filter_cut - tw.filter.b) <= 1e-6 * abs(tw.filter.b)           # same probe -> same damped interval
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(np.abs(D - Dt) / np.abs(Dt)) < 1e-8
    assert np.all(np.abs(D)[:-1] >= np.abs(D)[1:])
    assert _resid(A, D, V, 12.0) < 1e-6
    assert st.max_residual / 12.0 < 1e-6 and abs(st.max_residual / 12.0 - _resid(A, D, V, 12.0)) < 1e-7
    assert abs(st.iterations - tw.block_steps) <= 8
    assert st.iterations < 60                                                     # plain operator: 88 block steps
    G = V.T @ V
    assert np.max(np.abs(G - np.eye(k))) < (1e-5 if precision == "mixed" else 1e-9)


def test_filtered_two_sided_er(gpu):
    """Largest |lambda| on BOTH ends of the spectrum: odd filter on [-cut, cut]."""
    n, k, b = 6000, 12, 32
    A = matrices.erdos_renyi_sym(n, 32, seed=3)
    Om = np.random.default_rng(3).standard_normal((n, b))
    D, V, st = gpu.RBL_gpu(A, k, b, Omega=Om, max_kryl_sz=6400, filter_degree=4, return_stats=True)
    Dt, Vt, tw = rr.RBL_restarted(A, k, b, Om, max_blocks=200, filter_degree=4, return_details=True)
    w = np.linalg.eigvalsh(A.toarray())
    ref = w[np.argsort(-np.abs(w))][:k]
    assert st.converged and st.filter_two_sided == 1 and st.filter_degree == 5 == tw.filter.degree
    assert (D > 0).any() and (D < 0).any()
    assert np.max(np.abs(D - ref) / np.abs(ref)) < 1e-8
    assert np.max(np.abs(D - Dt) / np.abs(Dt)) < 1e-8
    assert _resid(A, D, V, np.max(np.abs(ref))) < 1e-6
    assert abs(st.iterations - tw.block_steps) <= 8


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
def test_restart_with_locking_when_the_cap_binds(gpu, precision):
    """40 blocks of room where the plain solve needs 88: lock what converged, restart from the best others."""
    N, k, b = 24, 30, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    Om = np.random.default_rng(2).standard_normal((N ** 3, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=40 * b, restart=True,
                           return_stats=True)
    Dt, Vt, tw = rr.RBL_restarted(A, k, b, Om, max_blocks=40, return_details=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and tw.converged
    assert st.restarts >= 1 and st.locked >= 1
    assert abs(st.restarts + 1 - tw.cycles) <= 2
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(np.abs(D - Dt) / np.abs(Dt)) < 1e-8
    assert _resid(A, D, V, 12.0) < 1e-6
    assert np.max(np.abs(V.T @ V - np.eye(k))) < (1e-5 if precision == "mixed" else 1e-8)
    # without restart the same cap ends in NOT_CONVERGED
    D2, V2, st2 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=40 * b, return_stats=True,
                              allow_not_converged=True)
    assert not st2.converged


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
def test_filtered_restart_replaces_the_filter_and_never_locks(gpu, precision):
    """The regime of BASELINE config 5 in small: the probe places the filter badly, the cap binds.  Settling cycles re-place
    the filter from their own Ritz values, nothing is locked (locking behind a re-placed high-degree filter returned ghost
    pairs), the dynamic range of p stays capped - device and CPU twin follow the same procedure."""
    N, k, b = 24, 40, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    Om = np.random.default_rng(1).standard_normal((N ** 3, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=16 * b, restart=True, filter_degree=16,
                           return_stats=True)
    Dt, Vt, tw = rr.RBL_restarted(A, k, b, Om, max_blocks=16, filter_degree=16, return_details=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and tw.converged
    assert st.locked == 0 and tw.locked == 0 and st.restarts >= 1
    assert abs(st.restarts + 1 - tw.cycles) <= 2
    assert st.filter_cut < exact[-1] and tw.filter.b < exact[-1]            # no wanted eigenvalue inside the damped interval
    assert abs(st.filter_cut - tw.filter.b) <= 2e-2 * (12.0 - tw.filter.b)
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(np.abs(D - Dt) / np.abs(Dt)) < 1e-8
    assert _resid(A, D, V, 12.0) < 1e-6 and st.max_residual / 12.0 < 1e-6
    assert np.max(np.abs(V.T @ V - np.eye(k))) < (1e-5 if precision == "mixed" else 1e-8)


def test_restart_driven_by_device_memory(gpu):
    """The cap comes from the memory plan (mem_limit_mb) instead of max_kryl_sz - the regime of configs 3 and 5."""
    N, k, b = 24, 30, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    Om = np.random.default_rng(2).standard_normal((N ** 3, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision="mixed", max_kryl_sz=100000, restart=True,
                           mem_limit_mb=64, return_stats=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and st.restarts >= 1
    assert st.buffer_blocks < 88
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert _resid(A, D, V, 12.0) < 1e-6


def test_restart_plus_filter_two_sided_small_block(gpu):
    n, k, b = 6000, 20, 8
    A = matrices.erdos_renyi_sym(n, 32, seed=3)
    Om = np.random.default_rng(3).standard_normal((n, b))
    D, V, st = gpu.RBL_gpu(A, k, b, Omega=Om, max_kryl_sz=30 * b, restart=True, filter_degree=5, return_stats=True)
    Dt, Vt, tw = rr.RBL_restarted(A, k, b, Om, max_blocks=30, filter_degree=5, return_details=True)
    w = np.linalg.eigvalsh(A.toarray())
    ref = w[np.argsort(-np.abs(w))][:k]
    assert st.converged and tw.converged
    assert np.max(np.abs(D - ref) / np.abs(ref)) < 1e-8
    assert np.max(np.abs(D - Dt) / np.abs(Dt)) < 1e-8
    assert _resid(A, D, V, np.max(np.abs(ref))) < 1e-6
    assert abs(st.restarts + 1 - tw.cycles) <= 2


def test_reference_fixture_with_restart(gpu):
    """The reference's own slow-decay fixture (Unit Testing/test.jl:31-37) through the restarted path, b = 1 like
    RBL_gpu_restarted (restarted.jl:98-146), bar 1e-13 like slow_dec.jl:5."""
    A, eig = matrices.slow_decay(500, 5)
    Om = np.random.default_rng(500).standard_normal((500, 1))
    d, V, st = gpu.RBL_gpu(A, 5, 1, Omega=Om, max_kryl_sz=40, restart=True, return_stats=True)
    assert st.converged and st.restarts >= 1
    assert np.linalg.norm((d - eig) / eig) < 1e-13

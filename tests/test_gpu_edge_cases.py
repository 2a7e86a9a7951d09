"""Edge cases of the device path (through the C ABI): the reference's own warm-up call, b = 1, k = 1, Krylov
space exhaustion (exact breakdown), tiny problems, ragged block sizes, and argument validation."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import matrices, rbl_oracle

pytestmark = pytest.mark.gpu


def test_reference_warmup_call(gpu):
    """benchmark.jl:58  `RBL_gpu(sprandn(DOUBLE,50,50,0.5),1,1)` (symmetrised: the solver requires A = A')."""
    rng = np.random.default_rng(0)
    M = sp.random(50, 50, density=0.5, random_state=1, data_rvs=rng.standard_normal).tocsr()
    A = (M + M.T).tocsr()
    d, V = gpu.RBL_gpu(A, 1, 1, Omega=rng.standard_normal((50, 1)))
    w = np.linalg.eigvalsh(A.toarray())
    lam = w[np.argmax(np.abs(w))]
    assert abs(d[0] - lam) < 1e-8 * abs(lam)
    assert np.linalg.norm(A @ V[:, 0] - d[0] * V[:, 0]) < 1e-6 * abs(lam)


@pytest.mark.parametrize("n,k,b", [(64, 1, 1), (200, 3, 2), (37, 2, 3), (500, 7, 7), (300, 4, 13), (1000, 20, 32)])
def test_small_and_ragged_shapes(gpu, n, k, b):
    rng = np.random.default_rng(n)
    A = matrices.erdos_renyi_sym(n, 8, seed=n) + sp.diags(np.linspace(1.0, 40.0, n))
    A = sp.csr_matrix(A)
    Om = rng.standard_normal((n, b))
    d, V, st = gpu.RBL_gpu(A, k, b, Omega=Om, max_kryl_sz=max(1200, 4 * n), return_stats=True, allow_not_converged=True)
    w = np.linalg.eigvalsh(A.toarray())
    ref = w[np.argsort(-np.abs(w))][:k]
    assert np.max(np.abs(d - ref) / np.abs(ref)) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, d, V)) < 1e-6


def test_krylov_space_exhaustion_is_deflated(gpu):
    """n = 24 with b = 4: after 6 blocks the Krylov space is the whole space and the residual block is exactly
    rank deficient - the block QR must deflate (no NaN), T decouples, and the Ritz values are the eigenvalues."""
    rng = np.random.default_rng(3)
    M = rng.standard_normal((24, 24))
    A = M + M.T
    d, V, st = gpu.RBL_gpu(A, 5, 4, Omega=rng.standard_normal((24, 4)), return_stats=True, allow_not_converged=True)
    w = np.linalg.eigvalsh(A)
    ref = w[np.argsort(-np.abs(w))][:5]
    assert np.all(np.isfinite(d)) and np.all(np.isfinite(V))
    assert np.max(np.abs(d - ref) / np.abs(ref)) < 1e-8
    assert st.deflated >= 1


def test_fewer_distinct_eigenvalues_than_block_columns(gpu):
    """step_dec.jl's situation in miniature: 3 distinct eigenvalues, b = 4."""
    a = np.r_[np.full(10, 5.0), np.full(10, -3.0), np.full(30, 1.0)]
    A = sp.diags(a, format="csr")
    d, V = gpu.RBL_gpu(A, 2, 4, Omega=np.random.default_rng(1).standard_normal((50, 4)))
    assert np.allclose(np.sort(np.abs(d))[::-1], [5.0, 5.0], rtol=1e-12) or np.allclose(np.abs(d), [5.0, 5.0], rtol=1e-12)


def test_argument_validation(gpu):
    A = matrices.laplacian_2d(6)
    with pytest.raises(gpu.RblError) as e:
        gpu.RBL_gpu(A, 0, 2)
    assert e.value.status == 4
    with pytest.raises(gpu.RblError):
        gpu.RBL_gpu(A, 2, 33)          # b > 32 is not supported by the block kernels
    with pytest.raises(gpu.RblError):
        gpu.RBL_gpu(A, 100, 2)         # k > n
    with pytest.raises(gpu.RblError):
        gpu.RBL_gpu(A, 2, 16, precision="mixed", Omega=np.zeros((36, 16)), max_kryl_sz=8)  # cap below k..: no pairs


def test_zero_start_block_is_reported(gpu):
    A = matrices.laplacian_2d(10)
    with pytest.raises(gpu.RblError):
        gpu.RBL_gpu(A, 2, 2, Omega=np.zeros((100, 2)), max_kryl_sz=64)


def test_repeated_solves_reuse_the_workspace(gpu):
    """Two different problems through the same process: the parked workspace must not leak state."""
    for n, b in ((400, 4), (150, 8), (900, 4)):
        A, eig = matrices.slow_decay(n, 5)
        d, V = gpu.RBL_gpu(A, 5, b, Omega=np.random.default_rng(n).standard_normal((n, b)))
        assert np.linalg.norm((d - eig) / eig) < 1e-12
    assert gpu.lib().rbl_release_cached_memory() == 0

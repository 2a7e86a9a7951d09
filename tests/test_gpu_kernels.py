"""Kernel-level parity on the GPU, through the C ABI, against NumPy/SciPy on the same seeded inputs."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import matrices

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b", [1, 2, 4, 5, 8, 16, 32])
def test_spmm_matches_scipy(gpu, b):
    A = matrices.laplacian_3d(14).tocsr()
    Q = np.random.default_rng(b).standard_normal((A.shape[0], b))
    U = gpu.k_spmm(A, Q)
    ref = A @ Q
    # summation in CSR order with fma: a few ulp of the row sum of |a||q|
    bound = 8 * np.finfo(float).eps * (abs(A) @ np.abs(Q))
    assert np.all(np.abs(U - ref) <= bound + 1e-300)


def test_spmm_shift_and_irregular_rows(gpu):
    A = matrices.erdos_renyi_sym(3000, 24, seed=1)
    A = A + sp.diags(np.arange(3000.0))          # also rows with many entries and an empty-ish structure
    A = sp.csr_matrix(A)
    A[17, :] = 0
    A[:, 17] = 0
    A.eliminate_zeros()                           # an empty row
    Q = np.random.default_rng(2).standard_normal((3000, 16))
    U = gpu.k_spmm(A, Q, op=1, sigma=3.5)
    ref = 3.5 * Q - A @ Q
    assert np.max(np.abs(U - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.array_equal(U[17], 3.5 * Q[17])


def test_spmm_dense_operator(gpu):
    rng = np.random.default_rng(3)
    M = rng.standard_normal((150, 150)); M = M + M.T
    Q = rng.standard_normal((150, 4))
    U = gpu.k_spmm(M, Q)
    assert np.max(np.abs(U - M @ Q)) < 1e-12 * np.max(np.abs(M @ Q))


@pytest.mark.parametrize("n,b", [(1, 4), (127, 3), (4096, 16), (50001, 16), (20000, 32), (9999, 8), (777, 1)])
def test_gram(gpu, n, b):
    rng = np.random.default_rng(n + b)
    X = rng.standard_normal((n, b)); Y = rng.standard_normal((n, b))
    C = gpu.k_gram(X, Y)
    ref = X.T @ Y
    assert np.max(np.abs(C - ref)) <= 1e-13 * n ** 0.5 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("n,b", [(5000, 16), (20011, 4), (3000, 5), (8192, 32), (1000, 1)])
def test_block_qr_well_conditioned(gpu, n, b):
    rng = np.random.default_rng(n)
    U = rng.standard_normal((n, b)) * (10.0 ** rng.uniform(-3, 3, b))[None, :]
    Q, R, d = gpu.k_block_qr(U)
    assert not d.any()
    assert np.allclose(np.triu(R), R)
    assert np.max(np.abs(Q.T @ Q - np.eye(b))) < 1e-13
    assert np.max(np.abs(Q @ R - U)) < 1e-12 * np.max(np.abs(U))
    # same factor as Householder QR up to column signs
    Rh = np.linalg.qr(U, mode="r")
    assert np.allclose(np.abs(np.diag(R)), np.abs(np.diag(Rh)), rtol=1e-10)


def test_block_qr_ill_conditioned_and_deflation(gpu):
    rng = np.random.default_rng(0)
    n, b = 6000, 8
    Qo = np.linalg.qr(rng.standard_normal((n, b)))[0]
    U = Qo @ np.diag(10.0 ** -np.arange(0, 16, 2.0)) @ np.linalg.qr(rng.standard_normal((b, b)))[0]  # cond 1e14
    Q, R, d = gpu.k_block_qr(U)
    keep = ~d.astype(bool)
    assert keep.sum() >= 6
    G = Q[:, keep].T @ Q[:, keep]
    assert np.max(np.abs(G - np.eye(keep.sum()))) < 1e-10
    assert np.max(np.abs(Q @ R - U)) < 1e-11
    # exact rank deficiency: duplicated and zero columns are deflated to exact zeros (SURVEY H4 / step_dec fixture)
    U2 = rng.standard_normal((n, 5))
    U2[:, 3] = U2[:, 1] * 2.0 - U2[:, 0]
    U2[:, 4] = 0.0
    Q2, R2, d2 = gpu.k_block_qr(U2)
    assert d2.tolist() == [0, 0, 0, 1, 1]
    assert np.all(Q2[:, 3] == 0) and np.all(Q2[:, 4] == 0)
    assert np.max(np.abs(Q2 @ R2 - U2)) < 1e-10 * np.max(np.abs(U2))
    assert np.max(np.abs(Q2[:, :3].T @ Q2[:, :3] - np.eye(3))) < 1e-13


def _krylov_like(n, b, m, rng):
    Qall = np.linalg.qr(rng.standard_normal((n, (m + 2) * b)))[0]
    blocks = Qall[:, :m * b].reshape(n, m, b).transpose(1, 0, 2).copy()
    W0 = Qall[:, m * b:(m + 1) * b] + 1e-3 * Qall[:, :b] @ rng.standard_normal((b, b))
    W1 = Qall[:, (m + 1) * b:] + 1e-3 * Qall[:, b:2 * b] @ rng.standard_normal((b, b))
    return blocks, W0.copy(), W1.copy()


@pytest.mark.parametrize("n,b,m,fp32", [(4000, 16, 5, False), (4000, 16, 5, True), (10007, 16, 70, True),
                                         (2500, 4, 300, True), (2500, 4, 30, False), (3001, 8, 40, True),
                                         (3001, 32, 12, True), (1500, 32, 9, False), (900, 5, 7, False), (64, 16, 1, True)])
def test_reorth_block_cgs(gpu, n, b, m, fp32):
    _reorth_case(gpu, n, b, m, fp32, impl=0)


@pytest.mark.parametrize("impl", [1, 3, 4])
@pytest.mark.parametrize("n,b,m", [(4000, 16, 5), (10007, 16, 70), (700, 16, 33), (70000, 16, 3), (5000, 13, 9), (64, 16, 1),
                                   (20000, 16, 130)])
def test_reorth_simt_and_tensor_core_paths(gpu, impl, n, b, m):
    """All implementations of K5 (SIMT fp32 FMA / scaled 2-term FP16 MMA on an fp32 buffer (3) and on the
    pre-split buffer format the solver uses (4)) meet the same fp32-grade bars."""
    _reorth_case(gpu, n, b, m, True, impl=impl)


@pytest.mark.parametrize("impl", [1, 3, 4])
@pytest.mark.parametrize("n,b,m", [(4000, 32, 5), (10007, 32, 37), (700, 32, 17), (70000, 32, 3), (5000, 27, 9), (128, 32, 1),
                                   (20000, 32, 66)])
def test_reorth_b32_simt_and_fp16_split_paths(gpu, impl, n, b, m):
    """Config 3's block size: the scaled 2-term FP16 MMA kernels (impl 3, the default) against the SIMT kernels."""
    _reorth_case(gpu, n, b, m, True, impl=impl)


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("n,b,m", [(4000, 16, 5), (10007, 16, 70), (700, 16, 33), (70000, 16, 3), (5000, 13, 9), (64, 16, 1)])
def test_reorth_fp64_simt_and_dmma_paths(gpu, impl, n, b, m):
    """All-fp64 mode (the reference as shipped): SIMT DFMA kernels (impl 1) and FP64 tensor-core kernels (impl 0)."""
    _reorth_case(gpu, n, b, m, False, impl=impl)


def _reorth_case(gpu, n, b, m, fp32, impl):
    rng = np.random.default_rng(n + m)
    blocks, W0, W1 = _krylov_like(n, b, m, rng)
    w0, w1, C = gpu.k_reorth(blocks, W0, W1, fp32, impl=impl)
    dt = np.float32 if fp32 else np.float64
    Qs = blocks.astype(dt).astype(np.float64)              # what the buffer holds
    Qm = Qs.transpose(1, 0, 2).reshape(n, m * b)
    W = np.hstack([W0, W1])
    Cref = Qm.T @ W
    tolC = (3e-6 if fp32 else 1e-13) * max(1.0, np.max(np.abs(Cref))) * 10
    assert np.max(np.abs(C.astype(np.float64) - Cref)) < tolC
    Wref = W - Qm @ Cref
    tolW = 2e-6 if fp32 else 1e-13
    assert np.max(np.abs(np.hstack([w0, w1]) - Wref)) < tolW
    # orthogonality against the stored blocks after the pass
    assert np.max(np.abs(Qm.T @ np.hstack([w0, w1]))) < (5e-6 if fp32 else 1e-12)


@pytest.mark.parametrize("n,b,m,k,fp32", [(3000, 16, 6, 100, True), (3000, 16, 6, 100, False), (5001, 4, 50, 10, True),
                                           (2000, 32, 5, 50, False), (1000, 8, 9, 64, True), (700, 5, 4, 7, False)])
def test_ritz(gpu, n, b, m, k, fp32):
    _ritz_case(gpu, n, b, m, k, fp32, 0)


@pytest.mark.parametrize("impl", [1, 4])
@pytest.mark.parametrize("n,b,m,k", [(3000, 16, 6, 100), (5003, 16, 41, 37), (2000, 32, 9, 64), (777, 13, 5, 3), (4100, 27, 20, 130),
                                     (300, 16, 1, 16)])
def test_ritz_simt_and_tensor_core_paths(gpu, impl, n, b, m, k):
    """K6 on an fp32 buffer: SIMT kernel (1) and the FP16-split tensor-core kernel on the pre-split format (4)."""
    _ritz_case(gpu, n, b, m, k, True, impl)


def _ritz_case(gpu, n, b, m, k, fp32, impl):
    rng = np.random.default_rng(k)
    blocks = rng.standard_normal((m, n, b))
    S = rng.standard_normal((m * b, k))
    V = gpu.k_ritz(blocks, S, fp32, impl=impl)
    dt = np.float32 if fp32 else np.float64
    ref = blocks.astype(dt).astype(np.float64).transpose(1, 0, 2).reshape(n, m * b) @ S.astype(dt).astype(np.float64)
    tol = (2e-5 if fp32 else 1e-12) * np.max(np.abs(ref))
    assert np.max(np.abs(V.astype(np.float64) - ref)) < tol


def _spmm_case(case):
    kw = {}
    if case.startswith("lap3d"):
        A = matrices.laplacian_3d(30).tocsr()
        if "shifted" in case:
            kw = dict(op=1, sigma=12.0)
    elif case == "lap2d-150":
        A = matrices.laplacian_2d(150).tocsr()
    elif case == "image-140":
        A = matrices.image_graph_laplacian(140, 140, seed=1).tocsr()
    else:
        n = 20000
        rng = np.random.default_rng(5)
        A = sp.diags([rng.standard_normal(n - abs(o)) for o in (-700, -3, -1, 0, 1, 3, 700)], (-700, -3, -1, 0, 1, 3, 700), format="lil")
        for _ in range(300):                      # stray entries far from the bands (plus long rows)
            i, j = rng.integers(0, n, 2)
            A[i, j] = rng.standard_normal()
        A[5, :40] = 1.0
        A = sp.csr_matrix(A)
    A.sort_indices()
    return A, kw


@pytest.mark.parametrize("b", [16, 13, 32, 4])
@pytest.mark.parametrize("case", ["lap3d-30", "lap2d-150", "lap3d-30 shifted", "banded+stray", "image-140"])
def test_spmm_structured_paths(gpu, case, b, monkeypatch):
    """K1 on stencil / banded matrices large enough for the structure planners: the gather kernel (default) and the TMA-staged
    window kernel (spmm.cu, RBL_SPMM_WINDOW=1, B = 16 / 32) against SciPy."""
    A, kw = _spmm_case(case)
    Q = np.random.default_rng(b).standard_normal((A.shape[0], b))
    ref = (12.0 * Q - A @ Q) if kw else A @ Q
    bound = 16 * np.finfo(float).eps * ((abs(A) @ np.abs(Q)) + (12.0 * np.abs(Q) if kw else 0))
    for path, env in (("gather", {}), ("window", {"RBL_SPMM_WINDOW": "1"})):
        monkeypatch.delenv("RBL_SPMM_WINDOW", raising=False)
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        U = gpu.k_spmm(A, Q, **kw)
        assert np.all(np.abs(U - ref) <= bound + 1e-300), path


@pytest.mark.parametrize("case", ["lap3d-30", "image-140"])
def test_spmm_laboratory_variants_are_bit_identical(gpu, case, monkeypatch):
    """csrc/spmm_lab.cu: the candidate K1 kernels (software-pipelined CSR, ELL, both also in patch-schedule order) repeat the
    product kernel's per-row arithmetic: every element equal bit for bit, with and without the Chebyshev Z term."""
    import ctypes as C
    from rbl_b200 import binding as B
    A, _ = _spmm_case(case)
    monkeypatch.setenv("RBL_SPMM_SCHED", "1")
    with B.Solver(A, options=B.default_options(precision=B.PRECISION_MIXED)) as s:
        for with_z in (0, 1):
            for variant in (0, 1, 1 | 256, 1 | 512, 2, 2 | 256, 16 | 1, 16 | 2, 16 | 2 | 256):
                us, bad = C.c_double(), C.c_int64(-1)
                rc = gpu.lib().rbl_spmm_bench(s._h, 16, variant, 16, 1, 0, with_z, C.byref(us), C.byref(bad))
                assert rc == 0, (variant, gpu.lib().rbl_last_error())
                assert bad.value == 0 and us.value > 0, (variant, with_z, bad.value)

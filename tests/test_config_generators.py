"""The fast matrix builders of tools/run_config.py (used for BASELINE configs 3-5 at full size) produce exactly the
matrices of oracle/matrices.py."""
import numpy as np
import scipy.sparse as sp

from oracle import matrices
from tools import run_config


def _same(A, B, tol=0.0):
    A = sp.csr_matrix(A); B = sp.csr_matrix(B)
    A.sort_indices(); B.sort_indices()
    A.eliminate_zeros(); B.eliminate_zeros()
    return A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and \
        np.max(np.abs(A.data - B.data), initial=0.0) <= tol


def test_er_fast_equals_oracle_generator():
    for n, seed in ((500, 3), (5000, 7)):
        assert _same(run_config.er_sym_fast(n, 32, seed=seed), matrices.erdos_renyi_sym(n, 32, seed=seed), tol=1e-15)


def test_laplacian_rows_equal_oracle_generator():
    N = 7
    L = matrices.laplacian_3d(N).tocsr()
    L.sort_indices()
    for r0, r1 in ((0, N ** 3), (50, 200), (300, 343)):
        rp, ci, va = run_config.laplacian_3d_rows(N, r0, r1)
        sub = L[r0:r1]
        assert np.array_equal(rp, sub.indptr) and np.array_equal(ci, sub.indices) and np.array_equal(va, sub.data)


def test_image_laplacian_fast_equals_oracle_generator():
    A = run_config.image_laplacian_fast(13, 17, seed=0)
    Bm = matrices.image_graph_laplacian(13, 17, seed=0)
    assert _same(A, Bm, tol=1e-13)
    assert abs(A - A.T).max() < 1e-15

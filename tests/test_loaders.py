"""N4: the loaders of the reference's benchmark driver (Julia/benchmark.jl:3-4,21-28: MatrixMarket.jl `mmread`, MAT.jl
`Problem["A"]`) - host only, checked against scipy.io on files written to a temporary directory."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from oracle import matrices


def _same(A, B):
    A = sp.csc_matrix(A); B = sp.csc_matrix(B)
    A.sort_indices(); B.sort_indices()
    return A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and \
        np.array_equal(A.data, B.data)


@pytest.mark.parametrize("symmetry", ["symmetric", "general"])
def test_matrix_market_reader_matches_scipy(rbl, tmp_path, symmetry):
    A = matrices.erdos_renyi_sym(300, 8, seed=4) + sp.diags(np.arange(300.0) - 150)
    path = os.path.join(tmp_path, "m.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(A), symmetry=symmetry, precision=17)
    M = rbl.load_matrix_market(path)
    ref = sp.csc_matrix(scipy.io.mmread(path))
    assert _same(M, ref)
    assert abs(M - M.T).max() == 0          # full symmetric matrix, ready for rbl_create
    assert _same(rbl.load_matrix(path), ref)


def test_matrix_market_pattern_integer_duplicates_and_errors(rbl, tmp_path):
    p = os.path.join(tmp_path, "p.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n% comment\n\n4 4 4\n1 1\n3 1\n4 2\n4 4\n")
    M = rbl.load_matrix_market(p).toarray()
    assert np.array_equal(M, np.array([[1, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 1.0]]))
    q = os.path.join(tmp_path, "q.mtx")
    with open(q, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n3 3 4\n1 2 5\n2 1 5\n1 2 -2\n3 3 7\n")
    M = rbl.load_matrix_market(q).toarray()
    assert np.array_equal(M, np.array([[0, 3, 0], [5, 0, 0], [0, 0, 7.0]]))    # duplicate (1,2) entries add up
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate real general\n2 3 1\n1 1 1.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
                "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1.0 0.0\n",
                "hello\n"):
        with open(q, "w") as f:
            f.write(bad)
        with pytest.raises(rbl.RblError):
            rbl.load_matrix_market(q)
    with pytest.raises(rbl.RblError):
        rbl.load_matrix_market(os.path.join(tmp_path, "missing.mtx"))


def test_one_based_arrays_are_what_julia_would_pass(rbl, tmp_path):
    """index_base = 1 returns colptr / rowval exactly as a SparseMatrixCSC holds them."""
    A = matrices.laplacian_2d(6)
    path = os.path.join(tmp_path, "l.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(A), symmetry="symmetric")
    L = rbl.lib()
    h = C.c_void_p(); n = C.c_int64(); nnz = C.c_int64()
    assert L.rbl_matrix_market_read(os.fsencode(path), 1, C.byref(h), C.byref(n), C.byref(nnz)) == 0
    cp = C.POINTER(C.c_int64)(); rv = C.POINTER(C.c_int64)(); nz = C.POINTER(C.c_double)()
    assert L.rbl_matrix_arrays(h, C.byref(cp), C.byref(rv), C.byref(nz)) == 0
    colptr = np.ctypeslib.as_array(cp, shape=(n.value + 1,)).copy()
    rowval = np.ctypeslib.as_array(rv, shape=(nnz.value,)).copy()
    L.rbl_matrix_free(h)
    ref = sp.csc_matrix(A); ref.sort_indices()
    assert np.array_equal(colptr, ref.indptr + 1) and np.array_equal(rowval, ref.indices + 1)


def test_mat_file_problem_struct(rbl, tmp_path):
    """SuiteSparse .mat layout: Problem.A (benchmark.jl:25-27)."""
    A = sp.csc_matrix(matrices.laplacian_2d(5))
    path = os.path.join(tmp_path, "s.mat")
    scipy.io.savemat(path, {"Problem": {"A": A, "name": "test"}})
    assert _same(rbl.load_matrix(path), A)

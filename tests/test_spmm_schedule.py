"""Host logic of the patch-scheduled SpMM (spmm_sched.cu): the row schedule is a permutation of the rows grouped into compact
patches of the grid implied by the stencil offsets; matrices without that structure get none."""
import numpy as np
import pytest

from oracle import matrices


def _rows_fetched(A, order, info, sample=64):
    slots, npatch = info["slots"], info["npatch"]
    tot, cnt = 0.0, 0
    for p in range(0, npatch, max(1, npatch // sample)):
        rows = order[p * slots:(p + 1) * slots]
        rows = rows[rows >= 0]
        cols = np.unique(np.concatenate([A.indices[A.indptr[r]:A.indptr[r + 1]] for r in rows]))
        tot += len(cols) / len(rows)
        cnt += 1
    return tot / cnt


@pytest.mark.parametrize("case,dims,limit", [("lap3d", 3, 2.1), ("lap2d", 2, 1.35), ("image", 2, 1.35)])
def test_schedule_is_a_permutation_with_compact_patches(rbl, case, dims, limit):
    from rbl_b200 import binding as B
    A = {"lap3d": lambda: matrices.laplacian_3d(40), "lap2d": lambda: matrices.laplacian_2d(200),
         "image": lambda: matrices.image_graph_laplacian(160, 160, seed=2)}[case]().tocsr()
    A.sort_indices()
    n = A.shape[0]
    order, info = B.spmm_schedule(A.indptr, A.indices)
    assert order is not None and info["dims"] == dims
    assert len(order) == info["npatch"] * info["slots"] and len(order) <= 1.3 * n + 65536
    rows = order[order >= 0]
    assert len(rows) == n and np.array_equal(np.sort(rows), np.arange(n))      # every row exactly once
    # distinct rows of Q a patch touches per row it computes: 5.06 (3-D) / 3.1 (2-D) for 32 consecutive rows
    assert _rows_fetched(A, order, info) < limit


def test_schedule_of_a_row_shard_ignores_halo_columns(rbl):
    import rbl_b200
    from rbl_b200 import binding as B
    N = 36
    A = matrices.laplacian_3d(N).tocsr()
    A.sort_indices()
    n = A.shape[0]
    rs = rbl_b200.partition_rows(n, 2)
    r0, r1 = int(rs[1]), int(rs[2])
    S = A[r0:r1]
    halo, optr, loc = B.halo_plan(n, 2, rs, 1, S.indptr, S.indices)
    order, info = B.spmm_schedule(S.indptr, loc, nown=r1 - r0)
    assert order is not None and info["dims"] == 3 and info["stride1"] == N and info["stride2"] == N * N
    rows = order[order >= 0]
    assert np.array_equal(np.sort(rows), np.arange(r1 - r0))


def test_no_schedule_without_stencil_structure(rbl):
    from rbl_b200 import binding as B
    A = matrices.erdos_renyi_sym(30000, 16, seed=1).tocsr()
    assert B.spmm_schedule(A.indptr, A.indices) == (None, None)
    small = matrices.laplacian_2d(40).tocsr()                 # below the size where the schedule pays
    assert B.spmm_schedule(small.indptr, small.indices) == (None, None)

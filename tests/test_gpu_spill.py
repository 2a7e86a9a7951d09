"""N3: the host tier of the Krylov buffer - the reference's "hybrid" mode (hybrid_part_reorth!,
RBL_gpu.jl:59-81; CPU half of recover_eigvec, :127-130; the pinned host mirror, :168-169).  Blocks that do not fit the
device budget live in pinned host memory and are streamed back for every Gram / update / Ritz pass.  The arithmetic is
the same as in the all-HBM solve, so the results must agree with it (and with the oracle) to roundoff."""
import numpy as np
import pytest

from oracle import matrices, rbl_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,b", [("mixed", 16), ("fp64", 16), ("mixed", 4), ("fp64", 8)])
def test_spilled_solve_equals_resident_solve(gpu, precision, b):
    N, k = 24, 30
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    n = N ** 3
    Om = np.random.default_rng(2).standard_normal((n, b))
    cap = 6000
    D0, V0, st0 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=cap, return_stats=True)
    assert st0.converged and st0.spilled_blocks == 0
    # a device budget that holds only ~40% of the blocks the solve needs (11 blocks of it are staging buffers)
    from rbl_b200 import binding as B
    limit = None
    for mb in range(24, 600, 2):
        with B.Solver(L, options=B.default_options(op=B.OP_SHIFT_MINUS_A, sigma=12.0, mem_limit_mb=mb, max_kryl_sz=cap,
                                                   precision=B.PRECISION_MIXED if precision == "mixed" else B.PRECISION_FP64)) as sp:
            if sp.plan_blocks(k, b) - 11 >= max(4, int(0.4 * st0.iterations)):
                limit = mb
                break
    assert limit is not None
    D1, V1, st1, s = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=cap, spill=True,
                                 mem_limit_mb=limit, return_solver=True)
    try:
        assert st1.converged
        assert st1.buffer_blocks < st0.iterations, (st1.buffer_blocks, st0.iterations)     # the device really was too small
        assert st1.spilled_blocks >= st1.iterations - st1.buffer_blocks - 1
        assert abs(st1.iterations - st0.iterations) <= 4     # same algorithm; a borderline check may flip on roundoff
        assert np.max(np.abs(D1 - D0) / np.abs(D0)) < 1e-10
        if b >= 16:   # (a block narrower than the multiplicities of this spectrum misses copies - in the reference too)
            exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
            assert np.max(np.abs(D1 - exact) / exact) < 1e-8
        assert np.max(rbl_oracle.ritz_residuals(A, D1, V1, norm_a=12.0)) < 1e-6
        # the basis (HBM part + host part) is as orthogonal as the resident one
        Q = s.krylov_basis()
        assert Q.shape[1] == st1.iterations * b
        G = Q.T @ Q
        keep = np.diag(G) > 0
        E = G[np.ix_(keep, keep)] - np.eye(int(keep.sum()))
        assert np.linalg.norm(E, 2) < (1e-6 if precision == "mixed" else 1e-13)
    finally:
        s.close()


def test_spill_combined_with_restart(gpu):
    """Both tiers full: the cap (max_kryl_sz) is reached with part of the basis on the host, then restart + locking."""
    N, k, b = 24, 30, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    Om = np.random.default_rng(2).standard_normal((N ** 3, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision="mixed", max_kryl_sz=40 * b, restart=True, spill=True,
                           mem_limit_mb=64, return_stats=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged and st.restarts >= 1 and st.spilled_blocks > 0
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=12.0)) < 1e-6

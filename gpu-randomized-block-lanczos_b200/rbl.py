"""Host-side mirror of the reference's Julia entry points, on top of the C ABI.

    RBL_gpu(A, k, b) -> (D, V)        Julia/RBL_gpu.jl:205-221
    RBL(A, k, b)     -> (D, V)        Julia/RBL.jl:119-142   (same device path; the reference's CPU twin)

Positional call and return values are the reference's: D holds the k eigenvalues of largest magnitude in
descending |lambda| order (RBL.jl:116), V the n x k Ritz vectors (host, column-major).  Keywords expose the
constants the reference hard-codes, with the reference's values as defaults.
"""
from __future__ import annotations

import numpy as np

from . import binding as _b


def RBL_gpu(A, k: int, b: int, *, Omega=None, max_kryl_sz: int = 1200, tol: float = 1e-7, reorth_period: int = 2,
            check_period: int = 4, precision: str = "fp64", shift: float | None = None, device: int = -1,
            async_check: bool = True, host_threads: int = 0, v_fp32: bool = False, verbose: int = 0,
            ngpus: int = 1, filter_degree: int = 0, restart: bool = False, spill: bool = False, probe_steps: int = 0,
            mem_limit_mb: int = 0, seed: int = 0, reorth_impl: int = 0, index_base: int = 0,
            return_stats: bool = False, allow_not_converged: bool = False, return_solver: bool = False):
    """Drop-in for RBL_gpu(A,k,b).  `shift=sigma` solves for the largest |lambda| of sigma*I - A (i.e. the
    lowest eigenpairs of A when sigma >= lambda_max); D is then reported for the shifted operator, exactly
    as if the caller had passed sigma*I - A to the reference.

    Beyond the reference's constants: `ngpus` (row-sharded over the GPUs of this process), `restart` (lock + restart
    at the Krylov cap, restarted.jl), `filter_degree` (Chebyshev-filtered operator), `spill` (host tier of the
    Krylov buffer), `mem_limit_mb` (device-memory budget), `index_base=1` (hand over Julia's 1-based arrays)."""
    opts = _b.default_options(max_kryl_sz=int(max_kryl_sz), tol=float(tol), reorth_period=int(reorth_period),
                              check_period=int(check_period),
                              precision=_b.PRECISION_MIXED if precision in ("mixed", "fp32") else _b.PRECISION_FP64,
                              op=_b.OP_SHIFT_MINUS_A if shift is not None else _b.OP_A,
                              sigma=float(shift) if shift is not None else 0.0, device=int(device),
                              async_check=int(async_check), host_threads=int(host_threads),
                              v_fp32=int(bool(v_fp32)), verbose=int(verbose), ngpus=int(ngpus),
                              filter_degree=int(filter_degree), restart=int(bool(restart)), spill=int(bool(spill)),
                              probe_steps=int(probe_steps), mem_limit_mb=int(mem_limit_mb), seed=int(seed),
                              reorth_impl=int(reorth_impl))
    s = _b.Solver(A, options=opts, index_base=index_base)
    try:
        D, V, st = s.solve(int(k), int(b), Omega, allow_not_converged=allow_not_converged)
    except Exception:
        s.close()
        raise
    if return_solver:
        return D, V, st, s
    s.close()
    if return_stats:
        return D, V, st
    return D, V


def RBL(A, k: int, b: int, **kw):
    """RBL(A,k,b) of RBL.jl:119 - same contract; the CPU cap of 1400 columns (RBL.jl:133) is the default."""
    kw.setdefault("max_kryl_sz", 1400)
    return RBL_gpu(A, k, b, **kw)


def rbl_solve_sharded(A_local_rows, n: int, row0: int, k: int, b: int, *, rank: int, world: int, uid: bytes,
                      Omega_local=None, **opt_kw):
    """One rank of the row-sharded solve: `A_local_rows` is the CSR slice [row0,row0+nloc) x n (global columns)."""
    import scipy.sparse as sp
    M = sp.csr_matrix(A_local_rows)
    M.sort_indices()
    opts = _b.default_options(**opt_kw)
    shard = dict(n=n, row0=row0, rowptr=M.indptr.astype(np.int64), colidx=M.indices.astype(np.int64), vals=M.data,
                 rank=rank, world=world, uid=uid)
    with _b.Solver(options=opts, shard=shard) as s:
        return s.solve(k, b, Omega_local)

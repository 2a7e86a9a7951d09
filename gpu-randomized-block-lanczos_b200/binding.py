"""ctypes binding of include/rbl_b200.h - exactly what a Julia `ccall` wrapper binds (julia/RBL_b200.jl)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS = {0: "OK", 1: "NOT_CONVERGED", 2: "BREAKDOWN", 3: "OOM", 4: "INVALID", 5: "CUDA_ERROR", 6: "NCCL_ERROR",
          7: "NO_DEVICE"}
PRECISION_FP64, PRECISION_MIXED = 0, 1
OP_A, OP_SHIFT_MINUS_A = 0, 1


class RblError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"rbl_b200 status {status} ({STATUS.get(status, '?')}): {msg}")
        self.status = status


class RblOptions(C.Structure):
    _fields_ = [("max_kryl_sz", C.c_int64), ("tol", C.c_double), ("reorth_period", C.c_int32),
                ("check_period", C.c_int32), ("precision", C.c_int32), ("op", C.c_int32), ("sigma", C.c_double),
                ("device", C.c_int32), ("async_check", C.c_int32), ("host_threads", C.c_int32), ("v_fp32", C.c_int32),
                ("verbose", C.c_int32), ("reorth_impl", C.c_int32), ("seed", C.c_int32), ("ngpus", C.c_int32),
                ("filter_degree", C.c_int32), ("restart", C.c_int32), ("spill", C.c_int32), ("probe_steps", C.c_int32),
                ("mem_limit_mb", C.c_int32)]


class RblStats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("kryl_sz", C.c_int64), ("iterations_run", C.c_int64),
                ("converged", C.c_int32), ("checks", C.c_int32), ("full_checks", C.c_int32), ("deflated", C.c_int32),
                ("t_total", C.c_double), ("t_spmm", C.c_double), ("t_3term", C.c_double), ("t_qr", C.c_double),
                ("t_part_reorth", C.c_double), ("t_loc_reorth", C.c_double), ("t_eig", C.c_double),
                ("t_ritz", C.c_double), ("t_h2d", C.c_double), ("t_d2h", C.c_double), ("t_eig_wait", C.c_double),
                ("bytes_part_reorth", C.c_double), ("bytes_spmm", C.c_double), ("kernel_launches", C.c_int64),
                ("t_reorth_gram", C.c_double), ("t_reorth_update", C.c_double), ("bytes_reorth_gram", C.c_double),
                ("bytes_reorth_update", C.c_double), ("launches_reorth_gram", C.c_int64),
                ("launches_reorth_update", C.c_int64), ("launches_spmm", C.c_int64), ("t_ritz_kernel", C.c_double),
                ("bytes_ritz", C.c_double), ("flops_ritz", C.c_double), ("host_factorizations", C.c_int64),
                ("restarts", C.c_int64), ("locked", C.c_int64), ("spilled_blocks", C.c_int64), ("buffer_blocks", C.c_int64),
                ("max_residual", C.c_double), ("t_host_blocked", C.c_double), ("filter_cut", C.c_double),
                ("filter_degree", C.c_int32), ("filter_two_sided", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if f != "reserved"}


def lib_path() -> str:
    return os.path.join(_HERE, "lib", "librbl_b200.so")


_P64 = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)
_P32 = C.POINTER(C.c_int32)

# every symbol declared in include/rbl_b200.h (tests check that the library exports all of them)
SIGNATURES = {
    "rbl_last_error": (C.c_char_p, []),
    "rbl_version": (C.c_char_p, []),
    "rbl_device_count": (C.c_int, []),
    "rbl_options_default": (C.c_int, [C.POINTER(RblOptions)]),
    "rbl_struct_sizes": (C.c_int, [_P64, _P64]),
    "rbl_create": (C.c_int, [C.c_int64, C.c_int64, _P64, _P64, _PD, C.c_int, C.POINTER(RblOptions), C.POINTER(C.c_void_p)]),
    "rbl_create_dense": (C.c_int, [C.c_int64, _PD, C.POINTER(RblOptions), C.POINTER(C.c_void_p)]),
    "rbl_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "rbl_create_sharded": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P64, _P64, _PD, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.POINTER(RblOptions), C.POINTER(C.c_void_p)]),
    "rbl_destroy": (C.c_int, [C.c_void_p]),
    "rbl_release_cached_memory": (C.c_int, []),
    "rbl_solve": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _PD, _PD, C.c_void_p, C.POINTER(RblStats)]),
    "rbl_solve_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, _PD, C.c_void_p, C.POINTER(RblStats)]),
    "rbl_buffer_blocks": (C.c_int, [C.c_void_p, C.c_int64, _P64]),
    "rbl_plan_blocks": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _P64]),
    "rbl_krylov_info": (C.c_int, [C.c_void_p, _P64, _P64]),
    "rbl_krylov_block": (C.c_int, [C.c_void_p, C.c_int64, _PD]),
    "rbl_orthogonality": (C.c_int, [C.c_void_p, _PD, _PD]),
    "rbl_query_memory": (C.c_int, [C.c_int, _P64, _P64]),
    "rbl_spmm": (C.c_int, [C.c_void_p, C.c_int64, _PD, _PD]),
    "rbl_gram": (C.c_int, [C.c_int64, C.c_int64, _PD, _PD, _PD]),
    "rbl_block_qr": (C.c_int, [C.c_int64, C.c_int64, _PD, _PD, _P32]),
    "rbl_reorth": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, _PD, _PD, C.c_void_p, C.c_int]),
    "rbl_ritz": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, _PD, C.c_void_p, C.c_int]),
    "rbl_band_eig_topk": (C.c_int, [C.c_int64, C.c_int64, _PD, C.c_int64, _PD, C.c_int64, C.c_double, C.c_int, _PD,
                                    _PD, _PD, _P32]),
    "rbl_checker_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rbl_checker_check": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _PD, C.c_int64, _PD, C.c_int64, C.c_double, C.c_int,
                                    _PD, _PD, _PD, _P32, _P32, _P64]),
    "rbl_checker_set_seeds": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _PD, _PD, _PD]),
    "rbl_checker_set_need_seeds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rbl_checker_destroy": (C.c_int, [C.c_void_p]),
    "rbl_band_count_below": (C.c_int, [C.c_int64, C.c_int64, _PD, C.c_double, _P64]),
    "rbl_partition_rows": (C.c_int, [C.c_int64, C.c_int, _P64]),
    "rbl_halo_plan": (C.c_int, [C.c_int64, C.c_int, _P64, C.c_int, C.c_int64, C.c_int64, _P64, _P64, _P64, _P64, _P64,
                                _P32]),
    "rbl_spmm_schedule": (C.c_int, [C.c_int64, C.c_int64, _P32, _P32, C.c_int, _P64, _P32, _P64]),
    "rbl_matrix_market_read": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), _P64, _P64]),
    "rbl_matrix_arrays": (C.c_int, [C.c_void_p, C.POINTER(_P64), C.POINTER(_P64), C.POINTER(_PD)]),
    "rbl_matrix_free": (C.c_int, [C.c_void_p]),
    "rbl_spmm_bench": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _PD, _P64]),
    "rbl_microbench": (C.c_int, [C.c_int, C.c_int64, C.c_int, _PD]),
}


def load_library(path: str | None = None):
    """Loads the C-ABI library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or lib_path()
    if not os.path.exists(p):
        raise RblError(7, f"{p} not found - build it with `python -m __graft_entry__` / build.py; there is no CPU fallback")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def lib():
    return load_library()


def _check(status, allow=(0,)):
    if status not in allow:
        raise RblError(status, lib().rbl_last_error().decode())
    return status


def _pd(a):
    return a.ctypes.data_as(_PD)


def _p64(a):
    return a.ctypes.data_as(_P64)


def default_options(**kw) -> RblOptions:
    o = RblOptions()
    _check(lib().rbl_options_default(C.byref(o)))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


class Solver:
    """Owns an rbl_handle (device-resident A).  `A` is any SciPy sparse matrix or a dense ndarray (symmetric).

    index_base=1 hands the library Julia's SparseMatrixCSC arrays exactly as `julia/RBL_b200.jl` does (Int64, 1-based
    colptr / rowval); ngpus > 1 (option) makes it a single-process multi-GPU group handle."""

    def __init__(self, A=None, *, options: RblOptions | None = None, shard=None, index_base: int = 0, **opt_kw):
        import scipy.sparse as sp
        self._h = C.c_void_p()
        self.options = options if options is not None else default_options(**opt_kw)
        L = lib()
        if shard is not None:
            # shard = dict(n, row0, rowptr, colidx, vals, rank, world, uid[, index_base])
            base = int(shard.get("index_base", 0))
            rp = np.ascontiguousarray(shard["rowptr"], dtype=np.int64)
            ci = np.ascontiguousarray(shard["colidx"], dtype=np.int64)
            va = np.ascontiguousarray(shard["vals"], dtype=np.float64)
            self.n = int(shard["n"])
            self.nloc = len(rp) - 1
            uid = shard.get("uid")
            uid_buf = C.create_string_buffer(bytes(uid), 128) if uid is not None else None
            _check(L.rbl_create_sharded(self.n, int(shard["row0"]), self.nloc, len(ci), _p64(rp), _p64(ci), _pd(va), base,
                                        int(shard["rank"]), int(shard["world"]), uid_buf, C.byref(self.options),
                                        C.byref(self._h)))
            return
        if sp.issparse(A):
            # a symmetric matrix: CSC of A is CSR of A; Julia hands colptr/rowval/nzval 1-based (index_base=1)
            M = sp.csr_matrix(A)
            M.sort_indices()
            self.n = M.shape[0]
            self.nloc = self.n
            rp = M.indptr.astype(np.int64) + index_base
            ci = M.indices.astype(np.int64) + index_base
            va = np.ascontiguousarray(M.data, dtype=np.float64)
            _check(L.rbl_create(self.n, len(ci), _p64(rp), _p64(ci), _pd(va), index_base, C.byref(self.options),
                                C.byref(self._h)))
        else:
            D = np.asfortranarray(A, dtype=np.float64)
            self.n = D.shape[0]
            self.nloc = self.n
            _check(L.rbl_create_dense(self.n, _pd(D), C.byref(self.options), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().rbl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def solve(self, k: int, b: int, Omega=None, allow_not_converged: bool = False):
        """(D, V, stats): host arrays in, host arrays out (H2D of Omega and D2H of V inside the call)."""
        D = np.zeros(k, dtype=np.float64)
        vdt = np.float32 if self.options.v_fp32 else np.float64
        V = np.zeros((self.nloc, k), dtype=vdt, order="F")
        st = RblStats()
        om = None
        if Omega is not None:
            om = np.asfortranarray(Omega, dtype=np.float64)
            assert om.shape == (self.nloc, b), (om.shape, (self.nloc, b))
        rc = lib().rbl_solve(self._h, k, b, _pd(om) if om is not None else None, _pd(D), V.ctypes.data_as(C.c_void_p),
                             C.byref(st))
        _check(rc, allow=(0, 1) if allow_not_converged else (0,))
        return D, V, st

    def solve_device(self, k: int, b: int, omega_ptr: int, v_ptr: int, allow_not_converged: bool = False):
        """Device-resident variant: omega_ptr / v_ptr are raw device addresses (column-major fp64 / V dtype)."""
        D = np.zeros(k, dtype=np.float64)
        st = RblStats()
        rc = lib().rbl_solve_device(self._h, k, b, C.c_void_p(omega_ptr), _pd(D), C.c_void_p(v_ptr), C.byref(st))
        _check(rc, allow=(0, 1) if allow_not_converged else (0,))
        return D, st

    def spmm(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        U = np.zeros_like(Q)
        _check(lib().rbl_spmm(self._h, Q.shape[1], _pd(Q), _pd(U)))
        return U

    def buffer_blocks(self, b: int) -> int:
        out = C.c_int64()
        _check(lib().rbl_buffer_blocks(self._h, b, C.byref(out)))
        return out.value

    def plan_blocks(self, k: int, b: int) -> int:
        out = C.c_int64()
        _check(lib().rbl_plan_blocks(self._h, k, b, C.byref(out)))
        return out.value

    def krylov_basis(self):
        """The Krylov basis the last solve left in the slab, decoded to fp64: n x (blocks*b) (locked vectors first)."""
        nb, b = C.c_int64(), C.c_int64()
        _check(lib().rbl_krylov_info(self._h, C.byref(nb), C.byref(b)))
        Q = np.zeros((self.nloc, nb.value * b.value), order="F")
        blk = np.zeros((self.nloc, b.value), order="F")
        for j in range(nb.value):
            _check(lib().rbl_krylov_block(self._h, j, _pd(blk)))
            Q[:, j * b.value:(j + 1) * b.value] = blk
        return Q

    def orthogonality(self):
        """(max |Q'Q - I|, ||Q'Q - I||_F) of the stored basis, computed on the device."""
        mx, fro = C.c_double(), C.c_double()
        _check(lib().rbl_orthogonality(self._h, C.byref(mx), C.byref(fro)))
        return mx.value, fro.value


# ---- kernel-level wrappers (parity tests) -------------------------------------------------------------
def k_spmm(A, Q, **opt_kw):
    with Solver(A, **opt_kw) as s:
        return s.spmm(Q)


def k_gram(X, Y):
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    n, b = X.shape
    Cm = np.zeros((b, b))
    _check(lib().rbl_gram(n, b, _pd(X), _pd(Y), _pd(Cm)))
    return Cm


def k_block_qr(U):
    Q = np.array(U, dtype=np.float64, order="C", copy=True)
    n, b = Q.shape
    R = np.zeros((b, b))
    d = np.zeros(b, dtype=np.int32)
    _check(lib().rbl_block_qr(n, b, _pd(Q), _pd(R), d.ctypes.data_as(_P32)))
    return Q, R, d


def k_reorth(Qbuf, W0, W1, fp32: bool, impl: int = 0):
    """Qbuf: (m, n, b) array; returns updated (W0, W1, C)."""
    dt = np.float32 if fp32 else np.float64
    Qb = np.ascontiguousarray(Qbuf, dtype=dt)
    m, n, b = Qb.shape
    w0 = np.array(W0, dtype=np.float64, order="C", copy=True)
    w1 = np.array(W1, dtype=np.float64, order="C", copy=True)
    Cm = np.zeros((m * b, 2 * b), dtype=dt)
    _check(lib().rbl_reorth(n, b, m, int(fp32), Qb.ctypes.data_as(C.c_void_p), _pd(w0), _pd(w1),
                            Cm.ctypes.data_as(C.c_void_p), impl))
    return w0, w1, Cm


def k_ritz(Qbuf, S, fp32: bool, impl: int = 0):
    dt = np.float32 if fp32 else np.float64
    Qb = np.ascontiguousarray(Qbuf, dtype=dt)
    m, n, b = Qb.shape
    S = np.ascontiguousarray(S, dtype=np.float64)
    k = S.shape[1]
    V = np.zeros((n, k), dtype=dt, order="F")
    _check(lib().rbl_ritz(n, b, m, k, int(fp32), Qb.ctypes.data_as(C.c_void_p), _pd(S), V.ctypes.data_as(C.c_void_p), impl))
    return V


# ---- host-only exports -----------------------------------------------------------------------------------
def band_eig_topk(ab, k: int, Bi=None, tol: float = 1e-7, threads: int = 1):
    """ab: (kd+1, N) LAPACK lower band (as built by insertA!/insertB!).  Returns D, S, resid, converged."""
    ab = np.asfortranarray(ab, dtype=np.float64)
    kd, N = ab.shape[0] - 1, ab.shape[1]
    D = np.zeros(k)
    S = np.zeros((N, k), order="F")
    res = np.zeros(k)
    conv = C.c_int32(0)
    b = 0
    bi = None
    if Bi is not None:
        bi = np.asfortranarray(Bi, dtype=np.float64)
        b = bi.shape[0]
    _check(lib().rbl_band_eig_topk(N, kd, _pd(ab), k, _pd(bi) if bi is not None else None, b, tol, threads, _pd(D),
                                   _pd(S), _pd(res), C.byref(conv)))
    return D, S, res, bool(conv.value)


class Checker:
    """Stateful host eigen-check (the object the solver keeps between convergence checks)."""

    def __init__(self, threads: int = 1):
        self._c = C.c_void_p()
        _check(lib().rbl_checker_create(threads, C.byref(self._c)))

    def check(self, ab, k: int, Bi, tol: float = 1e-7, force_full: bool = False):
        ab = np.asfortranarray(ab, dtype=np.float64)
        kd, N = ab.shape[0] - 1, ab.shape[1]
        bi = np.asfortranarray(Bi, dtype=np.float64)
        D = np.zeros(k); S = np.zeros((N, k), order="F"); res = np.zeros(k)
        conv = C.c_int32(); have = C.c_int32(); st = np.zeros(2, dtype=np.int64)
        _check(lib().rbl_checker_check(self._c, N, kd, _pd(ab), k, _pd(bi), bi.shape[0], tol, int(force_full), _pd(D),
                                       _pd(S), _pd(res), C.byref(conv), C.byref(have), _p64(st)))
        return dict(converged=bool(conv.value), have_all=bool(have.value), D=D, S=S, resid=res,
                    factorizations=int(st[0]), full=bool(st[1]))

    def set_seeds(self, D, S, resid=None):
        """k Ritz pairs (D, S: n_seed x k) of an earlier T: starting points of the next full check (the tracker's hand-over);
        resid: their residual bounds at that time (the worst few become witnesses)."""
        D = np.ascontiguousarray(D, dtype=np.float64)
        S = np.asfortranarray(S, dtype=np.float64)
        r = None if resid is None else np.ascontiguousarray(resid, dtype=np.float64)
        _check(lib().rbl_checker_set_seeds(self._c, S.shape[0], S.shape[1], _pd(D), _pd(S), _pd(r) if r is not None else None))

    def set_need_seeds(self, fn):
        """fn(N) is called when a full check is about to start without usable seeds (it may call set_seeds)."""
        self._need_cb = C.CFUNCTYPE(None, C.c_void_p, C.c_int64)(lambda user, N: fn(int(N))) if fn else None
        _check(lib().rbl_checker_set_need_seeds(self._c, C.cast(self._need_cb, C.c_void_p) if fn else None, None))

    def close(self):
        if self._c:
            lib().rbl_checker_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def band_count_below(ab, x: float) -> int:
    ab = np.asfortranarray(ab, dtype=np.float64)
    out = C.c_int64()
    _check(lib().rbl_band_count_below(ab.shape[1], ab.shape[0] - 1, _pd(ab), float(x), C.byref(out)))
    return out.value


def partition_rows(n: int, world: int):
    rs = np.zeros(world + 1, dtype=np.int64)
    _check(lib().rbl_partition_rows(n, world, _p64(rs)))
    return rs


def halo_plan(n: int, world: int, row_starts, rank: int, rowptr, colidx):
    rs = np.ascontiguousarray(row_starts, dtype=np.int64)
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colidx, dtype=np.int64)
    nloc, nnz = len(rp) - 1, len(ci)
    nh = C.c_int64()
    _check(lib().rbl_halo_plan(n, world, _p64(rs), rank, nloc, nnz, _p64(rp), _p64(ci), C.byref(nh), None, None, None))
    halo = np.zeros(nh.value, dtype=np.int64)
    optr = np.zeros(world + 1, dtype=np.int64)
    loc = np.zeros(nnz, dtype=np.int32)
    _check(lib().rbl_halo_plan(n, world, _p64(rs), rank, nloc, nnz, _p64(rp), _p64(ci), C.byref(nh), _p64(halo),
                               _p64(optr), loc.ctypes.data_as(_P32)))
    return halo, optr, loc


def spmm_schedule(rowptr, colidx, nown=None, slots: int = 0):
    """Row schedule of the patch-scheduled SpMM for 0-based CSR arrays: (order, info) or (None, None) when the matrix has
    no stencil structure.  info: dims, stride1, stride2, ext0..2, patch0..2, slots, npatch, halo0."""
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    n = len(rp) - 1
    cnt = C.c_int64()
    args = (n, n if nown is None else nown, rp.ctypes.data_as(_P32), ci.ctypes.data_as(_P32), slots)
    _check(lib().rbl_spmm_schedule(*args, C.byref(cnt), None, None))
    if cnt.value == 0:
        return None, None
    order = np.zeros(cnt.value, dtype=np.int32)
    info = np.zeros(12, dtype=np.int64)
    _check(lib().rbl_spmm_schedule(*args, C.byref(cnt), order.ctypes.data_as(_P32), _p64(info)))
    keys = ("dims", "stride1", "stride2", "ext0", "ext1", "ext2", "patch0", "patch1", "patch2", "slots", "npatch", "halo0")
    return order, dict(zip(keys, (int(v) for v in info)))


def microbench(which: int, size: int = 1 << 30, iters: int = 10) -> float:
    out = C.c_double()
    _check(lib().rbl_microbench(which, size, iters, C.byref(out)))
    return out.value


# ---- loaders (benchmark.jl:3-4,12-28: MatrixMarket.jl / MAT.jl) -----------------------------------------------------------
def load_matrix_market(path: str):
    """`mmread(path)` through the library's native reader -> scipy.sparse.csc_matrix (symmetric storage expanded)."""
    import scipy.sparse as sp
    h = C.c_void_p()
    n, nnz = C.c_int64(), C.c_int64()
    _check(lib().rbl_matrix_market_read(os.fsencode(path), 0, C.byref(h), C.byref(n), C.byref(nnz)))
    try:
        cp, rv, nz = _P64(), _P64(), _PD()
        _check(lib().rbl_matrix_arrays(h, C.byref(cp), C.byref(rv), C.byref(nz)))
        colptr = np.ctypeslib.as_array(cp, shape=(n.value + 1,)).copy()
        rowval = np.ctypeslib.as_array(rv, shape=(max(nnz.value, 1),))[:nnz.value].copy()
        nzval = np.ctypeslib.as_array(nz, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    finally:
        lib().rbl_matrix_free(h)
    return sp.csc_matrix((nzval, rowval, colptr), shape=(n.value, n.value))


def load_matrix(path: str):
    """SuiteSparse downloads as the reference's driver reads them: `.mtx` (Matrix Market, native reader) or `.mat`
    (MATLAB v5 `Problem.A`, benchmark.jl:25-27, through scipy.io.loadmat; v7.3 / HDF5 files are not supported)."""
    if path.lower().endswith(".mat"):
        import scipy.io
        import scipy.sparse as sp
        d = scipy.io.loadmat(path)
        A = d["Problem"]["A"][0, 0] if "Problem" in d else next(v for v in d.values() if sp.issparse(v))
        return sp.csc_matrix(A, dtype=np.float64)
    return load_matrix_market(path)

"""CPU-side tests: the C-ABI library loads and exports every declared symbol; host-only entry points
(band eigen-check, partition / halo plan) against SciPy / the NumPy oracle; no device compute calls."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.linalg import eig_banded

from oracle import matrices, partition_oracle, rbl_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(rbl):
    hdr = open(os.path.join(ROOT, "include", "rbl_b200.h")).read()
    declared = set(re.findall(r"\b(rbl_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rbl_options_default" if False else ""}
    L = rbl.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in rbl_b200.h but not exported"
    assert declared >= set(rbl.binding.SIGNATURES), "binding lists a symbol the header does not declare"


def test_struct_sizes_match_header(rbl):
    """The ctypes mirrors (and therefore the Julia structs of julia/RBL_b200.jl, which list the same fields in the same
    order) have exactly the compiled sizes."""
    import ctypes as C
    a, b = C.c_int64(), C.c_int64()
    assert rbl.lib().rbl_struct_sizes(C.byref(a), C.byref(b)) == 0
    assert C.sizeof(rbl.RblOptions) == a.value == 96
    assert C.sizeof(rbl.RblStats) == b.value
    o = rbl.binding.default_options()
    assert (o.max_kryl_sz, o.tol, o.reorth_period, o.check_period) == (1200, 1e-7, 2, 4)  # RBL_gpu.jl:211,189,164,186
    assert o.precision == 0 and o.op == 0
    assert (o.ngpus, o.filter_degree, o.restart, o.spill, o.mem_limit_mb, o.seed) == (1, 0, 0, 0, 0, 0)
    # the Julia wrapper's struct field lists mirror the header's
    jl = open(os.path.join(ROOT, "gpu-randomized-block-lanczos_b200", "julia", "RBL_b200.jl")).read()
    hdr = open(os.path.join(ROOT, "include", "rbl_b200.h")).read()
    for struct_name, jl_name in (("rbl_options", "RblOptions"), ("rbl_stats", "RblStats")):
        body = hdr[hdr.index("typedef struct {", hdr.index("Options.") if struct_name == "rbl_options" else hdr.index("Per-solve statistics")):]
        body = body[:body.index("} " + struct_name)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_fields = re.findall(r"\b(?:int64_t|int32_t|double)\s+([a-z_0-9]+)\s*(?:\[\d+\])?;", body)
        jbody = jl[jl.index("struct " + jl_name):]
        jbody = jbody[:jbody.index("\nend")]
        j_fields = re.findall(r"^\s+([a-z_0-9]+)::", jbody, flags=re.M)
        assert c_fields == j_fields, (struct_name, c_fields, j_fields)
        py_fields = [f for f, _ in (rbl.RblOptions if struct_name == "rbl_options" else rbl.RblStats)._fields_]
        assert c_fields == py_fields


def test_no_device_is_a_loud_error(rbl):
    if rbl.lib().rbl_device_count() > 0:
        pytest.skip("a device is present")
    A = matrices.laplacian_2d(8)
    with pytest.raises(rbl.RblError) as e:
        rbl.RBL_gpu(A, 2, 2)
    assert e.value.status == 7  # RBL_NO_DEVICE: no CPU fallback


def test_rejects_bad_arguments(rbl):
    import ctypes as C
    L = rbl.lib()
    h = C.c_void_p()
    rp = np.array([0, 1], dtype=np.int64)
    assert L.rbl_create(0, 0, rp.ctypes.data_as(C.POINTER(C.c_int64)), None, None, 0, None, C.byref(h)) == 4
    assert L.rbl_create(1 << 31, 1, rp.ctypes.data_as(C.POINTER(C.c_int64)), rp.ctypes.data_as(C.POINTER(C.c_int64)),
                        np.zeros(1).ctypes.data_as(C.POINTER(C.c_double)), 0, None, C.byref(h)) == 4
    assert b"2^31" in L.rbl_last_error()


def _band(M, kd):
    N = M.shape[0]
    ab = np.zeros((kd + 1, N))
    for d in range(kd + 1):
        ab[d, :N - d] = np.diag(M, -d)
    return ab


def _topk_ref(ab, k):
    w = eig_banded(ab, lower=True, eigvals_only=True)
    return w[np.argsort(-np.abs(w), kind="stable")][:k]


@pytest.mark.parametrize("N,kd,k,seed", [(40, 2, 5, 0), (200, 4, 16, 1), (600, 16, 60, 2), (333, 5, 333 // 3, 3)])
def test_band_eig_topk_random(rbl, N, kd, k, seed):
    rng = np.random.default_rng(seed)
    M = np.zeros((N, N))
    for d in range(kd + 1):
        v = rng.standard_normal(N - d)
        M += np.diag(v, d) + (np.diag(v, -d) if d else 0)
    ab = _band(M, kd)
    D, S, res, conv = rbl.band_eig_topk(ab, k, threads=2)
    ref = _topk_ref(ab, k)
    sc = np.max(np.abs(ref))
    assert np.max(np.abs(np.abs(D) - np.abs(ref))) < 1e-12 * sc
    assert np.all(np.abs(D)[:-1] >= np.abs(D)[1:] - 1e-12 * sc)          # descending |lambda| (RBL.jl:116)
    assert np.max(np.linalg.norm(M @ S - S * D[None, :], axis=0)) < 1e-11 * sc
    assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-8


def test_band_eig_multiplicities_and_zero_rows(rbl):
    rng = np.random.default_rng(5)
    kd = 4
    blk = rng.standard_normal((50, 50))
    blk = np.triu(np.tril(blk + blk.T, kd), -kd)
    M = np.zeros((230, 230))
    for i in range(4):
        M[i * 50:(i + 1) * 50, i * 50:(i + 1) * 50] = blk          # every eigenvalue 4-fold; trailing rows zero
    ab = _band(M, kd)
    D, S, res, conv = rbl.band_eig_topk(ab, 24)
    ref = _topk_ref(ab, 24)
    assert np.max(np.abs(np.sort(np.abs(D)) - np.sort(np.abs(ref)))) < 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(S.T @ S - np.eye(24))) < 1e-9
    assert np.max(np.linalg.norm(M @ S - S * D[None, :], axis=0)) < 1e-11 * np.max(np.abs(ref))


def test_band_count_is_sturm_count(rbl):
    rng = np.random.default_rng(9)
    N, kd = 150, 6
    M = np.zeros((N, N))
    for d in range(kd + 1):
        v = rng.standard_normal(N - d)
        M += np.diag(v, d) + (np.diag(v, -d) if d else 0)
    ab = _band(M, kd)
    w = np.linalg.eigvalsh(M)
    for x in np.r_[np.linspace(w[0] - 1, w[-1] + 1, 23), 0.5 * (w[10] + w[11])]:
        assert rbl.band_count_below(ab, float(x)) == int(np.sum(w < x))


def test_check_decision_matches_dsbev_on_lanczos_T(rbl):
    """Same accept/reject decision as dsbev + sort_eig_abs + check_convergence (common.jl:36-65) at every
    check of an oracle run, on the oracle's own T and B_i."""
    N, k, b = 24, 5, 3
    A = matrices.shifted(matrices.laplacian_2d(N), 8.0)
    Om = np.random.default_rng(11).standard_normal((N * N, b))
    Q = []
    Qi = np.linalg.qr(A @ Om)[0]
    # replay lanczos_iteration and intercept every check
    decisions = []
    orig = rbl_oracle.check_convergence

    def spy(Bi, V, bb, kk, tol):
        ok = orig(Bi, V, bb, kk, tol)
        decisions.append((Bi.copy(), ok))
        return ok

    rbl_oracle.check_convergence = spy
    try:
        Ts = []
        orig_dsbev = rbl_oracle.dsbev

        def spy_dsbev(T):
            Ts.append(T.copy())
            return orig_dsbev(T)

        rbl_oracle.dsbev = spy_dsbev
        rbl_oracle.lanczos_iteration(A, k, b, 1400, Qi, Q)
    finally:
        rbl_oracle.check_convergence = orig
        rbl_oracle.dsbev = orig_dsbev
    assert len(Ts) == len(decisions) and decisions[-1][1]
    for T, (Bi, ok) in zip(Ts, decisions):
        D, S, res, conv = rbl.band_eig_topk(T, k, Bi=Bi, tol=1e-7)
        w, z = orig_dsbev(T)
        Dr, Vr = rbl_oracle.sort_eig_abs(w, z, k)
        assert np.max(np.abs(D - Dr[::-1])) < 1e-11 * 8
        rr = rbl_oracle.residual_bounds(Bi, Vr, b)[::-1]
        # decisions agree unless a residual bound sits within rounding of the tolerance
        if np.min(np.abs(rr - 1e-7)) > 1e-9:
            assert conv == ok


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_and_halo_plan_bit_exact(rbl, world):
    A = matrices.laplacian_3d(9).tocsr()
    A.sort_indices()
    n = A.shape[0]
    rs = rbl.partition_rows(n, world)
    assert np.array_equal(rs, partition_oracle.partition_rows(n, world))
    for rank in range(world):
        Al = A[rs[rank]:rs[rank + 1], :]
        halo, optr, loc = rbl.halo_plan(n, world, rs, rank, Al.indptr, Al.indices)
        h2, o2, l2 = partition_oracle.halo_plan(n, world, rs, rank, Al.indptr, Al.indices)
        assert np.array_equal(halo, h2) and np.array_equal(optr, o2) and np.array_equal(loc, l2)
        # the plan reproduces the SpMM: local rows times [own | halo] rows of Q equals the global product
        Q = np.random.default_rng(rank).standard_normal((n, 4))
        Qext = np.vstack([Q[rs[rank]:rs[rank + 1]], Q[halo]])
        Aloc = sp.csr_matrix((Al.data, loc, Al.indptr), shape=(Al.shape[0], Qext.shape[0]))
        assert np.array_equal(Aloc @ Qext, (A @ Q)[rs[rank]:rs[rank + 1]]) or np.allclose(Aloc @ Qext, (A @ Q)[rs[rank]:rs[rank + 1]], rtol=0, atol=1e-13)


def test_halo_plan_random_sparsity(rbl):
    A = matrices.erdos_renyi_sym(500, 8, seed=4)
    A.sort_indices()
    rs = rbl.partition_rows(500, 4)
    for rank in range(4):
        Al = A[rs[rank]:rs[rank + 1], :]
        out = rbl.halo_plan(500, 4, rs, rank, Al.indptr, Al.indices)
        ref = partition_oracle.halo_plan(500, 4, rs, rank, Al.indptr, Al.indices)
        for a, b_ in zip(out, ref):
            assert np.array_equal(a, b_)


def test_seeded_full_check_matches_dsbev(rbl):
    """The accepting check refines the pairs of an earlier full solve (seeds) instead of slicing from scratch:
    same eigenvalues as dsbev, orthonormal Ritz vectors, same residual bounds (a 3-D Laplacian has many
    degenerate Ritz values, the hard case for the seeded path)."""
    import sys
    sys.path.insert(0, ROOT)
    from tools.replay_checks import capture
    N, k, b = 12, 24, 8
    A = matrices.shifted(matrices.laplacian_3d(N), 12.0)
    Om = np.random.default_rng(2).standard_normal((N ** 3, b))
    Ts, Bs, oks = capture(A, k, b, Om)
    assert oks[-1] and len(Ts) >= 3
    ck = rbl.Checker(threads=2)
    for T, Bi in zip(Ts[-3:], Bs[-3:]):
        r = ck.check(T, k, Bi, force_full=True)       # 1st: slicing; 2nd and 3rd: from the seeds of the previous one
        assert r["have_all"]
        w, z = rbl_oracle.dsbev(T)
        Dr, Vr = rbl_oracle.sort_eig_abs(w, z, k)
        assert np.max(np.abs(r["D"] - Dr[::-1])) < 1e-11 * 12
        S = r["S"]
        assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-8
        M = rbl_oracle.dense_band_from_T(T)
        assert np.max(np.linalg.norm(M @ S - S * r["D"][None, :], axis=0)) < 1e-10 * 12
    assert r["converged"] == oks[-1]


@pytest.mark.parametrize("grid,k,b", [(16, 40, 8), (20, 60, 16)])
def test_stale_seeds_are_refined_repaired_or_rejected(rbl, grid, k, b):
    """Seeds taken from a much earlier (smaller) T: the seeded full check must return dsbev's eigenvalues and an
    orthonormal basis, whether it refines them, repairs missing entrants, or falls back to slicing
    (regression: two stale seeds once collapsed onto one eigenvector and were accepted)."""
    import sys
    sys.path.insert(0, ROOT)
    from tools.replay_checks import capture
    A = matrices.shifted(matrices.laplacian_3d(grid), 12.0)
    Om = np.random.default_rng(0).standard_normal((grid ** 3, b))
    Ts, Bs, oks = capture(A, k, b, Om)
    w, z = rbl_oracle.dsbev(Ts[-1])
    Dr, _ = rbl_oracle.sort_eig_abs(w, z, k)
    for frac in (0.5, 0.75, 0.86, 0.9, 0.95):
        i0 = int(frac * len(Ts))
        ck = rbl.Checker(threads=2)
        assert ck.check(Ts[i0], k, Bs[i0], force_full=True)["have_all"]
        r = ck.check(Ts[-1], k, Bs[-1], force_full=True)
        assert r["have_all"]
        assert np.max(np.abs(r["D"] - Dr[::-1])) < 1e-11 * 12
        S = r["S"]
        assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-8


def test_rejected_seeds_assist_the_slicing(rbl):
    """In the middle of a run many Ritz values enter the wanted set between two full solves: the refined seeds fail the
    count validation, but they are still eigenpairs of T and the slicing only has to find what is missing.  The result
    must be dsbev's at every stage (eigenvalues, orthonormal vectors, small residuals)."""
    import sys
    sys.path.insert(0, ROOT)
    from tools.replay_checks import capture
    grid, k, b = 20, 60, 16
    A = matrices.shifted(matrices.laplacian_3d(grid), 12.0)
    Om = np.random.default_rng(0).standard_normal((grid ** 3, b))
    Ts, Bs, oks = capture(A, k, b, Om)
    ck = rbl.Checker(threads=2)
    L = len(Ts)
    picks = sorted({max(1, int(f * L)) for f in (0.3, 0.4, 0.5, 0.62, 0.75, 0.9)} | {L - 1})
    for i in picks:
        T = Ts[i]
        if T.shape[1] < 2 * k:
            continue
        r = ck.check(T, k, Bs[i], force_full=True)
        assert r["have_all"]
        w, z = rbl_oracle.dsbev(T)
        Dr, _ = rbl_oracle.sort_eig_abs(w, z, k)
        assert np.max(np.abs(r["D"] - Dr[::-1])) < 1e-11 * 12, i
        S = r["S"]
        assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-8, i
        M = rbl_oracle.dense_band_from_T(T)
        assert np.max(np.linalg.norm(M @ S - S * r["D"][None, :], axis=0)) < 1e-10 * 12, i


def _random_band(rng, N, kd, kind):
    """kind 0: plain random band (two-sided spectrum); 1: decaying off-diagonals; 2: decoupled identical blocks (exact
    multiplicities, half of them scaled by 0.9) plus a small diagonal tail."""
    M = np.zeros((N, N))
    for d in range(kd + 1):
        v = rng.standard_normal(N - d) * (0.3 ** d if kind else 1.0)
        M += np.diag(v, -d)
        if d:
            M += np.diag(v, d)
    if kind == 2:
        w = kd * 3
        blk = M[:w, :w].copy()
        M[:, :] = 0
        nb = N // w
        for q in range(nb):
            M[q * w:(q + 1) * w, q * w:(q + 1) * w] = blk if q % 2 == 0 else blk * 0.9
        for i in range(nb * w, N):
            M[i, i] = 0.01 * i
    return M


def _band_of(M, kd):
    N = M.shape[0]
    ab = np.zeros((kd + 1, N))
    for d in range(kd + 1):
        ab[d, :N - d] = np.diag(M, -d)
    return ab


@pytest.mark.parametrize("trial", range(15))
def test_seeded_paths_on_random_bands(rbl, trial):
    """A full solve of a leading principal submatrix provides the seeds for the full solve of the whole band matrix
    (refined, rejected or assisting the slicing - whatever the counts say): always the k eigenvalues of largest
    magnitude of a dense eigh, orthonormal vectors, small residuals."""
    rng = np.random.default_rng(1000 + trial)
    kd = int(rng.integers(2, 9))
    N2 = int(rng.integers(150, 420))
    N1 = int(N2 * rng.uniform(0.7, 0.97))
    k = int(rng.integers(5, 40))
    M2 = _random_band(rng, N2, kd, trial % 3)
    ck = rbl.Checker(threads=2)
    Bi = np.triu(rng.standard_normal((kd, kd)))
    assert ck.check(_band_of(M2[:N1, :N1], kd), k, Bi, force_full=True)["have_all"]
    r = ck.check(_band_of(M2, kd), k, Bi, force_full=True)
    assert r["have_all"]
    w = np.linalg.eigvalsh(M2)
    tn = np.max(np.abs(w))
    ref = np.sort(np.abs(w))[::-1][:k]
    assert np.max(np.abs(np.sort(np.abs(r["D"]))[::-1] - ref)) < 1e-10 * tn
    S = r["S"]
    assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-8
    assert np.max(np.linalg.norm(M2 @ S - S * r["D"][None, :], axis=0)) < 1e-9 * tn


@pytest.mark.parametrize("threads", [1, 4])
def test_band_eig_clusters_large_threaded(rbl, threads):
    """Degenerate clusters at a size where the threaded slicing pre-splits its roots and pulls narrow intervals out by
    deflated block inverse iteration (N >= 400): 3-fold eigenvalues, compared with LAPACK."""
    rng = np.random.default_rng(21)
    kd, nb, copies = 8, 180, 3
    blk = rng.standard_normal((nb, nb))
    blk = np.triu(np.tril(blk + blk.T, kd), -kd)
    N = nb * copies
    M = np.zeros((N, N))
    for i in range(copies):
        M[i * nb:(i + 1) * nb, i * nb:(i + 1) * nb] = blk
    ab = _band(M, kd)
    k = 45
    D, S, res, conv = rbl.band_eig_topk(ab, k, threads=threads)
    ref = _topk_ref(ab, k)
    sc = np.max(np.abs(ref))
    assert np.max(np.abs(np.sort(np.abs(D)) - np.sort(np.abs(ref)))) < 1e-12 * sc
    assert np.max(np.abs(S.T @ S - np.eye(k))) < 1e-9
    assert np.max(np.linalg.norm(M @ S - S * D[None, :], axis=0)) < 1e-11 * sc


def test_tracker_seeds_and_witnesses_keep_the_decisions(rbl):
    """The solver's arrangement in miniature: a main checker that sees every check and a tracker that hands it, every
    third check, the k pairs of an earlier T together with their residual bounds (the worst of which become witnesses).
    Every decision must equal dsbev + sort_eig_abs + check_convergence, and stale hand-overs must not displace fresher
    seeds."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from tools.replay_checks import capture
    N, k, b = 14, 30, 8
    A = matrices.shifted(matrices.laplacian_3d(N), 12.0)
    Om = np.random.default_rng(4).standard_normal((N ** 3, b))
    Ts, Bs, oks = capture(A, k, b, Om)
    assert oks[-1] and len(Ts) > 8
    main = rbl.Checker(threads=4)
    tracker = rbl.Checker(threads=3)
    handed = []
    for n, (T, Bi, ok) in enumerate(zip(Ts, Bs, oks)):
        if n >= 3 and n % 3 == 0:
            Told, Bold = Ts[n - 2], Bs[n - 2]
            q = tracker.check(Told, k, Bold, tol=0.0, force_full=True)
            assert q["have_all"] and not q["converged"]
            main.set_seeds(q["D"], q["S"], q["resid"])
            handed.append(Told.shape[1])
            if n % 6 == 0:      # an older hand-over arriving late: must be ignored
                main.set_seeds(q["D"][:k], q["S"][: Ts[n - 3].shape[1], :k].copy(), None)
        r = main.check(T, k, Bi)
        w, z = rbl_oracle.dsbev(T)
        Dr, Vr = rbl_oracle.sort_eig_abs(w, z, k)
        rr = rbl_oracle.residual_bounds(Bi, Vr, b)
        if np.min(np.abs(rr - 1e-7)) > 1e-9:
            assert r["converged"] == ok, (n, T.shape[1])
        if r["have_all"]:
            assert np.max(np.abs(np.sort(r["D"]) - np.sort(Dr))) < 1e-11 * 12
    assert handed


def test_check_timeline_tools_run_end_to_end(tmp_path, rbl):
    """tools/make_T_dump.py -> tools/check_timeline_sim.py (the CPU-side tuning loop of the host check): a small solve's T is
    generated, replayed against a virtual device clock, and accepted at a check point at which dsbev agrees."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from tools import check_timeline_sim, make_T_dump
    from tools.replay_dump import band, load
    path = str(tmp_path / "T.bin")
    make_T_dump.run(12, 8, 96, path, seed=3)                       # converges around block step 64
    # (a slow virtual device: every check is done before the next check point, whatever the load of the test machine)
    out = check_timeline_sim.simulate(path, step_ms=20.0, threads=2, k=24, verbose=False)
    assert out["accepted_step"] is not None and out["accepted_step"] % 4 == 0
    m, B, b, final_i, hA, hB = load(path)
    it = out["accepted_step"]
    w, z = rbl_oracle.dsbev(band(hA, hB, b, it))
    Dr, Vr = rbl_oracle.sort_eig_abs(w, z, 24)
    assert rbl_oracle.check_convergence(hB[it - 1][:b, :b], Vr, b, 24, 1e-7)
    assert out["idle_ms"] >= 0 and out["checks"] >= 1

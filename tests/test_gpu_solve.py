"""End-to-end parity of the device path (through the C ABI / RBL_gpu mirror) against the oracle:
the reference's own known-answer fixtures, and the north-star bars on reduced BASELINE configs
(eigenvalues rel 1e-8 vs the oracle, Ritz residual <= 1e-6 ||A||, orthogonality of V)."""
import numpy as np
import pytest

from oracle import matrices, rbl_oracle

pytestmark = pytest.mark.gpu
BAR = 1e-13  # Unit Testing/*_dec.jl:5


@pytest.mark.parametrize("gen,n", [(matrices.slow_decay, 100), (matrices.slow_decay, 500), (matrices.slow_decay, 900),
                                    (matrices.moderate_decay, 300), (matrices.moderate_decay, 700),
                                    (matrices.step_decay, 100000), (matrices.step_decay, 900000)])
def test_reference_fixtures_on_device(gpu, gen, n):
    """test.jl:10-50 with RBL replaced by the device path (k = b = 5, fp64 as shipped)."""
    A, eig = gen(n, 5)
    Om = np.random.default_rng(n).standard_normal((n, 5))
    d, V, st = gpu.RBL(A, 5, 5, Omega=Om, return_stats=True)
    assert st.converged
    assert np.linalg.norm((d - eig) / eig) < BAR
    assert np.max(rbl_oracle.ritz_residuals(A, d, V)) < 1e-6


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
@pytest.mark.parametrize("async_check", [False, True])
def test_config1_reduced_lowest_pairs(gpu, precision, async_check):
    """2-D 5-point Laplacian (config 1 at 60x60), 10 lowest pairs as the largest of 8I - A, b = 4."""
    N, k, b = 60, 10, 4
    L = matrices.laplacian_2d(N)
    A = matrices.shifted(L, 8.0)
    Om = np.random.default_rng(1).standard_normal((N * N, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=8.0, precision=precision, async_check=async_check,
                           max_kryl_sz=1400, return_stats=True)
    Do, Vo, det = rbl_oracle.RBL(A, k, b, Om, return_details=True)
    exact = 8.0 - matrices.laplacian_eigs(N, 2, k)
    assert st.converged
    assert np.max(np.abs(D - Do) / np.abs(Do)) < 1e-8            # north star: rel 1e-8 vs the reference restatement
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=8.0)) < 1e-6
    assert abs(st.iterations - det["stats"].iterations) <= 8      # same stopping check, +- two checks
    G = V.T.astype(np.float64) @ V.astype(np.float64)
    assert np.max(np.abs(G - np.eye(k))) < (1e-5 if precision == "mixed" else 1e-9)
    # subspace agreement with the oracle's Ritz vectors (degenerate pairs: compare projectors)
    P = Vo.T @ V
    assert np.min(np.linalg.svd(P, compute_uv=False)) > 1 - 1e-6


def test_async_equals_sync(gpu):
    N, k, b = 40, 8, 4
    L = matrices.laplacian_2d(N)
    Om = np.random.default_rng(5).standard_normal((N * N, b))
    r1 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=8.0, async_check=False, return_stats=True)
    r2 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=8.0, async_check=True, return_stats=True)
    assert r1[2].iterations == r2[2].iterations
    assert np.max(np.abs(r1[0] - r2[0])) < 1e-12
    assert r2[2].iterations_run >= r2[2].iterations


@pytest.mark.parametrize("precision", ["fp64", "mixed"])
def test_config2_reduced_3d(gpu, precision):
    """3-D 7-point Laplacian (config 2 at 24^3), 30 lowest pairs via 12I - A, b = 16."""
    N, k, b = 24, 30, 16
    L = matrices.laplacian_3d(N)
    A = matrices.shifted(L, 12.0)
    Om = np.random.default_rng(2).standard_normal((N ** 3, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=12.0, precision=precision, max_kryl_sz=3000, return_stats=True)
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    assert st.converged
    assert np.max(np.abs(D - exact) / exact) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=12.0)) < 1e-6
    Do, Vo = rbl_oracle.RBL(A, k, b, Om, max_kryl_sz=3000)
    assert np.max(np.abs(D - Do) / np.abs(Do)) < 1e-8


def test_config3_reduced_er_largest_magnitude(gpu):
    """Symmetric Erdos-Renyi (config 3 reduced): k extreme eigenvalues of both signs, b = 32."""
    n, k, b = 6000, 12, 32
    A = matrices.erdos_renyi_sym(n, 32, seed=3)
    Om = np.random.default_rng(3).standard_normal((n, b))
    D, V, st = gpu.RBL_gpu(A, k, b, Omega=Om, return_stats=True, max_kryl_sz=4000)
    Do, Vo = rbl_oracle.RBL(A, k, b, Om, max_kryl_sz=4000)
    assert st.converged
    assert (D > 0).any() and (D < 0).any()
    assert np.max(np.abs(D - Do) / np.abs(Do)) < 1e-8
    assert np.all(np.abs(D)[:-1] >= np.abs(D)[1:])
    assert np.max(rbl_oracle.ritz_residuals(A, D, V)) < 1e-6


def test_dense_operator_like_images_jl(gpu):
    """images.jl:14-30 style: eigenpairs of a dense B'B with block size 1."""
    rng = np.random.default_rng(8)
    Bm = rng.standard_normal((300, 120)) @ np.diag(0.9 ** np.arange(120))
    M = Bm.T @ Bm
    D, V = gpu.RBL_gpu(M, 6, 1, Omega=rng.standard_normal((120, 1)))
    w = np.linalg.eigvalsh(M)[::-1][:6]
    assert np.max(np.abs(D - w) / w) < 1e-8


def test_not_converged_status_and_cap(gpu):
    L = matrices.laplacian_2d(40)
    with pytest.raises(gpu.RblError) as e:
        gpu.RBL_gpu(L, 10, 4, shift=8.0, max_kryl_sz=64)
    assert e.value.status == 1
    D, V, st = gpu.RBL_gpu(L, 10, 4, shift=8.0, max_kryl_sz=64, allow_not_converged=True, return_stats=True)
    assert not st.converged and st.kryl_sz <= 64


def test_device_rng_start_block(gpu):
    """Omega = nothing: the library draws the start block itself (CUDA.randn, RBL_gpu.jl:213)."""
    A, eig = matrices.slow_decay(400, 5)
    d, V = gpu.RBL_gpu(A, 5, 5)
    assert np.linalg.norm((d - eig) / eig) < 1e-12


def test_config4_reduced_image_graph_laplacian(gpu):
    """Image-grid graph Laplacian with 8-neighbour weights (config 4 at 64x64): 10 smallest eigenpairs via
    sigma*I - L, b = 16, mixed precision (tensor-core reorth path), against a dense eigensolve."""
    H = W = 64
    L = matrices.image_graph_laplacian(H, W, seed=1)
    sigma = float(2.0 * L.diagonal().max())               # Gershgorin: lambda_max(L) <= 2 max degree
    A = matrices.shifted(L, sigma)
    k, b = 10, 16
    Om = np.random.default_rng(4).standard_normal((H * W, b))
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, shift=sigma, precision="mixed", max_kryl_sz=4096, return_stats=True)
    assert st.converged
    w = np.linalg.eigvalsh(L.toarray())[:k]
    assert np.max(np.abs((sigma - D) - w)) < 1e-8 * sigma
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=sigma)) < 1e-6
    assert abs(sigma - D[0]) < 1e-8 * sigma                # the constant vector: lambda_min(L) = 0


def test_config3_reduced_er_b32_mixed(gpu):
    """Erdos-Renyi, b = 32 (config 3's block size), mixed precision (FP16-split tensor-core reorth, B = 32)."""
    n, k, b = 20000, 16, 32
    A = matrices.erdos_renyi_sym(n, 32, seed=7)
    Om = np.random.default_rng(7).standard_normal((n, b))
    D, V, st = gpu.RBL_gpu(A, k, b, Omega=Om, precision="mixed", max_kryl_sz=6000, return_stats=True)
    Do, Vo = rbl_oracle.RBL(A, k, b, Om, max_kryl_sz=6000)
    assert st.converged
    assert np.max(np.abs(D - Do) / np.abs(Do)) < 1e-8
    assert np.max(rbl_oracle.ritz_residuals(A, D, V)) < 1e-6


# ---- single-process multi-GPU: RBL_gpu(A,k,b; ngpus=N) from ONE caller (RBL_gpu.jl:205) ------------
def _need_gpus(gpu, n):
    if gpu.lib().rbl_device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")


@pytest.mark.parametrize("case", ["lap3d-20 fp64", "lap3d-24 mixed", "er-4000 fp64", "lap3d-24 mixed filtered", "lap3d-24 fp64 restart"])
def test_multi_gpu_group_handle_equals_single_gpu(gpu, case):
    """Row-sharded solve over 2 GPUs of this process == single-GPU solve == analytic spectrum; V comes back as one
    n x k host matrix like the single-GPU call (was tools/multi_gpu_check.py under torchrun)."""
    _need_gpus(gpu, 2)
    kw = dict(max_kryl_sz=4000)
    if case.startswith("lap3d-20"):
        L, sigma, k, b, prec = matrices.laplacian_3d(20), 12.0, 20, 16, "fp64"
    elif case.startswith("lap3d-24"):
        L, sigma, k, b, prec = matrices.laplacian_3d(24), 12.0, 30, 16, ("mixed" if "mixed" in case else "fp64")
        if "filtered" in case:
            kw["filter_degree"] = 8
        if "restart" in case:
            kw.update(max_kryl_sz=40 * 16, restart=True)
    else:
        L, sigma, k, b, prec = matrices.erdos_renyi_sym(4000, 16, seed=1), None, 8, 8, "fp64"
    n = L.shape[0]
    Om = np.random.default_rng(5).standard_normal((n, b))
    D1, V1, st1 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=sigma, precision=prec, return_stats=True, **kw)
    for ng in (2, 4, 8):
        if gpu.lib().rbl_device_count() < ng:
            break
        D2, V2, st2 = gpu.RBL_gpu(L, k, b, Omega=Om, shift=sigma, precision=prec, ngpus=ng, return_stats=True, **kw)
        A = matrices.shifted(L, sigma) if sigma is not None else L
        assert st2.converged
        # row-sharded solves never wait at a check point: the accepting check may belong to a slightly later step (restarted
        # solves: in every cycle, and the cycles then lock different numbers of pairs)
        if "restart" in case:
            assert abs(st2.iterations - st1.iterations) <= 0.3 * st1.iterations
        else:
            assert -4 <= st2.iterations - st1.iterations <= 40
        assert np.max(np.abs(D2 - D1) / np.abs(D1)) < 1e-8
        assert V2.shape == (n, k)
        assert np.max(rbl_oracle.ritz_residuals(A, D2, V2)) < 1e-6
        # same invariant subspace - for the pairs above the last eigenvalue cluster (the k-th eigenvalue of a Laplacian
        # usually cuts a degenerate cluster, whose basis inside the returned set is not unique)
        sep = np.abs(D1 - D1[-1]) > 1e-6 * np.abs(D1[0])
        if sep.any():
            assert np.min(np.linalg.svd(V1[:, sep].T @ V2, compute_uv=False)) > 1 - 1e-6


def test_non_waiting_check_points_accept_the_same_solution(gpu):
    """async_check=2: a check point never waits for the host; the accepted step is never earlier than the waiting form's and
    the results agree far below the tolerances."""
    L = matrices.laplacian_3d(20)
    Om = np.random.default_rng(2).standard_normal((8000, 16))
    D1, V1, st1 = gpu.RBL_gpu(L, 20, 16, Omega=Om, shift=12.0, precision="mixed", async_check=1, return_stats=True)
    D2, V2, st2 = gpu.RBL_gpu(L, 20, 16, Omega=Om, shift=12.0, precision="mixed", async_check=2, return_stats=True)
    assert st2.converged and st1.iterations <= st2.iterations <= st1.iterations + 40
    assert st2.iterations % 4 == 0 and st2.iterations_run >= st2.iterations
    assert np.max(np.abs(D2 - D1) / np.abs(D1)) < 1e-10
    assert np.max(rbl_oracle.ritz_residuals(matrices.shifted(L, 12.0), D2, V2, norm_a=12.0)) < 1e-6


def test_multi_gpu_halo_exchange_overlapped_with_interior_rows(gpu, monkeypatch):
    """Default on (RBL_HALO_OVERLAP=0 disables): the SpMM computes the rows without halo columns while the exchange runs on a second stream, the
    flagged rows afterwards - same block SpMM results, hence the same solve (plain and Chebyshev form, where Z aliases U)."""
    _need_gpus(gpu, 2)
    L = matrices.laplacian_3d(24)
    Om = np.random.default_rng(6).standard_normal((24 ** 3, 16))
    for kw in (dict(), dict(filter_degree=6)):
        res = {}
        for ov in ("0", "1"):
            monkeypatch.setenv("RBL_HALO_OVERLAP", ov)
            res[ov] = gpu.RBL_gpu(L, 20, 16, Omega=Om, shift=12.0, precision="fp64", ngpus=2, async_check=0, max_kryl_sz=3200,
                                  return_stats=True, **kw)
        assert res["0"][2].converged and res["1"][2].converged
        assert res["0"][2].iterations == res["1"][2].iterations
        assert np.max(np.abs(res["0"][0] - res["1"][0]) / np.abs(res["0"][0])) < 1e-13
        exact = 12.0 - matrices.laplacian_eigs(24, 3, 20)
        assert np.max(np.abs(res["1"][0] - exact) / exact) < 1e-8


def test_multi_gpu_group_handle_one_based_and_device_rng(gpu):
    _need_gpus(gpu, 2)
    L = matrices.laplacian_2d(60)
    D1, V1 = gpu.RBL_gpu(L, 8, 4, shift=8.0, seed=3)
    D2, V2 = gpu.RBL_gpu(L, 8, 4, shift=8.0, seed=3, ngpus=2, index_base=1)
    # the counter-based generator draws element (row, col) independently of the partition: same start block
    assert np.max(np.abs(D2 - D1) / np.abs(D1)) < 1e-10


# ---- N4: the reference's drivers on top of the drop-in (benchmark.jl, images.jl) ------------------------------------
def test_benchmark_driver_with_matrix_market_file_and_arpack(gpu, tmp_path):
    """benchmark.jl:21-45: mmread -> RBL_gpu(A, nd, 4) -> compare with ARPACK eigs(A, nev=nd, tol=1e-7, which=:LM)."""
    import io
    import os
    import scipy.io
    import scipy.sparse as sp
    from tools import benchmark
    A = matrices.erdos_renyi_sym(3000, 12, seed=8)
    path = os.path.join(tmp_path, "er.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(A), symmetry="symmetric", precision=17)
    M = gpu.load_matrix(path)
    buf = io.StringIO()
    r = benchmark.run(M, 10, 4, precision="fp64", arpack=True, out=buf)
    assert r["stats"].converged
    assert np.max(np.abs(r["d"] - r["d_arpack"]) / np.abs(r["d_arpack"])) < 1e-6      # ARPACK's own tol is 1e-7
    assert "Largest:" in buf.getvalue() and "part reorth" in buf.getvalue()


def test_images_driver_low_rank_approximation(gpu):
    """images.jl:28-32: eigenpairs of the dense B'B with b = 1 give the optimal rank-k approximation of B."""
    from tools import images
    Bm = images.synthetic_image(120, 80)
    k = 12
    U, s, V, st = images.low_rank(Bm, k)
    sv = np.linalg.svd(Bm, compute_uv=False)
    assert st.converged
    assert np.max(np.abs(s - sv[:k]) / sv[:k]) < 1e-8
    Blr = (U * s[None, :]) @ V.T
    assert np.linalg.norm(Bm - Blr) <= np.sqrt(np.sum(sv[k:] ** 2)) * (1 + 1e-6)

"""Mirror of the reference's benchmark driver (Julia/benchmark.jl:23-62) on top of the drop-in entry points:

    python tools/benchmark.py [matrix.mtx|matrix.mat] [--nd 100] [--b 4] [--lowest SIGMA] [--precision mixed] [--arpack]

* loads a SuiteSparse matrix (`mmread` / `matopen`, benchmark.jl:21,25-28) through rbl_b200.load_matrix, or builds the
  3-D Laplacian when no file is given;
* warm-up call on a tiny random matrix like benchmark.jl:58, then `d, v = RBL_gpu(A, nd, b)` timed (benchmark.jl:36-37);
* optional ARPACK comparison `eigs(A, nev=nd, tol=1e-7, which=:LM)` (benchmark.jl:42) through SciPy's ARPACK wrapper
  (scipy.sparse.linalg.eigsh), with BLAS threads limited to 1 like benchmark.jl:49;
* prints "Largest ... smallest ..." (benchmark.jl:45) and the phase table the reference shows with `show(to)` (:61).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def run(A, nd, b, *, shift=None, precision="mixed", arpack=False, max_kryl_sz=None, restart=False, filter_degree=0, ngpus=1,
        out=sys.stdout):
    import rbl_b200
    n = A.shape[0]
    rng = np.random.default_rng(0)
    W = sp.random(50, 50, 0.5, random_state=1, data_rvs=rng.standard_normal)
    rbl_b200.RBL_gpu((W + W.T).tocsc(), 1, 1)                                           # benchmark.jl:58 warm-up
    t0 = time.perf_counter()
    d, v, st = rbl_b200.RBL_gpu(A, nd, b, shift=shift, precision=precision, max_kryl_sz=max_kryl_sz or max(1200, 40 * nd),
                                restart=restart, filter_degree=filter_degree, ngpus=ngpus, return_stats=True,
                                allow_not_converged=True)
    t_rbl = time.perf_counter() - t0
    print(f"Iterations: {st.iterations} and kryl_sz: {st.kryl_sz}", file=out)               # RBL_gpu.jl:195
    print(f"Largest: {d[0]} and smallest {d[nd - 1]}", file=out)                           # benchmark.jl:45
    res = {"RBL_gpu": t_rbl, "d": d, "stats": st}
    rows = [("RBL_gpu", 1, t_rbl), ("  AQ", st.launches_spmm, st.t_spmm), ("  3-term", st.iterations_run, st.t_3term),
            ("  qr", st.iterations_run, st.t_qr), ("  part reorth", st.launches_reorth_gram, st.t_part_reorth),
            ("  loc reorth", st.iterations_run, st.t_loc_reorth), ("  eig", st.checks, st.t_eig), ("  Ritz vectors", 1, st.t_ritz)]
    if arpack:
        from threadpoolctl import threadpool_limits
        Aop = A if shift is None else (shift * sp.identity(n, format="csr") - A)
        with threadpool_limits(limits=1, user_api="blas"):                                # benchmark.jl:49
            t0 = time.perf_counter()
            w, _ = spla.eigsh(Aop.tocsr(), k=nd, which="LM", tol=1e-7)                     # benchmark.jl:42
            t_ar = time.perf_counter() - t0
        w = w[np.argsort(-np.abs(w))]
        res["Arpack"] = t_ar
        res["d_arpack"] = w
        rows.append(("Arpack", 1, t_ar))
        print(f"Arpack: largest {w[0]} and smallest {w[nd - 1]}; max rel difference {np.max(np.abs(w - d) / np.abs(w)):.2e}", file=out)
    print(f"{'Section':<18}{'ncalls':>10}{'time':>12}", file=out)                          # show(to), benchmark.jl:61
    for name, ncalls, sec in rows:
        print(f"{name:<18}{int(ncalls):>10}{sec * 1e3:>10.1f}ms", file=out)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix", nargs="?")
    ap.add_argument("--nd", type=int, default=100)
    ap.add_argument("--b", type=int, default=4)
    ap.add_argument("--lowest", type=float, default=None, help="sigma: lowest eigenpairs of A as the largest of sigma*I - A")
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--arpack", action="store_true")
    ap.add_argument("--grid", type=int, default=40)
    ap.add_argument("--restart", action="store_true")
    ap.add_argument("--filter-degree", type=int, default=0)
    ap.add_argument("--ngpus", type=int, default=1)
    a = ap.parse_args()
    import rbl_b200
    if a.matrix:
        A = rbl_b200.load_matrix(a.matrix)
    else:
        from oracle import matrices
        A = matrices.laplacian_3d(a.grid).tocsc()
        if a.lowest is None:
            a.lowest = 12.0
    run(A, a.nd, a.b, shift=a.lowest, precision=a.precision, arpack=a.arpack, restart=a.restart, filter_degree=a.filter_degree,
        ngpus=a.ngpus)


if __name__ == "__main__":
    main()

"""SpMM (K1) timing inside a short solve: python tools/spmm_bench.py [--matrix lap3d|image|lap3d-shard] [--size N] [--b 16] [--degree 0]
t_spmm / launches_spmm come from CUDA events inside the library; --degree > 0 times the Chebyshev-recurrence form
(alpha*A*Q + beta*Q + gamma*Z), the SpMM of the filtered operator that dominates configs 4 and 5."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import rbl_b200
from rbl_b200 import binding as B
from oracle import matrices
ap = argparse.ArgumentParser()
ap.add_argument("--matrix", default="lap3d")
ap.add_argument("--size", type=int, default=100)
ap.add_argument("--b", type=int, default=16)
ap.add_argument("--degree", type=int, default=0)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
b = a.b
if a.matrix == "lap3d":
    L = matrices.laplacian_3d(a.size).tocsr(); sigma = 12.0
else:
    from run_config import image_laplacian_fast
    L = image_laplacian_fast(a.size, a.size, seed=0).tocsr(); sigma = 2.0 * float(L.diagonal().max())
L.sort_indices()
n = L.shape[0]
opts = B.default_options(max_kryl_sz=a.steps * b, precision=B.PRECISION_MIXED, op=B.OP_SHIFT_MINUS_A, sigma=sigma, async_check=0,
                         filter_degree=a.degree, restart=0, verbose=1 if os.environ.get("RBL_VERBOSE") else 0)
with B.Solver(L, options=opts) as s:
    for _ in range(2):
        D, V, st = s.solve(b, b, np.random.default_rng(0).standard_normal((n, b)), allow_not_converged=True)
    per = st.t_spmm / st.launches_spmm
    tag = " ".join(f"{k}={os.environ[k]}" for k in ("RBL_SPMM_SCHED", "RBL_SPMM_COOP", "RBL_SPMM_MINB", "RBL_SPMM_CTA", "RBL_SPMM_PATCH", "RBL_SPMM_CARVEOUT", "RBL_SPMM_WINDOW") if k in os.environ)
    print(f"{a.matrix} {a.size} b={b} deg={a.degree} [{tag}]: {st.launches_spmm} SpMM launches, {per * 1e6:.1f} us each, "
          f"{st.bytes_spmm / st.t_spmm / 1e9:.0f} GB/s algorithmic", flush=True)

"""SpMM (K1) timing on the bench workload's matrix: python tools/spmm_bench.py [grid] [b]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rbl_b200
from rbl_b200 import binding as B
from oracle import matrices
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
L = matrices.laplacian_3d(N).tocsr(); L.sort_indices()
n = L.shape[0]
# a short solve: t_spmm / launches_spmm from CUDA events inside the library
with B.Solver(L, options=B.default_options(max_kryl_sz=40 * b, precision=B.PRECISION_MIXED, op=B.OP_SHIFT_MINUS_A, sigma=12.0, async_check=0)) as s:
    for _ in range(2):
        D, V, st = s.solve(b, b, np.random.default_rng(0).standard_normal((n, b)), allow_not_converged=True)
    per = st.t_spmm / st.launches_spmm
    print(f"grid {N}^3 b={b}: {st.launches_spmm} SpMM launches, {per * 1e6:.1f} us each, {st.bytes_spmm / st.t_spmm / 1e9:.0f} GB/s algorithmic "
          f"(3-term {st.t_3term / st.iterations_run * 1e6:.0f} us/step, qr {st.t_qr / st.iterations_run * 1e6:.0f}, loc {st.t_loc_reorth / st.iterations_run * 1e6:.0f})")

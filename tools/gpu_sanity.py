"""Short GPU sanity run after host-side changes (no torch): smoke(), then the bench workload (config 2) through the public
entry with host buffers - waiting checks (async_check=1) and the non-waiting form row-sharded solves use (async_check=2) -
with the gates evaluated and the host-check statistics printed.

  gpurun -- 'python tools/gpu_sanity.py > gpurun_out/sanity.log 2>&1'
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry
import rbl_b200
from oracle import matrices

t0 = time.time()
entry.smoke()
print(f"smoke ok ({time.time() - t0:.1f} s)", flush=True)

N, K, B, SIGMA = 100, 100, 16, 12.0
L = matrices.laplacian_3d(N).tocsr()
L.sort_indices()
n = L.shape[0]
Om = np.random.default_rng(20240607).standard_normal((n, B))
exact = SIGMA - matrices.laplacian_eigs(N, 3, K)
A = matrices.shifted(L, SIGMA)
for mode in (1, 2, 1, 2):
    t1 = time.time()
    D, V, st = rbl_b200.RBL_gpu(L, K, B, Omega=Om, shift=SIGMA, precision="mixed", max_kryl_sz=9600, async_check=mode,
                                verbose=int(os.environ.get("SANITY_VERBOSE", "1")), return_stats=True)
    wall = time.time() - t1
    err = float(np.max(np.abs(np.sort(D)[::-1] - np.sort(exact)[::-1]) / np.abs(exact).max()))
    R = A @ V - V * D[None, :]
    res = float(np.max(np.linalg.norm(R, axis=0)) / SIGMA)
    print(f"async_check={mode}: wall {wall:.3f} s  t_total {st.t_total:.3f}  steps {st.iterations} (run {st.iterations_run})  checks {st.checks} "
          f"full {st.full_checks}  factorisations {st.host_factorizations}  t_eig {st.t_eig:.3f}  device idle {st.t_eig_wait:.3f}  "
          f"host blocked {st.t_host_blocked:.3f}  eig err {err:.2e}  residual {res:.2e}", flush=True)
    assert st.converged and err < 1e-8 and res < 1e-6
print("sanity ok")

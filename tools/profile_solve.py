"""One config-2 solve (the bench workload) for ncu: `python tools/profile_solve.py [--cap COLUMNS] [--precision mixed|fp64] [--impl I] [--block B]`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import rbl_b200
from rbl_b200 import binding as B

ap = argparse.ArgumentParser()
ap.add_argument("--cap", type=int, default=bench.MAX_KRYL)
ap.add_argument("--precision", default="mixed")
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--block", type=int, default=bench.BLOCK)
a = ap.parse_args()
L = bench.problem()
n = L.shape[0]
Om = bench.omega(n, a.block)
dev = torch.device("cuda", 0)
opts = B.default_options(max_kryl_sz=a.cap, precision=B.PRECISION_MIXED if a.precision == "mixed" else B.PRECISION_FP64,
                         op=B.OP_SHIFT_MINUS_A, sigma=bench.SIGMA, device=0, async_check=1, reorth_impl=a.impl)
s = B.Solver(L, options=opts)
om_dev = torch.from_numpy(np.ascontiguousarray(np.asfortranarray(Om).T)).to(dev)
v_dev = torch.empty((bench.K_WANTED, n), dtype=torch.float64, device=dev)
D, st = s.solve_device(bench.K_WANTED, a.block, om_dev.data_ptr(), v_dev.data_ptr(), allow_not_converged=True)
torch.cuda.synchronize()
print("iterations", st.iterations, "converged", st.converged, "launches", st.kernel_launches, "t_total", round(st.t_total, 3),
      "gram GB/s", round(st.bytes_reorth_gram / max(st.t_reorth_gram, 1e-9) / 1e9, 1),
      "update GB/s", round(st.bytes_reorth_update / max(st.t_reorth_update, 1e-9) / 1e9, 1))
print("phases[s]", {k: round(getattr(st, k), 3) for k in ("t_spmm", "t_3term", "t_qr", "t_part_reorth", "t_loc_reorth", "t_eig",
                                                         "t_eig_wait", "t_ritz", "t_reorth_gram", "t_reorth_update")},
      "checks", st.checks, "full", st.full_checks, "factorizations", st.host_factorizations)
s.close()

"""SpMM laboratory driver: times the K1 candidates of csrc/spmm_lab.cu in one process on the bench matrices.
python tools/spmm_lab.py [--matrix lap3d|image] [--size N]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
os.environ.setdefault("RBL_SPMM_SCHED", "1")       # plan the row schedule at create (laboratory only)
import rbl_b200
from rbl_b200 import binding as B
from oracle import matrices
ap = argparse.ArgumentParser()
ap.add_argument("--matrix", default="lap3d")
ap.add_argument("--size", type=int, default=100)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
if a.matrix == "lap3d":
    L = matrices.laplacian_3d(a.size).tocsr()
elif a.matrix == "slab":          # a rank's shard of BASELINE config 5: size x size x 50 grid points (halo columns dropped)
    import scipy.sparse as sp
    from run_config import laplacian_3d_rows
    N = a.size
    r0, r1 = 100 * N * N, 150 * N * N
    rp, ci, va = laplacian_3d_rows(N, r0, r1)
    keep = (ci >= r0) & (ci < r1)
    rows = np.repeat(np.arange(r1 - r0), np.diff(rp))
    L = sp.csr_matrix((va[keep], (rows[keep], ci[keep] - r0)), shape=(r1 - r0, r1 - r0))
else:
    from run_config import image_laplacian_fast
    L = image_laplacian_fast(a.size, a.size, seed=0).tocsr()
L.sort_indices()
n, nnz = L.shape[0], L.nnz
names = {0: "gen1 gather", 1: "csr pipe", 2: "ell"}
with B.Solver(L, options=B.default_options(precision=B.PRECISION_MIXED)) as s:
    for with_z in (0, 1):
        alg = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n * 16 + (8.0 * n * 16 if with_z else 0.0)
        for flush in (1, 0):
            for kind in ((0, 1) if os.environ.get("LAB_QUICK") else (0, 1, 2)):
                for sched in ((0,) if kind == 0 else (0, 16)):
                    for pf in ((0,) if (kind == 0 or os.environ.get("LAB_QUICK")) else (0, 1, 2) if kind == 1 else (0, 1)):
                        for gm in ((16,) if (kind == 0 or os.environ.get("LAB_QUICK")) else (8, 16, 32)):
                            us, bad = C.c_double(), C.c_int64()
                            rc = rbl_b200.lib().rbl_spmm_bench(s._h, 16, kind | sched | (pf << 8), gm, a.iters, flush, with_z, C.byref(us), C.byref(bad))
                            if rc != 0:
                                print(f"variant {kind}|{sched} pf={pf}: rc={rc} {rbl_b200.lib().rbl_last_error().decode()}"); continue
                            print(f"{a.matrix} {a.size} z={with_z} flush={flush} {names[kind]:12s} sched={sched//16} pf={pf} grid={gm:2d}xSM: "
                                  f"{us.value:8.1f} us  {alg / us.value / 1e3:6.0f} GB/s  mismatches={bad.value}", flush=True)

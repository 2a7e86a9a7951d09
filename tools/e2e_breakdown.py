"""Wall-clock split of the reference-facing call (rbl_create / rbl_solve with host buffers / rbl_destroy) on config 2."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import rbl_b200
from rbl_b200 import binding as B

L = bench.problem()
n = L.shape[0]
Om = bench.omega(n, bench.BLOCK)
om_pin = torch.from_numpy(np.asfortranarray(Om).T.copy()).pin_memory()
v_pin = torch.empty((bench.K_WANTED, n), dtype=torch.float64).pin_memory()
opts = dict(max_kryl_sz=bench.MAX_KRYL, precision=B.PRECISION_MIXED, op=B.OP_SHIFT_MINUS_A, sigma=bench.SIGMA, device=0,
            async_check=1, verbose=int(os.environ.get("RBL_VERBOSE", "0")))
for it in range(int(os.environ.get("CALLS", "4"))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = B.Solver(L, options=B.default_options(**opts))
    t1 = time.perf_counter()
    Dh = np.zeros(bench.K_WANTED)
    st = B.RblStats()
    rc = rbl_b200.lib().rbl_solve(s._h, bench.K_WANTED, bench.BLOCK, C.cast(om_pin.data_ptr(), C.POINTER(C.c_double)),
                                  Dh.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(v_pin.data_ptr()), C.byref(st))
    t2 = time.perf_counter()
    s.close()
    t3 = time.perf_counter()
    print(f"rc={rc} create {t1 - t0:.3f}  solve {t2 - t1:.3f} (library t_total {st.t_total:.3f}, h2d {st.t_h2d:.3f}, d2h {st.t_d2h:.3f}, "
          f"ritz {st.t_ritz:.3f})  destroy {t3 - t2:.3f}  total {t3 - t0:.3f}")

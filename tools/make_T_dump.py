"""CPU generator of a block-tridiagonal T in the RBL_DUMP_T format (what tools/replay_dump.py reads), so that the host
eigen-check (`csrc/band_eig.cpp`) can be profiled and tuned on T matrices of the bench's size without a GPU.

  python tools/make_T_dump.py 100 16 500 /tmp/config2_T.bin      # 3-D Laplacian 100^3, b=16, 500 block steps

Same recurrence as the oracle (`oracle/rbl_oracle.py::lanczos_iteration`, RBL.jl:74-117): local reorth every step, full
reorth of the two newest blocks every 2nd step; the Krylov blocks live in ONE preallocated fp32 slab (the device path's
mixed mode) so that n = 1e6 with 7900 columns fits 32 GB.  Test infrastructure only.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import matrices


def run(N, b, steps, out, seed=0, final_i=None):
    n = N ** 3
    A = matrices.shifted(matrices.laplacian_3d(N), 12.0).tocsr()
    rng = np.random.default_rng(seed)
    slab = np.zeros((steps * b, n), dtype=np.float32)        # row j*b+c = column c of Q_{j+1}
    hA = np.zeros((steps, b, b))
    hB = np.zeros((steps, b, b))

    def proj(X, m):
        """X -= Q[0:m] (Q[0:m]' X) against the first m stored blocks, fp32 products like the device's mixed mode."""
        if m <= 0:
            return
        S = slab[: m * b]
        G = S @ X.astype(np.float32)
        X -= (S.T @ G).astype(np.float64)

    Q_prev = None
    Q_cur = np.linalg.qr(A @ rng.standard_normal((n, b)))[0]
    Bi = None
    t0 = time.time()
    for i in range(1, steps + 1):
        slab[(i - 1) * b: i * b] = Q_cur.T
        if i >= 3 and i % 2 == 0:
            X = np.hstack([Q_cur, Q_prev])
            proj(X, i - 2)
            Q_cur, Q_prev = X[:, :b].copy(), X[:, b:].copy()
            slab[(i - 1) * b: i * b] = Q_cur.T
            slab[(i - 2) * b: (i - 1) * b] = Q_prev.T
        if Q_prev is not None:
            Q_cur -= Q_prev @ (Q_prev.T @ Q_cur)
        U = A @ Q_cur
        if Q_prev is not None:
            U -= Q_prev @ Bi.T
        Ai = Q_cur.T @ U
        U -= Q_cur @ Ai
        Qn, R = np.linalg.qr(U)
        sgn = np.sign(np.diag(R))
        sgn[sgn == 0] = 1
        Qn, R = Qn * sgn, R * sgn[:, None]
        hA[i - 1] = 0.5 * (Ai + Ai.T)
        hB[i - 1] = R
        Q_prev, Q_cur, Bi = Q_cur, Qn, R
        if i % 20 == 0:
            print(f"step {i}/{steps}  {time.time() - t0:.0f} s", flush=True)
    hdr = np.array([steps, b, b, final_i or steps], dtype=np.int64)
    with open(out, "wb") as f:
        f.write(hdr.tobytes())
        f.write(hA.tobytes())
        f.write(hB.tobytes())


if __name__ == "__main__":
    N, b, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    run(N, b, steps, sys.argv[4], final_i=int(sys.argv[5]) if len(sys.argv) > 5 else None)

"""Accuracy of the K5a Gram / K5b update arithmetic against fp64 and against NumPy float32 (what the oracle's FLOAT=Float32
mode and the reference's cuBLAS sgemm use), on Krylov-like data: orthonormal stored blocks, targets almost orthogonal to them.
    python tools/reorth_accuracy.py [n] [m] [b]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rbl_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13824
m = int(sys.argv[2]) if len(sys.argv) > 2 else 80
b = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rng = np.random.default_rng(1)
Qall = np.linalg.qr(rng.standard_normal((n, (m + 2) * b)))[0]
blocks = Qall[:, :m * b].reshape(n, m, b).transpose(1, 0, 2).copy()
Qm = Qall[:, :m * b]
W = Qall[:, m * b:] + 1e-6 * Qm @ rng.standard_normal((m * b, 2 * b))
W0, W1 = W[:, :b].copy(), W[:, b:].copy()
Q32 = Qm.astype(np.float32)
Cex = Q32.astype(np.float64).T @ W                      # exact Gram of what the buffer holds
C32 = (Q32.T @ W.astype(np.float32)).astype(np.float64)  # fp32 sgemm (OpenBLAS)
Wex = W - Q32.astype(np.float64) @ Cex
print(f"n={n} m={m} b={b}  |C| rms {np.sqrt(np.mean(Cex**2)):.2e}")
print(f"numpy float32 sgemm : C err rms {np.sqrt(np.mean((C32-Cex)**2)):.3e} max {np.max(np.abs(C32-Cex)):.3e}")
for impl in (1, 3, 4):
    w0, w1, C = rbl_b200.k_reorth(blocks, W0, W1, True, impl=impl)
    C = C.astype(np.float64)
    # impl 4 rounds the stored blocks to split16: compare with the Gram of THOSE values too
    err = C - Cex
    Wd = np.hstack([w0, w1])
    left = Q32.astype(np.float64).T @ Wd
    print(f"impl {impl}: C err rms {np.sqrt(np.mean(err**2)):.3e} max {np.max(np.abs(err)):.3e} mean {np.mean(err):+.2e} | "
          f"after update ||Q'W||_2 {np.linalg.norm(left, 2):.3e} max {np.max(np.abs(left)):.3e}")

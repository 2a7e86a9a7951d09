"""Replays the convergence checks of an oracle run through the stateful host checker (CPU only)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import rbl_b200
from oracle import rbl_oracle, matrices

def capture(A, k, b, Om, cap=4000):
    Ts, Bs, oks = [], [], []
    od, oc = rbl_oracle.dsbev, rbl_oracle.check_convergence
    def sd(T): Ts.append(T.copy()); return od(T)
    def sc(Bi, V, bb, kk, tol):
        ok = oc(Bi, V, bb, kk, tol); Bs.append(Bi.copy()); oks.append(ok); return ok
    rbl_oracle.dsbev, rbl_oracle.check_convergence = sd, sc
    try:
        Q = []
        Qi = np.linalg.qr(A @ Om)[0]
        rbl_oracle.lanczos_iteration(A, k, b, cap, Qi, Q)
    finally:
        rbl_oracle.dsbev, rbl_oracle.check_convergence = od, oc
    return Ts, Bs, oks

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    A = matrices.shifted(matrices.laplacian_3d(N), 12.0)
    Om = np.random.default_rng(0).standard_normal((N**3, b))
    Ts, Bs, oks = capture(A, k, b, Om)
    print("checks", len(Ts), "final N", Ts[-1].shape[1])
    ck = rbl_b200.Checker(threads=1)
    tot = 0; nfull = 0; t0 = time.time(); bad = 0
    for T, Bi, ok in zip(Ts, Bs, oks):
        r = ck.check(T, k, Bi)
        tot += r["factorizations"]; nfull += r["full"]
        if r["converged"] != ok: bad += 1
        print(T.shape[1], "fac", r["factorizations"], "full", r["full"], "conv", r["converged"], ok)
    print("total factorizations", tot, "full checks", nfull, "mismatches", bad, "time", round(time.time() - t0, 2))

"""Mirror of the reference's image low-rank demo (Julia/images.jl:10-42): rank-k approximation of an image B through the
eigenpairs of the DENSE operator B'B computed with block size 1,

    D, V = RBL_gpu(transpose(B)*B, k, 1);  U = (B*V) ./ transpose(D);  Blr = U*diagm(D)*transpose(V)      (images.jl:28-32)

compared with a truncated SVD (`svds`, images.jl:35-41).  The image is synthetic (no image files / codecs in this sandbox).
    python tools/images.py [--h 300] [--w 200] [--k 30]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def synthetic_image(h, w, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.5 + 0.3 * np.sin(xx / w * 9) * np.cos(yy / h * 7) + 0.2 * ((xx // 23 + yy // 31) % 2)
    return np.clip(img + 0.02 * rng.standard_normal((h, w)), 0, 1)


def low_rank(Bm, k, **kw):
    import rbl_b200
    M = Bm.T @ Bm                                                      # images.jl:28 (dense Matrix{Float64} operator)
    D, V, st = rbl_b200.RBL_gpu(M, k, 1, max_kryl_sz=max(1200, 8 * k), return_stats=True, **kw)
    s = np.sqrt(np.maximum(D, 0))                                      # singular values of B
    U = (Bm @ V) / s[None, :]
    return U, s, V, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=300)
    ap.add_argument("--w", type=int, default=200)
    ap.add_argument("--k", type=int, default=30)
    a = ap.parse_args()
    Bm = synthetic_image(a.h, a.w)
    t0 = time.perf_counter()
    U, s, V, st = low_rank(Bm, a.k)
    t = time.perf_counter() - t0
    Blr = (U * s[None, :]) @ V.T
    sv = np.linalg.svd(Bm, compute_uv=False)
    best = np.sqrt(np.sum(sv[a.k:] ** 2))
    print(f"RBL_gpu: {t * 1e3:.1f} ms, {st.iterations} iterations; ||B - Blr||_F = {np.linalg.norm(Bm - Blr):.6f} "
          f"(optimal rank-{a.k}: {best:.6f}); max rel singular value error {np.max(np.abs(s - sv[:a.k]) / sv[:a.k]):.2e}")


if __name__ == "__main__":
    main()

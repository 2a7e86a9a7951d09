"""Generates tests/golden/config1_full.json: BASELINE.json configs[0] AT FULL SIZE (2-D 5-point Laplacian on a
100 x 100 grid, n = 10^4, the 10 lowest eigenpairs as the largest of 8I - A, b = 4) solved by the CPU oracle
(oracle/rbl_oracle.py, the restatement of RBL.jl pinned to the reference's known-answer tests) in both precision
modes: FLOAT = Float64 as shipped (common.jl:5-6) and FLOAT = Float32 (README.md:69).  Stored: eigenvalues, block
iterations, max Ritz residual and the orthogonality loss ||Q'Q - I||_2 of the oracle's Krylov basis - the three
north-star gates the device path is held to in tests/test_gpu_parity_gates.py.

    python tests/golden/make_golden_config1.py        (about 2 minutes)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import matrices, rbl_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GRID, K, B, SIGMA, SEED, CAP = 100, 10, 4, 8.0, 1234, 1400


def problem():
    L = matrices.laplacian_2d(GRID)
    A = matrices.shifted(L, SIGMA)
    Om = np.random.default_rng(SEED).standard_normal((GRID * GRID, B))
    return L, A, Om


def main():
    L, A, Om = problem()
    out = {"grid": GRID, "k": K, "b": B, "sigma": SIGMA, "omega_seed": SEED, "max_kryl_sz": CAP,
           "analytic": [float(x) for x in (SIGMA - matrices.laplacian_eigs(GRID, 2, K))]}
    for mode, FLOAT in (("fp64", np.float64), ("mixed", np.float32)):
        D, V, det = rbl_oracle.RBL(A, K, B, Om, max_kryl_sz=CAP, FLOAT=FLOAT, return_details=True)
        st = det["stats"]
        nb = st.iterations
        out[mode] = {"D": [float(x) for x in D], "iterations": int(nb),
                     "max_ritz_residual_over_normA": float(np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=SIGMA))),
                     "orthogonality_loss_2norm": rbl_oracle.orthogonality_loss(det["Q"][:nb]),
                     "v_orthogonality_maxabs": float(np.max(np.abs(V.T.astype(np.float64) @ V.astype(np.float64) - np.eye(K))))}
        print(mode, out[mode]["iterations"], out[mode]["orthogonality_loss_2norm"], out[mode]["max_ritz_residual_over_normA"])
    with open(os.path.join(HERE, "config1_full.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

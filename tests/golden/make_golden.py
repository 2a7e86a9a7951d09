"""Generates tests/golden/*.json: small seeded cases of the hot path with the results the CPU oracle returns
(oracle/rbl_oracle.py, pinned to the reference's known-answer tests by tests/test_oracle_kat.py), plus the expected
eigenvalues of the reference's own fixtures (Julia/Unit Testing/{slow,mod,step}_dec.jl:3-6, analytic).

    python tests/golden/make_golden.py

The reference is Julia and cannot be executed here, so these vectors come from the restatement, not from the reference
binary; they freeze its behaviour (eigenvalues, iteration counts) for the CPU suite and give the GPU suite a target that
does not need the oracle at run time.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import matrices, rbl_oracle

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (matrix builder, shift, k, b, seed of Omega)
CASES = {
    "lap2d_30_k6_b4": (lambda: matrices.laplacian_2d(30), 8.0, 6, 4, 11),
    "lap2d_48_k6_b16": (lambda: matrices.laplacian_2d(48), 8.0, 6, 16, 7),
    "lap3d_12_k12_b8": (lambda: matrices.laplacian_3d(12), 12.0, 12, 8, 3),
    "er_2000_k8_b8": (lambda: matrices.erdos_renyi_sym(2000, 16, seed=5), 0.0, 8, 8, 9),
    "er_3000_k10_b32": (lambda: matrices.erdos_renyi_sym(3000, 24, seed=2), 0.0, 10, 32, 4),
}


def build(name):
    mk, shift, k, b, seed = CASES[name]
    L = mk()
    A = matrices.shifted(L, shift) if shift else L
    Om = np.random.default_rng(seed).standard_normal((L.shape[0], b))
    return L, A, shift, k, b, Om


def main():
    out = {}
    for name in CASES:
        L, A, shift, k, b, Om = build(name)
        D, V, det = rbl_oracle.RBL(A, k, b, Om, return_details=True)
        res = rbl_oracle.ritz_residuals(A, D, V)
        out[name] = {"n": int(L.shape[0]), "k": k, "b": b, "shift": shift, "omega_seed": CASES[name][4],
                     "D": [float(x) for x in D], "iterations": int(det["stats"].iterations),
                     "max_ritz_residual": float(np.max(res))}
        print(name, out[name]["iterations"], out[name]["D"][:3])
    with open(os.path.join(HERE, "oracle_small_cases.json"), "w") as f:
        json.dump(out, f, indent=1)
    kat = {}
    for gen_name, gen, sizes in (("slow_decay", matrices.slow_decay, (100, 300, 500, 700, 900)),
                                 ("moderate_decay", matrices.moderate_decay, (100, 300, 500, 700, 900)),
                                 ("step_decay", matrices.step_decay, (100000, 300000, 500000, 700000, 900000))):
        for n in sizes:
            _, eig = gen(n, 5)
            kat[f"{gen_name}_{n}"] = [float(x) for x in eig]
    with open(os.path.join(HERE, "reference_kat_eigenvalues.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()

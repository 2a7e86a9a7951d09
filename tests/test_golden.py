"""Committed golden vectors (tests/golden/*.json, written by tests/golden/make_golden.py).

CPU part: the oracle still reproduces them (a change of the restatement cannot pass silently), and the reference's
known-answer eigenvalues in the fixture file are the analytic ones the oracle is pinned to.
GPU part: the device path through the C ABI meets the north-star bars against the golden eigenvalues without calling
the oracle at run time."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import matrices, rbl_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


with open(os.path.join(GOLD, "oracle_small_cases.json")) as _f:
    CASES = json.load(_f)
with open(os.path.join(GOLD, "reference_kat_eigenvalues.json")) as _f:
    KAT = json.load(_f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    mk = _maker()
    L, A, shift, k, b, Om = mk.build(name)
    g = CASES[name]
    assert (g["n"], g["k"], g["b"]) == (L.shape[0], k, b)
    D, V, det = rbl_oracle.RBL(A, k, b, Om, return_details=True)
    assert det["stats"].iterations == g["iterations"]
    assert np.max(np.abs(D - np.array(g["D"])) / np.abs(np.array(g["D"]))) < 1e-12
    assert np.max(rbl_oracle.ritz_residuals(A, D, V)) < 1e-6


def test_reference_kat_fixture_is_the_analytic_spectrum():
    for gen_name, gen in (("slow_decay", matrices.slow_decay), ("moderate_decay", matrices.moderate_decay),
                          ("step_decay", matrices.step_decay)):
        keys = [k for k in KAT if k.startswith(gen_name + "_")]
        assert len(keys) == 5
        for key in keys:
            n = int(key.rsplit("_", 1)[1])
            _, eig = gen(n, 5)
            assert np.array_equal(np.array(KAT[key]), eig)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp64", "mixed"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_path_matches_golden(gpu, name, precision):
    mk = _maker()
    L, A, shift, k, b, Om = mk.build(name)
    g = CASES[name]
    kw = dict(shift=shift) if shift else {}
    cap = min(2400, L.shape[0])     # the golden runs never reach the reference's default cap of 1200; leave headroom
    D, V, st = gpu.RBL_gpu(L, k, b, Omega=Om, precision=precision, max_kryl_sz=cap, return_stats=True, **kw)
    Dg = np.array(g["D"])
    assert st.converged
    assert np.max(np.abs(D - Dg) / np.abs(Dg)) < 1e-8                 # north star: eigenvalues rel 1e-8
    assert abs(st.iterations - g["iterations"]) <= 8                   # same stopping rule, +- two checks
    norm_a = shift if shift else float(np.max(np.abs(Dg)))
    assert np.max(rbl_oracle.ritz_residuals(A, D, V, norm_a=norm_a)) < 1e-6

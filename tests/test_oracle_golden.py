"""The oracle (numpy restatement) replayed against every golden vector produced by the unmodified reference.

Runs on CPU (``-m "not gpu"``).  In the container that generated the goldens the match is bit for bit (same
numpy/scipy/OpenBLAS calls in the same order); across machines OpenBLAS may split ``ddot`` differently, so the
assertion is a tight relative tolerance for CG/MrR/small k and count-level for large k."""
import os

import numpy as np
import pytest

import krylov_oracle as oracle
from golden_util import CASES, history_tolerance, inputs, load


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    gold = load(case)
    mat, b = inputs(case)
    kwargs = {"tol": case["tol"], "maxiter": case["maxiter"]}
    if case["k"] is not None:
        kwargs["k"] = case["k"]
    x, info = oracle.SOLVERS[case["solver"]](mat, b.copy(), **kwargs)
    tol = history_tolerance(case)
    if tol is not None:
        assert np.array_equal(info["nosl"], gold["nosl"])
        np.testing.assert_allclose(info["residual"], gold["residual"], rtol=max(1e-11, tol[0] * 0.1), atol=tol[1])
        if "x" in gold:
            np.testing.assert_allclose(x, gold["x"], rtol=1e-9, atol=1e-12 * np.abs(gold["x"]).max())
        else:
            np.testing.assert_allclose(np.linalg.norm(x), gold["x_norm"], rtol=1e-10)
    else:
        assert abs(int(info["nosl"][-1]) - int(gold["nosl"][-1])) <= max(2, 0.05 * gold["nosl"][-1]) or \
            np.array_equal(info["nosl"], gold["nosl"])
    if "khistory" in gold and tol is not None:
        assert np.array_equal(info["khistory"], gold["khistory"])
    if case["final_residual"] < case["tol"]:
        assert oracle.true_relres(mat, b, x) < 1.05 * case["tol"]


def test_oracle_is_bitwise_on_generating_machine():
    """Same machine, same libraries ⇒ identical bits (guards against the restatement drifting)."""
    same = 0
    for case in CASES[:40]:
        gold = load(case)
        mat, b = inputs(case)
        kwargs = {"tol": case["tol"], "maxiter": case["maxiter"]}
        if case["k"] is not None:
            kwargs["k"] = case["k"]
        _, info = oracle.SOLVERS[case["solver"]](mat, b.copy(), **kwargs)
        if len(info["residual"]) == len(gold["residual"]) and np.array_equal(info["residual"], gold["residual"]):
            same += 1
    # OpenBLAS threading differs between hosts; require the overwhelming majority, not all.
    assert same >= 30, same


def test_c_restatement_of_csr_matvec_matches_scipy_bitwise():
    """oracle/oracle_kernels.c (plain-C csr_matvec, stencil generator) against scipy / problems.py."""
    import host_kernels as hk
    from parallel_krylov_b200 import problems
    for gen, args in ((problems.poisson3d, (9, 7, 5)), (problems.poisson2d, (31,)), (problems.banded_spd, (999, 13, 0))):
        rowptr, col, val, n = gen(*args)
        x = np.random.default_rng(0).standard_normal(n)
        assert np.array_equal(hk.csr_matvec(rowptr, col, val, x), problems.to_scipy(rowptr, col, val, n).dot(x))
    for dims in ((9, 7, 5), (31, 31, 1), (4, 1, 1)):
        r, c, v, n = hk.stencil_csr(*dims)
        hr, hc, hv, hn = problems.poisson3d(*dims) if dims[2] > 1 else problems._stencil_csr(dims[:2], 4.0)
        assert n == hn and np.array_equal(r, hr) and np.array_equal(c, hc) and np.array_equal(v, hv)


# ---- Chronopoulos-Gear CG: the oracle against the repaired reference text ------------------------
import json as _json

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cgcg_manifest.json")) as _fh:
    CGCG_CASES = _json.load(_fh)["cases"]


@pytest.mark.parametrize("case", CGCG_CASES, ids=[c["id"] for c in CGCG_CASES])
def test_cgcg_oracle_reproduces_the_repaired_reference_bitwise(case):
    from parallel_krylov_b200 import problems
    kind, args = case["matrix"]
    A = problems.to_scipy(*getattr(problems, kind)(*args))
    b = problems.rhs(A.shape[0], "randn", 0)
    M = A.diagonal().copy() if case["precond"] == "jacobi" else None
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case["id"] + ".npz")) as z:
        gold = z["residual"]
    x, info = oracle.cgcg(A, b.copy(), tol=1e-8, M=M)
    assert np.array_equal(info["residual"], gold)
    assert info["converged"] and int(info["nosl"][-1]) == case["iterations"]
    assert oracle.true_relres(A, b, x) < 1e-8 * (1 + 1e-6)

"""Opt-in Chebyshev basis of kskipmrr on the GPU: follows plain MrR — the reference's golden MrR
histories — to 1e-8 over the first 50 iterations at k = 8, 12, 16, where the reference's own monomial basis does not; the
default path is untouched (the golden parity tests cover it)."""
import os

import numpy as np
import pytest
import torch

import cheb_reference as cheb
import krylov_oracle as oracle
from golden_util import CASES, inputs, load
from parallel_krylov_b200 import problems

pytestmark = pytest.mark.gpu
os.environ.setdefault("PK_QUIET", "1")

_MRR = [c for c in CASES if c["solver"] == "mrr" and c["maxiter"] is None and c["tol"] == 1e-8
        and c["rhs"] == "randn" and c["matrix"] in ("p2d48", "p3d16", "p3d32", "p3d12x20x9", "p2d256")]


@pytest.mark.parametrize("k", [8, 12, 16])
@pytest.mark.parametrize("case", _MRR, ids=[c["id"] for c in _MRR])
def test_chebyshev_kskipmrr_matches_the_reference_mrr_history(case, k):
    import parallel_krylov_b200 as pk
    gold = load(case)                                   # plain MrR of the unmodified reference
    mat, b = inputs(case)
    x, info = pk.kskipmrr(mat, b, tol=1e-8, k=k, basis="chebyshev")
    nosl = info["nosl"].cpu().numpy()
    res = info["residual"].cpu().numpy()
    it_mrr = int(gold["nosl"][-1])
    assert info["converged"]
    assert it_mrr <= int(nosl[-1]) <= it_mrr + k + 1
    sel = nosl[nosl <= min(50, it_mrr)]
    np.testing.assert_allclose(res[:len(sel)], gold["residual"][sel], rtol=1e-8)
    assert oracle.true_relres(mat, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)


def test_chebyshev_matches_its_numpy_restatement_and_explicit_bounds():
    import parallel_krylov_b200 as pk
    A = problems.to_scipy(*problems.poisson3d(20))
    b = problems.rhs(A.shape[0], "randn", 0)
    xr, ir = cheb.kskipmrr_chebyshev(A, b, tol=1e-8, k=8)
    x, info = pk.kskipmrr(A, b, tol=1e-8, k=8, basis="chebyshev")
    assert np.array_equal(info["nosl"].cpu().numpy(), ir["nosl"])
    np.testing.assert_allclose(info["residual"].cpu().numpy(), ir["residual"], rtol=1e-9)
    lo, hi = cheb.gershgorin(A)
    x2, info2 = pk.kskipmrr(A, b, tol=1e-8, k=8, basis=("chebyshev", lo, hi), use_graph=False)
    assert torch.equal(info2["residual"], info["residual"]) and torch.equal(x2, x)
    with pytest.raises(pk.PkError):
        pk.mrr(A, b, basis="chebyshev")


_CGG = [c for c in CASES if c["solver"] == "cg" and c["maxiter"] is None and c["tol"] == 1e-8
        and c["rhs"] == "randn" and c["matrix"] in ("p2d48", "p3d16", "p3d32", "p3d12x20x9")]


@pytest.mark.parametrize("k", [4, 8, 12])
@pytest.mark.parametrize("case", _CGG, ids=[c["id"] for c in _CGG])
def test_chebyshev_kskipcg_matches_the_reference_cg_history(case, k):
    import parallel_krylov_b200 as pk
    gold = load(case)                                   # plain CG of the unmodified reference
    mat, b = inputs(case)
    x, info = pk.kskipcg(mat, b, tol=1e-8, k=k, basis="chebyshev")
    nosl = info["nosl"].cpu().numpy()
    res = info["residual"].cpu().numpy()
    it_cg = int(gold["nosl"][-1])
    assert info["converged"]
    assert it_cg <= int(nosl[-1]) <= it_cg + k + 1
    sel = nosl[nosl <= min(50, it_cg)]
    np.testing.assert_allclose(res[:len(sel)], gold["residual"][sel], rtol=1e-8)
    assert oracle.true_relres(mat, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)

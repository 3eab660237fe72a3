"""Blocks with 64-bit row pointers (nnz >= 2^31): applied as row segments of < 2^31 nonzeros with rebased 32-bit row
pointers (csrc/pk_solvers.cu: pk_mat_csr64).  Small matrices are forced through the same path with tiny segments
(PK_SEG_NNZ); one full-size case really exceeds 2^31 nonzeros."""
import os

import numpy as np
import pytest
import torch

import krylov_oracle as oracle
from parallel_krylov_b200 import problems

pytestmark = pytest.mark.gpu
os.environ.setdefault("PK_QUIET", "1")


def _wide(A):
    return (torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)),
            torch.from_numpy(A.data.astype(np.float64)), A.shape[0])


@pytest.mark.parametrize("name", ["p3d20", "band27", "ragged"])
def test_segmented_operator_is_bit_exact(monkeypatch, name):
    import parallel_krylov_b200 as pk
    monkeypatch.setenv("PK_SEG_NNZ", "3001")                      # odd size: segment bases are not multiples of 4
    if name == "p3d20":
        A = problems.to_scipy(*problems.poisson3d(20))
    elif name == "band27":
        A = problems.to_scipy(*problems.banded_spd(30011, 13, 0))
    else:
        import scipy.sparse as sp
        rng = np.random.default_rng(5)
        A = sp.random(5000, 5000, density=0.004, random_state=rng, format="csr") + sp.eye(5000, format="csr") * 3.0
        A = sp.csr_matrix(A)
        A.sort_indices()
    n = A.shape[0]
    op = pk.Operator.from_csr_tensors(*_wide(A))
    assert op.index64
    rng = np.random.default_rng(1)
    x0, x1, w = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    y = op.matvec(torch.from_numpy(x0))
    assert np.array_equal(y.cpu().numpy(), A.dot(x0))
    y0, y1 = op.matvec(torch.from_numpy(x0), x1=torch.from_numpy(x1))
    assert np.array_equal(y0.cpu().numpy(), A.dot(x0)) and np.array_equal(y1.cpu().numpy(), A.dot(x1))
    _, sums = op.matvec(torch.from_numpy(x0), dot_with=torch.from_numpy(w))
    ref = A.dot(x0)
    np.testing.assert_allclose(sums.cpu().numpy(), [np.dot(w, ref), np.dot(ref, ref), np.dot(w, w)], rtol=1e-13)


@pytest.mark.parametrize("solver,kw", [("cg", {}), ("mrr", {}), ("kskipcg", {"k": 2}), ("kskipmrr", {"k": 4}),
                                       ("adaptivekskipmrr", {"k": 4}), ("cgcg", {"M": "jacobi"})])
def test_solvers_on_segmented_operator(monkeypatch, solver, kw):
    import parallel_krylov_b200 as pk
    monkeypatch.setenv("PK_SEG_NNZ", "20000")
    A = problems.to_scipy(*problems.poisson3d(24))
    b = problems.rhs(A.shape[0], "randn", 0)
    okw = dict(kw)
    if "M" in okw:
        okw["M"] = A.diagonal().copy()
    xo, io = oracle.SOLVERS[solver](A, b.copy(), tol=1e-8, **okw)
    op = pk.Operator.from_csr_tensors(*_wide(A))
    x, info = getattr(pk, solver)(op, b, tol=1e-8, **kw)
    res = info["residual"].cpu().numpy()
    assert abs(int(info["nosl"][-1]) - int(io["nosl"][-1])) <= max(2, kw.get("k", 0) + 1)
    m = min(len(res), len(io["residual"]), int(np.searchsorted(io["nosl"], 50, side="right")))
    rtol, atol = (1e-10, 0.0) if kw.get("k", 0) <= 2 else (1e-8, 1e-11)
    np.testing.assert_allclose(res[:m], io["residual"][:m], rtol=rtol, atol=atol)
    assert oracle.true_relres(A, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)


def test_more_than_2_to_31_nonzeros():
    """3-D 7-point Poisson 700^3: n = 343 000 000 rows, nnz = 2 398 060 000 > 2^31, generated in HBM in three row pieces
    (each with 32-bit row pointers), concatenated under a 64-bit row pointer.  y = A x of the big operator must equal,
    bit for bit, what the three pieces give as separate operators; a short CG must report residuals that agree with
    b - A x recomputed through the operator."""
    import parallel_krylov_b200 as pk
    from parallel_krylov_b200 import device_problems as dp
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs ~70 GB of HBM")
    d = 700
    n = d ** 3
    cuts = [0, (n // 3) // 256 * 256 + 77, (2 * n // 3) // 256 * 256 + 130, n]      # deliberately not tile aligned
    pieces = [dp.stencil_csr(d, d, d, row0=cuts[i], n_rows=cuts[i + 1] - cuts[i]) for i in range(3)]
    x = dp.hash_normal(7, n)
    ys = []
    for rp, ci, va, _ in pieces:
        opp = pk.Operator.from_csr_tensors(rp, ci, va, n)
        ys.append(opp.matvec(x).clone())
        del opp
    nnz_piece = [int(p[2].numel()) for p in pieces]
    offs = np.concatenate([[0], np.cumsum(nnz_piece)])
    assert offs[-1] >= 2 ** 31
    rowptr = torch.cat([pieces[0][0].to(torch.int64)] + [pieces[i][0][1:].to(torch.int64) + int(offs[i]) for i in (1, 2)])
    col = torch.cat([p[1] for p in pieces])
    val = torch.cat([p[2] for p in pieces])
    del pieces
    torch.cuda.empty_cache()
    op = pk.Operator.from_csr_tensors(rowptr, col, val, n)
    assert op.index64 and op.nnz == int(offs[-1])
    y = op.matvec(x)
    for i in range(3):
        assert torch.equal(y[cuts[i]:cuts[i + 1]], ys[i]), f"piece {i}"
    del ys, y
    b = dp.hash_normal(0, n)
    xs, info = pk.cg(op, b, tol=1e-8, maxiter=8)
    r = b - op.matvec(xs)
    true = float(torch.linalg.norm(r) / torch.linalg.norm(b))
    assert info["iterations"] == 8
    assert abs(true - float(info["residual"][-1])) <= 1e-9 * true

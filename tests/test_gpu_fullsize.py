"""BASELINE.json configs at FULL size on the GPU: the oracle (CPU) is only run for the first iterations (seconds), the rest
is judged through size-independent properties — true residual of the returned x, linearity of the solve in b, iteration
counts consistent between solvers, determinism."""
import os

import numpy as np
import pytest
import torch

import krylov_oracle as oracle
import host_kernels as hk
from parallel_krylov_b200 import problems

pytestmark = pytest.mark.gpu
os.environ.setdefault("PK_QUIET", "1")


@pytest.fixture(scope="module")
def pk():
    import parallel_krylov_b200 as pk
    return pk


def _true_relres_gpu(op, b_t, x_t):
    r = b_t - op.matvec(x_t)
    return float(torch.linalg.norm(r) / torch.linalg.norm(b_t))


def test_config1_mrr_poisson3d_128(pk):
    """configs[1]: v3 mrr on 3-D 7-point Poisson 128^3 (n = 2 097 152)."""
    from parallel_krylov_b200 import device_problems as dp
    rowptr, col, val, n = dp.stencil_csr(128, 128, 128)
    op = pk.Operator.from_csr_tensors(rowptr, col, val, n)
    b = dp.hash_normal(0, n)
    x, info = pk.mrr(op, b, tol=1e-8)
    assert info["converged"] and 300 < info["iterations"] < 600
    assert _true_relres_gpu(op, b, x) < 1e-8 * (1 + 1e-6)
    # first 30 iterations against the oracle on the same system (host copy of the same arrays)
    A = problems.to_scipy(rowptr.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy(), n)
    _, io = oracle.mrr(A, b.cpu().numpy(), tol=1e-8, maxiter=30)
    np.testing.assert_allclose(info["residual"][:31].cpu().numpy(), io["residual"], rtol=1e-10)
    # determinism and linearity
    x2, info2 = pk.mrr(op, b, tol=1e-8)
    assert torch.equal(x, x2) and torch.equal(info["residual"], info2["residual"])
    x3, _ = pk.mrr(op, -2.5 * b, tol=1e-8)
    assert float(torch.linalg.norm(x3 + 2.5 * x) / torch.linalg.norm(x)) < 1e-6


def test_config2_kskipcg_k4_poisson3d_256(pk):
    """configs[2]: kskipcg k=4 on 3-D 7-point Poisson 256^3 (n = 16 777 216), matrix-powers (two-chain) basis."""
    from parallel_krylov_b200 import device_problems as dp
    rowptr, col, val, n = dp.stencil_csr(256, 256, 256)
    op = pk.Operator.from_csr_tensors(rowptr, col, val, n)
    b = dp.hash_normal(0, n)
    x, info = pk.kskipcg(op, b, tol=1e-8, k=4)
    xc, infoc = pk.cg(op, b, tol=1e-8)
    assert info["converged"] and infoc["converged"]
    # k-skip CG is CG in exact arithmetic: same count up to one trip / 5 %
    assert abs(info["iterations"] - infoc["iterations"]) <= max(5, 0.05 * infoc["iterations"])
    assert _true_relres_gpu(op, b, x) < 1e-8 * (1 + 1e-6)
    assert _true_relres_gpu(op, b, xc) < 1e-8 * (1 + 1e-6)
    # first two outer trips (10 iterations) against the oracle; host matrix from the C generator (same matrix)
    hr, hc, hv, hn = hk.stencil_csr(256, 256, 256)
    assert hn == n and np.array_equal(hr[:1000], rowptr[:1000].cpu().numpy())
    A = problems.to_scipy(hr, hc, hv, hn)
    _, io = oracle.kskipcg(A, b.cpu().numpy(), tol=1e-8, maxiter=10, k=4)
    m = len(io["residual"])
    np.testing.assert_allclose(info["residual"][:m].cpu().numpy(), io["residual"], rtol=1e-8, atol=1e-11)


def test_config3_kskipmrr_k8_banded_single_gpu_slice(pk):
    """configs[3] (banded SPD, 27 diagonals) at 2^23 rows on one GPU: k-skip MrR k=8 converges to the true residual and
    matches plain MrR's iteration count within one trip."""
    from parallel_krylov_b200 import device_problems as dp
    rowptr, col, val, n = dp.banded_csr(1 << 23, 13, 0)
    op = pk.Operator.from_csr_tensors(rowptr, col, val, n)
    b = dp.hash_normal(0, n)
    x, info = pk.kskipmrr(op, b, tol=1e-8, k=8)
    xm, infom = pk.mrr(op, b, tol=1e-8)
    assert info["converged"] and infom["converged"]
    assert abs(info["iterations"] - infom["iterations"]) <= 9
    assert _true_relres_gpu(op, b, x) < 1e-8 * (1 + 1e-6)

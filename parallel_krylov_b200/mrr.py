"""``mrr(A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None) -> (x, info)`` — drop-in for
/root/reference/v3/gpu/mrr.py:8 (same argument meaning; M, callback and atol are accepted and ignored exactly as the
reference ignores them).  x and info['nosl'/'residual'] are torch CUDA tensors (the reference returns cupy arrays)."""
from ._core import solve


def mrr(A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve("mrr", A, b, x=x, tol=tol, maxiter=maxiter, **kw)

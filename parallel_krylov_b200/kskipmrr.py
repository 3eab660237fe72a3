"""``kskipmrr(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None) -> (x, info)`` — drop-in for
/root/reference/v3/gpu/kskipmrr.py (same argument meaning; M, callback and atol are accepted and ignored exactly as the
reference ignores them).  Unlike the reference's CPU variant (numpy.dot(A, v): dense A only) sparse A is accepted."""
from ._core import solve


def kskipmrr(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve("kskipmrr", A, b, x=x, tol=tol, maxiter=maxiter, k=k, **kw)

"""``kskipcg(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None) -> (x, info)`` — drop-in for
/root/reference/v3/gpu/kskipcg.py (same argument meaning; M, callback and atol are accepted and ignored exactly as the
reference ignores them).  Unlike the reference's CPU variant (numpy.dot(A, v): dense A only) sparse A is accepted.

Opt-in, beyond the reference: ``basis="chebyshev"`` (or ``("chebyshev", lam_lo, lam_hi)``) builds the
trips on T_j((A - d)/c) r / p instead of A^j r / p — follows plain CG at k = 8, 12 where the monomial basis does not."""
from ._core import solve


def kskipcg(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve("kskipcg", A, b, x=x, tol=tol, maxiter=maxiter, k=k, **kw)

"""``kskipmrr(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None[, basis=None]) -> (x, info)`` — drop-in for
/root/reference/v3/gpu/kskipmrr.py (same argument meaning; M, callback and atol are accepted and ignored exactly as the
reference ignores them).  Unlike the reference's CPU variant (numpy.dot(A, v): dense A only) sparse A is accepted.

Opt-in, beyond the reference: ``basis="chebyshev"`` builds the trips on T_j((A - d)/c) r instead of
A^j r (Gershgorin bounds of the spectrum, or ``basis=("chebyshev", lam_lo, lam_hi)``) — same iterates as MrR in exact
arithmetic, and in fp64 too at k = 8, 12, 16, where the monomial basis of the reference has lost the history."""
from ._core import solve


def kskipmrr(A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve("kskipmrr", A, b, x=x, tol=tol, maxiter=maxiter, k=k, **kw)

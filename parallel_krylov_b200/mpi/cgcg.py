"""``cgcg(comm, local_A, b, x=None, tol=1e-05, maxiter=None, M=None, ...) -> (x, info)`` — row-partitioned
Chronopoulos–Gear CG (see parallel_krylov_b200/cgcg.py): one all-reduce per iteration instead of the two of ``cg``."""
from ._dist import solve_dist


def cgcg(comm, local_A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve_dist("cgcg", comm, local_A, b, x=x, tol=tol, maxiter=maxiter, M=M, **kw)

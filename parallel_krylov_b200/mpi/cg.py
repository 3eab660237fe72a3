"""``cg(comm, local_A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None) -> (x, info)`` — drop-in for
/root/reference/v3/gpu/mpi/cg.py:10 with a torch.distributed process group as ``comm`` (None = WORLD) and
``local_A`` this rank's contiguous row block (local_N x N; scipy CSR, ndarray, torch, or a DistOperator)."""
from ._dist import solve_dist


def cg(comm, local_A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve_dist("cg", comm, local_A, b, x=x, tol=tol, maxiter=maxiter, **kw)

"""``kskipmrr(comm, local_A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None) -> (x, info)`` —
drop-in for /root/reference/v3/gpu/mpi/kskipmrr.py:10 with a torch.distributed process group as ``comm`` (None = WORLD)
and ``local_A`` this rank's contiguous row block."""
from ._dist import solve_dist


def kskipmrr(comm, local_A, b, x=None, tol=1e-05, maxiter=None, k=0, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve_dist("kskipmrr", comm, local_A, b, x=x, tol=tol, maxiter=maxiter, k=k, **kw)

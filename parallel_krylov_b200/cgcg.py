"""``cgcg(A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None) -> (x, info)`` — Chronopoulos–Gear CG:
conjugate gradients with ONE reduction point per iteration (r.r, r.u and u.w are reduced together in the epilogue of the
SpMV), optionally Jacobi-preconditioned.  SURVEY.md §8f rank 4; restates the sketch
/root/reference/v1/threads/pipeline/chronopoulos_gear.py:7-56 (which cannot be imported and never updates ``old_gamma``;
both repaired, see oracle/gen_golden_cgcg.py) behind the v3 calling convention.

Opt-in: its recurrences round differently from ``cg`` (same iterates in exact arithmetic), so it is not a drop-in for a
v3 method; it halves the all-reduces per iteration, which is what limits ``cg`` at 8 GPUs.
``M``: None, ``"jacobi"`` (diag(A)), or the diagonal of M as a 1-D array (u = r / M)."""
from ._core import solve


def cgcg(A, b, x=None, tol=1e-05, maxiter=None, M=None, callback=None, atol=None, **kw) -> tuple:
    return solve("cgcg", A, b, x=x, tol=tol, maxiter=maxiter, M=M, **kw)

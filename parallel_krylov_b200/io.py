"""Input adapters: the file formats the reference's drivers fed to the solvers —
``*.npy`` (dense A or a vector), ``*.npz`` (scipy sparse), ``*.mtx`` (Matrix Market) — see the ignore list
/root/reference/.gitignore:14-17 and the "npy"/"npz" comments in /root/reference/v3/gpu/mpi/common.py:123-127.
Everything returned here is accepted by the solver entry points and by ``Operator.from_any``."""
from __future__ import annotations

import os

import numpy as np


def load_matrix(path: str):
    """Dense ``ndarray`` for .npy, scipy CSR for .npz / .mtx."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path)
        if a.ndim != 2:
            raise ValueError(f"{path}: expected a 2-D array, got shape {a.shape}")
        return np.ascontiguousarray(a, dtype=np.float64)
    if ext == ".npz":
        import scipy.sparse as sp
        return sp.load_npz(path).tocsr().astype(np.float64)
    if ext in (".mtx", ".mm"):
        import scipy.io
        import scipy.sparse as sp
        m = scipy.io.mmread(path)
        return sp.csr_matrix(m, dtype=np.float64) if sp.issparse(m) else np.ascontiguousarray(m, dtype=np.float64)
    raise ValueError(f"{path}: unsupported matrix file type {ext!r} (.npy, .npz, .mtx)")


def load_vector(path: str) -> np.ndarray:
    v = np.load(path) if path.lower().endswith(".npy") else np.loadtxt(path)
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())


def row_block(A, rank: int, world: int):
    """The contiguous row block rank ``rank`` of ``world`` owns (reference partition: N // world rows each,
    /root/reference/v3/gpu/mpi/common.py:104-131; here the last rank also takes the remainder)."""
    n = A.shape[0]
    base = n // world
    lo = rank * base
    hi = n if rank == world - 1 else lo + base
    return A[lo:hi]

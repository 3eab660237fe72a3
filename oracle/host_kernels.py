"""ctypes access to oracle/_build/liboracle.so (CPU oracle helpers — TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.pko_stencil_rowptr.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def csr_matvec(rowptr, col, val, x):
    """y = A x with the plain-C restatement of scipy's csr_matvec."""
    n = len(rowptr) - 1
    y = np.zeros(n, dtype=np.float64)
    lib().pko_csr_matvec(C.c_int64(n), _p(rowptr), _p(col), _p(val), _p(np.ascontiguousarray(x)), _p(y))
    return y


def stencil_csr(nx, ny, nz=1):
    """(rowptr, col, val, n) of the 5-/7-point Laplacian, identical to problems.poisson2d/3d, built in C."""
    n = nx * ny * nz
    rowptr = np.empty(n + 1, dtype=np.int32)
    nnz = lib().pko_stencil_rowptr(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), _p(rowptr))
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    lib().pko_stencil_fill(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), _p(rowptr), _p(col), _p(val))
    return rowptr, col, val, n

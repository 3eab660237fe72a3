"""TEST / MEASUREMENT INFRASTRUCTURE ONLY — the reference's own CPU files, staged for the GPU box.

The reference is pure Python, so "building" it means nothing more than staging the files of the path
(``v3/common.py``, ``v3/cpu/{common,cg,mrr,kskipcg,kskipmrr,adaptivekskipmrr}.py`` and the package ``__init__``s) from
``/root/reference`` into the git-ignored ``oracle/_ref/`` — done by ``__graft_entry__.build()`` in the build container,
where /root/reference exists.  ``oracle/_ref/`` is NOT gpurun-ignored, so it travels to the GPU box like a built
``.so``; nothing of it enters the git history.  ``bench.py --impl reference`` and bench.py's ``cpu_baseline`` leg then
time the UNMODIFIED reference functions (``kind: "reference"``) and fall back to the oracle port
(``oracle/krylov_oracle.py``, ``kind: "port"``) only when ``oracle/_ref`` is absent.

Shims applied at load time (the staged files are byte-identical to the reference):
  1. ``numpy.int = int``  (``np.int`` is used at v3/cpu/common.py:34; removed from numpy >= 1.24);
  2. ``sys.path`` gets ``oracle/_ref`` so that ``import v3.cpu.cg`` resolves the package-relative imports;
  3. ``kskipcg`` / ``adaptivekskipmrr`` call ``numpy.dot(A, v)`` (v3/cpu/kskipcg.py:21,37): a sparse ``A`` is wrapped in a
     proxy that routes ``np.dot(proxy, v)`` to the scipy ``csr_matvec`` the other solvers reach through ``A.dot(v)``.
"""
from __future__ import annotations

import contextlib
import io
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ["v3/__init__.py", "v3/common.py", "v3/cpu/__init__.py", "v3/cpu/common.py", "v3/cpu/cg.py", "v3/cpu/mrr.py",
         "v3/cpu/kskipcg.py", "v3/cpu/kskipmrr.py", "v3/cpu/adaptivekskipmrr.py"]
SOLVER_NAMES = ("cg", "mrr", "kskipcg", "kskipmrr", "adaptivekskipmrr")


def stage(src: str = "/root/reference") -> bool:
    """Copy the path's files from the reference tree into oracle/_ref (byte for byte).  False if there is no reference."""
    if not os.path.isdir(os.path.join(src, "v3", "cpu")):
        return False
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, rel)) for rel in FILES)


class DotProxy:
    """Lets ``numpy.dot(A, v)`` reach a scipy CSR mat-vec (shim 3)."""

    def __init__(self, mat):
        self.mat = mat
        self.shape = mat.shape

    def dot(self, v):
        return self.mat.dot(v)

    def __array_function__(self, func, types, args, kwargs):
        if func is np.dot and args and args[0] is self:
            return self.mat.dot(args[1])
        return NotImplemented


def load():
    """{name: callable(A, b, x=None, tol=..., maxiter=..., [k=...]) -> (x, info)} running the staged reference files,
    banner prints swallowed; None when oracle/_ref is absent."""
    if not available():
        return None
    if not hasattr(np, "int"):
        np.int = int                       # shim 1
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)        # shim 2
    import importlib
    fns = {}
    for name in SOLVER_NAMES:
        mod = importlib.import_module(f"v3.cpu.{name}")
        fns[name] = getattr(mod, name)

    def wrap(name):
        def call(A, b, x=None, **kw):
            a_arg = A
            if name in ("kskipcg", "adaptivekskipmrr") and not isinstance(A, np.ndarray):
                a_arg = DotProxy(A)        # shim 3
            with contextlib.redirect_stdout(io.StringIO()):
                return fns[name](a_arg, b, x, **kw)
        return call

    return {name: wrap(name) for name in SOLVER_NAMES}

"""Golden residual histories for the Chronopoulos–Gear CG entry point, produced by executing the
REFERENCE TEXT of /root/reference/v1/threads/pipeline/chronopoulos_gear.py with the two repairs it needs to run at all:

  1. its import ``from .common import start, end as finish, init`` points at a file that does not exist
     (v1/threads/pipeline/common.py); the names are bound to /root/reference/v1/threads/common.py's ``init`` (shim
     ``numpy.int = int``) and to silent ``start`` / ``finish``;
  2. ``old_gamma`` is assigned once before the loop (:31) and never again, so ``beta = gamma/old_gamma`` (:49) divides by
     the INITIAL gamma for ever; the line ``old_gamma = gamma`` is inserted after the ``alpha`` update (:50).

Nothing else is touched; the preconditioner argument ``ilu`` is an object whose ``solve(r)`` returns ``r / diag(A)``
(Jacobi) or ``r.copy()`` (none).  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/gen_golden_cgcg.py        # writes tests/golden/cgcg_*.npz + cgcg_manifest.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PK_REFERENCE", "/root/reference")

from parallel_krylov_b200 import problems  # noqa: E402

CASES = [("p2d48", "poisson2d", (48,), None), ("p3d16", "poisson3d", (16,), None),
         ("band27_20k", "banded_spd", (20000, 13, 0), None), ("band27_20k", "banded_spd", (20000, 13, 0), "jacobi"),
         ("band5_777", "banded_spd", (777, 2, 3), "jacobi"), ("p3d12x20x9", "poisson3d", (12, 20, 9), "jacobi")]


def load_fixed_reference():
    if not hasattr(np, "int"):
        np.int = int
    src = open(os.path.join(REF, "v1", "threads", "pipeline", "chronopoulos_gear.py")).read()
    bad_import = "from .common import start, end as finish, init\n"
    assert bad_import in src
    src = src.replace(bad_import, "")                                                       # repair 1
    line = "        alpha = gamma/(delta - beta*gamma/alpha)\n"
    assert src.count(line) == 1
    src = src.replace(line, line + "        old_gamma = gamma\n")                          # repair 2
    common = open(os.path.join(REF, "v1", "threads", "common.py")).read()
    common = common.replace("from ..common import _start, _end\n", "")
    ns = {}
    exec(compile(common, "v1/threads/common.py", "exec"), ns)
    env = {"init": ns["init"], "start": lambda method_name="", k=None: 0.0,
           "finish": lambda *a, **k: 0.0}
    exec(compile(src, "v1/threads/pipeline/chronopoulos_gear.py (2 repairs)", "exec"), env)
    return env["chronopoulos_gear"]


class Precond:
    def __init__(self, d):
        self.d = d

    def solve(self, r):
        return r.copy() if self.d is None else r / self.d


def main():
    fn = load_fixed_reference()
    out = os.path.join(ROOT, "tests", "golden")
    manifest = []
    for mname, kind, args, pre in CASES:
        A = problems.to_scipy(*getattr(problems, kind)(*args))
        n = A.shape[0]
        b = problems.rhs(n, "randn", 0)
        d = A.diagonal().copy() if pre == "jacobi" else None
        # the reference calls numpy.dot(A, x): give it a dense ndarray when small, else a proxy routing np.dot to scipy
        from gen_golden import DotProxy
        _, nosl, residual = fn(DotProxy(A), b.copy(), Precond(d), 1e-8)
        cid = f"cgcg__{mname}__{pre or 'none'}"
        np.savez_compressed(os.path.join(out, f"{cid}.npz"), residual=np.asarray(residual, dtype=np.float64),
                            iterations=np.int64(len(residual) - 1))
        manifest.append({"id": cid, "matrix": [kind, list(args)], "precond": pre, "iterations": int(len(residual) - 1),
                         "final_residual": float(residual[-1])})
        print(f"{cid:40s} it={len(residual) - 1:4d} res={residual[-1]:.3e}")
    with open(os.path.join(out, "cgcg_manifest.json"), "w") as fh:
        json.dump({"generator": "oracle/gen_golden_cgcg.py",
                   "reference": "v1/threads/pipeline/chronopoulos_gear.py with the two documented repairs", "cases": manifest}, fh, indent=1)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    main()

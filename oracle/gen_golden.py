"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference (``/root/reference/v3/cpu``).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/golden_<case>.npz + manifest.json

Shims applied before import — the reference is not modified:
  1. ``numpy.int = int``  (``np.int`` is used at v3/cpu/common.py:34 and was removed from numpy ≥ 1.24);
  2. ``sys.path`` gets /root/reference so that ``import v3.cpu.cg`` resolves the package-relative imports;
  3. for ``kskipcg`` / ``adaptivekskipmrr``, which call ``numpy.dot(A, v)`` (v3/cpu/kskipcg.py:21,37;
     v3/cpu/adaptivekskipmrr.py:22,48,78) a sparse ``A`` is wrapped in a proxy that routes
     ``np.dot(proxy, v)`` to the same scipy ``csr_matvec`` the other solvers reach through ``A.dot(v)``.
The reference prints a banner per solve; stdout is swallowed.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PK_REFERENCE", "/root/reference")

from parallel_krylov_b200 import problems  # noqa: E402  (input generators only)


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


def load_reference():
    if not hasattr(np, "int"):
        np.int = int  # shim 1
    if REF not in sys.path:
        sys.path.insert(0, REF)  # shim 2
    from v3.cpu.cg import cg
    from v3.cpu.mrr import mrr
    from v3.cpu.kskipcg import kskipcg
    from v3.cpu.kskipmrr import kskipmrr
    from v3.cpu.adaptivekskipmrr import adaptivekskipmrr
    return {"cg": cg, "mrr": mrr, "kskipcg": kskipcg, "kskipmrr": kskipmrr,
            "adaptivekskipmrr": adaptivekskipmrr}


def run_reference(fn_name, fns, mat, b, tol, maxiter, k=None):
    a_arg = mat
    if fn_name in ("kskipcg", "adaptivekskipmrr") and not isinstance(mat, np.ndarray):
        a_arg = DotProxy(mat)
    kwargs = {"tol": tol, "maxiter": maxiter}
    if k is not None:
        kwargs["k"] = k
    with contextlib.redirect_stdout(io.StringIO()):
        x, info = fns[fn_name](a_arg, b.copy(), **kwargs)
    return x, info


# (case name, matrix spec, rhs kind) — sizes the reference finishes in seconds.
MATRICES = {
    "p2d16": ("poisson2d", (16,)),
    "p2d48": ("poisson2d", (48,)),
    "p2d256": ("poisson2d", (256,)),          # BASELINE.json configs[0]
    "p3d16": ("poisson3d", (16,)),
    "p3d32": ("poisson3d", (32,)),
    "p3d12x20x9": ("poisson3d", (12, 20, 9)),  # ragged box: row counts not a multiple of any tile
    "band27_20k": ("banded_spd", (20000, 13, 0)),
    "band5_777": ("banded_spd", (777, 2, 3)),
    "dense256": ("dense_spd", (256, 0)),
    # the two systems bench.py solves through parallel_krylov_b200.mpi.* at every GPU count before its timed region
    "p3d48": ("poisson3d", (48,)),
    "band27_100k": ("banded_spd", (100000, 13, 0)),
}

# solver, k
SOLVER_SET = [("cg", None), ("mrr", None), ("kskipcg", 1), ("kskipcg", 2), ("kskipcg", 4),
              ("kskipmrr", 1), ("kskipmrr", 2), ("kskipmrr", 4), ("kskipmrr", 8),
              ("adaptivekskipmrr", 2), ("adaptivekskipmrr", 4), ("adaptivekskipmrr", 8)]

CASES = []
for mname in ("p2d16", "p2d48", "p3d16", "p3d32", "p3d12x20x9", "band27_20k", "band5_777", "dense256"):
    for rhs in ("randn", "ones"):
        if rhs == "ones" and mname not in ("p2d48", "p3d16"):
            continue
        for solver, k in SOLVER_SET:
            CASES.append((mname, rhs, solver, k, 1e-8, None))
# configs[0]: cg/mrr on 2-D 256² (b=ones and randn(0)) — the counts quoted in BASELINE.md §2
for rhs in ("ones", "randn"):
    for solver in ("cg", "mrr"):
        CASES.append(("p2d256", rhs, solver, None, 1e-8, None))
CASES.append(("p2d256", "randn", "kskipmrr", 4, 1e-8, None))
CASES.append(("p2d256", "randn", "kskipcg", 2, 1e-8, None))
# iteration-cap (non-convergence) semantics: maxiter hit, k-skip overshoot by up to k
CASES.append(("p2d48", "randn", "cg", None, 1e-8, 20))
CASES.append(("p2d48", "randn", "mrr", None, 1e-8, 20))
CASES.append(("p2d48", "randn", "kskipcg", 4, 1e-8, 22))
CASES.append(("p2d48", "randn", "kskipmrr", 4, 1e-8, 22))
CASES.append(("p2d48", "randn", "adaptivekskipmrr", 4, 1e-8, 22))
# residual-growth guard of adaptivekskipmrr actually firing (rollback + k lowered): needs k >= 10 on these systems
CASES.append(("p2d48", "randn", "adaptivekskipmrr", 12, 1e-8, 2000))
CASES.append(("p2d48", "randn", "adaptivekskipmrr", 16, 1e-8, 2000))
CASES.append(("p3d16", "randn", "adaptivekskipmrr", 12, 1e-8, 2000))
# bench.py's driver-visible parity set (N = 1, 2, 4, 8 GPUs): five solvers on two systems
for mname in ("p3d48", "band27_100k"):
    for solver, k in (("cg", None), ("mrr", None), ("kskipcg", 2), ("kskipmrr", 4), ("adaptivekskipmrr", 4)):
        CASES.append((mname, "randn", solver, k, 1e-8, None))
# loose tolerance: converges at the very first check
CASES.append(("p3d16", "randn", "cg", None, 10.0, None))
CASES.append(("p3d16", "randn", "mrr", None, 10.0, None))
CASES.append(("p3d16", "randn", "kskipmrr", 2, 10.0, None))


def build_matrix(mname):
    kind, args = MATRICES[mname]
    if kind == "dense_spd":
        return problems.dense_spd(*args)
    rowptr, col, val, n = getattr(problems, kind)(*args)
    return problems.to_scipy(rowptr, col, val, n)


def case_id(mname, rhs, solver, k, tol, maxiter):
    s = f"{solver}" + (f"_k{k}" if k is not None else "") + f"__{mname}__{rhs}"
    if tol != 1e-8:
        s += f"__tol{tol:g}"
    if maxiter is not None:
        s += f"__cap{maxiter}"
    return s


def main():
    fns = load_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    manifest = []
    mats = {}
    for mname, rhs, solver, k, tol, maxiter in CASES:
        if mname not in mats:
            mats[mname] = build_matrix(mname)
        mat = mats[mname]
        n = mat.shape[0]
        b = problems.rhs(n, rhs, 0)
        x, info = run_reference(solver, fns, mat, b, tol, maxiter, k)
        cid = case_id(mname, rhs, solver, k, tol, maxiter)
        true_res = float(np.linalg.norm(b - mat.dot(x)) / np.linalg.norm(b))
        payload = {
            "residual": np.asarray(info["residual"], dtype=np.float64),
            "nosl": np.asarray(info["nosl"], dtype=np.int64),
            "true_relres": np.float64(true_res),
            "x_norm": np.float64(np.linalg.norm(x)),
            "x_sum": np.float64(np.sum(x)),
        }
        if n <= 8192:
            payload["x"] = np.asarray(x, dtype=np.float64)
        else:
            payload["x_sample"] = np.asarray(x[:: max(1, n // 2048)], dtype=np.float64)
        if "khistory" in info:
            payload["khistory"] = np.asarray(info["khistory"], dtype=np.int64)
        np.savez_compressed(os.path.join(outdir, f"golden_{cid}.npz"), **payload)
        manifest.append({"id": cid, "matrix": mname, "rhs": rhs, "solver": solver, "k": k, "tol": tol,
                         "maxiter": maxiter, "iterations": int(info["nosl"][-1]),
                         "entries": int(len(info["residual"])),
                         "final_residual": float(info["residual"][-1]), "true_relres": true_res})
        print(f"{cid:60s} it={int(info['nosl'][-1]):5d} entries={len(info['residual']):4d} "
              f"res={info['residual'][-1]:.3e} true={true_res:.3e}")
    with open(os.path.join(outdir, "manifest.json"), "w") as fh:
        json.dump({"generator": "oracle/gen_golden.py", "reference": "5enxia/parallel-krylov v3/cpu (unmodified)",
                   "numpy": np.__version__, "matrices": {k: list(map(str, v)) for k, v in MATRICES.items()},
                   "cases": manifest}, fh, indent=1)
    print(f"{len(manifest)} cases written to {outdir}")


if __name__ == "__main__":
    main()

"""Helpers shared by the CPU and GPU parity tests: load golden fixtures, rebuild their inputs."""
import json
import os

import numpy as np

from parallel_krylov_b200 import problems

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN_DIR, "manifest.json")) as _fh:
    MANIFEST = json.load(_fh)
CASES = MANIFEST["cases"]
_MATRICES = MANIFEST["matrices"]
_cache = {}


def build_matrix(mname):
    """(scipy CSR | ndarray) for a manifest matrix name — same generators gen_golden.py used."""
    if mname not in _cache:
        kind, args = _MATRICES[mname][0], eval(_MATRICES[mname][1])
        if kind == "dense_spd":
            _cache[mname] = problems.dense_spd(*args)
        else:
            _cache[mname] = problems.to_scipy(*getattr(problems, kind)(*args))
    return _cache[mname]


def load(case):
    with np.load(os.path.join(GOLDEN_DIR, f"golden_{case['id']}.npz")) as z:
        return {k: z[k] for k in z.files}


def inputs(case):
    mat = build_matrix(case["matrix"])
    b = problems.rhs(mat.shape[0], case["rhs"], 0)
    return mat, b


def history_tolerance(case):
    """(rtol, atol) on the residual history over the first 50 solver iterations, or None.

    BASELINE.json north_star: 1e-10 relative.  That holds for cg / mrr / k <= 2 (measured on B200: <= 1e-11).  The
    k-skip recurrences rebuild every step's scalars from Gram sums of an unscaled monomial basis, so a different
    (equally valid) summation order moves the history by ~1e-9 relative at k = 4 and by O(1) at k = 8 — the
    reference disagrees with ITSELF by that much when only its mat-vec summation order changes.
    k = 3..4: 1e-8 relative plus 1e-11 absolute (residuals are relative to ||b||, i.e. start at 1);
    k >= 5: judged on iteration count (+-1 trip / 5 %), the final true residual, and (tests/test_gpu_solvers.py) the opening
    step plus the first two trips of the history at 1e-6."""
    k = case["k"] or 0
    if case["solver"] in ("cg", "mrr") or k <= 2:
        return (1e-10, 0.0)
    if k <= 4:
        return (1e-8, 1e-11)
    return None

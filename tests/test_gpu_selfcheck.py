"""The golden-vector self-check bench.py runs before its timed region (parallel_krylov_b200/selfcheck.py), on one GPU
through the row-partitioned entry points, plus the structural validation of CSR inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_selfcheck_single_rank():
    r = subprocess.run([sys.executable, "-m", "parallel_krylov_b200.selfcheck"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "SELFCHECK OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_malformed_csr_is_rejected():
    import parallel_krylov_b200 as pk
    from parallel_krylov_b200 import problems
    rowptr, col, val, n = problems.poisson2d(12)
    b = np.ones(n)
    bad_col = col.copy()
    bad_col[5] = n + 3                                             # column outside [0, n)
    with pytest.raises(pk.PkError, match="column index"):
        pk.cg((rowptr, bad_col, val, n), b)
    bad_rp = rowptr.copy()
    bad_rp[7], bad_rp[8] = rowptr[8], rowptr[7]                    # not monotone
    with pytest.raises(pk.PkError, match="monotone"):
        pk.cg((bad_rp, col, val, n), b)
    short_rp = rowptr.copy()
    short_rp[-1] -= 1                                              # rowptr[n] != nnz
    with pytest.raises(pk.PkError, match="nnz"):
        pk.cg((short_rp, col, val, n), b)
    wide = col.astype(np.int64)
    wide[3] = 2 ** 31 + 1                                          # would wrap when narrowed to int32
    with pytest.raises(pk.PkError, match="int32"):
        pk.cg((rowptr, wide, val, n), b)
    x, info = pk.cg((rowptr, col, val, n), b, tol=1e-8)            # the well-formed block still solves
    assert info["converged"]

"""Matrices arriving as files — .npz (scipy sparse), .mtx (Matrix Market), .npy (dense), the formats the reference's
drivers fed to the solvers (/root/reference/.gitignore:14-17) — loaded through parallel_krylov_b200.io and solved on the
GPU against the oracle."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

import krylov_oracle as oracle
from parallel_krylov_b200 import io as pkio
from parallel_krylov_b200 import problems

pytestmark = pytest.mark.gpu
os.environ.setdefault("PK_QUIET", "1")


@pytest.mark.parametrize("fmt", ["npz", "mtx", "npy"])
@pytest.mark.parametrize("solver,kw", [("cg", {}), ("kskipmrr", {"k": 2})])
def test_solve_from_matrix_file(tmp_path, fmt, solver, kw):
    import parallel_krylov_b200 as pk
    if fmt == "npy":
        A0 = problems.dense_spd(200, 1)
        np.save(tmp_path / "a.npy", A0)
    else:
        A0 = problems.to_scipy(*problems.poisson3d(11, 9, 10))
        if fmt == "npz":
            sp.save_npz(tmp_path / "a.npz", A0)
        else:
            scipy.io.mmwrite(str(tmp_path / "a.mtx"), A0)      # coordinate format: rows come back unsorted -> tocsr() sorts
    b0 = problems.rhs(A0.shape[0], "randn", 0)
    np.save(tmp_path / "b.npy", b0)
    A = pkio.load_matrix(str(tmp_path / f"a.{fmt}"))
    b = pkio.load_vector(str(tmp_path / "b.npy"))
    xo, io = oracle.SOLVERS[solver](A0, b0.copy(), tol=1e-8, **kw)
    x, info = getattr(pk, solver)(A, b, tol=1e-8, **kw)
    assert abs(int(info["nosl"][-1]) - int(io["nosl"][-1])) <= 3
    res = info["residual"].cpu().numpy()
    m = min(len(res), len(io["residual"]), 50)
    np.testing.assert_allclose(res[:m], io["residual"][:m], rtol=1e-10)
    assert oracle.true_relres(A0, b0, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)

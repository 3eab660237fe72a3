"""Input adapters: .npy / .npz / .mtx round trips and the block-row helper (CPU)."""
import numpy as np
import scipy.io
import scipy.sparse as sp

from parallel_krylov_b200 import io as pkio
from parallel_krylov_b200 import problems


def test_load_matrix_formats(tmp_path):
    A = problems.to_scipy(*problems.poisson2d(7))
    sp.save_npz(tmp_path / "a.npz", A)
    scipy.io.mmwrite(str(tmp_path / "a.mtx"), A)
    np.save(tmp_path / "a.npy", A.toarray())
    for name in ("a.npz", "a.mtx"):
        B = pkio.load_matrix(str(tmp_path / name))
        assert sp.issparse(B) and B.format == "csr" and abs(B - A).max() == 0
    D = pkio.load_matrix(str(tmp_path / "a.npy"))
    assert isinstance(D, np.ndarray) and np.array_equal(D, A.toarray())
    np.save(tmp_path / "b.npy", np.arange(5.0))
    assert np.array_equal(pkio.load_vector(str(tmp_path / "b.npy")), np.arange(5.0))


def test_row_block_covers_all_rows():
    A = problems.to_scipy(*problems.poisson2d(5))          # 25 rows, 3 ranks: 8 + 8 + 9
    blocks = [pkio.row_block(A, r, 3) for r in range(3)]
    assert [b.shape[0] for b in blocks] == [8, 8, 9]
    assert abs(sp.vstack(blocks) - A).max() == 0

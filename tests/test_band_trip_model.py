"""CPU model of the two claims the dense-band trips rest on (csrc/pk_matpow.cu), checked with numpy / scipy:

1. Trapezoid: computing m chained mat-vecs on a WINDOW of rows whose surroundings are unknown leaves every row at least
   m*bw inside the window exact — bit for bit — so a window of T rows finishes T - 2 m bw of them (k_matpow_band:
   m = k, k_mrr_steps_band: m = k + 1 with the element-wise MrR updates between the mat-vecs).
2. Row-partitioned form: a rank that runs the SAME computation on the extended operator [ghost rows | owned rows | ghost
   rows] (columns cut at its two ends, mpi/_dist.py::band_ext_csr) with ghost zones of the vectors that are exact only to
   depth (k+1) bw — and poisoned (NaN) beyond, as stale memory could be — gets its owned rows exactly as one GPU does.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from parallel_krylov_b200 import problems
from parallel_krylov_b200.mpi._dist import band_ext_csr


def _steps(A, r, ar, y, z, x, coef):
    """The k+1 steps of a k-skip MrR trip + closing mat-vec, as /root/reference/v3/cpu/kskipmrr.py:63-69, 87-93 and
    k_mrr_update / k_mrr_steps_band write them (separately rounded products and sums)."""
    r, ar, y, z, x = r.copy(), ar.copy(), y.copy(), z.copy(), x.copy()
    for zeta, eta in coef:
        y = eta * y + zeta * ar
        z = eta * z - zeta * r
        r = r - y
        x = x - z
        ar = A.dot(r)
    return r, ar, y, z, x


def _ext_operator(A, lo, hi, ghost):
    """Extended operator of the row block [lo, hi): rows [lo - ghost, hi + ghost) (clipped to the matrix), columns
    renumbered and cut, through the product's own band_ext_csr."""
    n = A.shape[0]
    e0, e1 = max(lo - ghost, 0), min(hi + ghost, n)
    pieces = []
    for a, b in ((e0, lo), (lo, hi), (hi, e1)):
        if b > a:
            blk = A[a:b].tocsr()
            blk.sort_indices()
            pieces.append((torch.from_numpy(blk.indptr.astype(np.int64)), torch.from_numpy(blk.indices.astype(np.int64)),
                           torch.from_numpy(blk.data.copy())))
    rp, col, val = band_ext_csr(pieces, e0, e1 - e0)
    return sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(e1 - e0, e1 - e0)), e0, e1


@pytest.mark.parametrize("n,bw,k,world", [(3000, 13, 8, 4), (2500, 13, 4, 3), (1201, 2, 6, 2), (4000, 6, 12, 5)])
def test_row_partitioned_trip_on_the_extended_operator_is_exact_on_owned_rows(n, bw, k, world):
    A = problems.to_scipy(*problems.banded_spd(n, bw, 3))
    rng = np.random.default_rng(k)
    r, y, z, x = (rng.standard_normal(n) for _ in range(4))
    ar = A.dot(r)
    coef = [(rng.uniform(0.1, 0.9), rng.uniform(-0.5, 0.5)) for _ in range(k + 1)]
    want = _steps(A, r, ar, y, z, x, coef)
    # matrix powers of both chains on the full problem
    lev_want = [(ar, y)]
    for _ in range(k):
        lev_want.append((A.dot(lev_want[-1][0]), A.dot(lev_want[-1][1])))
    ghost = 16 * bw                                    # ghost rows of A shipped at set-up
    depth = (k + 1) * bw + ((k + 1) * bw) % 2          # vector entries exchanged per trip (pk_band_ext_depth)
    assert depth <= ghost
    base = n // world
    for rank in range(world):
        lo, hi = rank * base, (n if rank == world - 1 else (rank + 1) * base)
        Ax, e0, e1 = _ext_operator(A, lo, hi, ghost)
        assert np.array_equal(Ax[lo - e0:hi - e0].toarray()[:, max(lo - bw, e0) - e0:min(hi + bw, e1) - e0],
                              A[lo:hi].toarray()[:, max(lo - bw, e0):min(hi + bw, e1)])      # owned rows are untouched

        def ext(v, exact_zone=True):
            out = np.full(e1 - e0, np.nan)             # stale pads: anything, here poison
            out[lo - e0:hi - e0] = v[lo:hi]
            if exact_zone:
                a, b = max(lo - depth, e0), min(hi + depth, e1)
                out[a - e0:b - e0] = v[a:b]            # what the one exchange per trip delivers
            return out

        got = _steps(Ax, ext(r), ext(ar), ext(y), ext(z, False), ext(x, False), coef)
        for g, w, name in zip(got, want, ("r", "Ar", "y", "z", "x")):
            assert np.array_equal(g[lo - e0:hi - e0], w[lo:hi]), (rank, name)
        u, v = ext(ar), ext(y)
        for l in range(1, k + 1):
            u, v = Ax.dot(u), Ax.dot(v)
            assert np.array_equal(u[lo - e0:hi - e0], lev_want[l][0][lo:hi]), (rank, "chain 0", l)
            assert np.array_equal(v[lo - e0:hi - e0], lev_want[l][1][lo:hi]), (rank, "chain 1", l)


@pytest.mark.parametrize("n,bw,k,window", [(5000, 13, 8, 768), (3000, 13, 3, 768), (2000, 5, 10, 768)])
def test_windows_finish_all_but_m_bw_rows_at_either_end(n, bw, k, window):
    """Tiling of k_mrr_steps_band: windows of `window` rows, each finishing window - 2 (k+1) bw rows, cover every row, and
    a window's finished rows are exact although nothing outside the window is known (poisoned)."""
    A = problems.to_scipy(*problems.banded_spd(n, bw, 1))
    rng = np.random.default_rng(1)
    r, y, z, x = (rng.standard_normal(n) for _ in range(4))
    ar = A.dot(r)
    coef = [(rng.uniform(0.1, 0.9), rng.uniform(-0.5, 0.5)) for _ in range(k + 1)]
    want = _steps(A, r, ar, y, z, x, coef)
    ghost = (k + 1) * bw
    t_out = window - 2 * ghost
    assert t_out >= window // 2                         # pk_mrr_steps_ok
    covered = np.zeros(n, dtype=bool)
    for tile in range((n + t_out - 1) // t_out):
        o0, o1 = tile * t_out, min((tile + 1) * t_out, n)
        s0, s1 = max(o0 - ghost, 0), min(o0 - ghost + window, n)
        W = A[s0:s1][:, s0:s1].tocsr()                  # what the window's threads hold: rows cut to the window
        cut = lambda v: v[s0:s1].copy()
        got = _steps(W, cut(r), cut(ar), cut(y), cut(z), cut(x), coef)
        for g, w in zip(got, want):
            assert np.array_equal(g[o0 - s0:o1 - s0], w[o0:o1]), tile
        covered[o0:o1] = True
    assert covered.all()


def test_band_ext_csr_rejects_blocks_that_do_not_add_up():
    from parallel_krylov_b200._lib import PkError
    A = problems.to_scipy(*problems.banded_spd(100, 2, 0))
    blk = A[10:30].tocsr()
    piece = (torch.from_numpy(blk.indptr.astype(np.int64)), torch.from_numpy(blk.indices.astype(np.int64)),
             torch.from_numpy(blk.data.copy()))
    with pytest.raises(PkError):
        band_ext_csr([piece], 10, 25)

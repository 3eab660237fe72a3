"""Synthetic SPD test systems (host side, numpy) — 2-D/3-D Poisson stencils and a banded SPD matrix.

These are the inputs BASELINE.json's configs name (SURVEY.md §8d "Synthetic inputs").  The reference
ships no generators (its inputs were git-ignored ``*.npz``/``*.mtx`` files, /root/reference/.gitignore:14-17),
so the definitions here are ours:

* ``poisson2d(n)``   — 5-point Laplacian on an n×n grid, Dirichlet, diag 4 / off-diag −1, row = y·n + x.
* ``poisson3d(nx,ny,nz)`` — 7-point Laplacian, diag 6 / off-diag −1, row = (z·ny + y)·nx + x.
* ``banded_spd(N, half_bw)`` — symmetric band with 2·half_bw+1 diagonals; off-diagonal (i, i+d) is
  ``−(0.1 + 0.9·u(seed, d, i))`` with ``u`` a counter-based hash (splitmix64) so that the CUDA generator in
  ``csrc/generators.cu`` produces the *same* matrix bit for bit; diagonal = Σ|off-diag of the row| + 1
  (strictly diagonally dominant ⇒ SPD).

All generators emit CSR directly (``rowptr`` int32 [n+1], ``col`` int32 [nnz], ``val`` float64 [nnz]) with
column indices sorted inside each row, which is what scipy's ``csr_matrix`` canonical format and our
kernels expect.
"""
from __future__ import annotations

import numpy as np

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 (wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
        return z ^ (z >> np.uint64(31))


def hash_uniform(seed: int, stream: np.ndarray | int, index: np.ndarray) -> np.ndarray:
    """u in [0,1) from (seed, stream, index): top 53 bits of two chained splitmix64 rounds."""
    with np.errstate(over="ignore"):
        s = splitmix64(np.uint64(seed) + np.uint64(0x632BE59BD9B4E019) * np.asarray(stream, dtype=np.uint64))
        h = splitmix64(s ^ np.asarray(index, dtype=np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_normal(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """Standard-normal vector from the hash stream (Box–Muller, cos branch); the device generator
    (``pk_fill_hash_normal``) reproduces it to libm rounding."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    u1 = hash_uniform(seed, 1, idx)
    u2 = hash_uniform(seed, 2, idx)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def _stencil_csr(dims, diag: float):
    """CSR of the (2·d+1)-point Laplacian on a box grid with Dirichlet boundaries."""
    dims = tuple(int(d) for d in dims)
    n = int(np.prod(dims))
    idx = np.arange(n, dtype=np.int64)
    # coordinates, fastest dimension first
    coords = []
    rem = idx
    for d in dims:
        coords.append(rem % d)
        rem = rem // d
    strides = [1]
    for d in dims[:-1]:
        strides.append(strides[-1] * d)
    # neighbour slots in ascending column order: -s_k ... -s_0, 0, +s_0 ... +s_k
    slots = []
    for k in reversed(range(len(dims))):
        slots.append((-strides[k], coords[k] > 0))
    slots.append((0, np.ones(n, dtype=bool)))
    for k in range(len(dims)):
        slots.append((strides[k], coords[k] < dims[k] - 1))
    counts = np.zeros(n, dtype=np.int64)
    for _, m in slots:
        counts += m
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    pos = rowptr[:-1].copy()
    for off, m in slots:
        p = pos[m]
        col[p] = (idx[m] + off).astype(np.int32)
        val[p] = diag if off == 0 else -1.0
        pos[m] += 1
    return rowptr.astype(np.int32), col, val, n


def poisson2d(n: int):
    """2-D 5-point Poisson, N = n² (BASELINE.json configs[0] uses n=256)."""
    return _stencil_csr((n, n), 4.0)


def poisson3d(nx: int, ny: int | None = None, nz: int | None = None):
    """3-D 7-point Poisson, N = nx·ny·nz (configs[1,2,4] use 128³, 256³, 512³)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return _stencil_csr((nx, ny, nz), 6.0)


def banded_spd(n: int, half_bw: int = 13, seed: int = 0):
    """Symmetric banded SPD matrix with 2·half_bw+1 diagonals (configs[3]: bandwidth 27, N = 2^25)."""
    n = int(n)
    idx = np.arange(n, dtype=np.int64)
    lo = np.maximum(idx - half_bw, 0)
    hi = np.minimum(idx + half_bw, n - 1)
    counts = hi - lo + 1
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    # entry (i, j) sits at rowptr[i] + (j - lo[i])
    for d in range(1, half_bw + 1):
        i = idx[: n - d]                       # pair (i, i+d), weight keyed by the smaller index
        w = 0.1 + 0.9 * hash_uniform(seed, d, i)
        up = rowptr[:-1][: n - d] + (i + d - lo[: n - d])
        dn = rowptr[:-1][d:] + (i - lo[d:])
        col[up] = (i + d).astype(np.int32)
        val[up] = -w
        col[dn] = i.astype(np.int32)
        val[dn] = -w
    # diagonal = 1 + Σ|off|, summed in ascending column order so host and device agree bit for bit
    dsum = np.zeros(n, dtype=np.float64)
    for d in range(-half_bw, half_bw + 1):
        if d == 0:
            continue
        j = idx + d
        m = (j >= 0) & (j < n)
        pos = rowptr[:-1][m] + (j[m] - lo[m])
        dsum[m] = dsum[m] + np.abs(val[pos])
    dpos = rowptr[:-1] + (idx - lo)
    col[dpos] = idx.astype(np.int32)
    val[dpos] = dsum + 1.0
    return rowptr.astype(np.int32), col, val, n


def dense_spd(n: int, seed: int = 0) -> np.ndarray:
    """Small dense SPD matrix for the GEMV path (the reference's ``np.ndarray`` A branch,
    /root/reference/v3/gpu/common.py:100-101)."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n, n))
    a = q @ q.T / n
    a[np.diag_indices(n)] += 1.0
    return np.ascontiguousarray(a)


def to_scipy(rowptr, col, val, n):
    import scipy.sparse as sp
    return sp.csr_matrix((val, col, rowptr), shape=(n, n))


def rhs(n: int, kind: str = "randn", seed: int = 0) -> np.ndarray:
    """Right-hand sides of SURVEY §8d: ``randn`` = default_rng(seed).standard_normal, ``ones``,
    ``hash`` = the device-reproducible hash stream."""
    if kind == "ones":
        return np.ones(n, dtype=np.float64)
    if kind == "randn":
        return np.random.default_rng(seed).standard_normal(n)
    if kind == "hash":
        return hash_normal(seed, n)
    raise ValueError(kind)

"""GPU parity of the individual kernels (through the C-ABI) against the oracle's numpy/scipy operations."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from parallel_krylov_b200 import problems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pk():
    import parallel_krylov_b200 as pk
    return pk


def _matrices():
    yield "p2d48", problems.poisson2d(48)
    yield "p3d_12x20x9", problems.poisson3d(12, 20, 9)
    yield "p3d40", problems.poisson3d(40)
    yield "band27_20k", problems.banded_spd(20000, 13, 0)
    yield "band5_777", problems.banded_spd(777, 2, 3)
    yield "band101_3000", problems.banded_spd(3000, 50, 1)     # long rows -> warp-per-row path
    yield "tiny1", problems.poisson2d(1)
    yield "tiny3", problems.poisson2d(3)


def _ragged(seed=0, n=5000):
    """Irregular rows incl. empty rows and a few very long ones (collision of both SpMV paths inside one matrix)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 12, size=n)
    lens[rng.integers(0, n, size=20)] = rng.integers(300, 3000, size=20)
    lens[:7] = 0
    lens[-3:] = 0
    rows = np.repeat(np.arange(n), lens)
    cols = rng.integers(0, n, size=rows.size)
    vals = rng.standard_normal(rows.size)
    m = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    m.sum_duplicates()
    m.sort_indices()
    return m


@pytest.mark.parametrize("name,csr", list(_matrices()), ids=[n for n, _ in _matrices()])
def test_spmv_bit_exact_vs_scipy(pk, name, csr):
    rowptr, col, val, n = csr
    A = problems.to_scipy(rowptr, col, val, n)
    x = np.random.default_rng(1).standard_normal(n)
    op = pk.Operator.from_any(A)
    y = op.matvec(torch.from_numpy(x)).cpu().numpy()
    ref = A.dot(x)
    info = op.kernel_info()
    if info["kernel"] == "csr-stream":
        # same products, same left-to-right accumulation as scipy's csr_matvec: identical bits
        assert np.array_equal(y, ref), (name, info, np.abs(y - ref).max())
    else:
        np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())


def test_spmv_ragged_rows(pk):
    A = _ragged()
    n = A.shape[0]
    x = np.random.default_rng(2).standard_normal(n)
    op = pk.Operator.from_any(A)
    y = op.matvec(torch.from_numpy(x)).cpu().numpy()
    ref = A.dot(x)
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    # rows short enough to be staged are bit exact; empty rows give +0.0
    lens = np.diff(A.indptr)
    assert np.all(y[lens == 0] == 0.0)


def test_spmv_two_chains_equals_two_single_passes(pk):
    rowptr, col, val, n = problems.poisson3d(24)
    A = problems.to_scipy(rowptr, col, val, n)
    rng = np.random.default_rng(3)
    x0, x1 = rng.standard_normal(n), rng.standard_normal(n)
    op = pk.Operator.from_any(A)
    y0, y1 = op.matvec(torch.from_numpy(x0), x1=torch.from_numpy(x1))
    assert np.array_equal(y0.cpu().numpy(), A.dot(x0))
    assert np.array_equal(y1.cpu().numpy(), A.dot(x1))


def test_spmv_fused_dots(pk):
    rowptr, col, val, n = problems.poisson3d(32)
    A = problems.to_scipy(rowptr, col, val, n)
    rng = np.random.default_rng(4)
    x, w = rng.standard_normal(n), rng.standard_normal(n)
    op = pk.Operator.from_any(A)
    y, sums = op.matvec(torch.from_numpy(x), dot_with=torch.from_numpy(w))
    ref = A.dot(x)
    s = sums.cpu().numpy()
    np.testing.assert_allclose(s, [np.dot(w, ref), np.dot(ref, ref), np.dot(w, w)], rtol=1e-13)
    # deterministic: a second launch gives the same bits
    _, sums2 = op.matvec(torch.from_numpy(x), dot_with=torch.from_numpy(w))
    assert np.array_equal(s, sums2.cpu().numpy())


def test_dense_gemv(pk):
    a = problems.dense_spd(517, 0)
    rng = np.random.default_rng(5)
    x, x1 = rng.standard_normal(517), rng.standard_normal(517)
    op = pk.Operator.from_any(a)
    assert op.kernel_info()["kernel"] == "dense-gemv"
    y0, y1 = op.matvec(torch.from_numpy(x), x1=torch.from_numpy(x1))
    np.testing.assert_allclose(y0.cpu().numpy(), a @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(y1.cpu().numpy(), a @ x1, rtol=1e-13, atol=1e-13)


def test_dense_gemv_two_chains_staged_x_is_bitwise_the_single_chain_kernel(pk):
    """Two right-hand sides on a wide dense block go through k_gemv_staged (x0 / x1 chunks staged in shared memory by
    bulk copies); per row it adds in the order of the one-vector kernel, so both must agree bit for bit.  4610 columns =
    two full chunks + a 514-column tail; 1003 rows = a ragged last row group."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((1003, 4610))
    x0, x1 = rng.standard_normal(4610), rng.standard_normal(4610)
    op = pk.Operator.from_dense_tensor(torch.from_numpy(a))
    y0, y1 = op.matvec(torch.from_numpy(x0), x1=torch.from_numpy(x1))
    s0 = op.matvec(torch.from_numpy(x0))
    s1 = op.matvec(torch.from_numpy(x1))
    assert torch.equal(y0, s0) and torch.equal(y1, s1)
    ref = a.dot(x0)
    np.testing.assert_allclose(y0.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    w = rng.standard_normal(1003)
    _, sums = op.matvec(torch.from_numpy(x0), dot_with=torch.from_numpy(w))
    np.testing.assert_allclose(sums.cpu().numpy(), [np.dot(w, ref), np.dot(ref, ref), np.dot(w, w)], rtol=1e-12)


@pytest.mark.parametrize("mode,k", [(0, 0), (0, 1), (0, 4), (0, 8), (0, 12), (1, 0), (1, 2), (1, 4), (1, 8), (1, 16)])
def test_gram_all_pairs(pk, mode, k):
    """The single-pass Gram kernel against the oracle's per-pair numpy.dot calls (kskipmrr.py:51-59, kskipcg.py:40-48)."""
    from parallel_krylov_b200 import _lib
    from parallel_krylov_b200._core import Context, _ptr
    ctx = Context.get()
    n, ld = 100003, 100032
    nu, nv = (k + 2, k + 1) if mode == 0 else (k + 1, k + 2)
    rng = np.random.default_rng(6)
    U = np.zeros((nu, ld)); V = np.zeros((nv, ld))
    U[:, :n] = rng.standard_normal((nu, n)); V[:, :n] = rng.standard_normal((nv, n))
    U[:, n:] = 7.0; V[:, n:] = -3.0          # padding must not be read
    Ud, Vd = torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
    njj = max(nu, nv)
    g = torch.zeros(6 * njj, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    _lib.check(ctx.lib.pk_gram(ctx.handle, mode, n, ld, _ptr(Ud), nu, _ptr(Vd), nv, _ptr(g)))
    ctx.sync()
    G = g.cpu().numpy().reshape(njj, 6)
    row = lambda M, j: M[j, :n] if j < M.shape[0] else np.zeros(n)
    for jj in range(njj):
        u0, u1, v0, v1 = row(U, jj), row(U, jj + 1), row(V, jj), row(V, jj + 1)
        cross = np.dot(v0, u1) if mode == 0 else np.dot(u0, v1)
        ref = [np.dot(u0, u0), np.dot(u0, u1), np.dot(u0, v0), cross, np.dot(v0, v0), np.dot(v0, v1)]
        np.testing.assert_allclose(G[jj], ref, rtol=1e-12, atol=1e-9)


def test_device_generators_match_host(pk):
    from parallel_krylov_b200 import device_problems as dp
    for args in [(16, 16, 1), (12, 20, 9), (33, 7, 5)]:
        rp, ci, va, n = dp.stencil_csr(*args)
        hr, hc, hv, hn = problems.poisson2d(args[0]) if args[2] == 1 else problems.poisson3d(*args)
        assert n == hn
        assert np.array_equal(rp.cpu().numpy(), hr) and np.array_equal(ci.cpu().numpy(), hc)
        assert np.array_equal(va.cpu().numpy(), hv)
    rp, ci, va, n = dp.banded_csr(5000, 13, 0)
    hr, hc, hv, hn = problems.banded_spd(5000, 13, 0)
    assert np.array_equal(rp.cpu().numpy(), hr) and np.array_equal(ci.cpu().numpy(), hc)
    assert np.array_equal(va.cpu().numpy(), hv)
    # a row slab of a bigger grid (what a rank generates for itself)
    rp, ci, va, n = dp.stencil_csr(12, 20, 9, row0=480, n_rows=960)
    sl = problems.to_scipy(*problems.poisson3d(12, 20, 9))[480:1440]
    assert np.array_equal(ci.cpu().numpy(), sl.indices) and np.array_equal(va.cpu().numpy(), sl.data)
    z = dp.hash_normal(0, 4096).cpu().numpy()
    np.testing.assert_allclose(z, problems.hash_normal(0, 4096), rtol=1e-12, atol=1e-14)


def test_row_pattern_compression_is_bit_exact(pk):
    """Opt-in lossless compression: same bits as scipy / the CSR kernels; matrices without repeating rows keep CSR."""
    rowptr, col, val, n = problems.poisson3d(20, 17, 23)
    A = problems.to_scipy(rowptr, col, val, n)
    rng = np.random.default_rng(8)
    x, x1, w = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    op = pk.Operator.from_any(A)
    assert op.compress_patterns() and op.kernel_info()["kernel"] == "row-pattern" and op.n_patterns == 27
    y, sums = op.matvec(torch.from_numpy(x), dot_with=torch.from_numpy(w))
    assert np.array_equal(y.cpu().numpy(), A.dot(x))
    np.testing.assert_allclose(sums.cpu().numpy(), [np.dot(w, A.dot(x)), np.dot(A.dot(x), A.dot(x)), np.dot(w, w)], rtol=1e-13)
    y0, y1 = op.matvec(torch.from_numpy(x), x1=torch.from_numpy(x1))
    assert np.array_equal(y0.cpu().numpy(), A.dot(x)) and np.array_equal(y1.cpu().numpy(), A.dot(x1))
    band = pk.Operator.from_any(problems.to_scipy(*problems.banded_spd(40000, 13, 0)))
    assert band.compress_patterns() is False and band.kernel_info()["kernel"] == "csr-stream"


@pytest.mark.parametrize("solver,k", [("cg", None), ("mrr", None), ("kskipcg", 3), ("kskipmrr", 4), ("adaptivekskipmrr", 2)])
def test_solvers_on_compressed_operator_match_csr_bitwise(pk, solver, k):
    rowptr, col, val, n = problems.poisson3d(18)
    A = problems.to_scipy(rowptr, col, val, n)
    b = problems.rhs(n, "randn", 0)
    kw = {"k": k} if k is not None else {}
    x0, i0 = getattr(pk, solver)(A, b, tol=1e-8, **kw)
    x1, i1 = getattr(pk, solver)(A, b, tol=1e-8, compress=True, **kw)
    # every SpMV is bit-identical; dots reduced in the SpMV epilogue (cg, mrr) are summed over a different grid, so
    # those trajectories agree to rounding; k-skip steps take their scalars from the (unchanged) Gram kernel
    assert torch.equal(i0["nosl"], i1["nosl"])
    np.testing.assert_allclose(i1["residual"].cpu().numpy(), i0["residual"].cpu().numpy(), rtol=1e-10)
    np.testing.assert_allclose(x1.cpu().numpy(), x0.cpu().numpy(), rtol=1e-8, atol=1e-12)
    if solver == "kskipcg":          # no dot product comes out of an SpMV epilogue in this solver: identical bits
        assert torch.equal(x0, x1)


def _band_with_holes(n, half_bw, seed):
    """The dense band with a third of its off-diagonal entries removed (symmetrically): same bandwidth, ragged rows."""
    import scipy.sparse as sp
    A = problems.to_scipy(*problems.banded_spd(n, half_bw, seed)).tocoo()
    lo, hi = np.minimum(A.row, A.col).astype(np.int64), np.maximum(A.row, A.col).astype(np.int64)
    keep = (A.row == A.col) | ((lo * 2654435761 + hi * 40503) % 3 != 0)
    B = sp.csr_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=A.shape)
    B.sort_indices()
    return B


@pytest.mark.parametrize("n,half_bw,k,structure", [
    (30011, 13, 8, "dense"), (5000, 2, 4, "dense"), (700, 13, 2, "dense"), (100003, 13, 5, "dense"),
    (27, 13, 2, "dense"), (1024, 13, 8, "dense"), (843, 13, 8, "dense"), (4099, 1, 8, "dense"), (60013, 13, 15, "dense"),
    (200003, 6, 12, "dense"),
    (30011, 13, 8, "holes"), (5000, 2, 4, "holes"), (700, 13, 2, "holes"), (100003, 13, 5, "holes")])
def test_matrix_powers_one_pass_is_bitwise_k_sequential_spmvs(pk, n, half_bw, k, structure):
    """k levels of both basis chains in ONE pass over A (csrc/pk_matpow.cu) == k sequential operator applications,
    bit for bit (same per-row left-to-right sums) — including windows cut by the matrix ends.  A full band runs the
    two-rows-per-thread kernel, a band with holes the general one."""
    import ctypes as C
    from parallel_krylov_b200 import _lib
    from parallel_krylov_b200._core import _ptr
    A = problems.to_scipy(*problems.banded_spd(n, half_bw, 1)) if structure == "dense" else _band_with_holes(n, half_bw, 1)
    op = pk.Operator.from_any(A)
    info = op.matpow_info(k)
    if os.environ.get("PK_MATPOW_BAND", "1") not in ("0", ""):
        assert info["kernel"] == ("dense-band" if structure == "dense" else "general"), info
    ctx, ld = op.ctx, op.ld
    rng = np.random.default_rng(k)
    u0, v0 = rng.standard_normal(n), rng.standard_normal(n)
    U = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
    V = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
    U[:n] = torch.from_numpy(u0).cuda()
    V[:n] = torch.from_numpy(v0).cuda()
    torch.cuda.synchronize()
    _lib.check(ctx.lib.pk_matpow(ctx.handle, op.handle, k, _ptr(U), _ptr(V)), "pk_matpow")
    ctx.sync()
    u, v = u0, v0
    for l in range(1, k + 1):
        u, v = A.dot(u), A.dot(v)               # == op.matvec, which is bit-identical to scipy (test above)
        assert np.array_equal(U[l * ld:l * ld + n].cpu().numpy(), u), f"chain 0 level {l}"
        assert np.array_equal(V[l * ld:l * ld + n].cpu().numpy(), v), f"chain 1 level {l}"


def test_matrix_powers_refuses_wide_operators(pk):
    import ctypes as C
    from parallel_krylov_b200 import _lib
    from parallel_krylov_b200._core import _ptr
    A = problems.to_scipy(*problems.poisson3d(20))         # half bandwidth 400: no one-pass basis
    op = pk.Operator.from_any(A)
    buf = torch.zeros(4 * op.ld, dtype=torch.float64, device="cuda")
    assert op.ctx.lib.pk_matpow(op.ctx.handle, op.handle, 2, _ptr(buf), _ptr(buf)) == -4      # PK_ERR_UNSUPPORTED


def _solve_in_subprocess(solver, k, n, env_extra):
    import subprocess
    import sys
    import tempfile
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import parallel_krylov_b200 as pk\n"
        "from parallel_krylov_b200 import problems\n"
        f"A = problems.to_scipy(*problems.banded_spd({n}, 13, 0)); b = problems.rhs(A.shape[0], 'randn', 0)\n"
        f"x, info = pk.{solver}(A, b, tol=1e-8, k={k})\n"
        "np.save(sys.argv[1], np.concatenate([x.cpu().numpy(), info['residual'].cpu().numpy()]))\n")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ref.npy")
        env = dict(os.environ, PK_QUIET="1", **env_extra)
        r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return np.load(out)


@pytest.mark.parametrize("solver,k,n", [("kskipmrr", 8, 20000), ("kskipmrr", 5, 20001), ("kskipmrr", 2, 3001),
                                        ("kskipcg", 4, 20000), ("adaptivekskipmrr", 8, 20000)])
def test_solvers_with_and_without_matrix_powers_agree_bitwise(pk, solver, k, n):
    """The one-pass basis (k_matpow*) changes no bit of a solve, and neither does running the k+1 steps of a k-skip MrR
    trip as one pass over a dense band (k_mrr_steps_band): same x to the last bit.  Only the trip-end r.r of the fused
    steps is summed over a different grid, so that residual history agrees to rounding, not bitwise.  Reference runs in
    subprocesses (the switches are read once per process)."""
    A = problems.to_scipy(*problems.banded_spd(n, 13, 0))
    b = problems.rhs(A.shape[0], "randn", 0)
    x, info = getattr(pk, solver)(A, b, tol=1e-8, k=k)
    got_x, got_res = x.cpu().numpy(), info["residual"].cpu().numpy()
    plain = _solve_in_subprocess(solver, k, n, {"PK_MATPOW": "0"})          # k two-chain SpMV passes, step-by-step trip
    assert np.array_equal(got_x, plain[:n])
    if solver == "kskipmrr":
        basis_only = _solve_in_subprocess(solver, k, n, {"PK_KSTEPS": "0"})  # one-pass basis, step-by-step trip
        assert np.array_equal(basis_only, plain)
        assert len(got_res) == len(plain) - n
        np.testing.assert_allclose(got_res, plain[n:], rtol=1e-12)
    else:
        assert np.array_equal(got_res, plain[n:])


def test_matrix_powers_dense_band_rejects_misaligned_level_vectors(pk):
    """The dense-band kernel stages level 0 by TMA bulk copies: level vectors that are not 16-byte aligned are refused
    with PK_ERR_ARG (the solvers check their work area up front and stay on the general kernel instead)."""
    A = problems.to_scipy(*problems.banded_spd(5000, 13, 1))
    op = pk.Operator.from_any(A)
    if op.matpow_info(2)["kernel"] != "dense-band":
        pytest.skip("dense-band kernel switched off")
    buf = torch.zeros(3 * op.ld + 2, dtype=torch.float64, device="cuda")
    p = C.c_void_p(buf.data_ptr() + 8)
    assert op.ctx.lib.pk_matpow(op.ctx.handle, op.handle, 2, p, p) == -2          # PK_ERR_ARG

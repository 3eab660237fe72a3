"""Opt-in Chebyshev basis of the k-skip MrR trips, on the CPU: the device's scalar engine
(csrc/pk_scalars.h: pk_kskipmrr_coef_cheb, host build) against the numpy restatement, and whole solves driven through it
against plain MrR — the method it is mathematically identical to — where the reference's monomial basis has long lost
the history (k = 8: third digit, k >= 12: chaotic)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cheb_reference as cheb
import krylov_oracle as oracle
from parallel_krylov_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheb") / "libhostcheb.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(out),
                    os.path.join(HERE, "host_scalars.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.host_kskipmrr_coef_cheb.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.host_kskipmrr_coef_cheb.restype = None
    lib.host_kskipcg_coef_cheb.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.host_kskipcg_coef_cheb.restype = None
    return lib


def _device_coefficients(lib, cg=False):
    def fn(G, k, c, d):
        G = np.ascontiguousarray(G)
        coef = np.zeros(2 * (k + 1))
        (lib.host_kskipcg_coef_cheb if cg else lib.host_kskipmrr_coef_cheb)(G.ctypes.data, k, c, d, coef.ctypes.data)
        return coef
    return fn


SYSTEMS = {"p2d48": ("poisson2d", (48,)), "p3d16": ("poisson3d", (16,)), "p3d12x20x9": ("poisson3d", (12, 20, 9))}


@pytest.mark.parametrize("k", [0, 1, 2, 4, 8, 12, 16])
def test_device_chebyshev_engine_is_bitwise_the_numpy_restatement(host_lib, k):
    A = problems.to_scipy(*problems.poisson3d(10, 9, 8))
    n = A.shape[0]
    rng = np.random.default_rng(k)
    lo, hi = cheb.gershgorin(A)
    c, d = 0.5 * (hi - lo), 0.5 * (hi + lo)
    U = np.zeros((k + 2, n)); V = np.zeros((k + 1, n))
    U[0] = rng.standard_normal(n); V[0] = 0.1 * rng.standard_normal(n)
    ah = lambda v: (A.dot(v) - d * v) / c
    U[1] = ah(U[0])
    if k >= 1:
        V[1] = ah(V[0])
    for j in range(1, k + 1):
        U[j + 1] = 2 * ah(U[j]) - U[j - 1]
        if j >= 2:
            V[j] = 2 * ah(V[j - 1]) - V[j - 2]
    G = cheb.gram_layout(U, V, k)
    want = cheb.coefficients(G, k, c, d)
    got = _device_coefficients(host_lib)(G, k, c, d)
    assert np.array_equal(got, want)
    # and the first pair is the plain MrR step from explicitly computed inner products
    Ar = A.dot(U[0])
    a1, a2, b1, d0 = np.dot(U[0], Ar), np.dot(Ar, Ar), np.dot(V[0], Ar), np.dot(V[0], V[0])
    dd = a2 * d0 - b1 * b1
    np.testing.assert_allclose(got[:2], [a1 * d0 / dd, -a1 * b1 / dd], rtol=1e-10)


@pytest.mark.parametrize("name", sorted(SYSTEMS))
@pytest.mark.parametrize("k", [4, 8, 12, 16])
def test_chebyshev_kskipmrr_follows_plain_mrr(host_lib, name, k):
    kind, args = SYSTEMS[name]
    A = problems.to_scipy(*getattr(problems, kind)(*args))
    b = problems.rhs(A.shape[0], "randn", 0)
    xm, im = oracle.mrr(A, b.copy(), tol=1e-8)
    x, info = cheb.kskipmrr_chebyshev(A, b, tol=1e-8, k=k, coef_fn=_device_coefficients(host_lib))
    assert info["converged"]
    it, it_mrr = int(info["nosl"][-1]), int(im["nosl"][-1])
    assert it_mrr <= it <= it_mrr + k + 1                       # MrR's count, rounded up to a whole trip
    sel = info["nosl"][info["nosl"] <= min(50, it_mrr)]         # plain MrR's residual at the iterations a trip ends on
    np.testing.assert_allclose(info["residual"][:len(sel)], im["residual"][sel], rtol=1e-8)
    assert oracle.true_relres(A, b, x) < 1e-8 * (1 + 1e-6)


def test_monomial_basis_has_lost_the_history_where_chebyshev_has_not():
    """The motivation, pinned: the reference's k = 12 trip is chaotic on the 2-D Poisson system, the Chebyshev one is not."""
    A = problems.to_scipy(*problems.poisson2d(48))
    b = problems.rhs(A.shape[0], "randn", 0)
    xm, im = oracle.mrr(A, b.copy(), tol=1e-8)
    xk, ik = oracle.kskipmrr(A, b.copy(), tol=1e-8, k=12, maxiter=3000)
    xc, ic = cheb.kskipmrr_chebyshev(A, b, tol=1e-8, k=12)
    assert int(ik["nosl"][-1]) > 5 * int(im["nosl"][-1])        # monomial: an order of magnitude more iterations
    assert int(ic["nosl"][-1]) <= int(im["nosl"][-1]) + 13


@pytest.mark.parametrize("name", sorted(SYSTEMS))
@pytest.mark.parametrize("k", [4, 8, 12])
def test_chebyshev_kskipcg_follows_plain_cg(host_lib, name, k):
    """Same for k-skip CG: through the device's engine (host build) the Chebyshev-basis trips follow plain CG where the
    reference's monomial k = 12 trip needs 5000 iterations instead of 65."""
    kind, args = SYSTEMS[name]
    A = problems.to_scipy(*getattr(problems, kind)(*args))
    b = problems.rhs(A.shape[0], "randn", 0)
    xc, ic = oracle.cg(A, b.copy(), tol=1e-8)
    x, info = cheb.kskipcg_chebyshev(A, b, tol=1e-8, k=k, coef_fn=_device_coefficients(host_lib, cg=True))
    x2, info2 = cheb.kskipcg_chebyshev(A, b, tol=1e-8, k=k)
    assert np.array_equal(info["residual"], info2["residual"]) and np.array_equal(x, x2)      # engine == numpy, bit for bit
    assert info["converged"]
    it, it_cg = int(info["nosl"][-1]), int(ic["nosl"][-1])
    assert it_cg <= it <= it_cg + k + 1
    sel = info["nosl"][info["nosl"] <= min(50, it_cg)]
    np.testing.assert_allclose(info["residual"][:len(sel)], ic["residual"][sel], rtol=1e-8)
    assert oracle.true_relres(A, b, x) < 1e-8 * (1 + 1e-6)

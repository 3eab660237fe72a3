"""The device's scalar engine for the k-skip solvers (csrc/pk_scalars.h: all k+1 coefficient pairs of a trip from the
Gram sums) compiled for the HOST and checked bit for bit against the oracle's scalar recurrences, on Gram values
taken from real oracle solves.  CPU test: covers the most order-sensitive arithmetic of the path without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import krylov_oracle as oracle
from parallel_krylov_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))


def _build(tmp_path_factory, name, extra):
    out = tmp_path_factory.mktemp(name) / f"lib{name}.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", *extra, "-o", str(out),
                    os.path.join(HERE, "host_scalars.cpp")], check=True)
    lib = C.CDLL(str(out))
    for fn in (lib.host_kskipcg_coef, lib.host_kskipmrr_coef):
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = None
    return lib


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    """The scalar engine exactly as the device compiles it (squares are products)."""
    return _build(tmp_path_factory, "hostscalars", [])


@pytest.fixture(scope="module")
def host_lib_pow(tmp_path_factory):
    """Same source, squares through libm pow() like the reference's `x ** 2` on numpy scalars."""
    return _build(tmp_path_factory, "hostscalarspow", ["-DPK_SQUARE_WITH_LIBM_POW", "-fno-builtin"])   # gcc folds pow(x, 2.0) into x*x otherwise


def _gram_layout(mode, U, V, k):
    """G[6*jj + t] exactly as pk_gram / k_gram_tma lay it out (absent rows contribute 0)."""
    njj = k + 2
    G = np.zeros(6 * njj)
    row = lambda M, j: M[j] if j < M.shape[0] else None
    dot = lambda a, b: 0.0 if a is None or b is None else float(np.dot(a, b))
    for jj in range(njj):
        u0, u1, v0, v1 = row(U, jj), row(U, jj + 1), row(V, jj), row(V, jj + 1)
        G[6 * jj:6 * jj + 6] = [dot(u0, u0), dot(u0, u1), dot(u0, v0),
                                dot(v0, u1) if mode == 0 else dot(u0, v1), dot(v0, v0), dot(v0, v1)]
    return G


@pytest.mark.parametrize("k", [0, 1, 2, 4, 8, 13])
def test_kskipmrr_coefficients_bitwise(host_lib, k):
    A = problems.to_scipy(*problems.poisson3d(10, 9, 8))
    n = A.shape[0]
    rng = np.random.default_rng(k)
    Ar = np.zeros((k + 2, n)); Ay = np.zeros((k + 1, n))
    Ar[0] = rng.standard_normal(n); Ay[0] = rng.standard_normal(n) * 0.1
    for j in range(1, k + 2):
        Ar[j] = A.dot(Ar[j - 1])
    for j in range(1, k + 1):
        Ay[j] = A.dot(Ay[j - 1])
    # oracle arrays, filled the way kskipmrr.py:51-59 fills them
    alpha = np.array([np.dot(Ar[j // 2], Ar[j // 2 + j % 2]) for j in range(2 * k + 3)])
    beta = np.zeros(2 * k + 2)
    for j in range(1, 2 * k + 2):
        beta[j] = np.dot(Ay[j // 2], Ar[j // 2 + j % 2])
    delta = np.array([np.dot(Ay[j // 2], Ay[j // 2 + j % 2]) for j in range(2 * k + 1)])
    want = oracle.kskipmrr_scalars(alpha.copy(), beta.copy(), delta.copy(), k)
    G = _gram_layout(0, Ar, Ay, k)
    coef = np.zeros(2 * (k + 1))
    host_lib.host_kskipmrr_coef(G.ctypes.data, k, coef.ctypes.data)
    assert np.array_equal(coef, np.array(want, dtype=np.float64).ravel())


@pytest.mark.parametrize("k", [0, 1, 2, 4, 8, 13])
def test_kskipcg_coefficients_bitwise(host_lib, k):
    A = problems.to_scipy(*problems.banded_spd(700, 5, 2))
    n = A.shape[0]
    rng = np.random.default_rng(100 + k)
    Ar = np.zeros((k + 2, n)); Ap = np.zeros((k + 3, n))
    Ar[0] = rng.standard_normal(n); Ap[0] = Ar[0] + 0.3 * rng.standard_normal(n)
    for j in range(1, k + 1):
        Ar[j] = A.dot(Ar[j - 1])
    for j in range(1, k + 2):
        Ap[j] = A.dot(Ap[j - 1])
    a = np.zeros(2 * k + 2); f = np.zeros(2 * k + 4); c = np.zeros(2 * k + 2)
    for j in range(2 * k + 1):
        a[j] = np.dot(Ar[j // 2], Ar[j // 2 + j % 2])
    for j in range(2 * k + 4):
        f[j] = np.dot(Ap[j // 2], Ap[j // 2 + j % 2])
    for j in range(2 * k + 2):
        c[j] = np.dot(Ar[j // 2], Ap[j // 2 + j % 2])
    want = oracle.kskipcg_scalars(a.copy(), f.copy(), c.copy(), k)
    # device layout: U = Ar rows 0..k, V = Ap rows 0..k+1 (row k+2 of the reference is never written: zeros)
    G = _gram_layout(1, Ar[: k + 1], Ap[: k + 2], k)
    coef = np.zeros(2 * (k + 1))
    host_lib.host_kskipcg_coef(G.ctypes.data, k, coef.ctypes.data)
    assert np.array_equal(coef, np.array(want, dtype=np.float64).ravel())


# ------------------------------------------------------------------------------------------------------------------
# Whole solves on the CPU: numpy does the vector work in the order the CUDA kernels do it, every scalar, the history
# and the stopping rule go through the device's own scalar engine (pk_state.h, host build).  Must equal the oracle —
# and therefore the reference — bit for bit, entry for entry.
import re

_SRC = open(os.path.join(HERE, "..", "parallel_krylov_b200", "csrc", "pk_state.h")).read()
_ENUM = re.search(r"enum PkEpi : int \{(.*?)\};", _SRC, flags=re.S).group(1)
EPI = {}
_next = 0
for _name, _val in re.findall(r"^\s*(EPI_[A-Z_0-9]+)(?:\s*=\s*(\d+))?\s*,", _ENUM, flags=re.M):
    _next = int(_val) if _val else _next
    EPI[_name] = _next
    _next += 1
ALPHA, BETA, GAMMA, ZETA, ETA, DONE, CONV, IT, IDX, COEF = range(10)


class Engine:
    def __init__(self, lib, maxiter, tol, k=0, with_khist=False):
        for fn, res, args in ((lib.hs_new, C.c_void_p, [C.c_longlong, C.c_double, C.c_int, C.c_longlong, C.c_void_p,
                                                         C.c_void_p, C.c_void_p]),
                              (lib.hs_free, None, [C.c_void_p]),
                              (lib.hs_epilogue, None, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                              (lib.hs_gram, None, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                              (lib.hs_get, C.c_double, [C.c_void_p, C.c_int, C.c_int])):
            fn.restype, fn.argtypes = res, args
        self.lib = lib
        n = maxiter + k + 3
        self.res = np.zeros(n)
        self.nosl = np.zeros(n, dtype=np.int64)
        self.khist = np.zeros(n, dtype=np.int64) if with_khist else None
        self.st = lib.hs_new(maxiter, tol, k, n, self.res.ctypes.data, self.nosl.ctypes.data,
                             self.khist.ctypes.data if with_khist else None)

    def epi(self, name, *sums):
        s = np.array(sums, dtype=np.float64)
        self.lib.hs_epilogue(self.st, EPI[name], s.ctypes.data, len(s))

    def gram(self, name, G):
        G = np.ascontiguousarray(G)
        self.lib.hs_gram(self.st, EPI[name], G.ctypes.data, len(G))

    def get(self, what, j=0):
        return self.lib.hs_get(self.st, what, j)

    def history(self):
        m = int(self.get(IDX)) + 1
        return self.res[:m].copy(), self.nosl[:m].copy()


def _emulate(lib, solver, A, b, tol, maxiter, k=0):
    """The launch sequence of csrc/pk_solvers.cu with numpy standing in for the vector kernels."""
    n = b.size
    x = np.zeros(n)
    e = Engine(lib, maxiter, tol, k)
    e.epi("EPI_BNORM", np.dot(b, b))
    if solver == "cg":
        r = b - A.dot(x); p = r.copy()
        e.epi("EPI_CG_INIT", np.dot(r, r))
        while not e.get(DONE):
            v = A.dot(p)
            e.epi("EPI_CG_ALPHA", np.dot(p, v), np.dot(v, v), np.dot(p, p))
            al = e.get(ALPHA)
            x = x + al * p
            r = r - al * v
            e.epi("EPI_CG_BETA", np.dot(r, r))
            p = r + e.get(BETA) * p
    elif solver == "mrr":
        r = b - A.dot(x)
        e.epi("EPI_RES0", np.dot(r, r))
        ar = A.dot(r)
        e.epi("EPI_MRR_FIRST", np.dot(r, ar), np.dot(ar, ar), np.dot(r, r))
        ze = e.get(ZETA)
        y = ze * ar; z = (-ze) * r; r = r - y; x = x - z
        e.epi("EPI_KS_FIRST", np.dot(r, r))
        while not e.get(DONE):
            ar = A.dot(r)
            e.epi("EPI_MRR_GAMMA", np.dot(y, ar), np.dot(ar, ar), np.dot(y, y))
            s = ar - e.get(GAMMA) * y
            e.epi("EPI_MRR_ZETA", np.dot(r, s), np.dot(s, s))
            ze, et = e.get(ZETA), e.get(ETA)
            y = et * y + ze * ar
            z = et * z - ze * r
            r = r - y
            x = x - z
            e.epi("EPI_MRR_STEP", np.dot(r, r))
    elif solver == "kskipcg":
        Ar = np.zeros((k + 1, n)); Ap = np.zeros((k + 2, n))
        Ar[0] = b - A.dot(x); Ap[0] = Ar[0]
        e.epi("EPI_CG_INIT", np.dot(Ar[0], Ar[0]))
        Ap[1] = A.dot(Ap[0])
        while not e.get(DONE):
            for j in range(1, k + 1):                      # two-chain pass
                Ar[j] = A.dot(Ar[j - 1]); Ap[j + 1] = A.dot(Ap[j])
            e.gram("EPI_GRAM_CG", _gram_layout(1, Ar, Ap, k))
            for j in range(k + 1):
                al, be = e.get(COEF, 2 * j), e.get(COEF, 2 * j + 1)
                x = x + al * Ap[0]
                Ar[0] = Ar[0] - al * Ap[1]
                Ap[0] = Ar[0] + be * Ap[0]
                if j == k:
                    e.epi("EPI_KS_TRIP_END", np.dot(Ar[0], Ar[0]))
                Ap[1] = A.dot(Ap[0])
    elif solver == "kskipmrr":
        Ar = np.zeros((k + 2, n)); Ay = np.zeros((k + 1, n))
        Ar[0] = b - A.dot(x)
        e.epi("EPI_RES0", np.dot(Ar[0], Ar[0]))
        Ar[1] = A.dot(Ar[0])
        e.epi("EPI_MRR_FIRST", np.dot(Ar[0], Ar[1]), np.dot(Ar[1], Ar[1]), np.dot(Ar[0], Ar[0]))
        ze = e.get(ZETA)
        Ay[0] = ze * Ar[1]; z = (-ze) * Ar[0]; Ar[0] = Ar[0] - Ay[0]; x = x - z
        e.epi("EPI_KS_FIRST", np.dot(Ar[0], Ar[0]))
        Ar[1] = A.dot(Ar[0])
        while not e.get(DONE):
            for j in range(1, k + 1):
                Ar[j + 1] = A.dot(Ar[j]); Ay[j] = A.dot(Ay[j - 1])
            e.gram("EPI_GRAM_MRR", _gram_layout(0, Ar, Ay, k))
            for j in range(k + 1):
                ze, et = e.get(COEF, 2 * j), e.get(COEF, 2 * j + 1)
                Ay[0] = et * Ay[0] + ze * Ar[1]
                z = et * z - ze * Ar[0]
                Ar[0] = Ar[0] - Ay[0]
                x = x - z
                if j == k:
                    e.epi("EPI_KS_TRIP_END", np.dot(Ar[0], Ar[0]))
                Ar[1] = A.dot(Ar[0])
    res, nosl = e.history()
    out = {"residual": res, "nosl": nosl, "converged": bool(e.get(CONV))}
    lib.hs_free(e.st)
    return x, out


@pytest.mark.parametrize("solver,k,cap", [("cg", 0, None), ("cg", 0, 17), ("mrr", 0, None), ("mrr", 0, 9),
                                          ("kskipcg", 0, None), ("kskipcg", 3, None), ("kskipcg", 4, 22),
                                          ("kskipmrr", 0, None), ("kskipmrr", 2, None), ("kskipmrr", 8, None),
                                          ("kskipmrr", 4, 22)])
def test_device_scalar_engine_drives_whole_solves_bitwise(host_lib, solver, k, cap):
    A = problems.to_scipy(*problems.poisson3d(9, 8, 7))
    b = problems.rhs(A.shape[0], "randn", 0)
    maxiter = A.shape[0] if cap is None else cap
    kw = {"k": k} if solver.startswith("kskip") else {}
    xo, io = oracle.SOLVERS[solver](A, b.copy(), tol=1e-8, maxiter=maxiter, **kw)
    x, info = _emulate(host_lib, solver, A, b, 1e-8, maxiter, k)
    assert np.array_equal(info["nosl"], io["nosl"])
    assert np.array_equal(info["residual"], io["residual"])
    assert np.array_equal(x, xo)
    assert info["converged"] == io["converged"]


def _emulate_adaptive(lib, A, b, tol, maxiter, k0):
    """Solve::adaptive() of csrc/pk_solvers.cu restated with numpy standing in for the vector kernels.  The guard, the
    rollback decision, the lowering of k and every stopping decision are taken by the DEVICE engine (EPI_ADAPT_GUARD /
    EPI_ADAPT_STEP / EPI_ADAPT_TRIP_END of pk_state.h); this driver only does what the predicated kernels do."""
    n = b.size
    x = np.zeros(n)
    e = Engine(lib, maxiter, tol, k0, with_khist=True)
    ROLLBACK, K = 10, 11
    Ar = np.zeros((k0 + 2, n)); Ay = np.zeros((k0 + 1, n))
    e.epi("EPI_BNORM", np.dot(b, b))
    Ar[0] = b - A.dot(x)
    e.epi("EPI_RES0", np.dot(Ar[0], Ar[0]))
    best_x = x.copy()

    def opening(epi_name):
        nonlocal x, z
        Ar[1] = A.dot(Ar[0])
        e.epi("EPI_MRR_FIRST", np.dot(Ar[0], Ar[1]), np.dot(Ar[1], Ar[1]), np.dot(Ar[0], Ar[0]))
        ze = e.get(ZETA)
        Ay[0] = ze * Ar[1]; z = (-ze) * Ar[0]; Ar[0] = Ar[0] - Ay[0]; x = x - z
        e.epi(epi_name, np.dot(Ar[0], Ar[0]))
        Ar[1] = A.dot(Ar[0])

    z = None
    opening("EPI_ADAPT_FIRST")
    while True:
        e.epi("EPI_ADAPT_GUARD")                       # k_scalar at the top of every trip
        if e.get(DONE):
            break
        if e.get(ROLLBACK):                            # k_adapt_save + the only_rollback kernels
            x = best_x.copy()
            Ar[0] = b - A.dot(x)
            opening("EPI_ADAPT_STEP")
            if e.get(DONE):
                break
        else:
            best_x = x.copy()
        k = int(e.get(K))                              # kernels read the current k from the device state
        for j in range(1, k + 1):
            Ar[j + 1] = A.dot(Ar[j]); Ay[j] = A.dot(Ay[j - 1])
        e.gram("EPI_GRAM_MRR", _gram_layout(0, Ar[: k + 2], Ay[: k + 1], k))
        for j in range(k + 1):
            ze, et = e.get(COEF, 2 * j), e.get(COEF, 2 * j + 1)
            Ay[0] = et * Ay[0] + ze * Ar[1]
            z = et * z - ze * Ar[0]
            Ar[0] = Ar[0] - Ay[0]
            x = x - z
            if j == k:
                e.epi("EPI_ADAPT_TRIP_END", np.dot(Ar[0], Ar[0]))
            Ar[1] = A.dot(Ar[0])
    m = int(e.get(IDX)) + 1
    out = {"residual": e.res[:m].copy(), "nosl": e.nosl[:m].copy(), "khistory": e.khist[:m].copy(),
           "converged": bool(e.get(CONV))}
    lib.hs_free(e.st)
    return x, out


@pytest.mark.parametrize("k,n2d", [(4, 24), (12, 48), (16, 48)])
def test_adaptive_guard_logic_bitwise(host_lib_pow, k, n2d):
    """The rollback / k-lowering flow (guard fires for k >= 10 on these systems) against the oracle, bit for bit.
    These runs are chaotic (k >= 10), so the build that squares with libm pow() like the reference is used: with it
    even the 372-iteration k=16 run with seven rollbacks is reproduced exactly."""
    host_lib = host_lib_pow
    A = problems.to_scipy(*problems.poisson2d(n2d))
    b = problems.rhs(A.shape[0], "randn", 0)
    xo, io = oracle.adaptivekskipmrr(A, b.copy(), tol=1e-8, maxiter=2000, k=k)
    x, info = _emulate_adaptive(host_lib, A, b, 1e-8, 2000, k)
    assert np.array_equal(info["nosl"], io["nosl"])
    assert np.array_equal(info["khistory"], io["khistory"])
    assert np.array_equal(info["residual"], io["residual"])
    assert np.array_equal(x, xo)
    if k >= 12:
        assert io["khistory"][-1] < k          # the guard really fired in this case


def test_product_and_pow_squares_differ_by_at_most_one_ulp(host_lib, host_lib_pow):
    """x*x (device) vs pow(x, 2) (reference): the coefficient sequences of a trip agree to ~1 ulp per square."""
    A = problems.to_scipy(*problems.poisson3d(10, 9, 8))
    n = A.shape[0]
    worst = 0.0
    for seed in range(40):
        rng = np.random.default_rng(seed)
        k = 6
        Ar = np.zeros((k + 2, n)); Ay = np.zeros((k + 1, n))
        Ar[0] = rng.standard_normal(n); Ay[0] = 0.1 * rng.standard_normal(n)
        for j in range(1, k + 2):
            Ar[j] = A.dot(Ar[j - 1])
        for j in range(1, k + 1):
            Ay[j] = A.dot(Ay[j - 1])
        G = _gram_layout(0, Ar, Ay, k)
        c1, c2 = np.zeros(2 * (k + 1)), np.zeros(2 * (k + 1))
        host_lib.host_kskipmrr_coef(G.ctypes.data, k, c1.ctypes.data)
        host_lib_pow.host_kskipmrr_coef(G.ctypes.data, k, c2.ctypes.data)
        worst = max(worst, float(np.max(np.abs(c1 - c2) / np.abs(c2))))
    assert worst < 1e-9        # the recurrence amplifies the odd ulp, it does not change the picture


@pytest.mark.parametrize("cap", [1, 2, 7, 22])
def test_adaptive_guard_iteration_caps_bitwise(host_lib, cap):
    """`while i < maxiter` is evaluated by the device guard at the top of every trip: tiny and mid-trip caps end exactly
    where the reference ends (k-skip trips may overshoot the cap by up to k)."""
    A = problems.to_scipy(*problems.poisson2d(24))
    b = problems.rhs(A.shape[0], "randn", 0)
    xo, io = oracle.adaptivekskipmrr(A, b.copy(), tol=1e-8, maxiter=cap, k=4)
    x, info = _emulate_adaptive(host_lib, A, b, 1e-8, cap, 4)
    assert np.array_equal(info["nosl"], io["nosl"])
    assert np.array_equal(info["residual"], io["residual"])
    assert np.array_equal(info["khistory"], io["khistory"])
    assert np.array_equal(x, xo)
    assert info["converged"] is False

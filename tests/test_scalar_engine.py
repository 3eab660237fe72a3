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


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostscalars") / "libhostscalars.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(out),
                    os.path.join(HERE, "host_scalars.cpp")], check=True)
    lib = C.CDLL(str(out))
    for fn in (lib.host_kskipcg_coef, lib.host_kskipmrr_coef):
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = None
    return lib


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

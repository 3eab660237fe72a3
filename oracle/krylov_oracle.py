"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by ``parallel_krylov_b200`` (the product path is
CUDA-only and fails loudly without its extension).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.

A numpy restatement of the five v3 CPU solvers of 5enxia/parallel-krylov, written from the algorithm
(not transliterated): the same floating-point operations in the same order on the same library calls
(``A.dot`` → scipy ``csr_matvec`` / OpenBLAS ``dgemv``; ``numpy.dot`` → OpenBLAS ``ddot``;
``numpy.linalg.norm`` → ``dnrm2``-style), so that histories agree with the reference bit for bit when run on
the same machine.  Third-party arithmetic the reference delegates to (not vendored, unpinned in
/root/reference/requirements.txt:1-2): numpy (2.3.5 here, OpenBLAS 0.3.30) and scipy (1.18.1 here).

Pinning status: the reference has NO tests, golden vectors or fixtures, so this oracle is pinned
against outputs of the reference itself: ``oracle/gen_golden.py`` imports the unmodified
``/root/reference/v3/cpu`` functions in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays this oracle against every fixture.

Each function cites the reference lines it follows.  Differences that are NOT arithmetic: no printing, no
wall-clock inside (callers time it), ``mat`` is anything with ``.dot(vector)`` (scipy CSR or ndarray) for every
solver (the reference's ``kskipcg``/``adaptivekskipmrr`` call ``numpy.dot(A, v)`` and hence take dense A only,
/root/reference/v3/cpu/kskipcg.py:21,37 — for an ndarray ``A.dot(v)`` is the same BLAS call).
"""
from __future__ import annotations

import time

import numpy as np
from numpy import dot
from numpy.linalg import norm

F64 = np.float64


class _Log:
    """residual / nosl bookkeeping of ``init`` (/root/reference/v3/cpu/common.py:22-36)."""

    def __init__(self, b, x, maxiter):
        self.bnorm = norm(b)
        self.n = b.size
        self.x = x if isinstance(x, np.ndarray) else np.zeros(self.n, dtype=F64)
        self.maxiter = self.n if maxiter is None else maxiter
        self.res = np.zeros(self.maxiter + 1, F64)
        self.nosl = np.zeros(self.maxiter + 1, np.int64)
        self.t0 = None

    def start(self):
        """Where the reference calls start() (e.g. v3/cpu/cg.py:18): after init and the initial residual."""
        self.t0 = time.perf_counter()

    def info(self, last, converged, **extra):
        out = {"nosl": self.nosl[: last + 1], "residual": self.res[: last + 1], "converged": converged,
               "time": (time.perf_counter() - self.t0) if self.t0 is not None else None}
        out.update(extra)
        return out


def cg(mat, b, x=None, tol=1e-05, maxiter=None):
    """Conjugate gradients — /root/reference/v3/cpu/cg.py:7-48."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    r = b - mat.dot(x)                       # cg.py:12
    p = r.copy()                             # :13
    gamma = dot(r, r)                        # :14
    it, ok = 0, False
    lg.start()                               # :18
    while it < lg.maxiter:                   # :19
        lg.res[it] = norm(r) / lg.bnorm      # :21
        if lg.res[it] < tol:                 # :22
            ok = True
            break
        v = mat.dot(p)                       # :27
        sigma = dot(p, v)                    # :28
        alpha = gamma / sigma                # :29
        x += alpha * p                       # :30
        r -= alpha * v                       # :31
        gamma_prev = gamma
        gamma = dot(r, r)                    # :33
        beta = gamma / gamma_prev            # :34
        p = r + beta * p                     # :35
        it += 1
        lg.nosl[it] = it                     # :37
    else:
        lg.res[it] = norm(r) / lg.bnorm      # :40
    return x, lg.info(it, ok)


def _mrr_first_step(mat, r):
    """The steepest-residual opening step shared by mrr / kskipmrr / adaptivekskipmrr
    (/root/reference/v3/cpu/mrr.py:18-23). Returns (Ar, zeta)."""
    ar = mat.dot(r)
    zeta = dot(r, ar) / dot(ar, ar)
    return ar, zeta


def mrr(mat, b, x=None, tol=1e-05, maxiter=None):
    """Minimised residual method based on the three-term recurrence — /root/reference/v3/cpu/mrr.py:7-61."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    r = b - mat.dot(x)                       # mrr.py:12
    lg.res[0] = norm(r) / lg.bnorm           # :13
    lg.start()                               # :17
    ar, zeta = _mrr_first_step(mat, r)       # :18-19
    y = zeta * ar                            # :20
    z = -zeta * r                            # :21
    r -= y                                   # :22
    x -= z                                   # :23
    lg.nosl[1] = 1
    it, ok = 1, False
    while it < lg.maxiter:                   # :28
        lg.res[it] = norm(r) / lg.bnorm      # :30
        if lg.res[it] < tol:
            ok = True
            break
        ar = mat.dot(r)                      # :36
        mu = dot(y, y)                       # :37
        nu = dot(y, ar)                      # :38
        gamma = nu / mu                      # :39
        s = ar - gamma * y                   # :40
        rs = dot(r, s)                       # :41
        ss = dot(s, s)                       # :42
        zeta = rs / ss                       # :43
        eta = -zeta * gamma                  # :44
        y = eta * y + zeta * ar              # :45
        z = eta * z - zeta * r               # :46
        r -= y                               # :47
        x -= z                               # :48
        it += 1
        lg.nosl[it] = it
    else:
        lg.res[it] = norm(r) / lg.bnorm      # :53
    return x, lg.info(it, ok)


def kskipcg_scalars(a, f, c, k):
    """All k+1 (alpha, beta) pairs of one outer trip from the Gram scalars alone — the scalar part of
    /root/reference/v3/cpu/kskipcg.py:51-52 and :59-68 (the recurrence never touches a vector).
    ``a``, ``f``, ``c`` are updated in place exactly as the reference does."""
    out = []
    alpha = a[0] / f[1]                                  # :51
    beta = alpha ** 2 * f[2] / a[0] - 1                  # :52
    out.append((alpha, beta))
    for j in range(k):                                   # :59
        for l in range(2 * (k - j) + 1):                 # :60
            a[l] += alpha * (alpha * f[l + 2] - 2 * c[l + 1])   # :61
            d = c[l] - alpha * f[l + 1]                  # :62
            c[l] = a[l] + d * beta                       # :63
            f[l] = c[l] + beta * (d + beta * f[l])       # :64
        alpha = a[0] / f[1]                              # :67
        beta = alpha ** 2 * f[2] / a[0] - 1              # :68
        out.append((alpha, beta))
    return out


def kskipcg(mat, b, x=None, tol=1e-05, maxiter=None, k=0):
    """k-skip CG — /root/reference/v3/cpu/kskipcg.py:8-87. One history entry per outer trip (= k+1 CG steps)."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    n = lg.n
    Ar = np.zeros((k + 2, n), F64)           # :14
    Ap = np.zeros((k + 3, n), F64)           # :15  (row k+2 is never written → f[2k+3] == 0)
    a = np.zeros(2 * k + 2, F64)
    f = np.zeros(2 * k + 4, F64)
    c = np.zeros(2 * k + 2, F64)
    Ar[0] = b - mat.dot(x)                   # :21
    Ap[0] = Ar[0]                            # :22
    it, idx, ok = 0, 0, False
    lg.start()                               # :27
    while it < lg.maxiter:                   # :28
        lg.res[idx] = norm(Ar[0]) / lg.bnorm  # :30
        if lg.res[idx] < tol:
            ok = True
            break
        for j in range(1, k + 1):            # :36-37
            Ar[j] = mat.dot(Ar[j - 1])
        for j in range(1, k + 2):            # :38-39
            Ap[j] = mat.dot(Ap[j - 1])
        for j in range(2 * k + 1):           # :40-42
            a[j] = dot(Ar[j // 2], Ar[j // 2 + j % 2])
        for j in range(2 * k + 4):           # :43-45
            f[j] = dot(Ap[j // 2], Ap[j // 2 + j % 2])
        for j in range(2 * k + 2):           # :46-48
            c[j] = dot(Ar[j // 2], Ap[j // 2 + j % 2])
        # scalar recurrences interleaved with the k+1 vector steps (:51-72); computing them up front is
        # the same arithmetic because they depend on a/f/c only.
        for alpha, beta in kskipcg_scalars(a, f, c, k):
            x += alpha * Ap[0]               # :53 / :69
            Ar[0] -= alpha * Ap[1]           # :54 / :70
            Ap[0] = Ar[0] + beta * Ap[0]     # :55 / :71
            Ap[1] = mat.dot(Ap[0])           # :56 / :72
        it += k + 1                          # :74
        idx += 1
        lg.nosl[idx] = it
    else:
        lg.res[idx] = norm(Ar[0]) / lg.bnorm
    return x, lg.info(idx, ok)


def kskipmrr_scalars(alpha, beta, delta, k):
    """All k+1 (zeta, eta) pairs of one outer trip — scalar part of
    /root/reference/v3/cpu/kskipmrr.py:62-64 and :72-88; arrays updated in place like the reference."""
    out = []
    d = alpha[2] * delta[0] - beta[1] ** 2               # :62
    zeta = alpha[1] * delta[0] / d                       # :63
    eta = -alpha[1] * beta[1] / d                        # :64
    out.append((zeta, eta))
    for j in range(k):                                   # :72
        delta[0] = zeta ** 2 * alpha[2] + eta * zeta * beta[1]        # :73
        alpha[0] -= zeta * alpha[1]                                   # :74
        delta[1] = eta ** 2 * delta[1] + 2 * eta * zeta * beta[2] + zeta ** 2 * alpha[3]   # :75-76
        beta[1] = eta * beta[1] + zeta * alpha[2] - delta[1]          # :77
        alpha[1] = -beta[1]                                           # :78
        for l in range(2, 2 * (k - j) + 1):                           # :79
            delta[l] = eta ** 2 * delta[l] + 2 * eta * zeta * beta[l + 1] + zeta ** 2 * alpha[l + 2]   # :80-81
            tau = eta * beta[l] + zeta * alpha[l + 1]                 # :82
            beta[l] = tau - delta[l]                                  # :83
            alpha[l] -= tau + beta[l]                                 # :84
        d = alpha[2] * delta[0] - beta[1] ** 2                        # :86
        zeta = alpha[1] * delta[0] / d                                # :87
        eta = -alpha[1] * beta[1] / d                                 # :88
        out.append((zeta, eta))
    return out


def _kskipmrr_trip(mat, Ar, Ay, z, x, alpha, beta, delta, k):
    """Basis + Gram + k+1 MrR steps of one outer trip (kskipmrr.py:45-93 ≡ adaptivekskipmrr.py:77-128)."""
    for j in range(1, k + 2):                # :45-46
        Ar[j] = mat.dot(Ar[j - 1])
    for j in range(1, k + 1):                # :47-48
        Ay[j] = mat.dot(Ay[j - 1])
    for j in range(2 * k + 3):               # :51-53
        alpha[j] = dot(Ar[j // 2], Ar[j // 2 + j % 2])
    for j in range(1, 2 * k + 2):            # :54-56
        beta[j] = dot(Ay[j // 2], Ar[j // 2 + j % 2])
    for j in range(2 * k + 1):               # :57-59
        delta[j] = dot(Ay[j // 2], Ay[j // 2 + j % 2])
    for zeta, eta in kskipmrr_scalars(alpha, beta, delta, k):
        Ay[0] = eta * Ay[0] + zeta * Ar[1]   # :65 / :89
        z = eta * z - zeta * Ar[0]           # :66 / :90
        Ar[0] -= Ay[0]                       # :67 / :91
        Ar[1] = mat.dot(Ar[0])               # :68 / :92
        x -= z                               # :69 / :93
    return z


def kskipmrr(mat, b, x=None, tol=1e-05, maxiter=None, k=0):
    """k-skip MrR — /root/reference/v3/cpu/kskipmrr.py:8-108."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    n = lg.n
    Ar = np.zeros((k + 2, n), F64)
    Ay = np.zeros((k + 1, n), F64)
    alpha = np.zeros(2 * k + 3, F64)
    beta = np.zeros(2 * k + 2, F64)
    delta = np.zeros(2 * k + 1, F64)
    Ar[0] = b - mat.dot(x)                   # :21
    lg.res[0] = norm(Ar[0]) / lg.bnorm       # :22
    lg.start()                               # :25
    Ar[1], zeta = _mrr_first_step(mat, Ar[0])   # :26-27
    Ay[0] = zeta * Ar[1]                     # :28
    z = -zeta * Ar[0]                        # :29
    Ar[0] -= Ay[0]                           # :30
    x -= z                                   # :31
    lg.nosl[1] = 1
    it, idx, ok = 1, 1, False
    while it < lg.maxiter:                   # :37
        lg.res[idx] = norm(Ar[0]) / lg.bnorm
        if lg.res[idx] < tol:
            ok = True
            break
        z = _kskipmrr_trip(mat, Ar, Ay, z, x, alpha, beta, delta, k)
        it += k + 1                          # :95
        idx += 1
        lg.nosl[idx] = it
    else:
        lg.res[idx] = norm(Ar[0]) / lg.bnorm
    return x, lg.info(idx, ok)


def adaptivekskipmrr(mat, b, x=None, tol=1e-05, maxiter=None, k=0):
    """Adaptive k-skip MrR — normative variant /root/reference/v3/cpu/adaptivekskipmrr.py:8-141:
    kskipmrr plus a residual-growth guard that rolls x back to the best iterate, takes one plain MrR step
    and lowers k (never below 1)."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    n = lg.n
    Ar = np.zeros((k + 3, n), F64)           # :13
    Ay = np.zeros((k + 2, n), F64)           # :14
    alpha = np.zeros(2 * k + 3, F64)
    beta = np.zeros(2 * k + 2, F64)
    delta = np.zeros(2 * k + 1, F64)
    khist = np.zeros(n + 1, np.int64)        # :18 (length N+1, not maxiter+1)
    khist[0] = k
    Ar[0] = b - mat.dot(x)                   # :22
    lg.res[0] = norm(Ar[0]) / lg.bnorm
    best_res = lg.res[0]                     # :24
    best_x = None
    lg.start()                               # :27
    Ar[1], zeta = _mrr_first_step(mat, Ar[0])   # :28-31
    Ay[0] = zeta * Ar[1]
    z = -zeta * Ar[0]
    Ar[0] -= Ay[0]
    x -= z
    lg.nosl[1] = 1
    khist[1] = k
    it, idx, ok = 1, 1, False
    while it < lg.maxiter:                   # :42
        lg.res[idx] = norm(Ar[0]) / lg.bnorm
        if lg.res[idx] > best_res:           # :45 residual grew → roll back
            x = best_x.copy()                # :47
            Ar[0] = b - mat.dot(x)           # :48
            Ar[1], zeta = _mrr_first_step(mat, Ar[0])   # :49-52
            Ay[0] = zeta * Ar[1]
            z = -zeta * Ar[0]
            Ar[0] -= Ay[0]
            x -= z
            it += 1                          # :58
            idx += 1
            lg.res[idx] = norm(Ar[0]) / lg.bnorm
            lg.nosl[idx] = it
            if k > 1:                        # :64
                k -= 1
            khist[idx] = k
        else:
            best_res = lg.res[idx]           # :68
            best_x = x.copy()                # :69
        if lg.res[idx] < tol:                # :72
            ok = True
            break
        z = _kskipmrr_trip(mat, Ar, Ay, z, x, alpha, beta, delta, k)
        it += k + 1                          # :130
        idx += 1
        lg.nosl[idx] = it
        khist[idx] = k
    else:
        lg.res[idx] = norm(Ar[0]) / lg.bnorm
    return x, lg.info(idx, ok, khistory=khist[: idx + 1], final_k=k)


def cgcg(mat, b, x=None, tol=1e-05, maxiter=None, M=None):
    """Chronopoulos–Gear CG (one reduction point per iteration), optionally preconditioned —
    /root/reference/v1/threads/pipeline/chronopoulos_gear.py:7-56, SURVEY.md §8f rank 4.

    The reference file is a sketch that cannot be imported (``from .common import ...`` — there is no
    ``pipeline/common.py``) and never updates ``old_gamma`` inside its loop (:30-31 vs :49), which makes ``beta`` wrong from
    the second iteration on.  Restated here with those two repairs and nothing else changed in the arithmetic:
    ``oracle/gen_golden_cgcg.py`` executes the reference text with exactly these repairs and pins this function on its
    residual histories.  ``M``: None (u = r) or the 1-D diagonal of a Jacobi preconditioner (u = r / M; the reference's
    ``ilu.solve(r)``, :25, :45).  v3 conventions for the interface: ``maxiter`` caps the number of solution updates,
    ``nosl[i] = i`` for every recorded entry (the reference leaves the last entry unset when it breaks, :41-52)."""
    lg = _Log(b, x, maxiter)
    x = lg.x
    minv = (lambda v: v.copy()) if M is None else (lambda v: v / M)
    r = b - mat.dot(x)                        # :22
    lg.res[0] = norm(r) / lg.bnorm            # :23
    u = minv(r)                               # :25
    w = mat.dot(u)                            # :26
    alpha = dot(r, u) / dot(w, u)             # :28
    beta = 0.0                                # :29
    gamma = dot(r, u)                         # :30
    p = np.zeros(lg.n, F64)                   # :33
    s = np.zeros(lg.n, F64)                   # :34
    it, ok = 0, False
    lg.start()
    while it < lg.maxiter:                    # :36
        p = u + beta * p                      # :37
        s = w + beta * s                      # :38
        x += alpha * p                        # :39
        r -= alpha * s                        # :40
        it += 1
        lg.nosl[it] = it
        lg.res[it] = norm(r) / lg.bnorm       # :41
        if lg.res[it] < tol:                  # :42
            ok = True
            break
        u = minv(r)                           # :45
        w = mat.dot(u)                        # :46
        gamma_new = dot(r, u)                 # :47
        delta = dot(w, u)                     # :48
        beta = gamma_new / gamma              # :49 (with old_gamma kept up to date)
        alpha = gamma_new / (delta - beta * gamma_new / alpha)   # :50
        gamma = gamma_new
    return x, lg.info(it, ok)


SOLVERS = {
    "cgcg": cgcg,
    "cg": cg,
    "mrr": mrr,
    "kskipcg": kskipcg,
    "kskipmrr": kskipmrr,
    "adaptivekskipmrr": adaptivekskipmrr,
}


def true_relres(mat, b, x):
    """‖b − A x‖ / ‖b‖ — the acceptance quantity of BASELINE.json's north_star."""
    return float(norm(b - mat.dot(x)) / norm(b))

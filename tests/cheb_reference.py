"""TEST INFRASTRUCTURE: numpy restatement of the opt-in Chebyshev-basis k-skip MrR (csrc/pk_scalars.h:
pk_kskipmrr_coef_cheb, csrc/pk_solvers.cu: kskipmrr_chebyshev).  There is no reference implementation of this variant
(SURVEY.md §8f rank 3 is a follow-up the reference never wrote): in exact arithmetic its iterates are those of plain MrR
(/root/reference/v3/cpu/mrr.py), which is what the tests compare it with."""
import numpy as np
from numpy import dot
from numpy.linalg import norm


def gershgorin(A):
    d = A.diagonal()
    R = np.asarray(abs(A).sum(axis=1)).ravel() - np.abs(d)
    return float((d - R).min()), float((d + R).max())


def mul_a(m, L, c, d):
    """(u, T_l A w) for l < L from m_l = (u, T_l w), l <= L."""
    out = np.zeros(L)
    out[0] = c * m[1] + d * m[0]
    for l in range(1, L):
        out[l] = (c * (m[l + 1] + m[l - 1])) / 2.0 + d * m[l]
    return out


def coefficients(G, k, c, d):
    """(zeta_j, eta_j), j = 0..k, from the Gram sums G[6*jj + t] of the Chebyshev basis (layout of pk_gram, MrR mode)."""
    al = np.zeros(2 * k + 3); be = np.zeros(2 * k + 2); de = np.zeros(2 * k + 1)
    al[0], al[1] = G[0], G[1]
    for j in range(2, 2 * k + 3):
        al[j] = 2.0 * G[6 * (j >> 1) + (j & 1)] - al[j & 1]
    be[0], be[1] = G[2], G[3]
    for j in range(2, 2 * k + 2):
        be[j] = 2.0 * G[6 * (j >> 1) + 2 + (j & 1)] - be[j & 1]
    de[0] = G[4]
    if k >= 1:
        de[1] = G[5]
    for j in range(2, 2 * k + 1):
        de[j] = 2.0 * G[6 * (j >> 1) + 4 + (j & 1)] - de[j & 1]
    Ld = 2 * k + 1
    coef = np.zeros(2 * (k + 1))
    for j in range(k + 1):
        Aal = mul_a(al, Ld + 1, c, d)
        AAal = mul_a(Aal, Ld, c, d)
        Abe = mul_a(be, Ld, c, d)
        a1, a2, b1, d0 = Aal[0], AAal[0], Abe[0], de[0]
        dd = a2 * d0 - b1 * b1
        zeta = (a1 * d0) / dd
        eta = ((-a1) * b1) / dd
        coef[2 * j], coef[2 * j + 1] = zeta, eta
        if j == k:
            break
        for l in range(Ld):
            den = ((eta * eta) * de[l] + ((2.0 * eta) * zeta) * Abe[l]) + (zeta * zeta) * AAal[l]
            tau = eta * be[l] + zeta * Aal[l]
            ben = tau - den
            al[l] = al[l] - (tau + ben)
            be[l] = ben
            de[l] = den
        Ld -= 2
    return coef


def gram_layout(U, V, k):
    G = np.zeros(6 * (k + 2))
    row = lambda M, j: M[j] if j < M.shape[0] else None
    dt = lambda a, b: 0.0 if a is None or b is None else float(dot(a, b))
    for jj in range(k + 2):
        u0, u1, v0, v1 = row(U, jj), row(U, jj + 1), row(V, jj), row(V, jj + 1)
        G[6 * jj:6 * jj + 6] = [dt(u0, u0), dt(u0, u1), dt(u0, v0), dt(v0, u1), dt(v0, v0), dt(v0, v1)]
    return G


def kskipmrr_chebyshev(A, b, tol=1e-8, maxiter=None, k=8, bounds=None, coef_fn=coefficients):
    """Launch order of Solve::kskipmrr_chebyshev with numpy standing in for the vector kernels."""
    n = b.size
    maxiter = n if maxiter is None else maxiter
    x = np.zeros(n)
    bn = norm(b)
    lo, hi = bounds if bounds is not None else gershgorin(A)
    c, d = 0.5 * (hi - lo), 0.5 * (hi + lo)
    U = np.zeros((k + 2, n)); V = np.zeros((k + 1, n))
    res, nosl = [], []
    r = b - A.dot(x)
    res.append(norm(r) / bn); nosl.append(0)
    AR = A.dot(r)
    zeta = dot(r, AR) / dot(AR, AR)
    y = zeta * AR; z = (-zeta) * r; r = r - y; x = x - z
    AR = A.dot(r)
    it = 1
    res.append(norm(r) / bn); nosl.append(1)
    converged = False
    while True:
        if it < maxiter:
            if res[-1] < tol:
                converged = True
                break
        else:
            break
        U[0], V[0] = r, y
        U[1] = (1.0 / c) * AR + (-d / c) * U[0]
        for j in range(1, k + 1):
            U[j + 1] = ((2.0 / c) * A.dot(U[j]) + (-2.0 * d / c) * U[j]) + (-1.0) * U[j - 1]
            if j == 1:
                V[1] = ((1.0 / c) * A.dot(V[0]) + (-d / c) * V[0])
            else:
                V[j] = ((2.0 / c) * A.dot(V[j - 1]) + (-2.0 * d / c) * V[j - 1]) + (-1.0) * V[j - 2]
        coef = coef_fn(gram_layout(U, V, k), k, c, d)
        for j in range(k + 1):
            ze, et = coef[2 * j], coef[2 * j + 1]
            y = et * y + ze * AR
            z = et * z - ze * r
            r = r - y
            x = x - z
            AR = A.dot(r)
        it += k + 1
        res.append(norm(r) / bn); nosl.append(it)
    return x, {"residual": np.array(res), "nosl": np.array(nosl), "converged": converged}


# ---- k-skip CG on the Chebyshev basis (mirror of the MrR variant; pk_scalars.h: pk_kskipcg_coef_cheb) ------------------
def coefficients_cg(G, k, c, d):
    """(alpha_j, beta_j), j = 0..k, from the Gram sums of the Chebyshev basis (layout of pk_gram, CG mode)."""
    a = np.zeros(2 * k + 1); f = np.zeros(2 * k + 3); cc = np.zeros(2 * k + 2)
    a[0] = G[0]
    if k >= 1:
        a[1] = G[1]
    for j in range(2, 2 * k + 1):
        a[j] = 2.0 * G[6 * (j >> 1) + (j & 1)] - a[j & 1]
    cc[0], cc[1] = G[2], G[3]
    for j in range(2, 2 * k + 2):
        cc[j] = 2.0 * G[6 * (j >> 1) + 2 + (j & 1)] - cc[j & 1]
    f[0], f[1] = G[4], G[5]
    for j in range(2, 2 * k + 3):
        f[j] = 2.0 * G[6 * (j >> 1) + 4 + (j & 1)] - f[j & 1]
    coef = np.zeros(2 * (k + 1))
    for j in range(k + 1):
        L = 2 * (k - j) + 1
        Af = mul_a(f, L + 1, c, d)
        AAf = mul_a(Af, L, c, d)
        Ac = mul_a(cc, L, c, d)
        alpha = a[0] / Af[0]
        beta = ((alpha * alpha) * AAf[0]) / a[0] - 1.0
        coef[2 * j], coef[2 * j + 1] = alpha, beta
        if j == k:
            break
        for l in range(L):
            a[l] = a[l] + alpha * (alpha * AAf[l] - 2.0 * Ac[l])
            dd = cc[l] - alpha * Af[l]
            cc[l] = a[l] + dd * beta
            f[l] = cc[l] + beta * (dd + beta * f[l])
    return coef


def gram_layout_cg(U, V, k):
    G = np.zeros(6 * (k + 2))
    row = lambda M, j: M[j] if j < M.shape[0] else None
    dt = lambda a, b: 0.0 if a is None or b is None else float(dot(a, b))
    for jj in range(k + 2):
        u0, u1, v0, v1 = row(U, jj), row(U, jj + 1), row(V, jj), row(V, jj + 1)
        G[6 * jj:6 * jj + 6] = [dt(u0, u0), dt(u0, u1), dt(u0, v0), dt(u0, v1), dt(v0, v0), dt(v0, v1)]
    return G


def kskipcg_chebyshev(A, b, tol=1e-8, maxiter=None, k=8, bounds=None, coef_fn=coefficients_cg):
    """Launch order of Solve::kskipcg_chebyshev with numpy standing in for the vector kernels."""
    n = b.size
    maxiter = n if maxiter is None else maxiter
    x = np.zeros(n)
    bn = norm(b)
    lo, hi = bounds if bounds is not None else gershgorin(A)
    c, d = 0.5 * (hi - lo), 0.5 * (hi + lo)
    r = b - A.dot(x)
    p = r.copy()
    AP = A.dot(p)
    U = np.zeros((k + 1, n)); V = np.zeros((k + 2, n))
    res, nosl, it = [], [0], 0
    converged = False
    while it < maxiter:
        res.append(norm(r) / bn)
        if res[-1] < tol:
            converged = True
            break
        U[0], V[0] = r, p
        V[1] = (1.0 / c) * AP + (-d / c) * V[0]
        for j in range(1, k + 1):
            if j == 1:
                U[1] = (1.0 / c) * A.dot(U[0]) + (-d / c) * U[0]
            else:
                U[j] = ((2.0 / c) * A.dot(U[j - 1]) + (-2.0 * d / c) * U[j - 1]) + (-1.0) * U[j - 2]
            V[j + 1] = ((2.0 / c) * A.dot(V[j]) + (-2.0 * d / c) * V[j]) + (-1.0) * V[j - 1]
        coef = coef_fn(gram_layout_cg(U, V, k), k, c, d)
        for j in range(k + 1):
            al, be = coef[2 * j], coef[2 * j + 1]
            x = x + al * p
            r = r - al * AP
            p = r + be * p
            AP = A.dot(p)
        it += k + 1
        nosl.append(it)
    else:
        res.append(norm(r) / bn)
    return x, {"residual": np.array(res), "nosl": np.array(nosl[:len(res)]), "converged": converged}

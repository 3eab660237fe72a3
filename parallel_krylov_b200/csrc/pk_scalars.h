// Scalar recurrences of the k-skip solvers: all k+1 coefficient pairs of one outer trip from the Gram sums alone.
// Plain C++ so that the same source runs on the device (one thread, after the Gram reduction) and on the host
// (tests/test_scalar_engine.py compiles it with g++ and checks it bit for bit against the oracle).
// Built with -fmad=false (nvcc) / -ffp-contract=off (g++): every product and sum rounds separately, like numpy.
#pragma once

#ifdef __CUDACC__
#define PK_HD __host__ __device__
#else
#define PK_HD
#endif

#ifndef PK_KMAX
#define PK_KMAX 32
#endif

// Squares.  The reference writes `zeta ** 2` on numpy float64 scalars, which goes through libm pow(); glibc's pow is
// accurate to < 1 ulp but not correctly rounded, so about one square in 10^4 differs from the correctly rounded product
// by one ulp.  The device multiplies (deterministic, correctly rounded).  The CPU tests also build this header with
// -DPK_SQUARE_WITH_LIBM_POW to show that this is the ONLY difference to the oracle (bit-identical solves with it).
#ifdef PK_SQUARE_WITH_LIBM_POW
#include <math.h>
#define PK_SQ(x) pow((x), 2.0)
#else
#define PK_SQ(x) ((x) * (x))
#endif

// Gram layout (see pkrylov.h: pk_gram): G[6*jj + t], t = {U[jj].U[jj], U[jj].U[jj+1], U[jj].V[jj],
// (MrR: V[jj].U[jj+1] | CG: U[jj].V[jj+1]), V[jj].V[jj], V[jj].V[jj+1]}.

// k-skip CG: (alpha_j, beta_j), j = 0..k  — /root/reference/v3/cpu/kskipcg.py:51-52 and :59-68
PK_HD inline void pk_kskipcg_coef(const double* G, int k, double* coef) {
    double a[2 * PK_KMAX + 2], f[2 * PK_KMAX + 4], c[2 * PK_KMAX + 2];
    for (int j = 0; j < 2 * k + 1; ++j) a[j] = G[6 * (j >> 1) + (j & 1)];
    a[2 * k + 1] = 0.0;
    for (int j = 0; j < 2 * k + 4; ++j) f[j] = G[6 * (j >> 1) + 4 + (j & 1)];
    for (int j = 0; j < 2 * k + 2; ++j) c[j] = G[6 * (j >> 1) + 2 + (j & 1)];
    double alpha = a[0] / f[1];
    double beta = (PK_SQ(alpha) * f[2]) / a[0] - 1.0;
    coef[0] = alpha;
    coef[1] = beta;
    for (int j = 0; j < k; ++j) {
        for (int l = 0; l < 2 * (k - j) + 1; ++l) {
            a[l] = a[l] + alpha * (alpha * f[l + 2] - 2.0 * c[l + 1]);
            double d = c[l] - alpha * f[l + 1];
            c[l] = a[l] + d * beta;
            f[l] = c[l] + beta * (d + beta * f[l]);
        }
        alpha = a[0] / f[1];
        beta = (PK_SQ(alpha) * f[2]) / a[0] - 1.0;
        coef[2 * (j + 1)] = alpha;
        coef[2 * (j + 1) + 1] = beta;
    }
}

// k-skip MrR: (zeta_j, eta_j), j = 0..k  — /root/reference/v3/cpu/kskipmrr.py:62-64 and :72-88
PK_HD inline void pk_kskipmrr_coef(const double* G, int k, double* coef) {
    double al[2 * PK_KMAX + 3], be[2 * PK_KMAX + 2], de[2 * PK_KMAX + 1];
    for (int j = 0; j < 2 * k + 3; ++j) al[j] = G[6 * (j >> 1) + (j & 1)];
    be[0] = 0.0;
    for (int j = 1; j < 2 * k + 2; ++j) be[j] = G[6 * (j >> 1) + 2 + (j & 1)];
    for (int j = 0; j < 2 * k + 1; ++j) de[j] = G[6 * (j >> 1) + 4 + (j & 1)];
    double d = al[2] * de[0] - PK_SQ(be[1]);
    double zeta = (al[1] * de[0]) / d;
    double eta = ((-al[1]) * be[1]) / d;
    coef[0] = zeta;
    coef[1] = eta;
    for (int j = 0; j < k; ++j) {
        de[0] = PK_SQ(zeta) * al[2] + (eta * zeta) * be[1];
        al[0] = al[0] - zeta * al[1];
        de[1] = (PK_SQ(eta) * de[1] + ((2.0 * eta) * zeta) * be[2]) + PK_SQ(zeta) * al[3];
        be[1] = (eta * be[1] + zeta * al[2]) - de[1];
        al[1] = -be[1];
        for (int l = 2; l < 2 * (k - j) + 1; ++l) {
            de[l] = (PK_SQ(eta) * de[l] + ((2.0 * eta) * zeta) * be[l + 1]) + PK_SQ(zeta) * al[l + 2];
            double tau = eta * be[l] + zeta * al[l + 1];
            be[l] = tau - de[l];
            al[l] = al[l] - (tau + be[l]);
        }
        d = al[2] * de[0] - PK_SQ(be[1]);
        zeta = (al[1] * de[0]) / d;
        eta = ((-al[1]) * be[1]) / d;
        coef[2 * (j + 1)] = zeta;
        coef[2 * (j + 1) + 1] = eta;
    }
}


// ---- k-skip MrR on a CHEBYSHEV basis (opt-in, SURVEY.md §8f rank 3: numerically safer k-skip) -------------------------
// The reference builds its trip from the moments (r, A^l r), (y, A^l r), (y, A^l y) of the unscaled monomial basis
// (/root/reference/v3/cpu/kskipmrr.py:45-59), whose entries span rho(A)^(2k) and whose vectors become parallel: at k = 8
// its history already differs from plain MrR in the third digit and k >= 12 is chaotic.  Here the basis vectors are
// U_j = T_j(Ah) r, V_j = T_j(Ah) y with Ah = (A - d I) / c mapping [lam_lo, lam_hi] (Gershgorin bounds of A) to [-1, 1];
// the SAME six near-diagonal Gram sums per level give the Chebyshev moments through T_a T_b = (T_{a+b} + T_{|a-b|}) / 2
// (A symmetric):   m_{2j} = 2 (T_j u, T_j w) - m_0,   m_{2j+1} = 2 (T_j u, T_{j+1} w) - m_1,
// and the reference's step recurrences (kskipmrr.py:72-84) carry over with "multiply by A" — an index shift on monomial
// moments — replaced by   (u, T_l A w) = c (m_{l+1} + m_{|l-1|}) / 2 + d m_l.
// Same iterates as MrR in exact arithmetic; in fp64 the k = 8 and k = 12 histories of the Poisson systems agree with plain
// MrR to ~1e-15 (tests/test_chebyshev_basis.py).  G has the layout of pk_kskipmrr_coef (U = T_j r rows, V = T_j y rows).
PK_HD inline void pk_cheb_mulA(const double* m, int L, double c, double d, double* out) {   // out[0..L) from m[0..L]
    out[0] = c * m[1] + d * m[0];
    for (int l = 1; l < L; ++l) out[l] = (c * (m[l + 1] + m[l - 1])) / 2.0 + d * m[l];
}

PK_HD inline void pk_kskipmrr_coef_cheb(const double* G, int k, double c, double d, double* coef) {
    double al[2 * PK_KMAX + 3], be[2 * PK_KMAX + 2], de[2 * PK_KMAX + 1];
    double Aal[2 * PK_KMAX + 3], AAal[2 * PK_KMAX + 3], Abe[2 * PK_KMAX + 2];
    // Chebyshev moments from the Gram sums: al_l = (r, T_l r), be_l = (y, T_l r), de_l = (y, T_l y)
    al[0] = G[0];                          // U0.U0
    al[1] = G[1];                          // U0.U1
    for (int j = 2; j < 2 * k + 3; ++j) al[j] = 2.0 * G[6 * (j >> 1) + (j & 1)] - al[j & 1];
    be[0] = G[2];                          // U0.V0
    be[1] = G[3];                          // V0.U1
    for (int j = 2; j < 2 * k + 2; ++j) be[j] = 2.0 * G[6 * (j >> 1) + 2 + (j & 1)] - be[j & 1];
    de[0] = G[4];                          // V0.V0
    if (k >= 1) de[1] = G[5];              // V0.V1
    for (int j = 2; j < 2 * k + 1; ++j) de[j] = 2.0 * G[6 * (j >> 1) + 4 + (j & 1)] - de[j & 1];
    int Ld = 2 * k + 1;                    // valid lengths: al Ld + 2, be Ld + 1, de Ld
    for (int j = 0; j <= k; ++j) {
        pk_cheb_mulA(al, Ld + 1, c, d, Aal);        // (r, T_l A r),    l < Ld + 1
        pk_cheb_mulA(Aal, Ld, c, d, AAal);          // (r, T_l A^2 r),  l < Ld
        pk_cheb_mulA(be, Ld, c, d, Abe);            // (y, T_l A r),    l < Ld
        const double a1 = Aal[0], a2 = AAal[0], b1 = Abe[0], d0 = de[0];
        const double dd = a2 * d0 - b1 * b1;        // kskipmrr.py:62-64 / :86-88
        const double zeta = (a1 * d0) / dd;
        const double eta = ((-a1) * b1) / dd;
        coef[2 * j] = zeta;
        coef[2 * j + 1] = eta;
        if (j == k) break;
        for (int l = 0; l < Ld; ++l) {              // kskipmrr.py:72-84 on Chebyshev moments
            const double den = (PK_SQ(eta) * de[l] + ((2.0 * eta) * zeta) * Abe[l]) + PK_SQ(zeta) * AAal[l];
            const double tau = eta * be[l] + zeta * Aal[l];
            const double ben = tau - den;
            al[l] = al[l] - (tau + ben);
            be[l] = ben;
            de[l] = den;
        }
        Ld -= 2;                                    // al keeps Ld + 2, be Ld + 1, de Ld valid entries
    }
}

// k-skip CG on the Chebyshev basis: U_j = T_j(Ah) r (j <= k), V_j = T_j(Ah) p (j <= k+1); moments a_l = (r, T_l r),
// f_l = (p, T_l p), c_l = (r, T_l p) from the same six Gram sums per level (CG layout of pk_gram); the reference's step
// recurrences (/root/reference/v3/cpu/kskipcg.py:51-52, :59-68) with f[l+1], f[l+2], c[l+1] replaced by (A f)_l, (A^2 f)_l,
// (A c)_l.  Same iterates as CG in exact arithmetic; follows plain CG to ~1e-12 at k = 8, 12 in fp64.
PK_HD inline void pk_kskipcg_coef_cheb(const double* G, int k, double c, double d, double* coef) {
    double a[2 * PK_KMAX + 1], f[2 * PK_KMAX + 3], cc[2 * PK_KMAX + 2];
    double Af[2 * PK_KMAX + 3], AAf[2 * PK_KMAX + 3], Ac[2 * PK_KMAX + 2];
    a[0] = G[0];
    if (k >= 1) a[1] = G[1];
    for (int j = 2; j < 2 * k + 1; ++j) a[j] = 2.0 * G[6 * (j >> 1) + (j & 1)] - a[j & 1];
    cc[0] = G[2];                          // U0.V0
    cc[1] = G[3];                          // U0.V1
    for (int j = 2; j < 2 * k + 2; ++j) cc[j] = 2.0 * G[6 * (j >> 1) + 2 + (j & 1)] - cc[j & 1];
    f[0] = G[4];                           // V0.V0
    f[1] = G[5];                           // V0.V1
    for (int j = 2; j < 2 * k + 3; ++j) f[j] = 2.0 * G[6 * (j >> 1) + 4 + (j & 1)] - f[j & 1];
    for (int j = 0; j <= k; ++j) {
        const int L = 2 * (k - j) + 1;     // entries the recurrence updates at this step; f has L + 2, cc L + 1, a L valid
        pk_cheb_mulA(f, L + 1, c, d, Af);
        pk_cheb_mulA(Af, L, c, d, AAf);
        pk_cheb_mulA(cc, L, c, d, Ac);
        const double alpha = a[0] / Af[0];                               // kskipcg.py:51 / :67
        const double beta = (PK_SQ(alpha) * AAf[0]) / a[0] - 1.0;        // :52 / :68
        coef[2 * j] = alpha;
        coef[2 * j + 1] = beta;
        if (j == k) break;
        for (int l = 0; l < L; ++l) {                                     // :60-64
            a[l] = a[l] + alpha * (alpha * AAf[l] - 2.0 * Ac[l]);
            const double dd = cc[l] - alpha * Af[l];
            cc[l] = a[l] + dd * beta;
            f[l] = cc[l] + beta * (dd + beta * f[l]);
        }
    }
}

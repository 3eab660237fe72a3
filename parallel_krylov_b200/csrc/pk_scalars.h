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

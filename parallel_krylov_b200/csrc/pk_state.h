// Device-resident solver state and the scalar engine (every scalar recurrence, the residual history and the stopping
// rule of the five reference loops).  Plain C++: compiled by nvcc for the device (run by ONE thread on the reduced sums)
// and by g++ for the CPU tests (tests/test_scalar_engine.py drives whole solves through it against the oracle).
// Rounding discipline: -fmad=false (nvcc) / -ffp-contract=off (g++): a*b+c rounds twice exactly like numpy's
// temporaries (`x += alpha * p` is a multiply then an add, /root/reference/v3/cpu/cg.py:30).
#pragma once
#include <math.h>

#include "pk_scalars.h"

#ifndef PK_KMAX
#define PK_KMAX 32
#endif

// device-resident solver state: every scalar of the reference loops lives here, so no iteration needs the host.
constexpr int PK_MAX_SUMS = 64;                      // sums one reducing kernel may produce
constexpr int PK_GRAM_MAX = 6 * (PK_KMAX + 2) + 8;   // Gram buffer entries

struct PkState {
    // control
    int done;             // 1: every later kernel of the stream is a no-op (stopping rule already fired)
    int converged;        // isConverged
    int guard;            // < 0: an in-kernel wait for a peer timed out (-1 all-reduce, -2 halo)
    int rollback;         // adaptive: the residual grew — this trip starts with the rollback branch (set by EPI_ADAPT_GUARD)
    long long it;         // `i` of the reference loops (number of solution updates)
    long long idx;        // `index` (history position); == it for cg/mrr
    long long maxiter;
    double tol;
    double bnorm;         // ||b||
    // cg / mrr scalars
    double gamma, alpha, beta, zeta, eta, mu, nu;
    double rr;            // ||r||^2 of the newest residual
    double best_res;      // adaptive: `pre_residual`, the smallest residual seen at a trip start
    double cheb_c, cheb_d; // Chebyshev basis (opt-in): A = c Ah + d I maps [lam_lo, lam_hi] onto [-1, 1]
    // k-skip: coefficient pairs of the k+1 steps of one trip: (alpha_j, beta_j) or (zeta_j, eta_j)
    double coef[2 * (PK_KMAX + 1)];
    double gram[PK_GRAM_MAX];
    // reduced sums of the last reducing kernel (multi-GPU: all-reduced in place, then the scalar kernel runs)
    double red[PK_MAX_SUMS];
    // history (device arrays owned by the caller)
    double* res;
    long long* nosl;
    long long* khist;
    int k;                // current k (adaptive)
    int pad1;
    long long hist_len;   // capacity of res/nosl/khist: entries beyond it are dropped, never written
};

// scalar epilogues ("scalar engine"): run by ONE thread on the fully reduced sums.
enum PkEpi : int {
    EPI_NONE = 0,        // sums -> st->red only
    EPI_BNORM,           // red[0] = b.b            -> bnorm
    EPI_CG_INIT,         // red[0] = r.r            -> gamma, res[0], stop test
    EPI_CG_ALPHA,        // red[0] = p.Ap           -> alpha = gamma / sigma
    EPI_CG_BETA,         // red[0] = r.r            -> beta, gamma, it++, res[it], stop test
    EPI_RES0,            // red[0] = r.r            -> res[0] (mrr family: no stop test before the first step)
    EPI_MRR_FIRST,       // red[0] = r.Ar, red[1] = Ar.Ar -> zeta
    EPI_MRR_STEP,        // red[0] = r.r            -> it++, res[it], stop test   (cg/mrr style: idx == it)
    EPI_MRR_GAMMA,       // red[0] = y.Ar, red[2] = y.y -> gamma(nu/mu)
    EPI_MRR_ZETA,        // red[0] = r.s, red[1] = s.s -> zeta, eta
    EPI_KS_FIRST,        // like EPI_MRR_STEP for the opening step of the k-skip MrR family (it=idx=1)
    EPI_KS_TRIP_END,     // red[0] = r.r            -> it += k+1, idx++, res[idx], stop test
    EPI_KS_STEP,         // (no scalar work; intermediate steps of a trip)
    EPI_GRAM_CG,         // gram[] complete         -> coef[] = (alpha_j, beta_j)
    EPI_GRAM_MRR,        // gram[] complete         -> coef[] = (zeta_j, eta_j)
    EPI_GRAM_PART,       // a Gram window that is not the last: copy sums into gram[] only
    EPI_ADAPT_STEP,      // adaptive rollback step: it++, idx++, res[idx], k = max(k-1, 1), khist[idx], convergence test
    EPI_ADAPT_FIRST,     // adaptive opening step: it = idx = 1 recorded (the loop-top logic is EPI_ADAPT_GUARD)
    EPI_ADAPT_TRIP_END,  // red[0] = r.r            -> it += k+1, idx++, res[idx], khist[idx]  (tested by the next guard)
    EPI_ADAPT_GUARD,     // (no sums) loop condition, residual-growth guard, convergence test at the top of a trip
    EPI_CGCG_INIT,       // red[0] = u.w, red[3] = r.u          -> gamma, alpha = gamma/delta, beta = 0
    EPI_CGCG,            // red[0] = u.w, red[3] = r.u, red[4] = r.r -> it++, res[it], stop test, beta, alpha (one reduction/iteration)
    EPI_GRAM_MRR_CHEB,   // gram[] of the Chebyshev basis complete -> coef[] = (zeta_j, eta_j)
    EPI_GRAM_CG_CHEB,    // gram[] of the Chebyshev basis complete -> coef[] = (alpha_j, beta_j)
};


PK_HD inline void pk_record(PkState* st, long long idx, long long it, double res, bool with_k) {
    if (idx < st->hist_len) {
        st->res[idx] = res;
        st->nosl[idx] = it;
        if (with_k && st->khist) st->khist[idx] = st->k;
    }
}

PK_HD inline void pk_stop_test(PkState* st, double res) {
    // `while i < maxiter:` is evaluated before `if residual[i] < tol` (v3/cpu/cg.py:19-24): reaching the cap ends
    // the loop as "not converged" even when the last residual is below tol.
    if (st->it < st->maxiter) {
        if (res < st->tol) {
            st->converged = 1;
            st->done = 1;
        }
    } else {
        st->converged = 0;
        st->done = 1;
    }
}

// The O(k^2) k-skip recurrences live in pk_scalars.h (plain C++, also compiled for the host by the CPU tests).
PK_HD inline void pk_kskipcg_scalars(PkState* st) { pk_kskipcg_coef(st->gram, st->k, st->coef); }
PK_HD inline void pk_kskipmrr_scalars(PkState* st) { pk_kskipmrr_coef(st->gram, st->k, st->coef); }

// The scalar engine.  `st->red` (or st->gram) already holds the fully reduced sums.
template <bool GRAM>
PK_HD inline void pk_epilogue(int epi, PkState* st) {
    const double* s = st->red;
    switch (epi) {
        case EPI_BNORM:
            st->bnorm = sqrt(s[0]);
            break;
        case EPI_CG_INIT: {               // v3/cpu/cg.py:14, :21-24 (first pass)
            st->gamma = s[0];
            st->rr = s[0];
            double res = sqrt(s[0]) / st->bnorm;
            st->it = 0;
            st->idx = 0;
            pk_record(st, 0, 0, res, false);
            pk_stop_test(st, res);
            break;
        }
        case EPI_CG_ALPHA:                // v3/cpu/cg.py:28-29
            st->alpha = st->gamma / s[0];
            break;
        case EPI_CG_BETA: {               // v3/cpu/cg.py:32-37, then :21-24 of the next pass
            double g = s[0];
            st->beta = g / st->gamma;
            st->gamma = g;
            st->rr = g;
            st->it += 1;
            st->idx = st->it;
            double res = sqrt(g) / st->bnorm;
            pk_record(st, st->it, st->it, res, false);
            pk_stop_test(st, res);
            break;
        }
        case EPI_RES0: {                  // v3/cpu/mrr.py:13 — recorded, not tested
            st->rr = s[0];
            st->it = 0;
            st->idx = 0;
            st->best_res = sqrt(s[0]) / st->bnorm;          // adaptivekskipmrr.py:24  pre_residual = residual[0]
            st->rollback = 0;
            pk_record(st, 0, 0, st->best_res, true);
            break;
        }
        case EPI_MRR_FIRST:               // v3/cpu/mrr.py:19   zeta = (r.Ar)/(Ar.Ar)
            st->zeta = s[0] / s[1];
            break;
        case EPI_KS_FIRST:                // opening step done: i = 1, index = 1 (v3/cpu/mrr.py:24-25, kskipmrr.py:32-34)
        case EPI_MRR_STEP: {              // v3/cpu/mrr.py:49-50 then :30-33
            st->rr = s[0];
            st->it += 1;
            st->idx = st->it;
            double res = sqrt(s[0]) / st->bnorm;
            pk_record(st, st->it, st->it, res, true);
            pk_stop_test(st, res);
            break;
        }
        case EPI_MRR_GAMMA:               // v3/cpu/mrr.py:37-39  mu = y.y, nu = y.Ar
            st->nu = s[0];
            st->mu = s[2];
            st->gamma = s[0] / s[2];
            break;
        case EPI_MRR_ZETA:                // v3/cpu/mrr.py:41-44
            st->zeta = s[0] / s[1];
            st->eta = (-st->zeta) * st->gamma;
            break;
        case EPI_KS_TRIP_END: {           // v3/cpu/kskipcg.py:74-76 / kskipmrr.py:95-97, then the loop-top test
            st->rr = s[0];
            st->it += st->k + 1;
            st->idx += 1;
            double res = sqrt(s[0]) / st->bnorm;
            pk_record(st, st->idx, st->it, res, true);
            pk_stop_test(st, res);
            break;
        }
        case EPI_ADAPT_STEP: {            // v3/cpu/adaptivekskipmrr.py:58-66 (end of the rollback branch), then :72-74
            st->rr = s[0];
            st->it += 1;
            st->idx += 1;
            double res = sqrt(s[0]) / st->bnorm;
            pk_record(st, st->idx, st->it, res, false);
            if (st->k > 1) st->k -= 1;                      // :64-65
            if (st->khist && st->idx < st->hist_len) st->khist[st->idx] = st->k;   // :66
            if (res < st->tol) {                            // :72 (no `i < maxiter` test here: that is the loop top)
                st->converged = 1;
                st->done = 1;
            }
            break;
        }
        case EPI_ADAPT_FIRST: {           // v3/cpu/adaptivekskipmrr.py:36-40  i = index = 1
            st->rr = s[0];
            st->it += 1;
            st->idx = st->it;
            pk_record(st, st->idx, st->it, sqrt(s[0]) / st->bnorm, true);
            break;
        }
        case EPI_ADAPT_TRIP_END: {        // v3/cpu/adaptivekskipmrr.py:125-128; residual[index] of :44 recorded here
            st->rr = s[0];
            st->it += st->k + 1;
            st->idx += 1;
            pk_record(st, st->idx, st->it, sqrt(s[0]) / st->bnorm, true);
            break;
        }
        case EPI_ADAPT_GUARD: {           // v3/cpu/adaptivekskipmrr.py:43-47 and :67-74 — the top of a trip, on the device
            st->rollback = 0;
            if (!(st->it < st->maxiter)) {                  // `while i < maxiter` failed: not converged (:129-131)
                st->converged = 0;
                st->done = 1;
                break;
            }
            double res = sqrt(st->rr) / st->bnorm;
            if (res > st->best_res) {                       // :46 the residual grew: this trip opens with the rollback
                st->rollback = 1;
            } else {
                st->best_res = res;                         // :68 (pre_x is saved by k_adapt_save)
                if (res < st->tol) {                        // :72
                    st->converged = 1;
                    st->done = 1;
                }
            }
            break;
        }
        case EPI_CGCG_INIT: {             // v1/threads/pipeline/chronopoulos_gear.py:28-31
            st->gamma = s[3];
            st->alpha = s[3] / s[0];
            st->beta = 0.0;
            break;
        }
        case EPI_CGCG: {                  // chronopoulos_gear.py:41-50 (old_gamma kept up to date), v3 stopping conventions
            st->rr = s[4];
            st->it += 1;
            st->idx = st->it;
            double res = sqrt(s[4]) / st->bnorm;
            pk_record(st, st->it, st->it, res, false);
            if (res < st->tol) {          // :42 — tested right after the update, before the loop condition
                st->converged = 1;
                st->done = 1;
            } else if (!(st->it < st->maxiter)) {
                st->converged = 0;
                st->done = 1;
            }
            double g = s[3];
            st->beta = g / st->gamma;                                   // :49
            st->alpha = g / (s[0] - (st->beta * g) / st->alpha);        // :50
            st->gamma = g;
            break;
        }
        case EPI_GRAM_CG:
            if (GRAM) pk_kskipcg_scalars(st);   // only the Gram kernel / scalar kernel carry the recurrence stack
            break;
        case EPI_GRAM_MRR:
            if (GRAM) pk_kskipmrr_scalars(st);
            break;
        case EPI_GRAM_MRR_CHEB:
            if (GRAM) pk_kskipmrr_coef_cheb(st->gram, st->k, st->cheb_c, st->cheb_d, st->coef);
            break;
        case EPI_GRAM_CG_CHEB:
            if (GRAM) pk_kskipcg_coef_cheb(st->gram, st->k, st->cheb_c, st->cheb_d, st->coef);
            break;
        default:
            break;
    }
}


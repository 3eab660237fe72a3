// Host build of the device's scalar engine (csrc/pk_scalars.h, csrc/pk_state.h) for tests/test_scalar_engine.py
// (g++ -O2 -ffp-contract=off -shared).  Nothing here is CUDA: the same source runs in the last block of the reducing
// kernels on the GPU.
#include <string.h>

#include "../parallel_krylov_b200/csrc/pk_state.h"

extern "C" {
void host_kskipcg_coef(const double* G, int k, double* coef) { pk_kskipcg_coef(G, k, coef); }
void host_kskipmrr_coef(const double* G, int k, double* coef) { pk_kskipmrr_coef(G, k, coef); }
void host_kskipcg_coef_cheb(const double* G, int k, double c, double d, double* coef) { pk_kskipcg_coef_cheb(G, k, c, d, coef); }
void host_kskipmrr_coef_cheb(const double* G, int k, double c, double d, double* coef) { pk_kskipmrr_coef_cheb(G, k, c, d, coef); }

void* hs_new(long long maxiter, double tol, int k, long long hist_len, double* res, long long* nosl, long long* khist) {
    PkState* st = new PkState();
    memset(st, 0, sizeof(PkState));
    st->maxiter = maxiter;
    st->tol = tol;
    st->k = k;
    st->hist_len = hist_len;
    st->res = res;
    st->nosl = nosl;
    st->khist = khist;
    return st;
}
void hs_free(void* p) { delete (PkState*)p; }
void hs_epilogue(void* p, int epi, const double* sums, int n) {
    PkState* st = (PkState*)p;
    for (int i = 0; i < n; ++i) st->red[i] = sums[i];
    pk_epilogue<true>(epi, st);
}
void hs_gram(void* p, int epi, const double* G, int n) {
    PkState* st = (PkState*)p;
    for (int i = 0; i < n; ++i) st->gram[i] = G[i];
    pk_epilogue<true>(epi, st);
}
void hs_set_k(void* p, int k) { ((PkState*)p)->k = k; }
double hs_get(void* p, int what, int j) {
    PkState* st = (PkState*)p;
    switch (what) {
        case 0: return st->alpha;
        case 1: return st->beta;
        case 2: return st->gamma;
        case 3: return st->zeta;
        case 4: return st->eta;
        case 5: return (double)st->done;
        case 6: return (double)st->converged;
        case 7: return (double)st->it;
        case 8: return (double)st->idx;
        case 9: return st->coef[j];
        case 10: return (double)st->rollback;
        case 11: return (double)st->k;
        default: return 0.0;
    }
}
}

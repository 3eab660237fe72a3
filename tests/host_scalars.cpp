// Host build of csrc/pk_scalars.h for tests/test_scalar_engine.py (g++ -O2 -ffp-contract=off -shared).
#include "../parallel_krylov_b200/csrc/pk_scalars.h"

extern "C" void host_kskipcg_coef(const double* G, int k, double* coef) { pk_kskipcg_coef(G, k, coef); }
extern "C" void host_kskipmrr_coef(const double* G, int k, double* coef) { pk_kskipmrr_coef(G, k, coef); }

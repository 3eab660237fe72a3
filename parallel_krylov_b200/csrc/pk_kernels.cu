// Fused vector kernels of the solver loops (K3 of SURVEY.md §2b), the single-pass Gram kernel (K4) and the
// scalar engine launch (K5).  All are HBM-bound streaming kernels: one pass, no temporaries, the dot products
// the next step needs are reduced in the same pass (warp shuffle -> block -> last block).
#include "pk_device.cuh"
#include "pk_launch.h"

#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>

namespace {

constexpr int EW_BLOCK = 256;

__device__ __forceinline__ bool pk_done(const PkState* st) { return *((volatile const int*)&st->done) != 0; }

// ---- dot ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_BLOCK) k_dot(long long n, const double* __restrict__ u,
                                                  const double* __restrict__ v, PkRedArgs ra, int ignore_done) {
    if (!ignore_done && pk_skip(ra)) return;
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) acc[0] += u[i] * v[i];
    pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- r = b - v (or r = b), optional p = r, sums[0] = r.r  — v3/cpu/cg.py:12-14, mrr.py:12-13 --------------------
__global__ void __launch_bounds__(EW_BLOCK) k_resid_init(long long n, const double* __restrict__ b,
                                                         const double* __restrict__ v, double* __restrict__ r,
                                                         double* __restrict__ p, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double ri = v ? (b[i] - v[i]) : b[i];
        r[i] = ri;
        if (p) p[i] = ri;
        acc[0] += ri * ri;
    }
    pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- CG: x += alpha p ; r -= alpha v ; sums[0] = r.r  — v3/cpu/cg.py:30-33 --------------------------------------
__global__ void __launch_bounds__(EW_BLOCK) k_cg_xr(long long n, double* __restrict__ x, double* __restrict__ r,
                                                    const double* __restrict__ p, const double* __restrict__ v,
                                                    PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double alpha = ra.st->alpha;
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        x[i] = x[i] + alpha * p[i];
        double ri = r[i] - alpha * v[i];
        r[i] = ri;
        acc[0] += ri * ri;
    }
    pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- CG with the x-update hiding the r.r all-reduce (multi-GPU): r -= alpha v ; r.r is POSTED to the peers by the last
//      block; x += alpha p runs while the sums are in flight; a one-block kernel collects them and runs the epilogue.
//      Same arithmetic per element as k_cg_xr (cg.py:30-33), two launches more, one all-reduce latency less.
__global__ void __launch_bounds__(EW_BLOCK) k_cg_r(long long n, double* __restrict__ r, const double* __restrict__ v,
                                                   PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double alpha = ra.st->alpha;
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double ri = r[i] - alpha * v[i];
        r[i] = ri;
        acc[0] += ri * ri;
    }
    pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

__global__ void __launch_bounds__(EW_BLOCK) k_cg_x(long long n, double* __restrict__ x, const double* __restrict__ p,
                                                   const PkState* st) {
    if (pk_done(st)) return;
    const double alpha = st->alpha;
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) x[i] = x[i] + alpha * p[i];
}

__global__ void __launch_bounds__(64) k_ar_wait(const PkP2P* pp, PkState* st, int n, int epi) {
    if (pk_done(st)) return;
    pk_mailbox_wait<64>(pp, st->red, n, st);
    if (threadIdx.x == 0) pk_epilogue<false>(epi, st);
}

// ---- CG: p = r + beta p  — v3/cpu/cg.py:35 -----------------------------------------------------------------------
__global__ void __launch_bounds__(EW_BLOCK) k_cg_p(long long n, double* __restrict__ p, const double* __restrict__ r,
                                                   const PkState* st) {
    if (pk_done(st)) return;
    const double beta = st->beta;
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) p[i] = r[i] + beta * p[i];
}

// ---- MrR opening step: y = zeta Ar ; z = -zeta r ; r -= y ; x -= z ; sums[0] = r.r  — v3/cpu/mrr.py:20-23 --------
__global__ void __launch_bounds__(EW_BLOCK) k_mrr_first(long long n, const double* __restrict__ ar,
                                                        double* __restrict__ r, double* __restrict__ x,
                                                        double* __restrict__ y, double* __restrict__ z, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double zeta = ra.st->zeta;
    const double nzeta = -zeta;
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double yi = zeta * ar[i];
        double zi = nzeta * r[i];
        double ri = r[i] - yi;
        y[i] = yi;
        z[i] = zi;
        r[i] = ri;
        x[i] = x[i] - zi;
        acc[0] += ri * ri;
    }
    pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- MrR: s = Ar - gamma y ; sums = {r.s, s.s}  — v3/cpu/mrr.py:40-42 (s is never stored) -----------------------
__global__ void __launch_bounds__(EW_BLOCK) k_mrr_s(long long n, const double* __restrict__ ar,
                                                    const double* __restrict__ y, const double* __restrict__ r,
                                                    PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double gamma = ra.st->gamma;
    double acc[2] = {0.0, 0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double s = ar[i] - gamma * y[i];
        acc[0] += r[i] * s;
        acc[1] += s * s;
    }
    pk_grid_reduce<2, EW_BLOCK>(acc, ra);
}

// ---- MrR step: y = eta y + zeta Ar ; z = eta z - zeta r ; r -= y ; x -= z ; sums[0] = r.r -----------------------
//      v3/cpu/mrr.py:45-48 ; kskipmrr.py:65-69 / :89-93 with (zeta, eta) = coef[2j], coef[2j+1]
__global__ void __launch_bounds__(EW_BLOCK) k_mrr_update(long long n, const double* __restrict__ ar,
                                                         double* __restrict__ y, double* __restrict__ z,
                                                         const double* r, double* r_out, double* r_alt,
                                                         double* __restrict__ x, int cj, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    bool reduce = ra.epi != EPI_KS_STEP;
    if (ra.dyn_cj >= 0) {
        // k lives on the device (adaptive): this is the last step of the trip iff cj == k; the residual ping-pongs
        // between its home (r_out) and the spare buffer (r_alt) so that the LAST step of the trip ends at home
        const int kk = ra.st->k;
        if (ra.dyn_last) reduce = (cj == kk);
        if (r_alt != nullptr && ((kk - cj) & 1)) r_out = r_alt;
    }
    const double zeta = (cj < 0) ? ra.st->zeta : ra.st->coef[2 * cj];
    const double eta = (cj < 0) ? ra.st->eta : ra.st->coef[2 * cj + 1];
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    // Nine streams (5 loads, 4 stores) per element: with one element per thread the kernel reached 76 % of HBM peak;
    // 128-bit accesses (two elements per thread) halve the number of memory instructions in flight per byte.
    const bool vec = ((((uintptr_t)ar | (uintptr_t)y | (uintptr_t)z | (uintptr_t)r | (uintptr_t)r_out | (uintptr_t)x) & 15) == 0);
    long long done_to = 0;
    if (vec) {
        const long long n2 = n >> 1;
        const double2* ar2 = reinterpret_cast<const double2*>(ar);
        const double2* r2 = reinterpret_cast<const double2*>(r);
        double2* y2 = reinterpret_cast<double2*>(y);
        double2* z2 = reinterpret_cast<double2*>(z);
        double2* ro2 = reinterpret_cast<double2*>(r_out);
        double2* x2 = reinterpret_cast<double2*>(x);
        for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n2; i += stride) {
            const double2 rv = r2[i], yv = y2[i], av = ar2[i], zv = z2[i], xv = x2[i];
            double2 yo, zo, ro, xo;
            yo.x = eta * yv.x + zeta * av.x;
            yo.y = eta * yv.y + zeta * av.y;
            zo.x = eta * zv.x - zeta * rv.x;
            zo.y = eta * zv.y - zeta * rv.y;
            ro.x = rv.x - yo.x;
            ro.y = rv.y - yo.y;
            xo.x = xv.x - zo.x;
            xo.y = xv.y - zo.y;
            y2[i] = yo;
            z2[i] = zo;
            ro2[i] = ro;
            x2[i] = xo;
            acc[0] += ro.x * ro.x;
            acc[0] += ro.y * ro.y;
        }
        done_to = n2 << 1;
    }
    for (long long i = done_to + (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double ri = r[i];
        double yi = eta * y[i] + zeta * ar[i];
        double zi = eta * z[i] - zeta * ri;
        ri = ri - yi;
        y[i] = yi;
        z[i] = zi;
        r_out[i] = ri;
        x[i] = x[i] - zi;
        acc[0] += ri * ri;
    }
    if (reduce) pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- adaptive: at the top of a trip either remember x (pre_x = x.copy(), adaptivekskipmrr.py:69) or roll x back to it
//      (x = pre_x.copy(), :48), as the guard decided (st->rollback)
__global__ void __launch_bounds__(EW_BLOCK) k_adapt_save(long long n, double* __restrict__ x, double* __restrict__ best_x,
                                                         const PkState* st) {
    if (pk_done(st)) return;
    const bool back = *((volatile const int*)&st->rollback) != 0;
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    if (back) for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) x[i] = best_x[i];
    else for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) best_x[i] = x[i];
}

// ---- Chronopoulos-Gear CG: p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = r / M ;
//      local sums {r.r, r.u} (all-reduced by the NEXT kernel, the SpMV w = A u, together with its u.w: ONE reduction
//      point per iteration) — v1/threads/pipeline/chronopoulos_gear.py:37-47.  init: only u = r / M and the sums.
//      Without a preconditioner u IS r (same buffer): the old r is read as u before it is overwritten.
__global__ void __launch_bounds__(EW_BLOCK) k_cgcg_update(long long n, double* __restrict__ x, double* r, double* u,
                                                          const double* __restrict__ w, double* __restrict__ p,
                                                          double* __restrict__ s, const double* __restrict__ mdiag,
                                                          int init, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double alpha = ra.st->alpha, beta = ra.st->beta;
    double acc[2] = {0.0, 0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double ri = r[i];
        if (!init) {
            const double pi = u[i] + beta * p[i];
            const double si = w[i] + beta * s[i];
            p[i] = pi;
            s[i] = si;
            x[i] = x[i] + alpha * pi;
            ri = ri - alpha * si;
            r[i] = ri;
        }
        const double ui = mdiag ? ri / mdiag[i] : ri;
        if (mdiag) u[i] = ui;
        acc[0] += ri * ui;        // -> red[3] = r.u (gamma)
        acc[1] += ri * ri;        // -> red[4] = r.r
    }
    pk_grid_reduce<2, EW_BLOCK>(acc, ra);
}

__global__ void k_csr_diag(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                           const double* __restrict__ val, long long n_rows, long long row0, double* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        double d = 0.0;
        for (int q = rowptr[r]; q < rowptr[r + 1]; ++q)
            if (col[q] == (int)(r + row0)) d = val[q];
        out[r] = d;
    }
}

// Gershgorin discs of the local rows: lo = min_i (a_ii - R_i), hi = max_i (a_ii + R_i), R_i = sum_{j != i} |a_ij|.
// Doubles are reduced with integer atomics on an order-preserving key.
__device__ __forceinline__ unsigned long long pk_ord_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_gershgorin_csr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const double* __restrict__ val, long long n_rows, long long row0,
                                 unsigned long long* __restrict__ keys) {
    double lo = 1e300, hi = -1e300;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        double d = 0.0, off = 0.0;
        for (int q = rowptr[r]; q < rowptr[r + 1]; ++q) {
            const double v = val[q];
            if (col[q] == (int)(r + row0)) d += v; else off += fabs(v);
        }
        lo = fmin(lo, d - off);
        hi = fmax(hi, d + off);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys, pk_ord_key(lo));
        atomicMax(keys + 1, pk_ord_key(hi));
    }
}
__global__ void k_gershgorin_dense(const double* __restrict__ a, long long lda, long long n_rows, long long n_cols,
                                   unsigned long long* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n_rows; r += nwarps) {
        double off = 0.0;
        for (long long c = lane; c < n_cols; c += 32)
            if (c != r) off += fabs(a[r * lda + c]);
        for (int o = 16; o > 0; o >>= 1) off += __shfl_down_sync(0xffffffffu, off, o);
        if (lane == 0) {
            const double d = a[r * lda + r];
            atomicMin(keys, pk_ord_key(d - off));
            atomicMax(keys + 1, pk_ord_key(d + off));
        }
    }
}

// out = s1 * a + s2 * b  (Chebyshev basis: U_1 = (A r - d r) / c from the A r the previous step left behind)
__global__ void __launch_bounds__(EW_BLOCK) k_axpby(long long n, double s1, const double* __restrict__ a, double s2,
                                                    const double* __restrict__ b, double* __restrict__ out,
                                                    const PkState* st) {
    if (pk_done(st)) return;
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) out[i] = s1 * a[i] + s2 * b[i];
}

__global__ void k_dense_diag(const double* __restrict__ a, long long lda, long long n_rows, double* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x)
        out[r] = a[r * lda + r];
}

// ---- k-skip CG step: x += a Ap0 ; Ar0 -= a Ap1 ; Ap0 = Ar0 + b Ap0 ; sums[0] = Ar0.Ar0  — kskipcg.py:53-55 -----
__global__ void __launch_bounds__(EW_BLOCK) k_kscg_update(long long n, double* __restrict__ x,
                                                          double* __restrict__ ar0, const double* ap0, double* ap0_out,
                                                          const double* __restrict__ ap1, int cj, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const double alpha = ra.st->coef[2 * cj];
    const double beta = ra.st->coef[2 * cj + 1];
    double acc[1] = {0.0};
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double p0 = ap0[i];
        x[i] = x[i] + alpha * p0;
        double ri = ar0[i] - alpha * ap1[i];
        ar0[i] = ri;
        ap0_out[i] = ri + beta * p0;
        acc[0] += ri * ri;
    }
    if (ra.epi != EPI_KS_STEP) pk_grid_reduce<1, EW_BLOCK>(acc, ra);
}

// ---- Gram window: all pair products for jj in [j0, j0+W) in one pass over the basis ------------------------------
// U rows [0,nu), V rows [0,nv), row stride ld.  Per jj six sums (see pkrylov.h: pk_gram).  Rows that do not exist
// contribute zeros (their loads are skipped; the predicate is uniform across the grid).
template <int W, int MODE>
__global__ void __launch_bounds__(EW_BLOCK) k_gram(long long n, long long ld, const double* __restrict__ U, int nu,
                                                   const double* __restrict__ V, int nv, int j0, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    if (ra.dyn_cj >= 0) { nu = ra.st->k + 2; nv = ra.st->k + 1; }      // adaptive (MrR layout): rows of the CURRENT k only
    double acc[6 * W];
#pragma unroll
    for (int t = 0; t < 6 * W; ++t) acc[t] = 0.0;
    const long long stride = (long long)gridDim.x * EW_BLOCK;
    for (long long i = (long long)blockIdx.x * EW_BLOCK + threadIdx.x; i < n; i += stride) {
        double u[W + 1], v[W + 1];
#pragma unroll
        for (int t = 0; t <= W; ++t) {
            u[t] = (j0 + t < nu) ? U[(long long)(j0 + t) * ld + i] : 0.0;
            v[t] = (j0 + t < nv) ? V[(long long)(j0 + t) * ld + i] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < W; ++t) {
            acc[6 * t + 0] += u[t] * u[t];
            acc[6 * t + 1] += u[t] * u[t + 1];
            acc[6 * t + 2] += u[t] * v[t];
            acc[6 * t + 3] += (MODE == 0) ? (v[t] * u[t + 1]) : (u[t] * v[t + 1]);
            acc[6 * t + 4] += v[t] * v[t];
            acc[6 * t + 5] += v[t] * v[t + 1];
        }
    }
    pk_grid_reduce<6 * W, EW_BLOCK, true>(acc, ra);
}

// ---- Gram window, TMA-pipelined: the (up to 2W+2) basis rows of a tile of T elements are brought into shared memory by
// 1-D bulk copies (ring of STAGES buffers, mbarrier per stage) while the previous tile is reduced, so HBM stays busy
// although the 6W register accumulators allow only one or two blocks per SM.  Same sums, same per-thread order of
// additions as k_gram (thread t owns elements t, t+T*grid, ...), hence bit-identical results for equal grids.
template <int W, int MODE, int STAGES>
__global__ void __launch_bounds__(EW_BLOCK) k_gram_tma(long long n, long long ld, const double* __restrict__ U, int nu,
                                                       const double* __restrict__ V, int nv, int j0, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    if (ra.dyn_cj >= 0) { nu = ra.st->k + 2; nv = ra.st->k + 1; }      // adaptive (MrR layout): rows of the CURRENT k only
    constexpr int T = EW_BLOCK;                     // elements per tile (one per thread)
    constexpr int ROWS = 2 * (W + 1);
    extern __shared__ __align__(128) unsigned char gsm_raw[];
    double* gsm = reinterpret_cast<double*>(gsm_raw);            // [STAGES][ROWS][T]
    __shared__ __align__(8) unsigned long long full[STAGES];
    const int tid = threadIdx.x;
    double acc[6 * W];
#pragma unroll
    for (int t = 0; t < 6 * W; ++t) acc[t] = 0.0;
    const long long n_full = n / T;                 // tiles that can be bulk-copied; the tail is read directly
    const long long my_tiles = n_full > blockIdx.x ? (n_full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int n_rows_present = 0;
#pragma unroll
    for (int t = 0; t <= W; ++t) n_rows_present += (j0 + t < nu) + (j0 + t < nv);
    // warp 0 issues the copies, one basis row per lane (a single thread issuing 2W+2 copies per tile serialises the
    // block); lane r < ROWS owns row r: even r = U[j0 + r/2], odd r = V[j0 + r/2]
    const double* my_row = nullptr;
    if (tid < ROWS) {
        const int t = tid >> 1;
        if ((tid & 1) == 0) { if (j0 + t < nu) my_row = U + (long long)(j0 + t) * ld; }
        else { if (j0 + t < nv) my_row = V + (long long)(j0 + t) * ld; }
    }
    auto issue = [&](long long i) {     // called by all lanes of warp 0
        const int s = (int)(i % STAGES);
        const long long e0 = (blockIdx.x + i * (long long)gridDim.x) * T;
        if (tid == 0) mbar_expect_tx(&full[s], (unsigned)(n_rows_present * T * sizeof(double)));
        __syncwarp();
        if (my_row) bulk_g2s(gsm + ((size_t)s * ROWS + tid) * T, my_row + e0, T * sizeof(double), &full[s]);
    };
    if (tid < 32)
        for (long long i = 0; i < STAGES - 1 && i < my_tiles; ++i) issue(i);
    for (long long i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % STAGES);
        if (tid < 32 && i + STAGES - 1 < my_tiles) issue(i + STAGES - 1);
        mbar_wait(&full[s], (unsigned)((i / STAGES) & 1));
        const double* src = gsm + (size_t)s * ROWS * T + tid;
        double u[W + 1], v[W + 1];
#pragma unroll
        for (int t = 0; t <= W; ++t) {
            u[t] = (j0 + t < nu) ? src[(size_t)(2 * t) * T] : 0.0;
            v[t] = (j0 + t < nv) ? src[(size_t)(2 * t + 1) * T] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < W; ++t) {
            acc[6 * t + 0] += u[t] * u[t];
            acc[6 * t + 1] += u[t] * u[t + 1];
            acc[6 * t + 2] += u[t] * v[t];
            acc[6 * t + 3] += (MODE == 0) ? (v[t] * u[t + 1]) : (u[t] * v[t + 1]);
            acc[6 * t + 4] += v[t] * v[t];
            acc[6 * t + 5] += v[t] * v[t + 1];
        }
        __syncthreads();      // stage s may be refilled by the next iteration's issue
    }
    // tail elements (n % T) by block 0, read directly
    if (blockIdx.x == 0) {
        const long long i = n_full * T + tid;
        if (i < n) {
            double u[W + 1], v[W + 1];
#pragma unroll
            for (int t = 0; t <= W; ++t) {
                u[t] = (j0 + t < nu) ? U[(long long)(j0 + t) * ld + i] : 0.0;
                v[t] = (j0 + t < nv) ? V[(long long)(j0 + t) * ld + i] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < W; ++t) {
                acc[6 * t + 0] += u[t] * u[t];
                acc[6 * t + 1] += u[t] * u[t + 1];
                acc[6 * t + 2] += u[t] * v[t];
                acc[6 * t + 3] += (MODE == 0) ? (v[t] * u[t + 1]) : (u[t] * v[t + 1]);
                acc[6 * t + 4] += v[t] * v[t];
                acc[6 * t + 5] += v[t] * v[t + 1];
            }
        }
    }
    pk_grid_reduce<6 * W, EW_BLOCK, true>(acc, ra);
}

__global__ void k_scalar(PkState* st, int epi, int ignore_done, int only_rollback, int dyn_cj, int dyn_last) {
    if (!ignore_done && pk_done(st)) return;
    if (only_rollback && st->rollback == 0) return;
    if (dyn_cj >= 0 && (dyn_cj > st->k || (dyn_last && dyn_cj != st->k))) return;
    pk_epilogue<true>(epi, st);
}

__global__ void k_set_k(PkState* st, int k) {
    st->k = k;
    if (st->khist && st->idx < st->hist_len) st->khist[st->idx] = k;   // adaptivekskipmrr.py:66
}

// persistent grid: one full wave of resident blocks (or fewer when the vector is short)
template <class K>
inline int ew_grid(pk_ctx* ctx, K kernel, long long n, int per_thread = 2) {
    long long want = (n + EW_BLOCK * per_thread - 1) / (EW_BLOCK * per_thread);
    long long cap = (long long)ctx->sm_count * pk_blocks_per_sm((const void*)kernel, EW_BLOCK, 0);
    if (cap > ctx->red.max_blocks) cap = ctx->red.max_blocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

inline PkRedArgs red_args(pk_ctx* ctx, int epi, int g_off = -1) {
    PkRedArgs ra;
    ra.partials = ctx->red.partials;
    ra.ticket = ctx->red.ticket;
    ra.max_blocks = ctx->red.max_blocks;
    ra.st = ctx->d_state;
    ra.epi = epi;
    ra.defer = (ctx->n_ranks > 1 && !ctx->d_p2p && !ctx->nocomm) ? 1 : 0;
    ra.g_off = g_off;
    ra.red_off = 0;
    ra.block_off = 0;
    ra.nb_total = 0;
    ra.store_only = 0;
    ra.p2p = ctx->nocomm ? nullptr : ctx->d_p2p;
    ra.ar_n = 0;          // set by the launcher: number of sums this kernel all-reduces
    ra.post_only = 0;
    ra.only_rollback = ctx->ctl_only_rollback;
    ra.dyn_cj = ctx->ctl_dyn_cj;
    ra.dyn_last = ctx->ctl_dyn_last;
    return ra;
}

inline PkRedArgs red_args_n(pk_ctx* ctx, int epi, int nsums) {
    PkRedArgs ra = red_args(ctx, epi);
    ra.ar_n = (epi == EPI_KS_STEP) ? 0 : nsums;
    return ra;
}

#define PK_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            pk_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return PK_ERR_CUDA;                                                                \
        }                                                                                      \
        ctx->launches++;                                                                       \
    } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
int pk_blocks_per_sm(const void* kernel, int block, size_t smem) {
    // caches are per (device, kernel): the opt-in attribute and the occupancy both belong to the current device
    static std::map<std::tuple<int, const void*, size_t>, int> cache;
    static std::map<std::pair<int, const void*>, bool> opted_in;
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 40 * 1024 && !opted_in[{dev, kernel}]) {
        // opt in ONCE to the device maximum: the attribute is a ceiling for later launches, so it must never be lowered
        int max_optin = 0;
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        int stat = 1024;
        if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - stat) != cudaSuccess)
            cudaGetLastError();   // do not leave a stale error for the next launch check
        opted_in[{dev, kernel}] = true;
    }
    auto key = std::make_tuple(dev, kernel, smem);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    cache[key] = per_sm;
    return per_sm;
}

int pk_launch_scalar(pk_ctx* ctx, int epi, int ignore_done) {
    k_scalar<<<1, 1, 0, ctx->stream>>>(ctx->d_state, epi, ignore_done, ctx->ctl_only_rollback, ctx->ctl_dyn_cj,
                                       ctx->ctl_dyn_last);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

int pk_launch_set_k(pk_ctx* ctx, int k) {
    k_set_k<<<1, 1, 0, ctx->stream>>>(ctx->d_state, k);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

// After a reducing kernel: single GPU -> nothing to do (the last block already ran the epilogue);
// multi GPU -> all-reduce the published sums over NVLink, then run the scalar engine.
int pk_finish_reduce(pk_ctx* ctx, int nsums, int epi, int g_off, int ignore_done) {
    if (ctx->n_ranks <= 1 || ctx->d_p2p || ctx->nocomm) return PK_OK;   // single GPU, or all-reduced inside the kernel
    double* buf = (g_off >= 0) ? (ctx->d_state->gram + g_off) : ctx->d_state->red;
    PK_CHECK(pk_comm_allreduce(ctx, buf, nsums, ctx->stream));
    if (epi != EPI_NONE && epi != EPI_GRAM_PART && epi != EPI_KS_STEP) return pk_launch_scalar(ctx, epi, ignore_done);
    return PK_OK;
}

int pk_launch_dot(pk_ctx* ctx, long long n, const double* u, const double* v, int epi, int ignore_done) {
    k_dot<<<ew_grid(ctx, k_dot, n), EW_BLOCK, 0, ctx->stream>>>(n, u, v, red_args_n(ctx, epi, 1), ignore_done);
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, epi, -1, ignore_done);
}

int pk_launch_resid_init(pk_ctx* ctx, long long n, const double* b, const double* v, double* r, double* p, int epi) {
    k_resid_init<<<ew_grid(ctx, k_resid_init, n), EW_BLOCK, 0, ctx->stream>>>(n, b, v, r, p, red_args_n(ctx, epi, 1));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, epi, -1, 0);
}

int pk_launch_cg_xr(pk_ctx* ctx, long long n, double* x, double* r, const double* p, const double* v) {
    k_cg_xr<<<ew_grid(ctx, k_cg_xr, n), EW_BLOCK, 0, ctx->stream>>>(n, x, r, p, v, red_args_n(ctx, EPI_CG_BETA, 1));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, EPI_CG_BETA, -1, 0);
}

// x += alpha p ; r -= alpha v ; r.r all-reduced ; beta, residual, stop test.  PK_CG_SPLIT=1 (multi-GPU, NVLink mailboxes):
// the all-reduce's flight is hidden behind the x-update (k_cg_r posts, k_cg_x runs, k_ar_wait collects).  Measured at 8
// GPUs on 512^3 (r02): 1860 it/s with the split, 1868 without — the wait it hides is not what limits the iteration
// there (the two extra launches cost what the overlap gains), so the single kernel stays the default.
int pk_launch_cg_xr_split(pk_ctx* ctx, long long n, double* x, double* r, const double* p, const double* v) {
    static int split = -1;
    if (split < 0) {
        const char* e = getenv("PK_CG_SPLIT");
        split = e ? atoi(e) : 0;
    }
    if (!split || ctx->n_ranks <= 1 || !ctx->d_p2p) return pk_launch_cg_xr(ctx, n, x, r, p, v);
    PkRedArgs ra = red_args_n(ctx, EPI_CG_BETA, 1);
    ra.post_only = 1;
    if (ra.p2p == nullptr) ra.post_only = 0;             // nocomm (compute-only timing): the epilogue runs in k_cg_r
    k_cg_r<<<ew_grid(ctx, k_cg_r, n), EW_BLOCK, 0, ctx->stream>>>(n, r, v, ra);
    PK_LAUNCH_CHECK();
    k_cg_x<<<ew_grid(ctx, k_cg_x, n), EW_BLOCK, 0, ctx->stream>>>(n, x, p, ctx->d_state);
    PK_LAUNCH_CHECK();
    if (ra.post_only) {
        k_ar_wait<<<1, 64, 0, ctx->stream>>>(ctx->d_p2p, ctx->d_state, 1, EPI_CG_BETA);
        PK_LAUNCH_CHECK();
    }
    return PK_OK;
}

int pk_launch_cg_p(pk_ctx* ctx, long long n, double* p, const double* r) {
    k_cg_p<<<ew_grid(ctx, k_cg_p, n), EW_BLOCK, 0, ctx->stream>>>(n, p, r, ctx->d_state);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

int pk_launch_mrr_first(pk_ctx* ctx, long long n, const double* ar, double* r, double* x, double* y, double* z,
                        int epi) {
    k_mrr_first<<<ew_grid(ctx, k_mrr_first, n), EW_BLOCK, 0, ctx->stream>>>(n, ar, r, x, y, z, red_args_n(ctx, epi, 1));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, epi, -1, 0);
}

int pk_launch_mrr_s(pk_ctx* ctx, long long n, const double* ar, const double* y, const double* r) {
    k_mrr_s<<<ew_grid(ctx, k_mrr_s, n), EW_BLOCK, 0, ctx->stream>>>(n, ar, y, r, red_args_n(ctx, EPI_MRR_ZETA, 2));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 2, EPI_MRR_ZETA, -1, 0);
}

int pk_launch_cgcg_update(pk_ctx* ctx, long long n, double* x, double* r, double* u, const double* w, double* p,
                          double* s, const double* mdiag, int init) {
    PkRedArgs ra = red_args(ctx, EPI_NONE);
    ra.red_off = 3;          // local r.r / r.u wait in red[3..4] for the SpMV's all-reduce
    ra.ar_n = 0;
    ra.defer = 1;            // publish only: no epilogue here
    ra.p2p = nullptr;
    k_cgcg_update<<<ew_grid(ctx, k_cgcg_update, n), EW_BLOCK, 0, ctx->stream>>>(n, x, r, u, w, p, s, mdiag, init, ra);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

extern "C" int pk_mat_diagonal(pk_mat* m, double* d_out) {
    PK_REQUIRE(m && d_out, "null argument");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    int grid = (int)((m->n_rows + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    if (grid < 1) grid = 1;
    if (m->kind == MAT_DENSE) k_dense_diag<<<grid, 256, 0, ctx->stream>>>(m->dense, m->lda, m->n_rows, d_out);
    else if (m->segs.empty()) k_csr_diag<<<grid, 256, 0, ctx->stream>>>(m->rowptr, m->col, m->val, m->n_rows, 0, d_out);
    else
        for (const PkSeg& sg : m->segs)
            k_csr_diag<<<grid, 256, 0, ctx->stream>>>(sg.rp32, m->col + sg.base, m->val + sg.base, sg.row_hi - sg.row_lo,
                                                      sg.row_lo, d_out + sg.row_lo);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

int pk_launch_axpby(pk_ctx* ctx, long long n, double s1, const double* a, double s2, const double* b, double* out) {
    k_axpby<<<ew_grid(ctx, k_axpby, n), EW_BLOCK, 0, ctx->stream>>>(n, s1, a, s2, b, out, ctx->d_state);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

extern "C" int pk_mat_gershgorin(pk_mat* m, double* h_out) {
    PK_REQUIRE(m && h_out, "null argument");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    unsigned long long h[2] = {~0ull, 0ull};
    PK_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    PK_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    int grid = (int)((m->n_rows + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    if (grid < 1) grid = 1;
    if (m->kind == MAT_DENSE) k_gershgorin_dense<<<grid, 256, 0, ctx->stream>>>(m->dense, m->lda, m->n_rows, m->n_cols, d);
    else if (m->segs.empty()) k_gershgorin_csr<<<grid, 256, 0, ctx->stream>>>(m->rowptr, m->col, m->val, m->n_rows, 0, d);
    else
        for (const PkSeg& sg : m->segs)
            k_gershgorin_csr<<<grid, 256, 0, ctx->stream>>>(sg.rp32, m->col + sg.base, m->val + sg.base,
                                                            sg.row_hi - sg.row_lo, sg.row_lo, d);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    for (int i = 0; i < 2; ++i) {
        const unsigned long long k = h[i];
        const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
        long long bits = (long long)b;
        memcpy(&h_out[i], &bits, sizeof(double));
    }
    return PK_OK;
}

int pk_launch_adapt_save(pk_ctx* ctx, long long n, double* x, double* best_x) {
    k_adapt_save<<<ew_grid(ctx, k_adapt_save, n), EW_BLOCK, 0, ctx->stream>>>(n, x, best_x, ctx->d_state);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

int pk_launch_mrr_update(pk_ctx* ctx, long long n, const double* ar, double* y, double* z, const double* r,
                         double* r_out, double* r_alt, double* x, int cj, int epi) {
    k_mrr_update<<<ew_grid(ctx, k_mrr_update, n), EW_BLOCK, 0, ctx->stream>>>(n, ar, y, z, r, r_out, r_alt, x, cj, red_args_n(ctx, epi, 1));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, epi, -1, 0);
}

int pk_launch_kscg_update(pk_ctx* ctx, long long n, double* x, double* ar0, const double* ap0, double* ap0_out,
                          const double* ap1, int cj, int epi) {
    k_kscg_update<<<ew_grid(ctx, k_kscg_update, n), EW_BLOCK, 0, ctx->stream>>>(n, x, ar0, ap0, ap0_out, ap1, cj, red_args_n(ctx, epi, 1));
    PK_LAUNCH_CHECK();
    return pk_finish_reduce(ctx, 1, epi, -1, 0);
}

namespace {
template <int W, int MODE>
int gram_window(pk_ctx* ctx, long long n, long long ld, const double* U, int nu, const double* V, int nv, int j0,
                int epi, int ar_n) {
    PkRedArgs ra = red_args(ctx, epi, 6 * j0);
    ra.ar_n = ar_n;
    static int use_tma = -1;
    if (use_tma < 0) {
        const char* e = getenv("PK_GRAM");
        use_tma = !(e && strcmp(e, "plain") == 0);
    }
    const bool aligned = (((uintptr_t)U | (uintptr_t)V) & 15) == 0 && (ld & 1) == 0;
    if (use_tma && aligned && n >= 4 * EW_BLOCK) {
        constexpr int STAGES = (W <= 5) ? 2 : 3;       // small windows fit two blocks per SM; large ones run one block
        const size_t smem = (size_t)STAGES * 2 * (W + 1) * EW_BLOCK * sizeof(double);
        auto kern = k_gram_tma<W, MODE, STAGES>;
        const int per_sm = pk_blocks_per_sm((const void*)kern, EW_BLOCK, smem);
        long long cap = (long long)ctx->sm_count * per_sm;
        if (cap > ctx->red.max_blocks) cap = ctx->red.max_blocks;
        long long want = n / EW_BLOCK;
        int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        kern<<<grid, EW_BLOCK, smem, ctx->stream>>>(n, ld, U, nu, V, nv, j0, ra);
        PK_LAUNCH_CHECK();
        return PK_OK;
    }
    int grid = ew_grid(ctx, k_gram<W, MODE>, n, 1);   // register-heavy (6W accumulators): occupancy decides the grid
    k_gram<W, MODE><<<grid, EW_BLOCK, 0, ctx->stream>>>(n, ld, U, nu, V, nv, j0, ra);
    PK_LAUNCH_CHECK();
    return PK_OK;
}

template <int MODE>
int gram_dispatch(pk_ctx* ctx, int w, long long n, long long ld, const double* U, int nu, const double* V, int nv,
                  int j0, int epi, int ar_n) {
    switch (w) {
        case 1: return gram_window<1, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 2: return gram_window<2, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 3: return gram_window<3, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 4: return gram_window<4, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 5: return gram_window<5, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 6: return gram_window<6, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 7: return gram_window<7, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 8: return gram_window<8, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 9: return gram_window<9, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        case 10: return gram_window<10, MODE>(ctx, n, ld, U, nu, V, nv, j0, epi, ar_n);
        default: pk_set_error("gram window %d unsupported", w); return PK_ERR_ARG;
    }
}
}  // namespace

// All Gram sums for jj = 0 .. njj-1 (njj = k+2), in ceil(njj/10) passes (one pass for k <= 8).
// final_epi runs once everything is reduced (EPI_GRAM_CG / EPI_GRAM_MRR / EPI_NONE).
int pk_launch_gram(pk_ctx* ctx, int mode, long long n, long long ld, const double* U, int nu, const double* V, int nv,
                   int njj, int final_epi) {
    static int WMAX = 0;
    if (WMAX == 0) {
        const char* e = getenv("PK_GRAM_W");
        WMAX = e ? atoi(e) : 10;
        if (WMAX < 1 || WMAX > 10) WMAX = 10;
    }
    int j0 = 0;
    while (j0 < njj) {
        int w = njj - j0 < WMAX ? njj - j0 : WMAX;
        bool last = (j0 + w >= njj);
        int epi = last ? final_epi : EPI_GRAM_PART;
        const int ar_n = last ? 6 * njj : 0;
        if (mode == 0) PK_CHECK((gram_dispatch<0>(ctx, w, n, ld, U, nu, V, nv, j0, epi, ar_n)));
        else PK_CHECK((gram_dispatch<1>(ctx, w, n, ld, U, nu, V, nv, j0, epi, ar_n)));
        j0 += w;
    }
    if (ctx->n_ranks > 1 && !ctx->d_p2p && !ctx->nocomm) {
        PK_CHECK(pk_comm_allreduce(ctx, ctx->d_state->gram, 6LL * njj, ctx->stream));
        if (final_epi != EPI_NONE) PK_CHECK(pk_launch_scalar(ctx, final_epi, 0));
    }
    return PK_OK;
}

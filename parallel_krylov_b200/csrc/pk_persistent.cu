// Small systems (everything L2-resident, a CG iteration costs less than its three kernel launches): the whole CG loop as
// ONE cooperative persistent kernel.  Three grid-wide barriers per iteration replace three launches; the scalars
// (gamma, alpha, beta, residual, stop test) are recomputed redundantly and identically by every block from the block
// partials, summed in block order, so no block waits for a "scalar engine".  Same operations as Solve::cg()
// (/root/reference/v3/cpu/cg.py:19-37); only the shape of the dot-product reduction tree differs.
#include <cooperative_groups.h>

#include "pk_device.cuh"
#include "pk_launch.h"

namespace cg = cooperative_groups;

namespace {

constexpr int PB = 256;

struct CgPersistArgs {
    const int32_t* rowptr;
    const int32_t* col;
    const double* val;
    long long n;
    double* x;
    double* r;
    double* p;
    double* v;
    double* partials;      // [2][gridDim.x]
    PkState* st;
    int iters;             // iterations this launch may perform
};

// deterministic: every block adds the same numbers in the same order
__device__ __forceinline__ double sum_partials(const double* part, int nb, double* bcast) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        double s = 0.0;
        for (int b = lane; b < nb; b += 32) s += __ldcg(part + b);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) *bcast = s;
    }
    __syncthreads();
    const double out = *bcast;
    __syncthreads();
    return out;
}

__device__ __forceinline__ void block_partial(double acc, double* sh, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = sh[0];
#pragma unroll
        for (int w = 1; w < PB / 32; ++w) s += sh[w];
        *out = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PB) k_cg_persistent(CgPersistArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[PB / 32];
    __shared__ double bcast;
    PkState* st = a.st;
    if (*((volatile int*)&st->done)) return;          // uniform: written before this launch
    const long long gsize = (long long)gridDim.x * PB;
    const long long gtid = (long long)blockIdx.x * PB + threadIdx.x;
    const int nb = gridDim.x;
    double* part0 = a.partials;
    double* part1 = a.partials + nb;
    double gamma = st->gamma;
    long long it = st->it;
    const long long maxiter = st->maxiter;
    const double tol = st->tol, bnorm = st->bnorm;
    int done = 0, converged = 0;

    for (int step = 0; step < a.iters && !done; ++step) {
        // ---- v = A p ; sigma = p.v ---------------------------------------------------------------------------
        double acc = 0.0;
        for (long long row = gtid; row < a.n; row += gsize) {
            double sum = 0.0;
            const int s = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
            for (int j = s; j < e; ++j) sum += __ldg(a.val + j) * __ldcg(a.p + __ldg(a.col + j));
            a.v[row] = sum;
            acc += __ldcg(a.p + row) * sum;
        }
        block_partial(acc, sh, part0 + blockIdx.x);
        grid.sync();
        const double sigma = sum_partials(part0, nb, &bcast);
        const double alpha = gamma / sigma;
        // ---- x += alpha p ; r -= alpha v ; gamma' = r.r ------------------------------------------------------
        acc = 0.0;
        for (long long row = gtid; row < a.n; row += gsize) {
            a.x[row] = a.x[row] + alpha * __ldcg(a.p + row);
            const double ri = a.r[row] - alpha * a.v[row];
            a.r[row] = ri;
            acc += ri * ri;
        }
        block_partial(acc, sh, part1 + blockIdx.x);
        grid.sync();
        const double g = sum_partials(part1, nb, &bcast);
        const double beta = g / gamma;
        gamma = g;
        it += 1;
        const double res = sqrt(g) / bnorm;
        if (it < maxiter) {
            if (res < tol) { converged = 1; done = 1; }
        } else {
            converged = 0;
            done = 1;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && it < st->hist_len) {
            st->res[it] = res;
            st->nosl[it] = it;
        }
        // ---- p = r + beta p -----------------------------------------------------------------------------------
        for (long long row = gtid; row < a.n; row += gsize) a.p[row] = a.r[row] + beta * __ldcg(a.p + row);
        grid.sync();           // p complete before the next mat-vec; partial buffers free again
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->gamma = gamma;
        st->rr = gamma;
        st->it = it;
        st->idx = it;
        st->converged = converged;
        st->done = done;
    }
}

}  // namespace

// iterations per launch: `iters`.  Returns PK_ERR_UNSUPPORTED when the device cannot co-schedule the grid.
int pk_launch_cg_persistent(pk_ctx* ctx, pk_mat* m, double* x, double* r, double* p, double* v, int iters) {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (!coop) return PK_ERR_UNSUPPORTED;
    int per_sm = pk_blocks_per_sm((const void*)k_cg_persistent, PB, 0);   // cached per device
    if (per_sm > 4) per_sm = 4;            // a barrier over fewer blocks is cheaper; the work is latency-bound anyway
    const int grid_cache = ctx->sm_count * per_sm;
    long long want = (m->n_rows + PB - 1) / PB;
    int grid = (int)(want < grid_cache ? (want < 1 ? 1 : want) : grid_cache);
    if (2 * grid > ctx->red.max_blocks * PK_MAX_SUMS) return PK_ERR_UNSUPPORTED;
    CgPersistArgs a{m->rowptr, m->col, m->val, m->n_rows, x, r, p, v, ctx->red.partials, ctx->d_state, iters};
    void* params[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_cg_persistent, dim3(grid), dim3(PB), params, 0, ctx->stream);
    if (e != cudaSuccess) {
        pk_set_error("cooperative CG launch (grid %d): %s", grid, cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    ctx->spmvs += iters;
    return PK_OK;
}

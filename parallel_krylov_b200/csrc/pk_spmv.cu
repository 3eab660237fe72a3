// Operator application y = A·x for a row block of A (K1/K2 of SURVEY.md §2b), optionally for two right-hand sides at
// once (the two chains A^j r / A^j p of the k-skip basis share ONE pass over A), with the dot products the solver
// needs next reduced in the epilogue.  Replaces MultiGpu.dot (/root/reference/v3/gpu/common.py:113-126) and the
// cupy.dot launches that follow it (/root/reference/v3/gpu/cg.py:31-32).
//
// CSR-stream kernel (short rows: stencils, bands): a block owns a tile of BLOCK consecutive rows.  Phase 1 streams the
// tile's nonzeros with 128-bit coalesced loads (int4 of column indices, 2 x double2 of values), gathers x through the
// read-only path and parks val*x in shared memory.  Phase 2: thread t adds up row t's products left to right — the
// accumulation order of scipy's csr_matvec, which makes y bit-identical to the oracle's A.dot(x) (products and sums
// are separately rounded: the library is built with -fmad=false).  Tiles whose nonzeros exceed the staging buffer
// (long rows) are processed warp-per-row with a shuffle reduction instead.
// Dense kernel: warp per row, 128-bit loads, no tensor cores (GEMV is HBM-bound).
#include "pk_device.cuh"
#include "pk_launch.h"

namespace {

__device__ __forceinline__ bool pk_done(const PkState* st) { return *((volatile const int*)&st->done) != 0; }

struct SpmvArgs {
    const int32_t* rowptr;
    const int32_t* col;
    const double* val;
    const double* x0;
    const double* x1;
    double* y0;
    double* y1;
    const double* w;        // fused dots against this vector (nullable)
    long long row_lo, row_hi;
    long long nnz_total;
    int cap;                // staging capacity (nonzeros) per right-hand side
    int reduce;             // 1: run the grid reduction (3 sums)
};

template <int NV, int BLOCK, bool VEC>
__global__ void __launch_bounds__(BLOCK) k_spmv_stream(SpmvArgs a, PkRedArgs ra) {
    if (pk_done(ra.st)) return;
    extern __shared__ double prod[];          // [NV][cap]
    __shared__ int rp[BLOCK + 1];
    constexpr int NW = BLOCK / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* prod0 = prod;
    double* prod1 = prod + a.cap;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long n_rows = a.row_hi - a.row_lo;
    const long long n_tiles = (n_rows + BLOCK - 1) / BLOCK;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = a.row_lo + tile * BLOCK;
        const int nr = (int)((a.row_hi - r0) < BLOCK ? (a.row_hi - r0) : BLOCK);
        for (int t = tid; t <= nr; t += BLOCK) rp[t] = a.rowptr[r0 + t];
        __syncthreads();
        const int base = rp[0], end = rp[nr];
        if (end - base <= a.cap) {
            // ---- phase 1: stream the tile's nonzeros, stage val * x[col] -------------------------------------
            if (VEC) {
                const int q0 = base & ~3;
                for (int q = q0 + 4 * tid; q < end; q += 4 * BLOCK) {
                    int c[4];
                    double v[4];
                    if ((long long)q + 4 <= a.nnz_total) {
                        const int4 c4 = __ldg(reinterpret_cast<const int4*>(a.col + q));
                        const double2 v01 = __ldg(reinterpret_cast<const double2*>(a.val + q));
                        const double2 v23 = __ldg(reinterpret_cast<const double2*>(a.val + q + 2));
                        c[0] = c4.x; c[1] = c4.y; c[2] = c4.z; c[3] = c4.w;
                        v[0] = v01.x; v[1] = v01.y; v[2] = v23.x; v[3] = v23.y;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool ok = (long long)q + e < a.nnz_total;
                            c[e] = ok ? a.col[q + e] : 0;
                            v[e] = ok ? a.val[q + e] : 0.0;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int idx = q + e;
                        if (idx >= base && idx < end) {
                            prod0[idx - base] = v[e] * __ldg(a.x0 + c[e]);
                            if (NV == 2) prod1[idx - base] = v[e] * __ldg(a.x1 + c[e]);
                        }
                    }
                }
            } else {
                for (int q = base + tid; q < end; q += BLOCK) {
                    const int c = a.col[q];
                    const double v = a.val[q];
                    prod0[q - base] = v * __ldg(a.x0 + c);
                    if (NV == 2) prod1[q - base] = v * __ldg(a.x1 + c);
                }
            }
            __syncthreads();
            // ---- phase 2: one thread per row, left-to-right sum (scipy csr_matvec order) -----------------------
            if (tid < nr) {
                const int s = rp[tid] - base, e = rp[tid + 1] - base;
                double sum0 = 0.0, sum1 = 0.0;
                for (int j = s; j < e; ++j) {
                    sum0 += prod0[j];
                    if (NV == 2) sum1 += prod1[j];
                }
                const long long row = r0 + tid;
                a.y0[row] = sum0;
                if (NV == 2) a.y1[row] = sum1;
                if (a.w) {
                    const double wi = a.w[row];
                    acc[0] += wi * sum0;
                    acc[1] += sum0 * sum0;
                    acc[2] += wi * wi;
                }
            }
        } else {
            // ---- long rows: warp per row, lanes stride the row, shuffle tree ------------------------------------
            for (int r = warp; r < nr; r += NW) {
                const int s = rp[r], e = rp[r + 1];
                double sum0 = 0.0, sum1 = 0.0;
                for (int q = s + lane; q < e; q += 32) {
                    const int c = a.col[q];
                    const double v = a.val[q];
                    sum0 += v * __ldg(a.x0 + c);
                    if (NV == 2) sum1 += v * __ldg(a.x1 + c);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    sum0 += __shfl_down_sync(0xffffffffu, sum0, off);
                    if (NV == 2) sum1 += __shfl_down_sync(0xffffffffu, sum1, off);
                }
                if (lane == 0) {
                    const long long row = r0 + r;
                    a.y0[row] = sum0;
                    if (NV == 2) a.y1[row] = sum1;
                    if (a.w) {
                        const double wi = a.w[row];
                        acc[0] += wi * sum0;
                        acc[1] += sum0 * sum0;
                        acc[2] += wi * wi;
                    }
                }
            }
        }
        __syncthreads();   // rp / prod are reused by the next tile
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

// Dense row-major block: warp per row.
struct GemvArgs {
    const double* A;
    long long lda;
    long long n_rows, n_cols;
    const double* x0;
    const double* x1;
    double* y0;
    double* y1;
    const double* w;
    int reduce;
    int vec;               // rows 16-byte aligned: double2 loads
};

template <int NV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_gemv(GemvArgs a, PkRedArgs ra) {
    if (pk_done(ra.st)) return;
    constexpr int NW = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long row = (long long)blockIdx.x * NW + warp; row < a.n_rows; row += (long long)gridDim.x * NW) {
        const double* ar = a.A + row * a.lda;
        double s0 = 0.0, s1 = 0.0;
        if (a.vec) {
            const long long n2 = a.n_cols >> 1;
            const double2* ar2 = reinterpret_cast<const double2*>(ar);
            const double2* x02 = reinterpret_cast<const double2*>(a.x0);
            const double2* x12 = reinterpret_cast<const double2*>(a.x1);
#pragma unroll 4
            for (long long c = lane; c < n2; c += 32) {
                const double2 av = __ldg(ar2 + c);
                const double2 xv = __ldg(x02 + c);
                s0 += av.x * xv.x;
                s0 += av.y * xv.y;
                if (NV == 2) {
                    const double2 xw = __ldg(x12 + c);
                    s1 += av.x * xw.x;
                    s1 += av.y * xw.y;
                }
            }
            if ((a.n_cols & 1) && lane == 0) {
                const long long c = a.n_cols - 1;
                s0 += ar[c] * a.x0[c];
                if (NV == 2) s1 += ar[c] * a.x1[c];
            }
        } else {
            for (long long c = lane; c < a.n_cols; c += 32) {
                const double av = ar[c];
                s0 += av * a.x0[c];
                if (NV == 2) s1 += av * a.x1[c];
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s0 += __shfl_down_sync(0xffffffffu, s0, off);
            if (NV == 2) s1 += __shfl_down_sync(0xffffffffu, s1, off);
        }
        if (lane == 0) {
            a.y0[row] = s0;
            if (NV == 2) a.y1[row] = s1;
            if (a.w) {
                const double wi = a.w[row];
                acc[0] += wi * s0;
                acc[1] += s0 * s0;
                acc[2] += wi * wi;
            }
        }
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

template <int NV, int BLOCK, bool VEC>
int stream_grid(pk_ctx* ctx, size_t smem, long long n_tiles) {
    static int per_sm_cache[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int per_sm = 0;
    auto kern = k_spmv_stream<NV, BLOCK, VEC>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    (void)per_sm_cache;
    long long g = (long long)ctx->sm_count * per_sm;
    if (g > n_tiles) g = n_tiles;
    if (g < 1) g = 1;
    return (int)g;
}

// mode 0: choose the grid and launch; 1: dry run (only report the grid); 2: launch with the grid in *grid_io
template <int NV, int BLOCK, bool VEC>
int launch_stream(pk_ctx* ctx, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int grid_cap, int mode) {
    const size_t smem = (size_t)NV * a.cap * sizeof(double);
    const long long n_rows = a.row_hi - a.row_lo;
    if (n_rows <= 0) { *grid_io = 0; return PK_OK; }
    const long long n_tiles = (n_rows + BLOCK - 1) / BLOCK;
    int grid = *grid_io;
    if (mode != 2) {
        grid = stream_grid<NV, BLOCK, VEC>(ctx, smem, n_tiles);
        if (grid > grid_cap) grid = grid_cap;
        *grid_io = grid;
        if (mode == 1) return PK_OK;
    }
    k_spmv_stream<NV, BLOCK, VEC><<<grid, BLOCK, smem, ctx->stream>>>(a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("spmv launch (grid %d, smem %zu): %s", grid, smem, cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    return PK_OK;
}

int launch_stream_any(pk_ctx* ctx, pk_mat* m, bool two, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap,
                      int mode) {
    const bool vec = m->vec_ok;
    if (m->tile_rows == 128) {
        if (two) return vec ? launch_stream<2, 128, true>(ctx, a, ra, grid_io, cap, mode)
                            : launch_stream<2, 128, false>(ctx, a, ra, grid_io, cap, mode);
        return vec ? launch_stream<1, 128, true>(ctx, a, ra, grid_io, cap, mode)
                   : launch_stream<1, 128, false>(ctx, a, ra, grid_io, cap, mode);
    }
    if (two) return vec ? launch_stream<2, 256, true>(ctx, a, ra, grid_io, cap, mode)
                        : launch_stream<2, 256, false>(ctx, a, ra, grid_io, cap, mode);
    return vec ? launch_stream<1, 256, true>(ctx, a, ra, grid_io, cap, mode)
               : launch_stream<1, 256, false>(ctx, a, ra, grid_io, cap, mode);
}

int launch_gemv(pk_ctx* ctx, pk_mat* m, bool two, const GemvArgs& a, PkRedArgs ra) {
    constexpr int BLOCK = 256;
    long long want = (a.n_rows + (BLOCK / 32) - 1) / (BLOCK / 32);
    long long cap = (long long)ctx->sm_count * 8;
    if (cap > ctx->red.max_blocks) cap = ctx->red.max_blocks;
    int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (two) k_gemv<2, BLOCK><<<grid, BLOCK, 0, ctx->stream>>>(a, ra);
    else k_gemv<1, BLOCK><<<grid, BLOCK, 0, ctx->stream>>>(a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("gemv launch: %s", cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    (void)m;
    return PK_OK;
}

}  // namespace

static int spmv_impl(pk_ctx* ctx, pk_mat* m, double* x, double* y, double* x1, double* y1, PkDots dots);

int pk_launch_spmv(pk_ctx* ctx, pk_mat* m, double* x, double* y, double* x1, double* y1, PkDots dots) {
    bool prof = false;
    if (ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size()) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(ctx->stream, &cs);
        prof = (cs == cudaStreamCaptureStatusNone);
    }
    if (prof) PK_CUDA(cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    int rc = spmv_impl(ctx, m, x, y, x1, y1, dots);
    if (prof) {
        PK_CUDA(cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    return rc;
}

static int spmv_impl(pk_ctx* ctx, pk_mat* m, double* x, double* y, double* x1, double* y1, PkDots dots) {
    const bool two = (x1 != nullptr);
    PkRedArgs ra;
    ra.partials = ctx->red.partials;
    ra.ticket = ctx->red.ticket;
    ra.max_blocks = ctx->red.max_blocks;
    ra.st = ctx->d_state;
    ra.epi = dots.epi;
    ra.defer = ctx->n_ranks > 1 ? 1 : 0;
    ra.g_off = -1;
    ra.block_off = 0;
    ra.nb_total = 0;
    ra.store_only = 0;
    ctx->spmvs += two ? 2 : 1;

    if (m->kind == MAT_DENSE) {
        if (m->distributed && m->n_halo > 0) {
            PK_CHECK(pk_comm_halo_start(ctx, m, x, x1));
            PK_CHECK(pk_comm_halo_wait(ctx));
        }
        GemvArgs g;
        g.A = m->dense; g.lda = m->lda; g.n_rows = m->n_rows; g.n_cols = m->n_cols;
        g.x0 = x; g.x1 = x1; g.y0 = y; g.y1 = y1; g.w = dots.w; g.reduce = dots.w ? 1 : 0;
        g.vec = (((uintptr_t)m->dense & 15) == 0 && (m->lda & 1) == 0 && ((uintptr_t)x & 15) == 0 &&
                 (!x1 || ((uintptr_t)x1 & 15) == 0)) ? 1 : 0;
        PK_CHECK(launch_gemv(ctx, m, two, g, ra));
        if (dots.w) return pk_finish_reduce(ctx, 3, dots.epi, -1, 0);
        return PK_OK;
    }

    SpmvArgs a;
    a.rowptr = m->rowptr; a.col = m->col; a.val = m->val;
    a.x0 = x; a.x1 = x1; a.y0 = y; a.y1 = y1; a.w = dots.w;
    a.nnz_total = m->nnz;
    a.cap = m->tile_cap;
    a.reduce = dots.w ? 1 : 0;
    int grid = 0;

    if (!m->distributed || m->n_halo == 0) {
        a.row_lo = 0; a.row_hi = m->n_rows;
        PK_CHECK(launch_stream_any(ctx, m, two, a, ra, &grid, ctx->red.max_blocks, 0));
    } else {
        // Interior rows (no halo column) run while the halo of x is in flight on the side stream; the boundary
        // rows follow the exchange.  All launches store partials into disjoint block slots of ONE reduction,
        // which the last launch finishes (fixed slot order => deterministic).
        PK_CHECK(pk_comm_halo_start(ctx, m, x, x1));
        const long long lo = m->interior_lo, hi = m->interior_hi;
        const long long rlo[3] = {lo, 0, hi}, rhi[3] = {hi, lo, m->n_rows};
        const int cap_each = ctx->red.max_blocks / 3;
        int grids[3] = {0, 0, 0}, last = -1, total = 0;
        for (int i = 0; i < 3; ++i) {
            if (rhi[i] <= rlo[i]) continue;
            a.row_lo = rlo[i]; a.row_hi = rhi[i];
            PK_CHECK(launch_stream_any(ctx, m, two, a, ra, &grids[i], cap_each, 1));
            total += grids[i];
            last = i;
        }
        int off = 0;
        bool waited = false;
        for (int i = 0; i < 3; ++i) {
            if (i >= 1 && !waited) { PK_CHECK(pk_comm_halo_wait(ctx)); waited = true; }
            if (grids[i] == 0) continue;
            a.row_lo = rlo[i]; a.row_hi = rhi[i];
            PkRedArgs r2 = ra;
            r2.block_off = off;
            r2.nb_total = total;
            r2.store_only = (i != last) ? 1 : 0;
            PK_CHECK(launch_stream_any(ctx, m, two, a, r2, &grids[i], cap_each, 2));
            off += grids[i];
        }
        if (!waited) PK_CHECK(pk_comm_halo_wait(ctx));
    }
    if (dots.w) return pk_finish_reduce(ctx, 3, dots.epi, -1, 0);
    return PK_OK;
}

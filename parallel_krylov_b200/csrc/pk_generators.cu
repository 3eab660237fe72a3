// Synthetic systems generated directly in HBM: the same matrices, bit for bit, as
// parallel_krylov_b200/problems.py builds on the host.  Row counts come out of a first kernel; the caller turns them
// into the row pointer (exclusive scan) and calls the fill kernel.  Column indices are GLOBAL; a distributed caller
// remaps them to [owned | halo] afterwards.
#include "pk_common.cuh"

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double hash_uniform(unsigned long long seed, unsigned long long stream, unsigned long long index) {
    const unsigned long long s = splitmix64(seed + 0x632BE59BD9B4E019ull * stream);
    const unsigned long long h = splitmix64(s ^ index);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_stencil_counts(long long nx, long long ny, long long nz, long long row0, long long n_rows,
                                 int32_t* __restrict__ counts) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (long long)gridDim.x * blockDim.x) {
        const long long i = row0 + t;
        const long long ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
        int c = 1 + (ix > 0) + (ix < nx - 1) + (iy > 0) + (iy < ny - 1);
        if (nz > 1) c += (iz > 0) + (iz < nz - 1);
        counts[t] = c;
    }
}

__global__ void k_stencil_fill(long long nx, long long ny, long long nz, long long row0, long long n_rows,
                               const int32_t* __restrict__ rowptr, int32_t* __restrict__ col, double* __restrict__ val) {
    const double diag = nz > 1 ? 6.0 : 4.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (long long)gridDim.x * blockDim.x) {
        const long long i = row0 + t;
        const long long ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
        long long p = rowptr[t];
        if (nz > 1 && iz > 0) { col[p] = (int32_t)(i - nx * ny); val[p++] = -1.0; }
        if (iy > 0) { col[p] = (int32_t)(i - nx); val[p++] = -1.0; }
        if (ix > 0) { col[p] = (int32_t)(i - 1); val[p++] = -1.0; }
        col[p] = (int32_t)i; val[p++] = diag;
        if (ix < nx - 1) { col[p] = (int32_t)(i + 1); val[p++] = -1.0; }
        if (iy < ny - 1) { col[p] = (int32_t)(i + nx); val[p++] = -1.0; }
        if (nz > 1 && iz < nz - 1) { col[p] = (int32_t)(i + nx * ny); val[p++] = -1.0; }
    }
}

__global__ void k_banded_counts(long long n, int hb, long long row0, long long n_rows, int32_t* __restrict__ counts) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (long long)gridDim.x * blockDim.x) {
        const long long i = row0 + t;
        const long long lo = i - hb > 0 ? i - hb : 0, hi = i + hb < n - 1 ? i + hb : n - 1;
        counts[t] = (int32_t)(hi - lo + 1);
    }
}

__global__ void k_banded_fill(long long n, int hb, unsigned long long seed, long long row0, long long n_rows,
                              const int32_t* __restrict__ rowptr, int32_t* __restrict__ col, double* __restrict__ val) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (long long)gridDim.x * blockDim.x) {
        const long long i = row0 + t;
        const long long lo = i - hb > 0 ? i - hb : 0, hi = i + hb < n - 1 ? i + hb : n - 1;
        const long long p0 = rowptr[t];
        double dsum = 0.0;
        for (long long j = lo; j <= hi; ++j) {
            if (j == i) continue;
            const long long d = j > i ? j - i : i - j;
            const long long key = j > i ? i : j;           // the pair is keyed by its smaller index
            const double w = 0.1 + 0.9 * hash_uniform(seed, (unsigned long long)d, (unsigned long long)key);
            col[p0 + (j - lo)] = (int32_t)j;
            val[p0 + (j - lo)] = -w;
            dsum = dsum + w;                               // ascending column order, like problems.banded_spd
        }
        col[p0 + (i - lo)] = (int32_t)i;
        val[p0 + (i - lo)] = dsum + 1.0;
    }
}

__global__ void k_hash_normal(unsigned long long seed, long long offset, long long n, double* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const unsigned long long i = (unsigned long long)(offset + t);
        const double u1 = hash_uniform(seed, 1, i), u2 = hash_uniform(seed, 2, i);
        out[t] = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
    }
}

inline int gen_grid(pk_ctx* ctx, long long n) {
    long long g = (n + 255) / 256;
    long long cap = (long long)ctx->sm_count * 16;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

}  // namespace

#define PK_GEN_PROLOG()                      \
    PK_REQUIRE(ctx != nullptr, "null context"); \
    PK_CUDA(cudaSetDevice(ctx->device))

extern "C" int pk_gen_stencil_counts(pk_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_rows,
                                     int32_t* d_counts) {
    PK_GEN_PROLOG();
    PK_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && nx * ny * nz < (1LL << 31), "grid must have < 2^31 points");
    k_stencil_counts<<<gen_grid(ctx, n_rows), 256, 0, ctx->stream>>>(nx, ny, nz, row0, n_rows, d_counts);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_gen_stencil_fill(pk_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_rows,
                                   const int32_t* d_rowptr, int32_t* d_col, double* d_val) {
    PK_GEN_PROLOG();
    k_stencil_fill<<<gen_grid(ctx, n_rows), 256, 0, ctx->stream>>>(nx, ny, nz, row0, n_rows, d_rowptr, d_col, d_val);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_gen_banded_counts(pk_ctx* ctx, int64_t n, int half_bw, int64_t row0, int64_t n_rows, int32_t* d_counts) {
    PK_GEN_PROLOG();
    PK_REQUIRE(n < (1LL << 31), "n must be < 2^31");
    k_banded_counts<<<gen_grid(ctx, n_rows), 256, 0, ctx->stream>>>(n, half_bw, row0, n_rows, d_counts);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_gen_banded_fill(pk_ctx* ctx, int64_t n, int half_bw, uint64_t seed, int64_t row0, int64_t n_rows,
                                  const int32_t* d_rowptr, int32_t* d_col, double* d_val) {
    PK_GEN_PROLOG();
    k_banded_fill<<<gen_grid(ctx, n_rows), 256, 0, ctx->stream>>>(n, half_bw, seed, row0, n_rows, d_rowptr, d_col, d_val);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_fill_hash_normal(pk_ctx* ctx, uint64_t seed, int64_t offset, int64_t n, double* d_out) {
    PK_GEN_PROLOG();
    k_hash_normal<<<gen_grid(ctx, n), 256, 0, ctx->stream>>>(seed, offset, n, d_out);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

/* CPU ORACLE HELPERS — TEST INFRASTRUCTURE ONLY (never linked into libpkrylov, never imported by the package).
 *
 *  pko_csr_matvec : plain-C restatement of the CSR mat-vec the reference reaches through `A.dot(v)`
 *                   (/root/reference/v3/cpu/cg.py:27 -> scipy.sparse csr_matrix.dot -> sparsetools `csr_matvec`,
 *                   scipy 1.18.1, not vendored): for each row, sum = Yx[i]; sum += Ax[jj]*Xx[Aj[jj]] left to right.
 *                   Built with -ffp-contract=off so products and sums round separately like the x86-64 scipy wheel.
 *                   tests/test_oracle_golden.py pins it against scipy bit for bit.
 *  pko_stencil_*  : host generator of the 5-/7-point Laplacian (same matrix as problems.py / pk_gen_stencil_*),
 *                   so bench.py can build the 256^3 / 512^3 CPU-baseline input in seconds (no numpy temporaries).
 */
#include <stdint.h>
#include <stdlib.h>

void pko_csr_matvec(int64_t n_row, const int32_t* Ap, const int32_t* Aj, const double* Ax, const double* Xx,
                    double* Yx) {
    for (int64_t i = 0; i < n_row; ++i) {
        double sum = Yx[i];
        for (int32_t jj = Ap[i]; jj < Ap[i + 1]; ++jj) sum += Ax[jj] * Xx[Aj[jj]];
        Yx[i] = sum;
    }
}

static inline int stencil_count(int64_t i, int64_t nx, int64_t ny, int64_t nz) {
    const int64_t ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
    int c = 1 + (ix > 0) + (ix < nx - 1) + (iy > 0) + (iy < ny - 1);
    if (nz > 1) c += (iz > 0) + (iz < nz - 1);
    return c;
}

/* rowptr[0..n] (int32); returns nnz */
int64_t pko_stencil_rowptr(int64_t nx, int64_t ny, int64_t nz, int32_t* rowptr) {
    const int64_t n = nx * ny * nz;
    int64_t acc = 0;
    rowptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        acc += stencil_count(i, nx, ny, nz);
        rowptr[i + 1] = (int32_t)acc;
    }
    return acc;
}

void pko_stencil_fill(int64_t nx, int64_t ny, int64_t nz, const int32_t* rowptr, int32_t* col, double* val) {
    const int64_t n = nx * ny * nz;
    const double diag = nz > 1 ? 6.0 : 4.0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
        int64_t p = rowptr[i];
        if (nz > 1 && iz > 0) { col[p] = (int32_t)(i - nx * ny); val[p++] = -1.0; }
        if (iy > 0) { col[p] = (int32_t)(i - nx); val[p++] = -1.0; }
        if (ix > 0) { col[p] = (int32_t)(i - 1); val[p++] = -1.0; }
        col[p] = (int32_t)i; val[p++] = diag;
        if (ix < nx - 1) { col[p] = (int32_t)(i + 1); val[p++] = -1.0; }
        if (iy < ny - 1) { col[p] = (int32_t)(i + nx); val[p++] = -1.0; }
        if (nz > 1 && iz < nz - 1) { col[p] = (int32_t)(i + nx * ny); val[p++] = -1.0; }
    }
}

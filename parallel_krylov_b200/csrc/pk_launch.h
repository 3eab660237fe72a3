// Internal launch API shared by the translation units of libpkrylov (not part of the C-ABI).
#pragma once
#include "pk_common.cuh"

// Resident blocks per SM of a kernel (occupancy API, cached): persistent grids are sized sm_count x this, so that a
// grid-stride kernel runs as exactly one full wave (a partial second wave costs up to 2x on these streaming kernels).
int pk_blocks_per_sm(const void* kernel, int block, size_t smem);

// pk_kernels.cu — fused vector kernels; every reducing launch is followed (multi-GPU) by all-reduce + scalar engine
int pk_launch_scalar(pk_ctx* ctx, int epi, int ignore_done);
int pk_launch_set_k(pk_ctx* ctx, int k);
int pk_finish_reduce(pk_ctx* ctx, int nsums, int epi, int g_off, int ignore_done);
int pk_launch_dot(pk_ctx* ctx, long long n, const double* u, const double* v, int epi, int ignore_done);
int pk_launch_resid_init(pk_ctx* ctx, long long n, const double* b, const double* v, double* r, double* p, int epi);
int pk_launch_cg_xr(pk_ctx* ctx, long long n, double* x, double* r, const double* p, const double* v);
int pk_launch_cg_xr_split(pk_ctx* ctx, long long n, double* x, double* r, const double* p, const double* v);
int pk_launch_cg_p(pk_ctx* ctx, long long n, double* p, const double* r);
int pk_launch_mrr_first(pk_ctx* ctx, long long n, const double* ar, double* r, double* x, double* y, double* z,
                        int epi);
int pk_launch_mrr_s(pk_ctx* ctx, long long n, const double* ar, const double* y, const double* r);
int pk_launch_mrr_update(pk_ctx* ctx, long long n, const double* ar, double* y, double* z, const double* r,
                         double* r_out, double* r_alt, double* x, int cj, int epi);
int pk_launch_cgcg_update(pk_ctx* ctx, long long n, double* x, double* r, double* u, const double* w, double* p,
                          double* s, const double* mdiag, int init);
int pk_launch_axpby(pk_ctx* ctx, long long n, double s1, const double* a, double s2, const double* b, double* out);
int pk_launch_adapt_save(pk_ctx* ctx, long long n, double* x, double* best_x);
int pk_launch_kscg_update(pk_ctx* ctx, long long n, double* x, double* ar0, const double* ap0, double* ap0_out,
                          const double* ap1, int cj, int epi);
int pk_launch_gram(pk_ctx* ctx, int mode, long long n, long long ld, const double* U, int nu, const double* V, int nv,
                   int njj, int final_epi);

// pk_spmv.cu — operator application y = A x (optionally two right-hand sides), fused dot epilogue
struct PkDots {
    const double* w = nullptr;   // sums: [0] = w.y, [1] = y.y, [2] = w.w   (nullptr: no reduction)
    int epi = EPI_NONE;
    int extra_sums = 0;          // st->red[3 .. 3+extra) hold local sums of an EARLIER kernel: all-reduce them with these
    // k-skip step fused into the row epilogue (see SpmvArgs in pk_spmv.cu); only the TMA CSR kernel supports it.
    // fuse = 3 (two right-hand sides): three-term recurrence of the Chebyshev basis in the epilogue,
    //   y0 = cs[0] (A x0) + cs[1] x0 + cs[2] f_a ;  y1 = cs[3] (A x1) + cs[4] x1 + cs[5] f_b   (f_a / f_b may be null)
    int fuse = 0, cj = 0;
    double cs[6] = {0, 0, 0, 0, 0, 0};
    double* f_a = nullptr;
    double* f_b = nullptr;
    double* f_x = nullptr;
    double* f_out = nullptr;
};
bool pk_mat_can_fuse(const pk_mat* m);
int pk_launch_spmv(pk_ctx* ctx, pk_mat* mat, double* x, double* y, double* x1, double* y1, PkDots dots);
int pk_csr_validate(pk_ctx* ctx, const void* rowptr, int rowptr64, const int32_t* col, long long n_rows,
                    long long n_cols, long long nnz, int* flags);
int pk_rebase_rowptr(pk_ctx* ctx, const int64_t* rp64, long long base, long long count, int32_t* out);
int pk_tile_max_nnz(pk_ctx* ctx, const int32_t* rowptr, long long n_rows, int tile_rows, int* result);

// pk_matpow.cu — k levels of both basis chains in one pass over A (small-bandwidth operators)
bool pk_matpow_ok(pk_ctx* ctx, pk_mat* m, int k);
int pk_launch_matpow(pk_ctx* ctx, pk_mat* m, int k, double* base0, double* base1, int dyn);
bool pk_mrr_steps_ok(pk_ctx* ctx, pk_mat* m, int k);
int pk_launch_mrr_steps(pk_ctx* ctx, pk_mat* m, int k, double* r, double* ar, double* y, double* z, double* x,
                        double* t0, double* t1, double* t2, int epi, long long own_lo = 0, long long own_hi = -1);
long long pk_band_ext_depth(pk_ctx* ctx, pk_mat* m, int k);   // > 0: ghost depth of the row-partitioned dense-band trip; 0: not applicable
int pk_comm_ghost_exchange_inplace(pk_ctx* ctx, double* const* own, int nvec, long long n_loc, long long depth,
                                   bool has_prev, bool has_next);

// pk_persistent.cu — whole CG loop as one cooperative kernel (small, L2-resident systems)
int pk_launch_cg_persistent(pk_ctx* ctx, pk_mat* m, double* x, double* r, double* p, double* v, int iters);

// pk_comm.cu — NCCL over NVLink
int pk_comm_allreduce(pk_ctx* ctx, double* buf, long long n, cudaStream_t s);
int pk_comm_allgather(pk_ctx* ctx, const double* send, double* recv, long long n, cudaStream_t s);
int pk_comm_ghost_exchange(pk_ctx* ctx, const double* v0, const double* v1, long long n_loc, long long depth, double* gin,
                           bool has_prev, bool has_next);
int pk_comm_halo_start(pk_ctx* ctx, pk_mat* mat, double* x, double* x1);   // pack + send/recv on the side stream
int pk_comm_halo_wait(pk_ctx* ctx);
void pk_mat_halo_p2p_close(pk_mat* m);                                        // main stream waits for the exchange

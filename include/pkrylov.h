/*
 * pkrylov.h — C-ABI of libpkrylov.so, the B200 (sm_100a) implementation of the solver inner loop of
 * 5enxia/parallel-krylov (v3 entry points cg / mrr / kskipcg / kskipmrr / adaptivekskipmrr).
 *
 * The reference has no FFI: its boundary is a plain Python call.  This header is the boundary
 * a maintainer binds with ctypes; every entry point cites the reference code it replaces.
 * Conventions: all pointers named d_* are DEVICE pointers (e.g. torch `tensor.data_ptr()`); buffers are
 * borrowed, never owned; every call is asynchronous on the context's stream unless stated; the return value is
 * 0 on success or a negative pk_status, with pk_last_error() giving the text.  No exceptions cross the ABI.
 * fp64 values, int32 column indices; row pointers int32 (pk_mat_csr) or int64 (pk_mat_csr64, nnz >= 2^31).
 */
#ifndef PKRYLOV_H
#define PKRYLOV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_VERSION 102
#define PK_KMAX 32            /* largest k of the k-skip solvers */
#define PK_NCCL_ID_BYTES 128
#define PK_IPC_HANDLE_BYTES 64

typedef enum {
    PK_OK = 0,
    PK_ERR_CUDA = -1,
    PK_ERR_ARG = -2,
    PK_ERR_NCCL = -3,
    PK_ERR_UNSUPPORTED = -4
} pk_status;

typedef enum {
    PK_CG = 0,               /* v3/gpu/cg.py:8            */
    PK_MRR = 1,              /* v3/gpu/mrr.py:8           */
    PK_KSKIPCG = 2,          /* v3/gpu/kskipcg.py:9       */
    PK_KSKIPMRR = 3,         /* v3/gpu/kskipmrr.py:9      */
    PK_ADAPTIVEKSKIPMRR = 4, /* v3/gpu/adaptivekskipmrr.py:10 (semantics of v3/cpu/adaptivekskipmrr.py) */
    PK_CGCG = 5              /* Chronopoulos-Gear CG, one reduction per iteration, optional Jacobi preconditioner
                                (v1/threads/pipeline/chronopoulos_gear.py:7) — opt-in, not a v3 method */
} pk_method;

typedef struct pk_ctx pk_ctx;   /* one per device/stream: scratch for reductions, SM count, optional communicator */
typedef struct pk_mat pk_mat;   /* a row block of A: CSR or dense, plus its halo plan when distributed           */

/* ------------------------------------------------------------------------------------------------------------ */
/* library                                                                                                      */
int pk_version(void);
const char* pk_last_error(void);

/* Context.  Replaces MultiGpu.init() (v3/gpu/common.py:62-79: per-device pool + stream + P2P enable): one
 * process per GPU here, so a context is one device and one stream (`stream` = cudaStream_t as void*, 0 = legacy
 * default stream). */
int pk_ctx_create(pk_ctx** out, int device, void* stream);
int pk_ctx_destroy(pk_ctx* ctx);
int pk_ctx_sync(pk_ctx* ctx);                     /* cudaStreamSynchronize */
int pk_ctx_sm_count(pk_ctx* ctx);
/* Per-launch timing of the operator kernel with CUDA event pairs on the context's stream (bench.py's roofline
 * numerator).  Only plain stream launches are timed (graph replays are not): total over the recorded launches. */
int pk_prof_begin(pk_ctx* ctx, int max_launches);
int pk_prof_end(pk_ctx* ctx, double* total_ms, int64_t* n_launches);

/* ------------------------------------------------------------------------------------------------------------ */
/* operator (row block of A).  Replaces MultiGpu.alloc() (v3/gpu/common.py:83-109, v3/gpu/mpi/common.py:102-134) */

/* CSR block: n_rows local rows; column indices address a vector of length n_cols_local (= n_rows + n_halo when
 * distributed: owned entries first, halo entries after, see pk_mat_set_halo).  Arrays are borrowed. */
int pk_mat_csr(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols_local, int64_t nnz,
               const int32_t* d_rowptr, const int32_t* d_col, const double* d_val);
/* The same with 64-bit row pointers, for blocks of nnz >= 2^31 (column indices stay int32: fewer than 2^31 columns).
 * The block is applied as row segments of < 2^31 nonzeros with 32-bit row pointers rebased per segment (built here, owned
 * by the operator).  Single-GPU operators and distributed blocks that need no halo exchange; a distributed block that
 * exchanges must have nnz < 2^31 (shard further). */
int pk_mat_csr64(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols_local, int64_t nnz,
                 const int64_t* d_rowptr, const int32_t* d_col, const double* d_val);
/* Dense row-major block (the reference's np.ndarray branch, v3/gpu/common.py:100-101 → cuBLAS dgemv). */
int pk_mat_dense(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols, const double* d_a, int64_t lda);
int pk_mat_destroy(pk_mat* mat);
/* Kernel choice made from the nnz distribution: 0 = csr-stream (rows staged in shared memory, thread per row),
 * 1 = csr-vector (warp per row), 2 = dense gemv.  *tile_rows / *tile_cap describe the stream tiling. */
int pk_mat_kernel_info(pk_mat* mat, int* kind, int* tile_rows, int* tile_cap);

/* Row-pattern compression (opt-in, lossless).  Rows with the same (column offset from the row, value) sequence share
 * one table entry; the block is then applied from one 16-bit id per row (constant-coefficient stencils: 2 bytes per row
 * instead of 12 per nonzero) with the same products in the same order as the CSR kernels.  pk_mat_row_hashes gives a
 * 64-bit hash per row to group rows by; pk_mat_set_patterns verifies on the device that EVERY row reproduces its pattern
 * bit for bit (else PK_ERR_ARG and the CSR kernels stay in use).  Table limits: <= 65535 patterns, 12*entries +
 * 4*(patterns+1) <= 64 KiB.  All arrays are borrowed device pointers. */
int pk_mat_row_hashes(pk_mat* mat, uint64_t* d_hash);
int pk_mat_set_patterns(pk_mat* mat, int n_pat, int n_entries, const uint16_t* d_id, const int32_t* d_ptr,
                        const int32_t* d_off, const double* d_val);

/* Halo plan for a distributed block (replaces the full-vector memcpyPeer broadcast v3/gpu/mpi/common.py:144 and
 * comm.Allgather :163).  For each peer p: this rank sends send_count[p] owned entries, listed (local indices) in
 * d_send_idx[send_off[p] .. send_off[p+1]), and receives recv_count[p] entries into the halo tail at
 * [n_rows + recv_off[p], ...).  If a peer's send list is one contiguous run the pack kernel is skipped.
 * interior_lo/hi: rows [lo,hi) reference no halo column (overlapped with the exchange). */
int pk_mat_set_halo(pk_mat* mat, int n_peers_total, const int64_t* send_off, const int64_t* recv_off,
                    const int32_t* d_send_idx, const int32_t* h_send_idx, int64_t interior_lo, int64_t interior_hi);

/* Halo exchange fused into the operator kernel over NVLink peer memory (the default when the peers' buffers can be
 * mapped; no NCCL, no side stream, one launch): each rank exports its halo receive buffer (pk_mat_halo_p2p_handle), the
 * handles are all-gathered, and pk_mat_halo_p2p_open maps the peers' buffers; dst_off[q] = where this rank's entries
 * start inside q's halo (q's recv_off[this rank]), peer_nhalo[q] = q's halo length.  The SpMV kernel then pushes this
 * rank's boundary entries with plain remote stores + a sequence flag at its start, runs the interior tiles while the
 * peers' entries are in flight and waits for the peers' flags right before its boundary tiles.
 * pk_mat_halo_p2p_disable returns the block to the ncclSend/ncclRecv exchange (all ranks must agree on the path). */
int pk_mat_halo_p2p_handle(pk_mat* mat, char handle[PK_IPC_HANDLE_BYTES]);
int pk_mat_halo_p2p_open(pk_mat* mat, const char* handles, const int64_t* dst_off, const int64_t* peer_nhalo);
int pk_mat_halo_p2p_disable(pk_mat* mat);

/* ------------------------------------------------------------------------------------------------------------ */
/* communicator (NCCL over NVLink).  Replaces MultiGpu.joint_mpi(comm) (v3/gpu/mpi/common.py:168-171).
 * libnccl is dlopen'ed from `nccl_path` (the copy torch already loaded); single-GPU use never touches NCCL. */
int pk_nccl_unique_id(const char* nccl_path, char id[PK_NCCL_ID_BYTES]);
int pk_comm_init(pk_ctx* ctx, const char* nccl_path, int n_ranks, int rank, const char id[PK_NCCL_ID_BYTES]);
int pk_comm_destroy(pk_ctx* ctx);
/* Measurement aid: with nocomm != 0 a solve launches the same kernels but skips halo exchange and all-reduces (its
 * results are meaningless); bench.py times it to report the exposed communication per iteration. */
int pk_ctx_set_nocomm(pk_ctx* ctx, int on);
/* Optional: all-reduce the dot products INSIDE the reducing kernels over NVLink peer memory (one mailbox per rank,
 * exported with CUDA IPC) instead of ncclAllReduce + a scalar kernel.  pk_p2p_handle returns this rank's 64-byte IPC
 * handle; after all-gathering the handles, pk_p2p_open(handles = n_ranks x 64 bytes) maps the peers and switches the
 * context to the fused path.  Returns PK_ERR_UNSUPPORTED (and stays on the NCCL path) if peer mapping fails. */
int pk_p2p_handle(pk_ctx* ctx, char handle[PK_IPC_HANDLE_BYTES]);
int pk_p2p_open(pk_ctx* ctx, int n_ranks, int rank, const char* handles);
int pk_allreduce_sum(pk_ctx* ctx, double* d_buf, int64_t n);           /* in place, fp64 sum */
int pk_allgather(pk_ctx* ctx, const double* d_send, double* d_recv, int64_t n_per_rank);

/* ------------------------------------------------------------------------------------------------------------ */
/* building blocks (each is one kernel launch; exposed so the parity tests can check them op by op)             */

/* y = A·x (and y1 = A·x1 when d_x1 != NULL: the two-chain pass of the matrix-powers basis).  Replaces
 * MultiGpu.dot (v3/gpu/common.py:113-126).  When d_w != NULL the epilogue also reduces, deterministically,
 * d_sums[0] = w·y, d_sums[1] = y·y, d_sums[2] = w·w  (replaces the separate cupy.dot launches, e.g.
 * v3/gpu/cg.py:32, v3/gpu/mrr.py:41-42).  Distributed blocks exchange the halo of x first. */
int pk_spmv(pk_ctx* ctx, pk_mat* mat, double* d_x, double* d_y, double* d_x1, double* d_y1,
            const double* d_w, double* d_sums);
/* Matrix-powers kernel: levels l = 1..k of BOTH k-skip basis chains, (A^l u, A^l v), in ONE pass over A, written to
 * d_base0 + l*ld and d_base1 + l*ld (ld = pk_mat_ld) from level 0 at d_base0 / d_base1 — bit-identical to k sequential
 * pk_spmv calls.  Replaces the basis loops v3/cpu/kskipmrr.py:45-48, v3/cpu/kskipcg.py:36-39 for operators of small
 * bandwidth (square CSR block, rows of <= 28 nonzeros, half bandwidth bw with W - 2(k-1)bw >= W/2 for the window W of
 * the kernel that applies: 640 rows in general, 768 when every row holds its full band of <= 27 diagonals);
 * PK_ERR_UNSUPPORTED otherwise (the solvers then use k two-chain SpMV passes). */
int pk_matpow(pk_ctx* ctx, pk_mat* mat, int k, double* d_base0, double* d_base1);
/* Which matrix-powers kernel pk_matpow / the k-skip solvers would run for this operator and k: *kind = 0 none (two-chain
 * SpMV passes), 1 general small-bandwidth kernel (one row per thread), 2 dense-band kernel (two rows per thread);
 * *window_rows = rows per block before the trapezoid overlap is taken off (blocking on first use: probes the structure). */
int pk_mat_matpow_info(pk_ctx* ctx, pk_mat* mat, int k, int* kind, int* window_rows);
/* Row-partitioned matrix powers: register copies of the neighbours' rows of A next to this block — the last `rows_above`
 * rows of the previous rank and the first `rows_below` rows of the next one, as CSR with GLOBAL int32 column indices
 * (device arrays, borrowed) — so that the basis of a trip needs ONE exchange of depth rows + half_bw instead of one per
 * level (the redundant ghost-zone scheme; replaces the 2k MultiGpu.dot exchanges of v3/gpu/mpi/kskipmrr.py:55-58).
 * row0 = global index of the first owned row, d_halo_global[h] = global index of local column n_rows + h (pk_mat_set_halo
 * order).  Every rank must own at least rows + half_bw rows; half_bw / max_row_nnz are the GLOBAL maxima. */
int pk_mat_set_matpow_ext(pk_mat* mat, int half_bw, int max_row_nnz, int64_t row0, int64_t n_global,
                          const int64_t* d_halo_global, int64_t rows_above, const int32_t* d_rowptr_above,
                          const int32_t* d_col_above, const double* d_val_above, int64_t rows_below,
                          const int32_t* d_rowptr_below, const int32_t* d_col_below, const double* d_val_below);
/* Row-partitioned DENSE band: `ext` is a single-GPU CSR operator of the same context over this rank's
 * [rows_above ghost rows | owned rows | rows_below ghost rows] (copies of the neighbours' boundary rows of A; columns
 * renumbered from the first of those rows, entries beyond the two ends dropped; both counts even; borrowed, must outlive
 * `mat`).  pk_solve(PK_KSKIPMRR) then runs each trip on it with the dense-band kernels and ONE ghost-zone exchange of depth
 * (k+1) bw per trip; the work area needs no more room than pk_work_doubles says, but pk_mat_ld(mat) grows to hold the ghost
 * zones.  ext = NULL detaches.  PK_ERR_UNSUPPORTED if `ext` is not a dense band of <= 27 diagonals. */
int pk_mat_set_band_ext(pk_mat* mat, pk_mat* ext, int64_t rows_above, int64_t rows_below);
/* Longest row and half bandwidth max |global column - global row| of a CSR block given as raw device arrays
 * (h_out[0], h_out[1]; blocking) — col holds GLOBAL indices, row0 is the global index of the block's first row. */
int pk_csr_band_info(pk_ctx* ctx, int64_t n_rows, int64_t row0, const int32_t* d_rowptr, const int32_t* d_col, int* h_out);
/* d_out[0] = u·v (local part; all-reduced when the context has a communicator). */
int pk_dot(pk_ctx* ctx, int64_t n, const double* d_u, const double* d_v, double* d_out);
/* All Gram sums of one k-skip outer trip in one pass (replaces the 6k+5 / 6k+7 cupy.dot calls,
 * v3/gpu/kskipmrr.py:53-61, v3/gpu/kskipcg.py:44-52).  U has nu rows, V has nv rows, row stride ld.
 * mode 0 (MrR): U=Ar, V=Ay; mode 1 (CG): U=Ar, V=Ap.  d_g[6*jj + t], t = {U[jj]·U[jj], U[jj]·U[jj+1],
 * U[jj]·V[jj], (mode0: V[jj]·U[jj+1] | mode1: U[jj]·V[jj+1]), V[jj]·V[jj], V[jj]·V[jj+1]}; absent rows give 0. */
int pk_gram(pk_ctx* ctx, int mode, int64_t n, int64_t ld, const double* d_u, int nu, const double* d_v, int nv,
            double* d_g);

/* ------------------------------------------------------------------------------------------------------------ */
/* solvers: the whole v3 loop, device resident (scalars, history and the convergence flag never leave the GPU;
 * the host polls the flag once per `check_every` iterations/trips).                                            */
typedef struct {
    int64_t maxiter;       /* iteration cap; reference default = N (v3/gpu/common.py:35-36)                     */
    double tol;            /* relative: ||r||/||b|| < tol, strict (v3/gpu/cg.py:26)                              */
    int32_t k;             /* k-skip depth, 0..PK_KMAX                                                           */
    int32_t check_every;   /* host polls the device flag every this many iterations (cg/mrr) or trips; <=0 auto */
    int32_t use_graph;     /* 1: replay a captured CUDA graph per batch; 0: plain stream launches                */
    int32_t x_is_zero;     /* 1: x0 == 0, skip the initial A·x (result identical: b - A·0 == b)                   */
    int64_t global_n;      /* global number of rows (== n_rows when not distributed)                             */
    const double* d_mdiag; /* PK_CGCG only: diagonal of the Jacobi preconditioner M (local rows), u = r / M; NULL = none  */
    int32_t basis;         /* PK_KSKIPCG / PK_KSKIPMRR: 0 = the reference's monomial basis A^j r (default, parity path);
                              1 = Chebyshev basis T_j((A - d)/c) r on [lam_lo, lam_hi] (opt-in)         */
    int32_t pad0;
    double lam_lo, lam_hi; /* basis = 1: bounds of the spectrum of A (pk_mat_gershgorin gives rigorous ones)                */
} pk_solve_opts;

typedef struct {
    int64_t iterations;    /* number of solution updates (`i` of the reference loop)                            */
    int64_t entries;       /* valid entries in residual[]/nosl[]/khistory[]  (index+1)                           */
    int32_t converged;     /* isConverged                                                                        */
    int32_t final_k;       /* adaptive: k at exit                                                                */
    double final_residual; /* residual[entries-1]                                                                */
    double elapsed_s;      /* loop time by CUDA events (the reference's `info['time']` placement, plus a sync)   */
    int64_t kernel_launches; /* kernels this solve launched (bench.py's gpu_launches)                           */
    int64_t spmv_count;    /* operator applications                                                              */
} pk_solve_result;

/* d_out[i] = A[i][i] for the local rows (0 where the row stores no diagonal entry): the Jacobi preconditioner of PK_CGCG. */
int pk_mat_diagonal(pk_mat* mat, double* d_out);

/* Gershgorin bounds of the spectrum from the local rows: h_out[0] = min_i (a_ii - sum_{j != i} |a_ij|),
 * h_out[1] = max_i (a_ii + sum_{j != i} |a_ij|) (host doubles; blocking).  Distributed: min / max over the ranks' results. */
int pk_mat_gershgorin(pk_mat* mat, double* h_out);

/* Number of doubles of scratch (d_work) the solver needs for vectors of padded length `ld`. */
int64_t pk_work_doubles(int method, int64_t ld, int k);
/* Padded vector length for a block (n_rows + n_halo, rounded up to 32 doubles). */
int64_t pk_mat_ld(pk_mat* mat);

/* d_b: right-hand side (local rows); d_x: initial guess in, solution out (local rows);
 * d_residual: double[hist_len]; d_nosl, d_khistory: int64[hist_len] (khistory may be NULL except adaptive);
 * hist_len >= maxiter + 2.  Blocks until the solve has finished (it must read the stop flag). */
int pk_solve(pk_ctx* ctx, int method, pk_mat* mat, const double* d_b, double* d_x, double* d_work,
             double* d_residual, int64_t* d_nosl, int64_t* d_khistory, int64_t hist_len,
             const pk_solve_opts* opts, pk_solve_result* result);

/* ------------------------------------------------------------------------------------------------------------ */
/* synthetic inputs generated in HBM: same matrices as parallel_krylov_b200/problems.py          */
/* Box stencil (2·dims+1 points; nz == 1 gives the 2-D 5-point operator, diag 4, else diag 6), rows
 * [row0, row0+n_rows) of the global grid nx·ny·nz.  Two steps: per-row counts, caller scans them into the row
 * pointer (relative to row0), then the fill.  Column indices are GLOBAL (int32: grids of < 2^31 points). */
int pk_gen_stencil_counts(pk_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_rows,
                          int32_t* d_counts);
int pk_gen_stencil_fill(pk_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t n_rows,
                        const int32_t* d_rowptr, int32_t* d_col, double* d_val);
/* Symmetric banded SPD matrix (2·half_bw+1 diagonals), identical to problems.banded_spd(n, half_bw, seed). */
int pk_gen_banded_counts(pk_ctx* ctx, int64_t n, int half_bw, int64_t row0, int64_t n_rows, int32_t* d_counts);
int pk_gen_banded_fill(pk_ctx* ctx, int64_t n, int half_bw, uint64_t seed, int64_t row0, int64_t n_rows,
                       const int32_t* d_rowptr, int32_t* d_col, double* d_val);
int pk_fill_hash_normal(pk_ctx* ctx, uint64_t seed, int64_t offset, int64_t n, double* d_out);

#ifdef __cplusplus
}
#endif
#endif /* PKRYLOV_H */

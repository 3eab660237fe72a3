// C-ABI entry points and the host side of the five solver loops.  The host only ENQUEUES: every scalar, the residual
// history and the stopping flag live on the device (PkState); the host polls the flag once per batch of iterations,
// one batch behind the GPU, so the device never idles waiting for the host (the reference syncs every iteration:
// `if residual[i] < tol` on a cupy scalar, /root/reference/v3/gpu/cg.py:26).  Kernels enqueued after the stopping
// rule fired are no-ops, so x, the history and the iteration count are exactly those of the reference loop.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#include "pk_common.cuh"
#include "pk_launch.h"

// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void pk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pk_last_error(void) { return g_err; }
extern "C" int pk_version(void) { return PK_VERSION; }

// ---------------------------------------------------------------------------------------------------------------
static int ctx_init(pk_ctx* c, int device, void* stream) {
    c->device = device;
    c->stream = (cudaStream_t)stream;
    if (c->stream == nullptr) {
        // A *blocking* stream: it orders itself against the legacy default stream (where torch runs by default), and,
        // unlike the legacy stream, it can be captured into a CUDA graph.
        PK_CUDA(cudaStreamCreate(&c->stream));
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    PK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        pk_set_error("libpkrylov is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
        return PK_ERR_UNSUPPORTED;
    }
    c->sm_count = prop.multiProcessorCount;
    c->red.max_blocks = c->sm_count * 16;
    PK_CUDA(cudaMalloc(&c->red.partials, sizeof(double) * (size_t)PK_MAX_SUMS * c->red.max_blocks));
    PK_CUDA(cudaMalloc(&c->red.ticket, sizeof(unsigned int)));
    PK_CUDA(cudaMemset(c->red.ticket, 0, sizeof(unsigned int)));
    PK_CUDA(cudaMalloc(&c->d_state, sizeof(PkState)));
    PK_CUDA(cudaMemset(c->d_state, 0, sizeof(PkState)));
    PK_CUDA(cudaMallocHost(&c->h_state, sizeof(PkState)));
    PK_CUDA(cudaMallocHost(&c->h_flags, sizeof(int) * 8));
    PK_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    PK_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    PK_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    PK_CUDA(cudaEventCreate(&c->ev_t0));
    PK_CUDA(cudaEventCreate(&c->ev_t1));
    PK_CUDA(cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming));
    PK_CUDA(cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming));
    return PK_OK;
}

extern "C" int pk_ctx_create(pk_ctx** out, int device, void* stream) {
    PK_REQUIRE(out != nullptr, "null out");
    PK_CUDA(cudaSetDevice(device));
    pk_ctx* c = new pk_ctx();
    const int rc = ctx_init(c, device, stream);
    if (rc != PK_OK) {
        pk_ctx_destroy(c);          // releases whatever the partial initialisation had already acquired
        return rc;
    }
    *out = c;
    return PK_OK;
}

extern "C" int pk_ctx_destroy(pk_ctx* c) {
    if (!c) return PK_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    pk_comm_destroy(c);
    if (c->red.partials) cudaFree(c->red.partials);
    if (c->red.ticket) cudaFree(c->red.ticket);
    if (c->d_state) cudaFree(c->d_state);
    if (c->my_mbox) cudaFree(c->my_mbox);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    for (cudaEvent_t e : {c->ev_a, c->ev_b, c->ev_t0, c->ev_t1, c->ev_poll[0], c->ev_poll[1]})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    cudaGetLastError();
    delete c;
    return PK_OK;
}

extern "C" int pk_ctx_sync(pk_ctx* c) {
    PK_REQUIRE(c != nullptr, "null context");
    PK_CUDA(cudaStreamSynchronize(c->stream));
    return PK_OK;
}

extern "C" int pk_ctx_sm_count(pk_ctx* c) { return c ? c->sm_count : 0; }

// Per-launch timing of the operator kernel: CUDA event pairs on the launching stream (not under a profiler).
extern "C" int pk_prof_begin(pk_ctx* c, int max_launches) {
    PK_REQUIRE(c != nullptr && max_launches > 0, "bad argument");
    PK_CUDA(cudaSetDevice(c->device));
    while ((int)c->prof_ev.size() < 2 * max_launches) {
        cudaEvent_t e;
        PK_CUDA(cudaEventCreate(&e));
        c->prof_ev.push_back(e);
    }
    c->prof_used = 0;
    c->prof_on = true;
    return PK_OK;
}

extern "C" int pk_prof_end(pk_ctx* c, double* total_ms, int64_t* n_launches) {
    PK_REQUIRE(c != nullptr, "null context");
    PK_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
        float ms = 0.f;
        PK_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[i], c->prof_ev[i + 1]));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = (int64_t)(c->prof_used / 2);
    c->prof_on = false;
    c->prof_used = 0;
    return PK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
static long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

static int csr_check(pk_ctx* ctx, const void* rowptr, int rowptr64, const int32_t* col, long long n_rows,
                     long long n_cols, long long nnz) {
    const char* e = getenv("PK_VALIDATE");        // PK_VALIDATE=0 skips the one-time structural check
    if (n_rows <= 0 || (e && atoi(e) == 0)) return PK_OK;
    int bad = 0;
    PK_CHECK(pk_csr_validate(ctx, rowptr, rowptr64, col, n_rows, n_cols, nnz, &bad));
    if (bad) {
        pk_set_error("malformed CSR block:%s%s%s%s", (bad & 1) ? " rowptr[0] != 0;" : "",
                     (bad & 2) ? " rowptr not monotone;" : "", (bad & 4) ? " rowptr[n_rows] != nnz;" : "",
                     (bad & 8) ? " column index outside [0, n_cols);" : "");
        return PK_ERR_ARG;
    }
    return PK_OK;
}

// Kernel choice from the nnz distribution (mean row length): tiles of 256 rows (128 for mid-length rows) staged in
// shared memory when a tile's nonzeros fit; otherwise the warp-per-row path handles the tile.
static int csr_tile_rows(long long n_rows, long long nnz) {
    const double mean = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
    return mean <= 12.0 ? 256 : 128;
}

static void csr_choose_kernel(pk_mat* m, int tile_max) {
    const double mean = m->n_rows > 0 ? (double)m->nnz / (double)m->n_rows : 0.0;
    long long heur = (long long)(m->tile_rows * mean * 1.25) + 64;      // room for moderately irregular rows
    long long cap = std::min<long long>(heur, (long long)tile_max + 3);  // +3: the staged window starts 16-byte aligned
    cap = std::max<long long>(round_up(cap, 64), 256);
    if (mean <= 40.0 && cap * 12 <= 96 * 1024) {
        m->tile_cap = (int)cap;
        m->kind = MAT_CSR_STREAM;
    } else {
        m->tile_cap = 256;        // practically every tile exceeds this: rows are reduced warp-per-row
        m->kind = MAT_CSR_VECTOR;
    }
    const char* e = getenv("PK_SPMV");            // "stream": plain-load kernel; default: TMA-pipelined kernel
    m->use_tma = m->vec_ok && !(e && strcmp(e, "stream") == 0);
    const char* st = getenv("PK_SPMV_STAGES");
    m->stages = st ? atoi(st) : 2;
    if (m->stages < 2 || m->stages > 4) m->stages = 2;
    m->ld = round_up(std::max<long long>(m->n_rows, m->n_cols), 32);
}

extern "C" int pk_mat_csr(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols_local, int64_t nnz,
                          const int32_t* d_rowptr, const int32_t* d_col, const double* d_val) {
    PK_REQUIRE(ctx && out, "null argument");
    PK_REQUIRE(n_rows >= 0 && nnz >= 0 && nnz < (1LL << 31), "pk_mat_csr takes 0 <= nnz < 2^31 (use pk_mat_csr64 beyond)");
    PK_REQUIRE(n_rows == 0 || (d_rowptr && (nnz == 0 || (d_col && d_val))), "null CSR arrays");
    PK_REQUIRE(n_cols_local >= 0 && n_cols_local < (1LL << 31), "column space of a block must be < 2^31 (int32 indices)");
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(csr_check(ctx, d_rowptr, 0, d_col, n_rows, n_cols_local, nnz));
    pk_mat* m = new pk_mat();
    m->ctx = ctx;
    m->n_rows = n_rows;
    m->n_cols = n_cols_local;
    m->nnz = nnz;
    m->rowptr = d_rowptr;
    m->col = d_col;
    m->val = d_val;
    m->vec_ok = (((uintptr_t)d_col & 15) == 0) && (((uintptr_t)d_val & 15) == 0) && (((uintptr_t)d_rowptr & 15) == 0);
    m->tile_rows = csr_tile_rows(n_rows, nnz);
    int tile_max = 0;
    if (n_rows > 0) {
        int rc = pk_tile_max_nnz(ctx, d_rowptr, n_rows, m->tile_rows, &tile_max);
        if (rc != PK_OK) { delete m; return rc; }
    }
    csr_choose_kernel(m, tile_max);
    *out = m;
    return PK_OK;
}

static void free_segs(pk_mat* m) {
    for (PkSeg& sg : m->segs)
        if (sg.rp32) cudaFree(sg.rp32);
    m->segs.clear();
}

extern "C" int pk_mat_csr64(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols_local, int64_t nnz,
                            const int64_t* d_rowptr, const int32_t* d_col, const double* d_val) {
    PK_REQUIRE(ctx && out, "null argument");
    PK_REQUIRE(n_rows >= 0 && nnz >= 0, "negative size");
    PK_REQUIRE(n_rows == 0 || (d_rowptr && (nnz == 0 || (d_col && d_val))), "null CSR arrays");
    PK_REQUIRE(n_cols_local >= 0 && n_cols_local < (1LL << 31), "column space of a block must be < 2^31 (int32 indices)");
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(csr_check(ctx, d_rowptr, 1, d_col, n_rows, n_cols_local, nnz));
    pk_mat* m = new pk_mat();
    m->ctx = ctx;
    m->n_rows = n_rows;
    m->n_cols = n_cols_local;
    m->nnz = nnz;
    m->col = d_col;
    m->val = d_val;
    m->vec_ok = (((uintptr_t)d_col & 15) == 0) && (((uintptr_t)d_val & 15) == 0);
    m->tile_rows = csr_tile_rows(n_rows, nnz);
    // segment size: below 2^31 with room for rounding the cut to a tile boundary (PK_SEG_NNZ: tests use tiny segments)
    long long seg_max = (1LL << 31) - (1LL << 24);
    if (const char* e = getenv("PK_SEG_NNZ")) seg_max = std::max<long long>(atoll(e), 64);
    const long long n_seg = std::max<long long>(1, (nnz + seg_max - 1) / seg_max);
    auto rp_at = [&](long long r, long long* v) -> int {      // one 8-byte D2H read (a handful per cut)
        PK_CUDA(cudaMemcpyAsync(v, d_rowptr + r, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        return PK_OK;
    };
    std::vector<long long> cuts{0};
    for (long long sgi = 1; sgi < n_seg; ++sgi) {
        const long long target = (long long)((double)nnz * (double)sgi / (double)n_seg);
        long long lo = cuts.back(), hi = n_rows;              // first row r with rowptr[r] >= target
        while (lo < hi) {
            const long long mid = (lo + hi) / 2;
            long long v = 0;
            int rc = rp_at(mid, &v);
            if (rc != PK_OK) { delete m; return rc; }
            if (v < target) lo = mid + 1; else hi = mid;
        }
        const long long cut = lo / m->tile_rows * m->tile_rows;   // tile boundary (keeps the staged windows 16-byte aligned)
        if (cut > cuts.back() && cut < n_rows) cuts.push_back(cut);
    }
    cuts.push_back(n_rows);
    int tile_max = 0;
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
        PkSeg sg;
        sg.row_lo = cuts[i];
        sg.row_hi = cuts[i + 1];
        long long b0 = 0, b1 = 0;
        int rc = rp_at(sg.row_lo, &b0);
        if (rc == PK_OK) rc = rp_at(sg.row_hi, &b1);
        if (rc == PK_OK && b1 - b0 >= (1LL << 31)) {
            pk_set_error("rows %lld..%lld hold %lld nonzeros: a segment of < 2^31 cannot be cut at a tile boundary",
                         sg.row_lo, sg.row_hi, b1 - b0);
            rc = PK_ERR_UNSUPPORTED;
        }
        if (rc == PK_OK && cudaMalloc(&sg.rp32, sizeof(int32_t) * (size_t)(sg.row_hi - sg.row_lo + 4)) != cudaSuccess) {
            pk_set_error("cudaMalloc of a segment row pointer failed");
            rc = PK_ERR_CUDA;
        }
        if (rc == PK_OK) {
            sg.base = b0;
            sg.nnz = b1 - b0;
            rc = pk_rebase_rowptr(ctx, d_rowptr + sg.row_lo, b0, sg.row_hi - sg.row_lo + 1, sg.rp32);
        }
        int tmax = 0;
        if (rc == PK_OK && sg.row_hi > sg.row_lo) rc = pk_tile_max_nnz(ctx, sg.rp32, sg.row_hi - sg.row_lo, m->tile_rows, &tmax);
        if (rc != PK_OK) {
            if (sg.rp32) cudaFree(sg.rp32);
            free_segs(m);
            delete m;
            return rc;
        }
        tile_max = std::max(tile_max, tmax);
        m->segs.push_back(sg);
    }
    csr_choose_kernel(m, tile_max);
    *out = m;
    return PK_OK;
}

extern "C" int pk_mat_dense(pk_ctx* ctx, pk_mat** out, int64_t n_rows, int64_t n_cols, const double* d_a, int64_t lda) {
    PK_REQUIRE(ctx && out && d_a, "null argument");
    PK_REQUIRE(lda >= n_cols, "lda < n_cols");
    pk_mat* m = new pk_mat();
    m->ctx = ctx;
    m->kind = MAT_DENSE;
    m->n_rows = n_rows;
    m->n_cols = n_cols;
    m->dense = d_a;
    m->lda = lda;
    m->nnz = n_rows * n_cols;
    m->ld = round_up(std::max<long long>(n_rows, n_cols), 32);
    *out = m;
    return PK_OK;
}

extern "C" int pk_mat_destroy(pk_mat* m) {
    if (!m) return PK_OK;
    if (m->d_sendbuf) cudaFree(m->d_sendbuf);
    pk_mat_halo_p2p_close(m);
    free_segs(m);
    if (m->mp_gin) cudaFree(m->mp_gin);
    delete m;
    return PK_OK;
}

extern "C" int pk_mat_kernel_info(pk_mat* m, int* kind, int* tile_rows, int* tile_cap) {
    PK_REQUIRE(m != nullptr, "null operator");
    if (kind) *kind = m->kind;
    if (tile_rows) *tile_rows = m->tile_rows;
    if (tile_cap) *tile_cap = m->tile_cap;
    return PK_OK;
}

extern "C" int64_t pk_mat_ld(pk_mat* m) { return m ? m->ld : 0; }

bool pk_mat_can_fuse(const pk_mat* m) {
    const char* e = getenv("PK_FUSE");            // PK_FUSE=0: keep update and SpMV as separate kernels (A/B, tests)
    const int allow = e ? atoi(e) : 1;
    return allow && m->kind != MAT_DENSE && m->use_tma;
}

extern "C" int pk_mat_set_halo(pk_mat* m, int n_peers_total, const int64_t* send_off, const int64_t* recv_off,
                               const int32_t* d_send_idx, const int32_t* h_send_idx, int64_t interior_lo,
                               int64_t interior_hi) {
    PK_REQUIRE(m && send_off && recv_off, "null argument");
    PK_REQUIRE(n_peers_total == m->ctx->n_ranks, "halo plan size != communicator size");
    PK_REQUIRE(m->segs.empty() || (recv_off[n_peers_total] == 0 && send_off[n_peers_total] == 0),
               "a distributed block that exchanges a halo must have nnz < 2^31 (shard it over more ranks)");
    const int P = n_peers_total;
    m->send_off.assign(send_off, send_off + P + 1);
    m->recv_off.assign(recv_off, recv_off + P + 1);
    m->n_halo = recv_off[P];
    PK_REQUIRE(m->kind == MAT_DENSE || m->n_rows + m->n_halo <= m->n_cols, "halo larger than the column space of the block");
    for (int p = 0; p < P; ++p)
        PK_REQUIRE(send_off[p + 1] >= send_off[p] && recv_off[p + 1] >= recv_off[p], "offsets must be non-decreasing");
    PK_REQUIRE(send_off[P] == 0 || (d_send_idx != nullptr && h_send_idx != nullptr), "null send index list");
    m->d_send_idx = d_send_idx;
    m->send_contig.assign(P, 0);
    m->send_first.assign(P, 0);
    for (int p = 0; p < P; ++p) {
        const long long a = send_off[p], b = send_off[p + 1];
        if (b <= a) { m->send_contig[p] = 1; continue; }
        bool contig = true;
        for (long long i = a + 1; i < b && contig; ++i) contig = (h_send_idx[i] == h_send_idx[i - 1] + 1);
        m->send_contig[p] = contig ? 1 : 0;
        m->send_first[p] = h_send_idx[a];
    }
    if (m->d_sendbuf) { cudaFree(m->d_sendbuf); m->d_sendbuf = nullptr; }
    if (send_off[P] > 0) PK_CUDA(cudaMalloc(&m->d_sendbuf, sizeof(double) * 2 * (size_t)send_off[P]));
    m->interior_lo = std::max<long long>(0, std::min<long long>(interior_lo, m->n_rows));
    m->interior_hi = std::max<long long>(m->interior_lo, std::min<long long>(interior_hi, m->n_rows));
    m->distributed = true;
    m->ld = round_up(std::max<long long>(m->n_rows + m->n_halo, m->ld), 32);
    return PK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// building blocks exposed for op-by-op parity tests
static int state_for_standalone(pk_ctx* ctx) {
    // stand-alone kernels must not be skipped by a stale `done` flag of a previous solve
    PK_CUDA(cudaMemsetAsync(&ctx->d_state->done, 0, sizeof(int), ctx->stream));
    return PK_OK;
}

extern "C" int pk_spmv(pk_ctx* ctx, pk_mat* mat, double* d_x, double* d_y, double* d_x1, double* d_y1,
                       const double* d_w, double* d_sums) {
    PK_REQUIRE(ctx && mat && d_x && d_y, "null argument");
    PK_REQUIRE((d_x1 == nullptr) == (d_y1 == nullptr), "x1/y1 must both be given or both be null");
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(state_for_standalone(ctx));
    PkDots dots;
    dots.w = d_w;
    dots.epi = EPI_NONE;
    PK_CHECK(pk_launch_spmv(ctx, mat, d_x, d_y, d_x1, d_y1, dots));
    if (d_w && d_sums)
        PK_CUDA(cudaMemcpyAsync(d_sums, ctx->d_state->red, 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return PK_OK;
}

extern "C" int pk_dot(pk_ctx* ctx, int64_t n, const double* d_u, const double* d_v, double* d_out) {
    PK_REQUIRE(ctx && d_u && d_v && d_out, "null argument");
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(pk_launch_dot(ctx, n, d_u, d_v, EPI_NONE, 1));
    PK_CUDA(cudaMemcpyAsync(d_out, ctx->d_state->red, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return PK_OK;
}

extern "C" int pk_gram(pk_ctx* ctx, int mode, int64_t n, int64_t ld, const double* d_u, int nu, const double* d_v,
                       int nv, double* d_g) {
    PK_REQUIRE(ctx && d_u && d_v && d_g, "null argument");
    PK_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (MrR) or 1 (CG)");
    const int njj = std::max(nu, nv);
    PK_REQUIRE(njj >= 1 && njj <= PK_KMAX + 2, "too many basis rows");
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(state_for_standalone(ctx));
    PK_CHECK(pk_launch_gram(ctx, mode, n, ld, d_u, nu, d_v, nv, njj, EPI_NONE));
    PK_CUDA(cudaMemcpyAsync(d_g, ctx->d_state->gram, sizeof(double) * 6 * njj, cudaMemcpyDeviceToDevice, ctx->stream));
    return PK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" int64_t pk_work_doubles(int method, int64_t ld, int k) {
    long long nvec = 0;
    switch (method) {
        case PK_CG: nvec = 3; break;                       // r, p, v
        case PK_MRR: nvec = 4; break;                      // r, Ar, y, z
        case PK_KSKIPCG: nvec = (k + 1) + (k + 2) + 2; break;  // Ar[0..k], Ap[0..k+1], spare Ap0 (fused steps), A p (Chebyshev basis)
        case PK_KSKIPMRR: nvec = (k + 2) + (k + 1) + 3; break;          // Ar[0..k+1], Ay[0..k], z, spare Ar0, A r (Chebyshev basis)
        case PK_ADAPTIVEKSKIPMRR: nvec = (k + 2) + (k + 1) + 3; break;  // + best_x
        case PK_CGCG: nvec = 5; break;                     // r, w, p, s, u (u aliases r without a preconditioner)
        default: return -1;
    }
    return nvec * ld;
}

namespace {

struct Solve {
    pk_ctx* ctx;
    pk_mat* A;
    const double* b;
    double* x;
    double* work;
    long long n, ld;
    pk_solve_opts o;
    int method;
    // row-partitioned dense band (pk_matpow.cu): the trip runs on the extended single-GPU operator; every vector of the
    // solve then has ext_off entries of ghost zone in front of its owned part (and as many behind)
    pk_mat* ext = nullptr;
    long long ext_off = 0, ext_depth = 0;
    // graph replay
    cudaGraphExec_t gexec = nullptr;
    long long graph_launches = 0, graph_spmvs = 0;

    double* vec(int i) const { return work + (size_t)i * ld; }

    int apply(double* xin, double* yout, const double* w = nullptr, int epi = EPI_NONE) {
        PkDots d;
        d.w = w;
        d.epi = epi;
        return pk_launch_spmv(ctx, A, xin, yout, nullptr, nullptr, d);
    }
    int apply2(double* x0, double* y0, double* x1, double* y1) {
        PkDots d;
        return pk_launch_spmv(ctx, A, x0, y0, x1, y1, d);
    }

    int poll(int slot) {
        PK_CUDA(cudaMemcpyAsync(ctx->h_flags + 4 * slot, &ctx->d_state->done, 4 * sizeof(int), cudaMemcpyDeviceToHost,
                                ctx->stream));
        PK_CUDA(cudaEventRecord(ctx->ev_poll[slot], ctx->stream));
        return PK_OK;
    }
    int wait(int slot, bool* done) {
        PK_CUDA(cudaEventSynchronize(ctx->ev_poll[slot]));
        *done = ctx->h_flags[4 * slot] != 0;
        return PK_OK;
    }
    int fetch_state() {   // blocking read of the whole device state
        PK_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(PkState), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        return PK_OK;
    }

    // Enqueue `body` batches until the device reports the stop flag.  unit = solver iterations one body() adds.
    template <class Body>
    int run_batches(long long unit, long long already, Body body, long long force_batch = 0) {
        long long batch = force_batch > 0 ? force_batch : o.check_every;
        if (batch <= 0) {
            // aim for >= ~0.5 ms of device work between polls (rough: 200 bytes per row per iteration at 6 TB/s).
            // The batch must be the SAME on every rank (each rank enqueues whole batches, and the host-enqueued
            // collectives of a batch need a matching peer), so it is derived from the global size, never from the
            // local row count: uneven blocks would otherwise give different batches and hang the NCCL paths.
            const double rows = (ctx->n_ranks > 1 && o.global_n > 0) ? (double)o.global_n / (double)ctx->n_ranks
                                                                      : (double)n;
            double per_it_us = std::max(8.0, rows * 200.0 / 6.0e6);
            batch = (long long)std::ceil(500.0 / (per_it_us * (double)unit));
            batch = std::max<long long>(1, std::min<long long>(batch, 64));
        }
        const bool graph = o.use_graph != 0;
        if (graph && !gexec) {
            // capture one batch (stream capture sees the side-stream halo exchange through the fork/join events)
            cudaGraph_t g = nullptr;
            const long long l0 = ctx->launches, s0 = ctx->spmvs;
            PK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            int rc = PK_OK;
            for (long long i = 0; i < batch && rc == PK_OK; ++i) rc = body();
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
            if (rc != PK_OK || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                if (rc != PK_OK) return rc;
            }
            PK_CUDA(ce);
            graph_launches = ctx->launches - l0;
            graph_spmvs = ctx->spmvs - s0;
            ctx->launches = l0;
            ctx->spmvs = s0;
            cudaError_t ie = cudaGraphInstantiate(&gexec, g, 0);
            cudaGraphDestroy(g);
            PK_CUDA(ie);
        }
        long long enq = already;
        int slot = 0, prev = -1;
        bool done = false;
        while (true) {
            if (graph) {
                PK_CUDA(cudaGraphLaunch(gexec, ctx->stream));
                ctx->launches += graph_launches;
                ctx->spmvs += graph_spmvs;
            } else {
                for (long long i = 0; i < batch; ++i) PK_CHECK(body());
            }
            enq += batch * unit;
            PK_CHECK(poll(slot));
            if (prev >= 0) {
                PK_CHECK(wait(prev, &done));
                if (done) break;
            }
            if (enq >= o.maxiter + unit) {   // the device has certainly hit the cap by the end of this batch
                PK_CHECK(wait(slot, &done));
                if (done) break;
            }
            prev = slot;
            slot ^= 1;
        }
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
        return PK_OK;
    }

    int init_state(double* d_res, int64_t* d_nosl, int64_t* d_khist, long long hist_len) {
        PkState* h = ctx->h_state;
        memset(h, 0, sizeof(PkState));
        h->maxiter = o.maxiter;
        h->tol = o.tol;
        h->res = d_res;
        h->nosl = (long long*)d_nosl;
        h->khist = (long long*)d_khist;
        h->k = o.k;
        h->hist_len = hist_len;
        PK_CUDA(cudaMemcpyAsync(ctx->d_state, h, sizeof(PkState), cudaMemcpyHostToDevice, ctx->stream));
        // ||b|| (global): v3/cpu/common.py:24
        PK_CHECK(pk_launch_dot(ctx, n, b, b, EPI_BNORM, 1));
        return PK_OK;
    }

    // r = b - A x  (v = scratch for A x);  p (nullable) = copy of r;  epilogue epi on r.r
    int initial_residual(double* r, double* p, double* scratch, int epi) {
        if (o.x_is_zero) return pk_launch_resid_init(ctx, n, b, nullptr, r, p, epi);
        PK_CHECK(apply(x, scratch));
        return pk_launch_resid_init(ctx, n, b, scratch, r, p, epi);
    }

    // ---- CG: /root/reference/v3/cpu/cg.py:7-48 -----------------------------------------------------------------
    int cg() {
        double *r = vec(0), *p = vec(1), *v = vec(2);
        PK_CHECK(initial_residual(r, p, v, EPI_CG_INIT));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        if (use_persistent()) {
            // small system: the loop runs inside one cooperative kernel, `chunk` iterations per launch
            const int chunk = 64;
            Solve* self = this;
            const bool graph_saved = o.use_graph != 0;
            o.use_graph = 0;                                       // cooperative launches are not captured
            int rc = run_batches(chunk, 0, [&]() -> int {
                return pk_launch_cg_persistent(self->ctx, self->A, self->x, r, p, v, chunk);
            }, 1);
            o.use_graph = graph_saved ? 1 : 0;
            return rc;
        }
        PK_CHECK(run_batches(1, 0, [&]() -> int {
            PK_CHECK(apply(p, v, p, EPI_CG_ALPHA));               // v = A p ; sigma = p.v ; alpha
            PK_CHECK(pk_launch_cg_xr_split(ctx, n, x, r, p, v));  // x += alpha p ; r -= alpha v ; gamma' ; beta ; test
            PK_CHECK(pk_launch_cg_p(ctx, n, p, r));               // p = r + beta p
            return PK_OK;
        }));
        return PK_OK;
    }

    // ---- Chronopoulos-Gear CG: /root/reference/v1/threads/pipeline/chronopoulos_gear.py:7-56 (with old_gamma kept up to
    // date).  Two kernels and ONE reduction point per iteration: the fused update leaves its local r.r / r.u in
    // PkState::red[3..4], the SpMV w = A u adds u.w and all-reduces the three together in its epilogue.
    int cgcg() {
        double *r = vec(0), *w = vec(1), *p = vec(2), *s = vec(3);
        const double* md = o.d_mdiag;
        double* u = md ? vec(4) : r;
        PK_CHECK(initial_residual(r, nullptr, w, EPI_RES0));                       // :22-23
        PK_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)ld * 2, ctx->stream));   // p, s (adjacent) = 0  :33-34
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        auto apply_u = [&](int epi) -> int {
            PkDots d;
            d.w = u;
            d.epi = epi;
            d.extra_sums = 2;
            return pk_launch_spmv(ctx, A, u, w, nullptr, nullptr, d);
        };
        PK_CHECK(pk_launch_cgcg_update(ctx, n, x, r, u, w, p, s, md, 1));          // :25  u = M^-1 r ; r.u
        PK_CHECK(apply_u(EPI_CGCG_INIT));                                          // :26-31
        PK_CHECK(run_batches(1, 0, [&]() -> int {
            PK_CHECK(pk_launch_cgcg_update(ctx, n, x, r, u, w, p, s, md, 0));      // :37-40, :45
            return apply_u(EPI_CGCG);                                              // :46-50, :41-42
        }));
        return PK_OK;
    }

    // one cooperative kernel for the whole CG loop: single GPU, CSR, small enough to be latency-bound
    bool use_persistent() const {
        const char* e = getenv("PK_PERSISTENT");                   // 0: never, 1: whenever possible, unset: by size
        const int mode = e ? atoi(e) : -1;
        if (mode == 0 || ctx->n_ranks > 1 || A->kind == MAT_DENSE || A->distributed || !A->segs.empty()) return false;
        if (mode == 1) return true;
        // measured on B200 (graph replay vs persistent): 2-D 256^2 (65 k rows) 62.7 k -> 109 k it/s; 128^3 (2.1 M rows)
        // 15.5 k vs 12.5 k it/s.  Below ~300 k rows the launches dominate, above the TMA-pipelined kernels win.
        return A->n_rows <= 300000 && A->nnz <= 8000000;
    }

    // ---- MrR: /root/reference/v3/cpu/mrr.py:7-61 -------------------------------------------------------------
    int mrr() {
        double *r = vec(0), *ar = vec(1), *y = vec(2), *z = vec(3);
        PK_CHECK(initial_residual(r, nullptr, ar, EPI_RES0));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(apply(r, ar, r, EPI_MRR_FIRST));                 // Ar ; zeta = (r.Ar)/(Ar.Ar)
        PK_CHECK(pk_launch_mrr_first(ctx, n, ar, r, x, y, z, EPI_KS_FIRST));
        PK_CHECK(run_batches(1, 1, [&]() -> int {
            PK_CHECK(apply(r, ar, y, EPI_MRR_GAMMA));             // Ar ; nu = y.Ar ; mu = y.y ; gamma
            PK_CHECK(pk_launch_mrr_s(ctx, n, ar, y, r));          // s = Ar - gamma y ; zeta, eta
            PK_CHECK(pk_launch_mrr_update(ctx, n, ar, y, z, r, r, nullptr, x, -1, EPI_MRR_STEP));
            return PK_OK;
        }));
        return PK_OK;
    }

    // ---- k-skip CG: /root/reference/v3/cpu/kskipcg.py:8-87 ---------------------------------------------------
    // ---- k-skip CG on a Chebyshev basis (opt-in; pk_scalars.h: pk_kskipcg_coef_cheb) — the mirror image of
    // kskipmrr_chebyshev(): U_j = T_j(Ah) r in the Ar rows, V_j = T_j(Ah) p in the Ap rows, A p in its own vector.
    int kskipcg_chebyshev() {
        const int k = o.k;
        PK_REQUIRE(o.lam_hi > o.lam_lo, "Chebyshev basis needs spectrum bounds lam_lo < lam_hi (pk_mat_gershgorin)");
        PK_REQUIRE(A->kind != MAT_DENSE && A->use_tma && !A->pat_on,
                   "the Chebyshev basis needs the TMA CSR kernel (16-byte aligned CSR arrays, no pattern compression)");
        const double c = 0.5 * (o.lam_hi - o.lam_lo), d = 0.5 * (o.lam_hi + o.lam_lo);
        auto U = [&](int j) { return vec(j); };                   // rows 0..k,   U(0) = r
        auto V = [&](int j) { return vec(k + 1 + j); };           // rows 0..k+1, V(0) = p
        double *spare = vec(2 * k + 3), *AP = vec(2 * k + 4);
        {
            PkState* h = ctx->h_state;
            h->cheb_c = c;
            h->cheb_d = d;
            PK_CUDA(cudaMemcpyAsync(&ctx->d_state->cheb_c, &h->cheb_c, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        PK_CHECK(initial_residual(U(0), V(0), AP, EPI_CG_INIT));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(apply(V(0), AP));                                // invariant: AP = A p at trip start
        PK_CHECK(run_batches(k + 1, 0, [&]() -> int {
            PK_CHECK(pk_launch_axpby(ctx, n, 1.0 / c, AP, -d / c, V(0), V(1)));        // V_1 = (A p - d p) / c
            for (int j = 1; j <= k; ++j) {                        // (U_j, V_{j+1}) from (U_{j-1}, V_j) and the levels before
                PkDots dd;
                dd.fuse = 3;
                if (j == 1) { dd.cs[0] = 1.0 / c; dd.cs[1] = -d / c; dd.cs[2] = 0.0; dd.f_a = nullptr; }
                else { dd.cs[0] = 2.0 / c; dd.cs[1] = -2.0 * d / c; dd.cs[2] = -1.0; dd.f_a = U(j - 2); }
                dd.cs[3] = 2.0 / c; dd.cs[4] = -2.0 * d / c; dd.cs[5] = -1.0;
                dd.f_b = V(j - 1);
                PK_CHECK(pk_launch_spmv(ctx, A, U(j - 1), U(j), V(j), V(j + 1), dd));
            }
            PK_CHECK(pk_launch_gram(ctx, 1, n, ld, U(0), k + 1, V(0), k + 2, k + 2, EPI_GRAM_CG_CHEB));
            double* cur = (k % 2 == 1) ? spare : V(0);
            PK_CHECK(pk_launch_kscg_update(ctx, n, x, U(0), V(0), cur, AP, 0, k == 0 ? EPI_KS_TRIP_END : EPI_KS_STEP));
            for (int j = 1; j <= k; ++j) {
                double* nxt = (cur == spare) ? V(0) : spare;
                PkDots ds;
                ds.epi = (j == k) ? EPI_KS_TRIP_END : EPI_KS_STEP;
                ds.fuse = 2; ds.cj = j; ds.f_a = U(0); ds.f_x = x; ds.f_out = nxt;
                PK_CHECK(pk_launch_spmv(ctx, A, cur, nullptr, nullptr, nullptr, ds));
                cur = nxt;
            }
            return apply(V(0), AP);                               // cur == V(0): A p for the next trip
        }));
        return PK_OK;
    }

    int kskipcg() {
        const int k = o.k;
        if (o.basis == 1) return kskipcg_chebyshev();
        auto Ar = [&](int j) { return vec(j); };                  // rows 0..k
        auto Ap = [&](int j) { return vec(k + 1 + j); };          // rows 0..k+1
        double* spare = vec(2 * k + 3);                           // second home of Ap[0] for the fused steps
        const bool fuse = pk_mat_can_fuse(A);
        const bool mp = pk_matpow_ok(ctx, A, k);
        PK_CHECK(initial_residual(Ar(0), Ap(0), Ap(1), EPI_CG_INIT));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(apply(Ap(0), Ap(1)));                            // invariant: Ap[1] = A Ap[0] at trip start
        PK_CHECK(run_batches(k + 1, 0, [&]() -> int {
            // basis: (Ar[j], Ap[j+1]) = A (Ar[j-1], Ap[j]), j = 1..k — all k levels of both chains in ONE pass over A when
            // the operator's bandwidth allows (pk_matpow.cu), else one pass over A per level
            if (mp) PK_CHECK(pk_launch_matpow(ctx, A, k, Ar(0), Ap(1), 0));
            else for (int j = 1; j <= k; ++j) PK_CHECK(apply2(Ar(j - 1), Ar(j), Ap(j), Ap(j + 1)));
            PK_CHECK(pk_launch_gram(ctx, 1, n, ld, Ar(0), k + 1, Ap(0), k + 2, k + 2, EPI_GRAM_CG));
            if (!fuse) {
                for (int j = 0; j <= k; ++j) {
                    PK_CHECK(pk_launch_kscg_update(ctx, n, x, Ar(0), Ap(0), Ap(0), Ap(1), j,
                                                   j == k ? EPI_KS_TRIP_END : EPI_KS_STEP));
                    PK_CHECK(apply(Ap(0), Ap(1)));
                }
                return PK_OK;
            }
            // All (alpha_j, beta_j) are known after the Gram kernel, so step j+1 rides in the epilogue of the SpMV that
            // follows step j: Ap1 = A Ap0 is consumed in registers and never stored.  Ap0 ping-pongs between its home
            // and `spare` (other rows still gather the old Ap0); step 0 picks the side so that it ends at home.
            double* cur = (k % 2 == 1) ? spare : Ap(0);
            PK_CHECK(pk_launch_kscg_update(ctx, n, x, Ar(0), Ap(0), cur, Ap(1), 0, k == 0 ? EPI_KS_TRIP_END : EPI_KS_STEP));
            for (int j = 1; j <= k; ++j) {
                double* nxt = (cur == spare) ? Ap(0) : spare;
                PkDots d;
                d.epi = (j == k) ? EPI_KS_TRIP_END : EPI_KS_STEP;
                d.fuse = 2; d.cj = j; d.f_a = Ar(0); d.f_x = x; d.f_out = nxt;
                PK_CHECK(pk_launch_spmv(ctx, A, cur, nullptr, nullptr, nullptr, d));
                cur = nxt;
            }
            return apply(Ap(0), Ap(1));                           // cur == Ap(0) here
        }));
        return PK_OK;
    }

    // ---- k-skip MrR: /root/reference/v3/cpu/kskipmrr.py:8-108 ------------------------------------------------
    int kskipmrr_open(double* Ar0, double* Ar1, double* Ay0, double* z, int first_epi) {
        PK_CHECK(apply(Ar0, Ar1, Ar0, EPI_MRR_FIRST));            // kskipmrr.py:26-27
        PK_CHECK(pk_launch_mrr_first(ctx, n, Ar1, Ar0, x, Ay0, z, first_epi));      // :28-34
        return apply(Ar0, Ar1);                                   // invariant: Ar[1] = A Ar[0] at trip start
    }
    // Launch-time predicates copied into every kernel launched inside the scope (pk_device.cuh: pk_skip).
    struct Ctl {
        pk_ctx* c;
        Ctl(pk_ctx* c_, int only_rollback, int dyn_cj, int dyn_last) : c(c_) {
            c->ctl_only_rollback = only_rollback;
            c->ctl_dyn_cj = dyn_cj;
            c->ctl_dyn_last = dyn_last;
        }
        ~Ctl() {
            c->ctl_only_rollback = 0;
            c->ctl_dyn_cj = -1;
            c->ctl_dyn_last = 0;
        }
    };

    // One outer trip.  dyn = false: k is known to the host (kskipmrr).  dyn = true (adaptivekskipmrr): the sequence is
    // enqueued for k = k_alloc, the CURRENT k lives in PkState: levels / steps beyond it are skipped by the kernels
    // themselves, the step with cj == k reduces r.r and runs the trip-end epilogue, and the Ar0 ping-pong picks its
    // side from the parity of the device k.
    int kskipmrr_trip(int k, int k_alloc, bool dyn = false) {
        auto Ar = [&](int j) { return vec(j); };                  // rows 0..k_alloc+1
        auto Ay = [&](int j) { return vec(k_alloc + 2 + j); };    // rows 0..k_alloc
        double* z = vec(2 * k_alloc + 3);
        double* spare = vec(2 * k_alloc + 4);                     // second home of Ar[0] for the fused steps
        const int end_epi = dyn ? EPI_ADAPT_TRIP_END : EPI_KS_TRIP_END;
        if (ext != nullptr && !dyn) {
            // ONE exchange per trip: r, A r, y of the neighbours' boundary rows into the pads of the home vectors; then the
            // single-GPU dense-band kernels on the extended operator (owned rows come out bit-identical to one GPU)
            const bool has_prev = ctx->rank > 0, has_next = ctx->rank + 1 < ctx->n_ranks;
            double* own[3] = {Ar(0), Ar(1), Ay(0)};
            PK_CHECK(pk_comm_ghost_exchange_inplace(ctx, own, 3, n, ext_depth, has_prev, has_next));
            PK_CHECK(pk_launch_matpow(ctx, ext, k, Ar(1) - ext_off, Ay(0) - ext_off, 0));
            PK_CHECK(pk_launch_gram(ctx, 0, n, ld, Ar(0), k + 2, Ay(0), k + 1, k + 2, EPI_GRAM_MRR));
            return pk_launch_mrr_steps(ctx, ext, k, Ar(0) - ext_off, Ar(1) - ext_off, Ay(0) - ext_off, z - ext_off,
                                       x - ext_off, Ar(2) - ext_off, Ar(3) - ext_off, Ay(1) - ext_off, EPI_KS_TRIP_END,
                                       ext_off, ext_off + n);
        }
        if (pk_matpow_ok(ctx, A, k)) {
            // all k levels of both chains in ONE pass over A (pk_matpow.cu); dyn: the kernel reads the current k itself
            PK_CHECK(pk_launch_matpow(ctx, A, k, Ar(1), Ay(0), dyn ? 1 : 0));
        } else {
            for (int j = 1; j <= k; ++j) {
                Ctl c(ctx, 0, dyn ? j : -1, 0);
                PK_CHECK(apply2(Ar(j), Ar(j + 1), Ay(j - 1), Ay(j)));
            }
        }
        {
            Ctl c(ctx, 0, dyn ? 0 : -1, 0);
            PK_CHECK(pk_launch_gram(ctx, 0, n, ld, Ar(0), k + 2, Ay(0), k + 1, k + 2, EPI_GRAM_MRR));
        }
        // Dense band on one GPU: the k+1 steps, the closing mat-vec and the trip-end reduction in ONE pass over A
        // (pk_matpow.cu: k_mrr_steps_band) — bit-identical to the step-by-step sequence below.  Scratch: basis slots
        // that are free once the Gram kernel has run.
        if (!dyn && pk_mrr_steps_ok(ctx, A, k) &&
            ((((uintptr_t)x | (uintptr_t)z | (uintptr_t)Ar(0) | (uintptr_t)Ay(0)) & 15) == 0) && (ld % 2 == 0))
            return pk_launch_mrr_steps(ctx, A, k, Ar(0), Ar(1), Ay(0), z, x, Ar(2), Ar(3), Ay(1), EPI_KS_TRIP_END);
        // The dynamic ping-pong needs the kernel to choose the vector it multiplies, which a host-enqueued halo exchange
        // (ncclSend/Recv of a host-known pointer) cannot follow: only the exchange fused into the SpMV kernel can, so
        // distributed adaptive solves on the NCCL halo path keep update and SpMV separate.
        const bool inkernel_halo = A->halo_p2p && A->use_tma && !A->pat_on;
        const bool fuse = pk_mat_can_fuse(A) && !(dyn && A->distributed && !inkernel_halo);
        if (!fuse) {
            for (int j = 0; j <= k; ++j) {
                {
                    Ctl c(ctx, 0, dyn ? j : -1, dyn ? 1 : 0);
                    PK_CHECK(pk_launch_mrr_update(ctx, n, Ar(1), Ay(0), z, Ar(0), Ar(0), nullptr, x, j,
                                                  (dyn || j == k) ? end_epi : EPI_KS_STEP));
                }
                Ctl c(ctx, 0, dyn ? j : -1, 0);
                PK_CHECK(apply(Ar(0), Ar(1)));
            }
            return PK_OK;
        }
        if (dyn) {
            {
                Ctl c(ctx, 0, 0, 1);
                PK_CHECK(pk_launch_mrr_update(ctx, n, Ar(1), Ay(0), z, Ar(0), Ar(0), spare, x, 0, end_epi));
            }
            for (int j = 1; j <= k; ++j) {
                Ctl c(ctx, 0, j, 1);
                PkDots d;
                d.epi = end_epi;
                d.fuse = 1; d.cj = j; d.f_a = Ay(0); d.f_b = z; d.f_x = x; d.f_out = spare;   // (home, spare): kernel picks
                PK_CHECK(pk_launch_spmv(ctx, A, Ar(0), nullptr, nullptr, nullptr, d));
            }
            return apply(Ar(0), Ar(1));
        }
        // Fused steps (see kskipcg): Ar1 = A Ar0 lives only in registers; Ar0 ping-pongs between home and `spare`.
        double* cur = (k % 2 == 1) ? spare : Ar(0);
        PK_CHECK(pk_launch_mrr_update(ctx, n, Ar(1), Ay(0), z, Ar(0), cur, nullptr, x, 0, k == 0 ? EPI_KS_TRIP_END : EPI_KS_STEP));
        for (int j = 1; j <= k; ++j) {
            double* nxt = (cur == spare) ? Ar(0) : spare;
            PkDots d;
            d.epi = (j == k) ? EPI_KS_TRIP_END : EPI_KS_STEP;
            d.fuse = 1; d.cj = j; d.f_a = Ay(0); d.f_b = z; d.f_x = x; d.f_out = nxt;
            PK_CHECK(pk_launch_spmv(ctx, A, cur, nullptr, nullptr, nullptr, d));
            cur = nxt;
        }
        return apply(Ar(0), Ar(1));                               // cur == Ar(0) here
    }
    int kskipmrr() {
        const int k = o.k;
        if (o.basis == 1) return kskipmrr_chebyshev();
        double *Ar0 = vec(0), *Ar1 = vec(1), *Ay0 = vec(k + 2), *z = vec(2 * k + 3);
        PK_CHECK(initial_residual(Ar0, nullptr, Ar1, EPI_RES0));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(kskipmrr_open(Ar0, Ar1, Ay0, z, EPI_KS_FIRST));
        PK_CHECK(run_batches(k + 1, 1, [&]() -> int { return kskipmrr_trip(k, k); }));
        return PK_OK;
    }

    // ---- k-skip MrR on a Chebyshev basis (opt-in; pk_scalars.h: pk_kskipmrr_coef_cheb) -------------
    // Same loop as kskipmrr(): only the basis of a trip (U_j = T_j(Ah) r, V_j = T_j(Ah) y instead of A^j r, A^j y; the
    // three-term recurrence rides in the epilogue of the two-chain SpMV) and the scalar engine's moment recurrences differ.
    // The k+1 steps of a trip are the same fused update + SpMV kernels, fed with (zeta_j, eta_j) computed from Chebyshev
    // moments.  A r of the current residual lives in its own vector (AR): level 1 of the basis is no longer A r.
    int kskipmrr_chebyshev() {
        const int k = o.k;
        PK_REQUIRE(o.lam_hi > o.lam_lo, "Chebyshev basis needs spectrum bounds lam_lo < lam_hi (pk_mat_gershgorin)");
        PK_REQUIRE(A->kind != MAT_DENSE && A->use_tma && !A->pat_on,
                   "the Chebyshev basis needs the TMA CSR kernel (16-byte aligned CSR arrays, no pattern compression)");
        const double c = 0.5 * (o.lam_hi - o.lam_lo), d = 0.5 * (o.lam_hi + o.lam_lo);
        auto U = [&](int j) { return vec(j); };                   // rows 0..k+1, U(0) = r
        auto V = [&](int j) { return vec(k + 2 + j); };           // rows 0..k,   V(0) = y
        double *z = vec(2 * k + 3), *spare = vec(2 * k + 4), *AR = vec(2 * k + 5);
        {   // c, d into the device state before anything reads them
            PkState* h = ctx->h_state;
            h->cheb_c = c;
            h->cheb_d = d;
            PK_CUDA(cudaMemcpyAsync(&ctx->d_state->cheb_c, &h->cheb_c, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        PK_CHECK(initial_residual(U(0), nullptr, AR, EPI_RES0));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(kskipmrr_open(U(0), AR, V(0), z, EPI_KS_FIRST));      // opening MrR step; leaves AR = A r
        PK_CHECK(run_batches(k + 1, 1, [&]() -> int {
            // basis: U_1 = (A r - d r) / c from the A r at hand, then per level one pass over A for both chains
            PK_CHECK(pk_launch_axpby(ctx, n, 1.0 / c, AR, -d / c, U(0), U(1)));
            for (int j = 1; j <= k; ++j) {
                PkDots dd;
                dd.fuse = 3;
                dd.cs[0] = 2.0 / c; dd.cs[1] = -2.0 * d / c; dd.cs[2] = -1.0;          // U_{j+1} = 2 Ah U_j - U_{j-1}
                dd.f_a = U(j - 1);
                if (j == 1) { dd.cs[3] = 1.0 / c; dd.cs[4] = -d / c; dd.cs[5] = 0.0; dd.f_b = nullptr; }   // V_1 = Ah V_0
                else { dd.cs[3] = 2.0 / c; dd.cs[4] = -2.0 * d / c; dd.cs[5] = -1.0; dd.f_b = V(j - 2); }
                PK_CHECK(pk_launch_spmv(ctx, A, U(j), U(j + 1), V(j - 1), V(j), dd));
            }
            PK_CHECK(pk_launch_gram(ctx, 0, n, ld, U(0), k + 2, V(0), k + 1, k + 2, EPI_GRAM_MRR_CHEB));
            // the k+1 steps: as in kskipmrr_trip (fused: A r of the step lives in registers; r ping-pongs home <-> spare)
            double* cur = (k % 2 == 1) ? spare : U(0);
            PK_CHECK(pk_launch_mrr_update(ctx, n, AR, V(0), z, U(0), cur, nullptr, x, 0, k == 0 ? EPI_KS_TRIP_END : EPI_KS_STEP));
            for (int j = 1; j <= k; ++j) {
                double* nxt = (cur == spare) ? U(0) : spare;
                PkDots ds;
                ds.epi = (j == k) ? EPI_KS_TRIP_END : EPI_KS_STEP;
                ds.fuse = 1; ds.cj = j; ds.f_a = V(0); ds.f_b = z; ds.f_x = x; ds.f_out = nxt;
                PK_CHECK(pk_launch_spmv(ctx, A, cur, nullptr, nullptr, nullptr, ds));
                cur = nxt;
            }
            return apply(U(0), AR);                                   // cur == U(0): A r for the next trip
        }));
        return PK_OK;
    }

    // ---- adaptive k-skip MrR: /root/reference/v3/cpu/adaptivekskipmrr.py:8-141 (normative variant) -----------
    // The residual-growth guard (:45-69) runs ON THE DEVICE: a one-thread kernel at the top of each trip decides
    // rollback / convergence / stop from the residual the previous trip recorded (EPI_ADAPT_GUARD), the rollback branch
    // is a fixed sequence of kernels predicated on that decision, and the trip itself is enqueued for the initial k
    // with the current k read from PkState by the kernels.  No host synchronisation per trip: the host polls the stop
    // flag once per batch like the other solvers, and the whole batch is graph-capturable.
    int adaptive() {
        const int k0 = o.k;
        double *Ar0 = vec(0), *Ar1 = vec(1), *Ay0 = vec(k0 + 2), *z = vec(2 * k0 + 3), *best_x = vec(2 * k0 + 5);
        PK_CHECK(initial_residual(Ar0, nullptr, Ar1, EPI_RES0));                  // :23-25 (pre_residual = residual[0])
        PK_CUDA(cudaMemcpyAsync(best_x, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        PK_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
        PK_CHECK(kskipmrr_open(Ar0, Ar1, Ay0, z, EPI_ADAPT_FIRST));               // :28-40
        PK_CHECK(run_batches(k0 + 1, 1, [&]() -> int {
            PK_CHECK(pk_launch_scalar(ctx, EPI_ADAPT_GUARD, 0));                  // :43-47, :67-68, :72-74
            PK_CHECK(pk_launch_adapt_save(ctx, n, x, best_x));                    // :48 x = pre_x  |  :69 pre_x = x
            {
                Ctl c(ctx, 1, -1, 0);                                             // the rollback branch, :49-66
                PK_CHECK(apply(x, Ar1));
                PK_CHECK(pk_launch_resid_init(ctx, n, b, Ar1, Ar0, nullptr, EPI_NONE));
                PK_CHECK(apply(Ar0, Ar1, Ar0, EPI_MRR_FIRST));
                PK_CHECK(pk_launch_mrr_first(ctx, n, Ar1, Ar0, x, Ay0, z, EPI_ADAPT_STEP));
                PK_CHECK(apply(Ar0, Ar1));
            }
            return kskipmrr_trip(k0, k0, true);                                   // :77-128 with the device's k
        }));
        return PK_OK;
    }
};

}  // namespace

extern "C" int pk_solve(pk_ctx* ctx, int method, pk_mat* mat, const double* d_b, double* d_x, double* d_work,
                        double* d_residual, int64_t* d_nosl, int64_t* d_khistory, int64_t hist_len,
                        const pk_solve_opts* opts, pk_solve_result* result) {
    PK_REQUIRE(ctx && mat && d_b && d_x && d_work && d_residual && d_nosl && opts && result, "null argument");
    PK_REQUIRE(opts->k >= 0 && opts->k <= PK_KMAX, "k out of range (0..PK_KMAX)");
    PK_REQUIRE(opts->maxiter >= 0, "maxiter < 0");
    PK_REQUIRE(hist_len >= 4, "history arrays too short");
    PK_REQUIRE(method != PK_ADAPTIVEKSKIPMRR || d_khistory != nullptr, "adaptivekskipmrr needs d_khistory");
    PK_REQUIRE(ctx->n_ranks == 1 || opts->global_n > 0, "distributed solve needs opts->global_n");
    PK_CUDA(cudaSetDevice(ctx->device));
    Solve s;
    s.ctx = ctx;
    s.A = mat;
    s.b = d_b;
    s.x = d_x;
    s.work = d_work;
    s.n = mat->n_rows;
    s.ld = mat->ld;
    s.o = *opts;
    s.method = method;
    ctx->launches = 0;
    ctx->spmvs = 0;
    // structure probe of the one-pass basis kernel: synchronises once per operator, so it must happen before any capture
    if (method == PK_KSKIPCG || method == PK_KSKIPMRR || method == PK_ADAPTIVEKSKIPMRR) {
        (void)pk_matpow_ok(ctx, mat, opts->k);
        // the dense-band kernels stage vectors by TMA bulk copies: with a work area or solution vector that is not
        // 16-byte aligned this operator stays on the general kernels (decided once, before anything is captured)
        if (mat->mp_dense && ((((uintptr_t)d_work | (uintptr_t)d_x) & 15) != 0 || (mat->ld & 1))) mat->mp_dense = false;
        if (method == PK_KSKIPMRR && opts->basis == 0 && mat->distributed && ((((uintptr_t)d_work | (uintptr_t)d_x) & 15) == 0)) {
            const long long depth = pk_band_ext_depth(ctx, mat, opts->k);
            if (depth > 0) {
                s.ext = mat->band_ext;
                s.ext_depth = depth;
                s.ext_off = mat->band_ra;               // rows of the extended operator in front of the owned ones
                s.work = d_work + s.ext_off;            // the last two work vectors (spare, Chebyshev A r) are unused here
            }
        }
    }
    PK_CHECK(s.init_state(d_residual, d_nosl, method == PK_ADAPTIVEKSKIPMRR ? d_khistory : nullptr, hist_len));
    int final_k = opts->k;
    int rc = PK_OK;
    switch (method) {
        case PK_CG: rc = s.cg(); break;
        case PK_MRR: rc = s.mrr(); break;
        case PK_KSKIPCG: rc = s.kskipcg(); break;
        case PK_KSKIPMRR: rc = s.kskipmrr(); break;
        case PK_ADAPTIVEKSKIPMRR: rc = s.adaptive(); break;
        case PK_CGCG: rc = s.cgcg(); break;
        default: pk_set_error("unknown method %d", method); return PK_ERR_ARG;
    }
    if (rc != PK_OK) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    PK_CUDA(cudaEventRecord(ctx->ev_t1, ctx->stream));
    PK_CHECK(s.fetch_state());
    float ms = 0.f;
    PK_CUDA(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    const PkState* h = ctx->h_state;
    if (h->guard < 0) {
        pk_set_error("a peer rank never delivered its %s (in-kernel wait timed out); solve aborted",
                     h->guard == -1 ? "all-reduce contribution" : "halo");
        return PK_ERR_NCCL;
    }
    result->iterations = h->it;
    result->entries = h->idx + 1;
    result->converged = h->converged;
    result->final_k = (method == PK_ADAPTIVEKSKIPMRR) ? h->k : final_k;
    result->final_residual = std::sqrt(h->rr) / h->bnorm;
    result->elapsed_s = (double)ms * 1e-3;
    result->kernel_launches = ctx->launches;
    result->spmv_count = ctx->spmvs;
    return PK_OK;
}

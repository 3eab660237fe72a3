// Multi-GPU exchange over NCCL / NVLink 5.  One process per GPU; vectors are SHARDED with the rows of A.
//   * halo exchange of x before an operator application: only the entries the local CSR block references
//     (ncclSend/ncclRecv group to the peers that need them) — replaces the reference's full-vector
//     memcpyPeer broadcast + gather + MPI_Allgather per mat-vec (/root/reference/v3/gpu/mpi/common.py:144,156,163);
//   * all-reduce of the packed partial dot products (the reference recomputes full-length dots redundantly on every
//     rank, /root/reference/v3/gpu/mpi/cg.py:45,50; v1 has the Allreduce precedent,
//     /root/reference/v1/processes/adaptivekskipmrr.py:114-116).
// libnccl is resolved at run time with dlopen (the copy torch ships), so single-GPU use has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>

#include "pk_common.cuh"
#include "pk_launch.h"

struct PkNcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct PkComm {
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
};

static PkNcclApi g_nccl;

static int pk_nccl_load(const char* path) {
    if (g_nccl.handle) return PK_OK;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        pk_set_error("cannot dlopen NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        return PK_ERR_NCCL;
    }
#define PK_SYM(field, name)                                             \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);              \
    if (!g_nccl.field) {                                                \
        pk_set_error("NCCL symbol %s missing", name);                   \
        return PK_ERR_NCCL;                                             \
    }
    PK_SYM(GetUniqueId, "ncclGetUniqueId");
    PK_SYM(CommInitRank, "ncclCommInitRank");
    PK_SYM(CommDestroy, "ncclCommDestroy");
    PK_SYM(AllReduce, "ncclAllReduce");
    PK_SYM(AllGather, "ncclAllGather");
    PK_SYM(Send, "ncclSend");
    PK_SYM(Recv, "ncclRecv");
    PK_SYM(GroupStart, "ncclGroupStart");
    PK_SYM(GroupEnd, "ncclGroupEnd");
    PK_SYM(GetErrorString, "ncclGetErrorString");
#undef PK_SYM
    g_nccl.handle = h;
    return PK_OK;
}

#define PK_NCCL(expr)                                                                             \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            pk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return PK_ERR_NCCL;                                                                   \
        }                                                                                         \
    } while (0)

extern "C" int pk_nccl_unique_id(const char* nccl_path, char id[PK_NCCL_ID_BYTES]) {
    PK_CHECK(pk_nccl_load(nccl_path));
    static_assert(sizeof(ncclUniqueId) == PK_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId uid;
    PK_NCCL(g_nccl.GetUniqueId(&uid));
    memcpy(id, &uid, PK_NCCL_ID_BYTES);
    return PK_OK;
}

extern "C" int pk_comm_init(pk_ctx* ctx, const char* nccl_path, int n_ranks, int rank, const char id[PK_NCCL_ID_BYTES]) {
    PK_REQUIRE(ctx != nullptr, "null context");
    PK_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank / n_ranks");
    PK_CHECK(pk_nccl_load(nccl_path));
    PK_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, PK_NCCL_ID_BYTES);
    PkComm* c = new PkComm();
    c->n_ranks = n_ranks;
    c->rank = rank;
    PK_NCCL(g_nccl.CommInitRank(&c->comm, n_ranks, uid, rank));
    ctx->comm = c;
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return PK_OK;
}

// ---- NVLink mailboxes for the in-kernel all-reduce ------------------------------------------------------------------
// Each rank allocates its mailbox with cudaMalloc, exports it with CUDA IPC; after the handles have been all-gathered
// (torch.distributed, host side) every rank maps its peers' mailboxes.  From then on the reducing kernels all-reduce
// their sums themselves (pk_device.cuh) and neither ncclAllReduce nor the 1-thread scalar kernel is launched.
extern "C" int pk_p2p_handle(pk_ctx* ctx, char handle[PK_IPC_HANDLE_BYTES]) {
    PK_REQUIRE(ctx != nullptr && handle != nullptr, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == PK_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    PK_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->my_mbox) {
        PK_CUDA(cudaMalloc(&ctx->my_mbox, PK_MBOX_DOUBLES * sizeof(double)));
        PK_CUDA(cudaMemset(ctx->my_mbox, 0, PK_MBOX_DOUBLES * sizeof(double)));
    }
    cudaIpcMemHandle_t h;
    PK_CUDA(cudaIpcGetMemHandle(&h, ctx->my_mbox));
    memcpy(handle, &h, PK_IPC_HANDLE_BYTES);
    return PK_OK;
}

extern "C" int pk_p2p_open(pk_ctx* ctx, int n_ranks, int rank, const char* handles) {
    PK_REQUIRE(ctx != nullptr && handles != nullptr, "null argument");
    PK_REQUIRE(n_ranks >= 2 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank / n_ranks");
    PK_REQUIRE(ctx->my_mbox != nullptr, "call pk_p2p_handle first");
    PK_CUDA(cudaSetDevice(ctx->device));
    PkP2P h;
    memset(&h, 0, sizeof(h));
    h.n_ranks = n_ranks;
    h.rank = rank;
    for (int p = 0; p < n_ranks; ++p) {
        if (p == rank) {
            h.mbox[p] = ctx->my_mbox;
            continue;
        }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles + (size_t)p * PK_IPC_HANDLE_BYTES, PK_IPC_HANDLE_BYTES);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            pk_set_error("cudaIpcOpenMemHandle(rank %d -> %d): %s", rank, p, cudaGetErrorString(e));
            for (void* q : ctx->peer_mbox) cudaIpcCloseMemHandle(q);
            ctx->peer_mbox.clear();
            return PK_ERR_UNSUPPORTED;
        }
        ctx->peer_mbox.push_back(ptr);
        h.mbox[p] = (double*)ptr;
    }
    PK_CUDA(cudaMalloc(&h.seq, sizeof(unsigned long long)));
    PK_CUDA(cudaMemset(h.seq, 0, sizeof(unsigned long long)));
    PkP2P* d = nullptr;
    PK_CUDA(cudaMalloc(&d, sizeof(PkP2P)));
    PK_CUDA(cudaMemcpy(d, &h, sizeof(PkP2P), cudaMemcpyHostToDevice));
    ctx->d_p2p = d;
    return PK_OK;
}

// ---- NVLink halo push: receive buffers ------------------------------------------------------------------------------
extern "C" int pk_mat_halo_p2p_handle(pk_mat* m, char handle[PK_IPC_HANDLE_BYTES]) {
    PK_REQUIRE(m != nullptr && handle != nullptr, "null argument");
    PK_REQUIRE(m->distributed, "pk_mat_set_halo first");
    PK_CUDA(cudaSetDevice(m->ctx->device));
    if (!m->d_recvbuf) {
        const size_t n = PK_HALO_HDR + 4 * (size_t)std::max<long long>(m->n_halo, 1);
        PK_CUDA(cudaMalloc(&m->d_recvbuf, n * sizeof(double)));
        PK_CUDA(cudaMemset(m->d_recvbuf, 0, n * sizeof(double)));
    }
    cudaIpcMemHandle_t h;
    PK_CUDA(cudaIpcGetMemHandle(&h, m->d_recvbuf));
    memcpy(handle, &h, PK_IPC_HANDLE_BYTES);
    return PK_OK;
}

// handles: n_ranks x 64 bytes; dst_off[q] = offset of my entries inside q's halo (q's recv_off[me]);
// peer_nhalo[q] = q's halo length.
extern "C" int pk_mat_halo_p2p_open(pk_mat* m, const char* handles, const int64_t* dst_off, const int64_t* peer_nhalo) {
    PK_REQUIRE(m && handles && dst_off && peer_nhalo, "null argument");
    PK_REQUIRE(m->d_recvbuf != nullptr, "call pk_mat_halo_p2p_handle first");
    pk_ctx* ctx = m->ctx;
    const int P = ctx->n_ranks, me = ctx->rank;
    PK_REQUIRE(P <= PK_MAX_RANKS, "too many ranks for the push path");
    PK_CUDA(cudaSetDevice(ctx->device));
    PkHaloPush& hp = m->push;
    memset(&hp, 0, sizeof(hp));
    hp.n_ranks = P;
    hp.me = me;
    hp.send_idx = m->d_send_idx;
    for (int q = 0; q <= P; ++q) hp.send_off[q] = m->send_off[q];
    for (int q = 0; q < P; ++q) {
        hp.dst_off[q] = dst_off[q];
        hp.peer_nhalo[q] = peer_nhalo[q];
        hp.send_first[q] = m->send_first[q];
        hp.send_contig[q] = m->send_contig[q];
        if (q == me) { hp.peer_recv[q] = m->d_recvbuf; continue; }
        // q is a peer if entries flow in EITHER direction (symmetric by construction: my send to q is q's receive from me)
        const bool needed = m->send_off[q + 1] > m->send_off[q] || m->recv_off[q + 1] > m->recv_off[q];
        if (!needed) continue;                      // no exchange with q: no mapping, no flags
        hp.peer_mask |= (1u << q);
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles + (size_t)q * PK_IPC_HANDLE_BYTES, PK_IPC_HANDLE_BYTES);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            pk_set_error("cudaIpcOpenMemHandle(halo, rank %d -> %d): %s", me, q, cudaGetErrorString(e));
            return PK_ERR_UNSUPPORTED;
        }
        m->peer_recv_maps.push_back(ptr);
        hp.peer_recv[q] = (double*)ptr;
    }
    PK_CUDA(cudaMalloc(&hp.seq, sizeof(unsigned long long)));
    PK_CUDA(cudaMemset(hp.seq, 0, sizeof(unsigned long long)));
    PK_CUDA(cudaMalloc(&hp.ticket, sizeof(unsigned int)));
    PK_CUDA(cudaMemset(hp.ticket, 0, sizeof(unsigned int)));
    PK_CUDA(cudaMalloc(&hp.ticket_done, sizeof(unsigned int)));
    PK_CUDA(cudaMemset(hp.ticket_done, 0, sizeof(unsigned int)));
    PK_CUDA(cudaMalloc(&m->d_push, sizeof(PkHaloPush)));
    PK_CUDA(cudaMemcpy(m->d_push, &hp, sizeof(PkHaloPush), cudaMemcpyHostToDevice));
    m->halo_p2p = true;
    return PK_OK;
}

extern "C" int pk_mat_halo_p2p_disable(pk_mat* m) {
    PK_REQUIRE(m != nullptr, "null operator");
    PK_CUDA(cudaSetDevice(m->ctx->device));
    pk_mat_halo_p2p_close(m);
    return PK_OK;
}

void pk_mat_halo_p2p_close(pk_mat* m) {
    for (void* q : m->peer_recv_maps) cudaIpcCloseMemHandle(q);
    m->peer_recv_maps.clear();
    if (m->push.seq) cudaFree(m->push.seq);
    if (m->push.ticket) cudaFree(m->push.ticket);
    if (m->push.ticket_done) cudaFree(m->push.ticket_done);
    m->push.ticket_done = nullptr;
    if (m->d_push) cudaFree(m->d_push);
    m->d_push = nullptr;
    if (m->d_recvbuf) cudaFree(m->d_recvbuf);
    m->d_recvbuf = nullptr;
    m->push.seq = nullptr;
    m->push.ticket = nullptr;
    m->halo_p2p = false;
}

// Measurement aid for bench.py ("exposed communication per iteration"): with nocomm set, a solve launches the same
// kernels but skips the halo exchange and the all-reduces, so its time is the compute-only time of this rank.
extern "C" int pk_ctx_set_nocomm(pk_ctx* ctx, int on) {
    PK_REQUIRE(ctx != nullptr, "null context");
    ctx->nocomm = on != 0;
    return PK_OK;
}

extern "C" int pk_comm_destroy(pk_ctx* ctx) {
    if (ctx && ctx->d_p2p) {
        cudaStreamSynchronize(ctx->stream);
        for (void* q : ctx->peer_mbox) cudaIpcCloseMemHandle(q);
        ctx->peer_mbox.clear();
        cudaFree(ctx->d_p2p);
        ctx->d_p2p = nullptr;
    }
    if (ctx && ctx->comm) {
        if (ctx->comm->comm) g_nccl.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
        ctx->n_ranks = 1;
        ctx->rank = 0;
    }
    return PK_OK;
}

int pk_comm_allreduce(pk_ctx* ctx, double* buf, long long n, cudaStream_t s) {
    if (!ctx->comm || ctx->n_ranks <= 1 || ctx->nocomm) return PK_OK;
    PK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, ctx->comm->comm, s));
    return PK_OK;
}

// Neighbour exchange of the matrix-powers ghost zone (pk_matpow.cu): `depth` entries of each of the two level-0 vectors
// to / from the previous and the next rank, one grouped send/recv on the solver's stream (graph-capturable).
int pk_comm_ghost_exchange(pk_ctx* ctx, const double* v0, const double* v1, long long n_loc, long long depth,
                           double* gin, bool has_prev, bool has_next) {
    if (!ctx->comm || ctx->n_ranks <= 1 || ctx->nocomm) return PK_OK;
    const int me = ctx->rank;
    const double* src[2] = {v0, v1};
    PK_NCCL(g_nccl.GroupStart());
    for (int c = 0; c < 2; ++c) {
        double* above = gin + (size_t)(2 * c + 0) * depth;
        double* below = gin + (size_t)(2 * c + 1) * depth;
        if (has_prev) {
            PK_NCCL(g_nccl.Send(src[c], (size_t)depth, ncclDouble, me - 1, ctx->comm->comm, ctx->stream));
            PK_NCCL(g_nccl.Recv(above, (size_t)depth, ncclDouble, me - 1, ctx->comm->comm, ctx->stream));
        }
        if (has_next) {
            PK_NCCL(g_nccl.Send(src[c] + (n_loc - depth), (size_t)depth, ncclDouble, me + 1, ctx->comm->comm, ctx->stream));
            PK_NCCL(g_nccl.Recv(below, (size_t)depth, ncclDouble, me + 1, ctx->comm->comm, ctx->stream));
        }
    }
    PK_NCCL(g_nccl.GroupEnd());
    return PK_OK;
}

// Ghost zones kept IN the vectors (row-partitioned dense-band trip): own[v] points at the n_loc owned entries of vector v,
// which has `depth` entries of room on either side; the first / last `depth` owned entries go to the previous / next rank
// and theirs land in the pads.  One grouped send/recv on the solver's stream (graph-capturable).
int pk_comm_ghost_exchange_inplace(pk_ctx* ctx, double* const* own, int nvec, long long n_loc, long long depth,
                                   bool has_prev, bool has_next) {
    if (!ctx->comm || ctx->n_ranks <= 1 || ctx->nocomm) return PK_OK;
    const int me = ctx->rank;
    PK_NCCL(g_nccl.GroupStart());
    for (int v = 0; v < nvec; ++v) {
        if (has_prev) {
            PK_NCCL(g_nccl.Send(own[v], (size_t)depth, ncclDouble, me - 1, ctx->comm->comm, ctx->stream));
            PK_NCCL(g_nccl.Recv(own[v] - depth, (size_t)depth, ncclDouble, me - 1, ctx->comm->comm, ctx->stream));
        }
        if (has_next) {
            PK_NCCL(g_nccl.Send(own[v] + (n_loc - depth), (size_t)depth, ncclDouble, me + 1, ctx->comm->comm, ctx->stream));
            PK_NCCL(g_nccl.Recv(own[v] + n_loc, (size_t)depth, ncclDouble, me + 1, ctx->comm->comm, ctx->stream));
        }
    }
    PK_NCCL(g_nccl.GroupEnd());
    return PK_OK;
}

int pk_comm_allgather(pk_ctx* ctx, const double* send, double* recv, long long n, cudaStream_t s) {
    if (!ctx->comm || ctx->n_ranks <= 1) {
        if (send != recv) PK_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        return PK_OK;
    }
    PK_NCCL(g_nccl.AllGather(send, recv, (size_t)n, ncclDouble, ctx->comm->comm, s));
    return PK_OK;
}

extern "C" int pk_allreduce_sum(pk_ctx* ctx, double* d_buf, int64_t n) {
    PK_REQUIRE(ctx != nullptr, "null context");
    return pk_comm_allreduce(ctx, d_buf, n, ctx->stream);
}

extern "C" int pk_allgather(pk_ctx* ctx, const double* d_send, double* d_recv, int64_t n_per_rank) {
    PK_REQUIRE(ctx != nullptr, "null context");
    return pk_comm_allgather(ctx, d_send, d_recv, n_per_rank, ctx->stream);
}

// gather the entries a peer needs into the contiguous send buffer (skipped when the list is one contiguous run)
__global__ void k_pack(const int32_t* __restrict__ idx, long long n, const double* __restrict__ x0,
                       const double* __restrict__ x1, double* __restrict__ out0, double* __restrict__ out1) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int32_t j = idx[i];
        out0[i] = x0[j];
        if (x1) out1[i] = x1[j];
    }
}

// Start the halo exchange of x (and x1) on the side stream: the main stream may run interior rows meanwhile.
int pk_comm_halo_start(pk_ctx* ctx, pk_mat* m, double* x, double* x1) {
    if (!m->distributed || ctx->n_ranks <= 1 || ctx->nocomm) return PK_OK;
    const int P = ctx->n_ranks;
    const long long n_send = m->send_off[P];
    // A rank whose rows are needed by a peer must send even when it references no remote column itself
    // (structurally nonsymmetric A, e.g. block triangular): the exchange is skipped only if BOTH directions are empty.
    if (m->n_halo == 0 && n_send == 0) return PK_OK;
    PK_REQUIRE(ctx->comm != nullptr, "distributed operator without a communicator");
    // x must be complete before it is sent
    PK_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    PK_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_a, 0));
    bool need_pack = false;
    for (int p = 0; p < P; ++p)
        if (m->send_off[p + 1] > m->send_off[p] && !m->send_contig[p]) need_pack = true;
    double* sb0 = m->d_sendbuf;
    double* sb1 = m->d_sendbuf + n_send;
    if (need_pack) {
        int grid = (int)((n_send + 255) / 256);
        if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
        k_pack<<<grid, 256, 0, ctx->side>>>(m->d_send_idx, n_send, x, x1, sb0, sb1);
        PK_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    PK_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < P; ++p) {
        const long long ns = m->send_off[p + 1] - m->send_off[p];
        const long long nr = m->recv_off[p + 1] - m->recv_off[p];
        if (ns > 0) {
            const double* s0 = m->send_contig[p] ? (x + m->send_first[p]) : (sb0 + m->send_off[p]);
            PK_NCCL(g_nccl.Send(s0, (size_t)ns, ncclDouble, p, ctx->comm->comm, ctx->side));
            if (x1) {
                const double* s1 = m->send_contig[p] ? (x1 + m->send_first[p]) : (sb1 + m->send_off[p]);
                PK_NCCL(g_nccl.Send(s1, (size_t)ns, ncclDouble, p, ctx->comm->comm, ctx->side));
            }
        }
        if (nr > 0) {
            PK_NCCL(g_nccl.Recv(x + m->n_rows + m->recv_off[p], (size_t)nr, ncclDouble, p, ctx->comm->comm, ctx->side));
            if (x1)
                PK_NCCL(g_nccl.Recv(x1 + m->n_rows + m->recv_off[p], (size_t)nr, ncclDouble, p, ctx->comm->comm,
                                    ctx->side));
        }
    }
    PK_NCCL(g_nccl.GroupEnd());
    PK_CUDA(cudaEventRecord(ctx->ev_b, ctx->side));
    return PK_OK;
}

int pk_comm_halo_wait(pk_ctx* ctx) {
    if (ctx->n_ranks <= 1 || ctx->nocomm) return PK_OK;
    PK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return PK_OK;
}

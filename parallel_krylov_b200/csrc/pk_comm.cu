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

extern "C" int pk_comm_destroy(pk_ctx* ctx) {
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
    if (!ctx->comm || ctx->n_ranks <= 1) return PK_OK;
    PK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, ctx->comm->comm, s));
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
    if (!m->distributed || m->n_halo == 0 || ctx->n_ranks <= 1) return PK_OK;
    PK_REQUIRE(ctx->comm != nullptr, "distributed operator without a communicator");
    const int P = ctx->n_ranks;
    const long long n_send = m->send_off[P];
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
    if (ctx->n_ranks <= 1) return PK_OK;
    PK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return PK_OK;
}

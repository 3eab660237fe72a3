// Device-side helpers: deterministic block/grid reductions with a last-block epilogue, and the scalar engine
// (every scalar recurrence of the reference loops, executed by one thread on the reduced sums).
#pragma once
#include "pk_common.cuh"
#include "pk_scalars.h"

// ---- TMA bulk-copy / mbarrier primitives (sm_90+; SASS: UBLKCP, SYNCS.*) --------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Bulk copy with an L2 evict-first hint: for streams that are read once per pass (the CSR arrays), so that the vectors
// the same pass re-reads (x, and the previous kernel's outputs) keep their L2 lines.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, unsigned bytes,
                                              unsigned long long* bar, unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- waiting for a peer's flag (NVLink mailboxes, halo push) ------------------------------------------------------------
// Poll a flag in this GPU's own memory until it holds `want`.  The first polls are back to back (the common case: the
// peer is a few microseconds behind); after that the thread backs off with nanosleep so that a long wait does not
// hammer L2, and the wall-clock budget (PK_SPIN_TIMEOUT_NS of %globaltimer, not a poll count) bounds a wait for a peer
// that died: the caller then stops the solve with an error instead of hanging the GPU.
// acquire/release fence at system scope (orders this thread's accesses to peer GPUs' memory).  __threadfence_system() is
// the sequentially-consistent flavour (SASS MEMBAR.SC.SYS), which every block of a persistent grid executing once costs
// hundreds of microseconds in total (measured, r02); the protocols here only need release / acquire ordering.
__device__ __forceinline__ void pk_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

constexpr unsigned long long PK_SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long pk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool pk_spin_until(const volatile unsigned long long* flag, unsigned long long want) {
    if (*flag == want) return true;
    unsigned int polls = 0, ns = 32;
    unsigned long long t0 = 0;
    while (*flag != want) {
        if (++polls < 256) continue;
        if (t0 == 0) t0 = pk_globaltimer();
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
        if ((polls & 63u) == 0 && pk_globaltimer() - t0 > PK_SPIN_TIMEOUT_NS) return false;
    }
    return true;
}

struct PkRedArgs {
    double* partials;       // [nsums][max_blocks]
    unsigned int* ticket;
    int max_blocks;
    PkState* st;
    int epi;                // PkEpi
    int defer;              // 1: only publish the sums (multi-GPU: all-reduce + pk_scalar_kernel follow)
    int g_off;              // >= 0: publish into st->gram[g_off + j] instead of st->red[j]
    int red_off;            // publish into st->red[red_off + j] (local sums a LATER kernel all-reduces with its own)
    // split launches that feed ONE reduction (interior rows, then boundary rows after the halo arrived):
    int block_off;          // this launch's partials start at this block slot
    int nb_total;           // slots the final reduction covers (0: gridDim.x)
    int store_only;         // 1: just store the partials (a later launch on the same stream finishes)
    // multi-GPU all-reduce fused into this kernel's last block (NVLink peer stores into per-rank mailboxes)
    const PkP2P* p2p;       // nullptr: single GPU, or NCCL path (defer = 1)
    int ar_n;               // doubles to all-reduce once the local sums are published (0: local publish only)
    int post_only;          // 1 (with p2p): post the sums to the peers and return; k_ar_wait collects and runs the epilogue
    // Launch sequences whose shape depends on device-resident state (adaptivekskipmrr: the guard and the current k live
    // in PkState, so the host enqueues the sequence for the initial k and kernels decide for themselves):
    int only_rollback;      // 1: the kernel runs only when st->rollback != 0 (the rollback branch of a trip)
    int dyn_cj;             // >= 0: index of a k-skip step / basis level (or 0 for the Gram kernel): skipped when > st->k
    int dyn_last;           // 1 (step kernels): the reduction + epilogue happen only when dyn_cj == st->k (last step of the trip)
};

// Should this kernel do nothing?  (stop flag raised, or a predicated launch whose condition is false.)  Uniform across
// the grid — and across ranks, because the state it reads was computed from all-reduced, bit-identical sums.
__device__ __forceinline__ bool pk_skip(const PkRedArgs& ra) {
    const volatile PkState* st = ra.st;
    if (st->done != 0) return true;
    if (ra.only_rollback && st->rollback == 0) return true;
    if (ra.dyn_cj >= 0 && ra.dyn_cj > st->k) return true;
    return false;
}

// --------------------------------------------------------------------------------------------------------------
// Rounding discipline: the library is compiled with -fmad=false so that a*b+c rounds twice exactly like numpy's
// temporaries (`x += alpha * p` is a multiply then an add, /root/reference/v3/cpu/cg.py:30).

// pk_record / pk_stop_test / pk_epilogue (the scalar engine) live in pk_state.h: plain C++ shared with the host tests.

// --------------------------------------------------------------------------------------------------------------
// The two halves of the mailbox all-reduce (one block; called by the last block of a reducing kernel, or — the wait —
// by the one-block kernel k_ar_wait when independent work is scheduled between post and wait).
template <int BLOCK>
__device__ __forceinline__ void pk_mailbox_post(const PkP2P* pp, const double* buf, int n) {
    const int P = pp->n_ranks, me = pp->rank;
    const unsigned long long seq = *pp->seq + 1ull;
    const int bank = (int)(seq & 1ull);
    for (int t = threadIdx.x; t < n * P; t += BLOCK) {
        const int p = t / n, j = t - p * n;
        pp->mbox[p][((size_t)(bank * PK_MAX_RANKS + me)) * PK_MBOX_STRIDE + j] = buf[j];
    }
    __syncthreads();
    if (threadIdx.x < P) {
        pk_fence_sys();                               // release: cumulative over the block's payload stores (barrier above)
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
            pp->mbox[threadIdx.x] + ((size_t)(bank * PK_MAX_RANKS + me)) * PK_MBOX_STRIDE + PK_MBOX_PAYLOAD);
        *f = seq;
    }
}

template <int BLOCK>
__device__ __forceinline__ void pk_mailbox_wait(const PkP2P* pp, double* buf, int n, PkState* st) {
    const int P = pp->n_ranks, me = pp->rank;
    const unsigned long long seq = *pp->seq + 1ull;
    const int bank = (int)(seq & 1ull);
    if (threadIdx.x < P) {
        volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(
            pp->mbox[me] + ((size_t)(bank * PK_MAX_RANKS + threadIdx.x)) * PK_MBOX_STRIDE + PK_MBOX_PAYLOAD);
        if (!pk_spin_until(mine, seq)) {              // a peer never arrived: stop the solve instead of hanging
            st->done = 1;
            st->converged = 0;
            st->guard = -1;
        }
        pk_fence_sys();                               // acquire: the payload behind the flag is visible ...
    }
    __syncthreads();                                  // ... to the whole block
    for (int j = threadIdx.x; j < n; j += BLOCK) {
        double v = 0.0;
        for (int p0 = 0; p0 < P; p0 += 4) {                 // four contributions requested at once ...
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                t[u] = (p0 + u < P) ? __ldcv(pp->mbox[me] + ((size_t)(bank * PK_MAX_RANKS + p0 + u)) * PK_MBOX_STRIDE + j) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)                     // ... added in rank order: identical bits on every rank
                if (p0 + u < P) v += t[u];
        }
        buf[j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) *pp->seq = seq;
}

// --------------------------------------------------------------------------------------------------------------
// Block reduction of NS running sums (fixed shape: shuffle tree, then warps in order), one partial per block,
// then the last block to finish (ticket) reduces the partials in a fixed order and runs the epilogue.
// Deterministic for a given grid size; no floating-point atomics anywhere.
template <int NS, int BLOCK, bool GRAM = false>
__device__ __forceinline__ void pk_grid_reduce(double (&acc)[NS], const PkRedArgs& ra) {
    constexpr int NW = BLOCK / 32;
    __shared__ double sh[NS][NW];
    __shared__ double tot[NS];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double v = acc[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sh[j][warp] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < NS; j += BLOCK) {
        double v = sh[j][0];
#pragma unroll
        for (int w = 1; w < NW; ++w) v += sh[j][w];
        ra.partials[(size_t)j * ra.max_blocks + ra.block_off + blockIdx.x] = v;
        __threadfence();
    }
    if (ra.store_only) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ra.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int nb = ra.nb_total > 0 ? ra.nb_total : (int)gridDim.x;
    for (int j = warp; j < NS; j += NW) {
        const double* p = ra.partials + (size_t)j * ra.max_blocks;
        double v = 0.0;
        // the partials of up to a few thousand blocks: 4 loads in flight per lane, added in the same fixed order as a
        // plain loop would (this serial tail of every reduction is on the critical path of each iteration: ~20 dependent
        // L2 round trips before, 5 now)
        int b = lane;
        for (; b + 3 * 32 < nb; b += 4 * 32) {
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = __ldcg(p + b + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) v += t[u];
        }
        for (; b < nb; b += 32) v += __ldcg(p + b);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) tot[j] = v;
    }
    __syncthreads();
    double* dst = (ra.g_off >= 0) ? (ra.st->gram + ra.g_off) : (ra.st->red + ra.red_off);
    if (threadIdx.x == 0) {
        *ra.ticket = 0u;
        for (int j = 0; j < NS; ++j) dst[j] = tot[j];
    }
    if (ra.p2p != nullptr && ra.ar_n > 0) {
        // ---- all-reduce over NVLink, inside this kernel --------------------------------------------------------------
        // Every rank's last block stores its local sums into every rank's mailbox (peer-mapped memory, plain stores
        // through NVSwitch), publishes a sequence flag, waits for the P flags addressed to it, and adds the P
        // contributions in RANK ORDER — so all ranks obtain bit-identical sums and take identical decisions.  Two
        // mailbox banks (sequence parity) suffice: a rank can run at most one reduction ahead of a peer.
        __syncthreads();
        double* buf = (ra.g_off >= 0) ? ra.st->gram : ra.st->red;   // Gram: the whole gram[] (all windows) is reduced
        pk_mailbox_post<BLOCK>(ra.p2p, buf, ra.ar_n);
        if (ra.post_only) return;           // a later kernel (k_ar_wait) collects: the flight overlaps the work in between
        pk_mailbox_wait<BLOCK>(ra.p2p, buf, ra.ar_n, ra.st);
    }
    if (threadIdx.x == 0 && !ra.defer) pk_epilogue<GRAM>(ra.epi, ra.st);
}

// Operator application y = A·x for a row block of A (K1/K2 of SURVEY.md §2b), optionally for two right-hand sides at
// once (the two chains A^j r / A^j p of the k-skip basis share ONE pass over A), with the dot products the solver
// needs next reduced in the epilogue.  Replaces MultiGpu.dot (/root/reference/v3/gpu/common.py:113-126) and the
// cupy.dot launches that follow it (/root/reference/v3/gpu/cg.py:31-32).
//
// CSR-stream kernel (short rows: stencils, bands): a block owns a tile of BLOCK consecutive rows.  Phase 1 streams the
// tile's nonzeros with 128-bit coalesced loads (int4 of column indices, 2 x double2 of values) into shared memory.
// Phase 2: thread t walks row t left to right, gathering x through the read-only path — neighbouring lanes are
// neighbouring rows, so on banded/stencil structure a warp's gather touches 2-3 cache lines instead of ~14 when the
// gather is done in nonzero order (ncu r01: that version was L1-wavefront bound at 64 % of HBM peak).  The
// accumulation order is that of scipy's csr_matvec, which makes y bit-identical to the oracle's A.dot(x) (products
// and sums are separately rounded: the library is built with -fmad=false).  Tiles whose nonzeros exceed the staging buffer
// (long rows) are processed warp-per-row with a shuffle reduction instead.
// Dense kernel: warp per row, 128-bit loads, no tensor cores (GEMV is HBM-bound).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "pk_device.cuh"
#include "pk_launch.h"

namespace {

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
    long long row_lo2, row_hi2;   // optional second and third row ranges processed by the same launch, after the first:
    long long row_lo3, row_hi3;   // the boundary rows above / below the interior
    // Halo exchange fused into this kernel (HALO variant): every block first pushes its share of the boundary entries of
    // x into the peers' receive buffers over NVLink, the tiles of the first range (interior rows: no halo column) run
    // while the peers' pushes are in flight, and a block waits for the peers' flags right before its first boundary tile
    // (ranges 2 and 3), whose halo columns (>= n_own) are then read from this rank's receive buffer.
    const double* hrecv;          // this rank's receive buffer (nullptr: no exchange, columns index x directly)
    const PkHaloPush* hp;         // push descriptor in device memory
    int halo_dry;                 // measurement aid (pk_ctx_set_nocomm): same kernel, but neither push nor flag waits
    long long n_own, n_halo;
    long long nnz_total;
    long long rowptr_len;   // entries of rowptr (n_rows_total + 1)
    int base_mis;           // (index of col[0]/val[0] in the caller's arrays) mod 4: a row segment of a block with 64-bit
                            // row pointers starts anywhere, and the staged windows must start 16-byte aligned in MEMORY
    int cap;                // staging capacity (nonzeros) per right-hand side
    int evict_first;        // CSR arrays are streamed with an L2 evict-first hint (keeps the vectors in L2)
    int blocked;            // tile -> block assignment: 1 contiguous chunks per block, 0 grid-strided
    // k-skip step fused into the row epilogue (coefficients of step cj are already in PkState::coef):
    //   fuse 1 (MrR, x0 = Ar0): Ay0 = eta Ay0 + zeta y ; z = eta z - zeta Ar0 ; Ar0' = Ar0 - Ay0 ; x -= z   (y = A Ar0 not stored)
    //   fuse 2 (CG,  x0 = Ap0): x += alpha Ap0 ; Ar0 -= alpha y ; Ap0' = Ar0 + beta Ap0                    (y = A Ap0 not stored)
    int fuse, cj;
    double cs[6];           // fuse 3: coefficients of the three-term recurrence (see PkDots)
    double* f_a;            // MrR: Ay0      | CG: Ar0 (in place)    | fuse 3: previous level of chain 0 (nullable)
    double* f_b;            // MrR: z        | CG: unused
    double* f_x;            // solution x
    double* f_out;          // MrR: Ar0'     | CG: Ap0'   (a buffer different from x0: other rows still gather x0)
    int reduce;             // 1: run the grid reduction (3 sums)
};

template <int NV, int BLOCK, bool VEC>
__global__ void __launch_bounds__(BLOCK) k_spmv_stream(SpmvArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sval = reinterpret_cast<double*>(smem_raw);              // [cap]
    int* scol = reinterpret_cast<int*>(sval + a.cap);                // [cap]
    __shared__ int rp[BLOCK + 1];
    constexpr int NW = BLOCK / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long n_rows = a.row_hi - a.row_lo;
    const long long n_tiles = (n_rows + BLOCK - 1) / BLOCK;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = a.row_lo + tile * BLOCK;
        const int nr = (int)((a.row_hi - r0) < BLOCK ? (a.row_hi - r0) : BLOCK);
        for (int t = tid; t <= nr; t += BLOCK) rp[t] = a.rowptr[r0 + t];
        __syncthreads();
        const int base = rp[0], end = rp[nr];
        const int q0 = VEC ? (((base + a.base_mis) & ~3) - a.base_mis) : base;   // 16-byte aligned start of the staged window
        if (end - q0 <= a.cap) {
            // ---- phase 1: stream the tile's (col, val) into shared memory, 128-bit coalesced ------------------
            if (VEC) {
                for (int q = q0 + 4 * tid; q < end; q += 4 * BLOCK) {
                    const int i = q - q0;
                    if (q >= 0 && (long long)q + 4 <= a.nnz_total) {
                        const int4 c4 = __ldg(reinterpret_cast<const int4*>(a.col + q));
                        const double2 v01 = __ldg(reinterpret_cast<const double2*>(a.val + q));
                        const double2 v23 = __ldg(reinterpret_cast<const double2*>(a.val + q + 2));
                        *reinterpret_cast<int4*>(scol + i) = c4;
                        *reinterpret_cast<double2*>(sval + i) = v01;
                        *reinterpret_cast<double2*>(sval + i + 2) = v23;
                    } else {
                        for (int e = 0; e < 4 && (long long)q + e < a.nnz_total; ++e) {
                            if (q + e < 0) continue;          // entries before the segment: never referenced by a row
                            scol[i + e] = a.col[q + e];
                            sval[i + e] = a.val[q + e];
                        }
                    }
                }
            } else {
                for (int q = base + tid; q < end; q += BLOCK) {
                    scol[q - q0] = a.col[q];
                    sval[q - q0] = a.val[q];
                }
            }
            __syncthreads();
            // ---- phase 2: one thread per row; x gathered here, so that neighbouring lanes (neighbouring rows) read
            //      neighbouring x entries on banded/stencil structure; left-to-right sum (scipy csr_matvec order) --
            if (tid < nr) {
                const int s = rp[tid] - q0, e = rp[tid + 1] - q0;
                const long long row = r0 + tid;
                double sum0 = 0.0, sum1 = 0.0;
                // batches of UNR entries: all gathers of a batch are in flight before the (ordered) accumulation
                constexpr int UNR = 8;
                for (int j = s; j < e; j += UNR) {
                    double vv[UNR], xa[UNR], xb[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const bool ok = j + u < e;
                        const int c = ok ? scol[j + u] : 0;
                        vv[u] = ok ? sval[j + u] : 0.0;
                        xa[u] = ok ? __ldg(a.x0 + c) : 0.0;
                        if (NV == 2) xb[u] = ok ? __ldg(a.x1 + c) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        if (j + u < e) {
                            sum0 += vv[u] * xa[u];
                            if (NV == 2) sum1 += vv[u] * xb[u];
                        }
                    }
                }
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
        __syncthreads();   // rp / staging buffers are reused by the next tile
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

// --------------------------------------------------------------------------------------------------------------
// TMA-pipelined variant (the default on B200).  Same tiling and the same per-row arithmetic as k_spmv_stream, but the
// tile's rowptr / col / val windows are brought in by 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) into a ring of
// STAGES shared-memory buffers, completion signalled on an mbarrier per stage.  One thread issues the copies for tile
// i+STAGES-1 while all threads run the row phase of tile i, so HBM requests stay in flight regardless of how long the
// row phase takes; the CSR arrays never pass through registers.  A tile that cannot be bulk-copied (window crosses the
// end of the arrays, or longer than the stage: long rows) is fetched/processed by the plain paths below.
constexpr int PK_PUSH_BLOCKS = 32;   // blocks of the fused-exchange SpMV that push the boundary entries to the peers

struct TileMeta {
    int q0;       // first staged nonzero (16-byte aligned index)
    int ra0;      // first staged rowptr entry (aligned row index, relative to the range: may be < r0)
    int bulk;     // 1: the stage holds rowptr/col/val of this tile; 0: plain path
    int pad;
};

// NB on register budgets (tests/test_build_resources.py pins them): the plain single-vector variant needs 64 registers
// = 4 resident blocks per SM.  The fused-exchange (HALO) variants reach the same budgets as their plain counterparts
// because the push phase runs before anything else is live and the boundary tiles use a plain (not unrolled) row loop;
// forcing them there with __launch_bounds__(256, 4) instead cost 30 % (spills + fewer gathers in flight), and a
// constraint of 1 on the plain kernel made ptxas spend 101 registers and halve its occupancy (both measured, r02).
template <int NV, int BLOCK, int STAGES, bool HALO, int FUSE>
__global__ void __launch_bounds__(BLOCK) k_spmv_tma(SpmvArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    __shared__ TileMeta meta[STAGES];
    constexpr int RP = BLOCK + 8;                                    // staged rowptr entries per tile (aligned window)
    const size_t stage_bytes = (((size_t)a.cap * 12 + RP * 4) + 127) / 128 * 128;
    auto st_val = [&](int s) { return reinterpret_cast<double*>(smem_raw + (size_t)s * stage_bytes); };
    auto st_col = [&](int s) { return reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes + (size_t)a.cap * 8); };
    auto st_rp = [&](int s) { return reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes + (size_t)a.cap * 12); };
    constexpr int NW = BLOCK / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[3] = {0.0, 0.0, 0.0};
    const double c0 = (FUSE == 1 || FUSE == 2) ? ra.st->coef[2 * a.cj] : 0.0;       // zeta | alpha
    const double c1 = (FUSE == 1 || FUSE == 2) ? ra.st->coef[2 * a.cj + 1] : 0.0;   // eta  | beta
    if ((FUSE == 1 || FUSE == 2) && ra.dyn_cj >= 0) {
        // k lives on the device (adaptive): the host passed x0 = home and f_out = spare of the ping-pong pair; step cj
        // reads the home buffer iff (k - cj + 1) is even, so that the last step of the trip (cj == k) writes home, and
        // only that last step reduces r.r and runs the trip-end epilogue
        const int kk = ra.st->k;
        if ((kk - a.cj + 1) & 1) {
            const double* t = a.x0;
            a.x0 = a.f_out;
            a.f_out = const_cast<double*>(t);
        }
        if (ra.dyn_last) a.reduce = (a.cj == kk) ? 1 : 0;
    }
    // ---- halo exchange, part 1: push (before anything else of the kernel is live in registers) ----------------------------
    __shared__ int push_last;
    if (HALO) {
        const PkHaloPush* hp = a.hp;
        // The first PK_PUSH_BLOCKS blocks push (a few thousand entries each: 2 x 2 MiB planes at 512^3 / 8 ranks); the
        // others go straight to their tiles.  Every pushing block pays one system-scope fence before it may count
        // itself done — with ALL ~600 blocks pushing, those fences alone cost ~250 us per SpMV (measured, r02).
        const unsigned int n_push = gridDim.x < (unsigned)PK_PUSH_BLOCKS ? gridDim.x : (unsigned)PK_PUSH_BLOCKS;
        if (blockIdx.x < n_push && !a.halo_dry) {
            const unsigned long long hseq = *hp->seq + 1ull;
            const int bank = (int)(hseq & 1ull);
            const int P = hp->n_ranks;
            const long long total = hp->send_off[P];
            const long long stride = (long long)n_push * BLOCK;
            constexpr int U = 4;                   // loads of U entries in flight before the remote stores
            for (long long i0 = (long long)blockIdx.x * BLOCK + tid; i0 < total; i0 += U * stride) {
                double v0[U], v1[U];
                double* dst[U];
                long long nh[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long i = i0 + u * stride;
                    dst[u] = nullptr;
                    if (i < total) {
                        int q = 0;
                        while (i >= hp->send_off[q + 1]) ++q;
                        const long long kq = i - hp->send_off[q];
                        const long long src = hp->send_contig[q] ? (long long)hp->send_first[q] + kq : (long long)hp->send_idx[i];
                        nh[u] = hp->peer_nhalo[q];
                        dst[u] = hp->peer_recv[q] + PK_HALO_HDR + (size_t)(bank * 2) * nh[u] + hp->dst_off[q] + kq;
                        v0[u] = a.x0[src];
                        if (NV == 2) v1[u] = a.x1[src];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (dst[u]) {
                        dst[u][0] = v0[u];
                        if (NV == 2) dst[u][nh[u]] = v1[u];
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                pk_fence_sys();          // release, cumulative over the block's remote stores (bar.sync above)
                push_last = (atomicAdd(hp->ticket, 1u) == n_push - 1) ? 1 : 0;
            }
            __syncthreads();
            if (push_last && tid < P && ((hp->peer_mask >> tid) & 1u)) {
                // every pushing block's entries are on their way: raise this exchange's flag at every peer I talk to (also
                // the ones that only send to me — the flag is their licence to reuse this bank two exchanges from now)
                pk_fence_sys();
                volatile unsigned long long* f =
                    reinterpret_cast<volatile unsigned long long*>(hp->peer_recv[tid]) + bank * PK_MAX_RANKS + hp->me;
                *f = hseq;
            }
        }
    }
    // tile index space: tiles of [row_lo,row_hi), then of [row_lo2,row_hi2), then of [row_lo3,row_hi3)
    const long long tiles_a = a.row_hi > a.row_lo ? (a.row_hi - a.row_lo + BLOCK - 1) / BLOCK : 0;
    const long long tiles_b = a.row_hi2 > a.row_lo2 ? (a.row_hi2 - a.row_lo2 + BLOCK - 1) / BLOCK : 0;
    const long long tiles_c = a.row_hi3 > a.row_lo3 ? (a.row_hi3 - a.row_lo3 + BLOCK - 1) / BLOCK : 0;
    const long long n_tiles = tiles_a + tiles_b + tiles_c;
    auto tile_rows = [&](long long tile, long long& r0, long long& rend) {
        if (tile < tiles_a) { r0 = a.row_lo + tile * BLOCK; rend = a.row_hi; }
        else if (tile < tiles_a + tiles_b) { r0 = a.row_lo2 + (tile - tiles_a) * BLOCK; rend = a.row_hi2; }
        else { r0 = a.row_lo3 + (tile - tiles_a - tiles_b) * BLOCK; rend = a.row_hi3; }
    };
    // Tile -> block assignment.  strided (default): tile = b + i*grid, i.e. the grid is one moving window over A.
    // blocked (PK_TILE_ORDER=blocked): block b owns the contiguous tiles [b*chunk, (b+1)*chunk), which turns the
    // +-(one grid line) gathers into L1 hits but splits HBM traffic into hundreds of streams — measured on B200 at
    // 256^3: 301 us vs 254 us strided, so it stays an experiment switch.
    const long long chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long first_tile = a.blocked ? (long long)blockIdx.x * chunk : (long long)blockIdx.x;
    const long long tile_step = a.blocked ? 1 : (long long)gridDim.x;
    const long long my_tiles = a.blocked
        ? (first_tile < n_tiles ? (first_tile + chunk <= n_tiles ? chunk : n_tiles - first_tile) : 0)
        : (n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned long long l2pol = l2_evict_first_policy();
    // thread 0: issue the bulk copies of my i-th tile into stage i % STAGES (endpoints nb/ne already loaded)
    auto issue = [&](long long i, int nb, int ne) {
        const int s = (int)(i % STAGES);
        const long long tile = first_tile + i * tile_step;
        long long r0, rend;
        tile_rows(tile, r0, rend);
        const int nr = (int)((rend - r0) < BLOCK ? (rend - r0) : BLOCK);
        const int q0 = ((nb + a.base_mis) & ~3) - a.base_mis;     // aligned in memory; may be -1..-3 at a segment start
        const int cnt = (ne - q0 + 3) & ~3;
        const long long ra0 = r0 & ~3LL;
        const int rcnt = (int)((r0 - ra0) + nr + 1 + 3) & ~3;
        const bool ok = q0 >= 0 && cnt <= a.cap && (long long)q0 + cnt <= a.nnz_total && ra0 + rcnt <= a.rowptr_len;
        meta[s].q0 = q0;
        meta[s].ra0 = (int)(r0 - ra0);
        meta[s].bulk = ok ? 1 : 0;
        if (ok) {
            mbar_expect_tx(&full[s], (unsigned)(cnt * 12 + rcnt * 4));
            if (a.evict_first) {
                bulk_g2s_hint(st_rp(s), a.rowptr + ra0, (unsigned)(rcnt * 4), &full[s], l2pol);
                if (cnt > 0) {
                    bulk_g2s_hint(st_col(s), a.col + q0, (unsigned)(cnt * 4), &full[s], l2pol);
                    bulk_g2s_hint(st_val(s), a.val + q0, (unsigned)(cnt * 8), &full[s], l2pol);
                }
            } else {
                bulk_g2s(st_rp(s), a.rowptr + ra0, (unsigned)(rcnt * 4), &full[s]);
                if (cnt > 0) {
                    bulk_g2s(st_col(s), a.col + q0, (unsigned)(cnt * 4), &full[s]);
                    bulk_g2s(st_val(s), a.val + q0, (unsigned)(cnt * 8), &full[s]);
                }
            }
        } else {
            mbar_arrive(&full[s]);
        }
    };
    auto endpoints = [&](long long i, int& nb, int& ne) {
        const long long tile = first_tile + i * tile_step;
        long long r0, rend;
        tile_rows(tile, r0, rend);
        const long long r1 = (r0 + BLOCK < rend) ? r0 + BLOCK : rend;
        nb = __ldg(a.rowptr + r0);
        ne = __ldg(a.rowptr + r1);
    };

    int nb = 0, ne = 0;          // endpoints of the next tile to issue (thread 0)
    if (tid == 0) {
        for (long long i = 0; i < STAGES - 1 && i < my_tiles; ++i) {
            endpoints(i, nb, ne);
            issue(i, nb, ne);
        }
        if (STAGES - 1 < my_tiles) endpoints(STAGES - 1, nb, ne);
    }

    // ---- halo exchange, part 1: push.  The bulk copies of the first tiles are already in flight.  Sequence number of
    // this exchange = completed exchanges + 1 (the counter is advanced by the last block to LEAVE the kernel, so every
    // block reads the same value); its parity selects the bank of the peers' receive buffers.
    bool halo_ready = true;            // false until this block has seen the peers' flags of this exchange
    // (the sequence number and the receive-buffer pointers are recomputed where they are needed instead of being kept
    // in registers across the tile loop: the interior tiles must not pay for the exchange)
    auto wait_halo = [&]() {           // all threads of the block
        const unsigned long long hseq = *a.hp->seq + 1ull;
        if (!a.halo_dry && tid < a.hp->n_ranks && ((a.hp->peer_mask >> tid) & 1u)) {
            const volatile unsigned long long* f =
                reinterpret_cast<const volatile unsigned long long*>(a.hrecv) + (int)(hseq & 1ull) * PK_MAX_RANKS + tid;
            if (!pk_spin_until(f, hseq)) { ra.st->done = 1; ra.st->converged = 0; ra.st->guard = -2; }
            pk_fence_sys();            // acquire: the entries the flag announces are visible to this thread ...
        }
        __syncthreads();               // ... and, through the barrier, to the whole block
        halo_ready = true;
    };
    if (HALO) halo_ready = false;
    const int n_own = (int)a.n_own;

    // what a thread does with its finished row: store y (+ dots), or the fused k-skip step
    auto finish_row = [&](long long row, double sum0, double sum1, double wi, double fa, double fb, double fx, double xr) {
        if (FUSE == 0) {
            a.y0[row] = sum0;
            if (NV == 2) a.y1[row] = sum1;
            if (a.w) {
                acc[0] += wi * sum0;
                acc[1] += sum0 * sum0;
                acc[2] += wi * wi;
            }
        } else if (FUSE == 3) {           // Chebyshev basis level: T_{j+1} = 2 Ah T_j - T_{j-1}, Ah = (A - d) / c
            // fa / fb: previous levels (0 when absent); xr / fx: the multiplied vectors' own entries of this row
            a.y0[row] = (a.cs[0] * sum0 + a.cs[1] * xr) + a.cs[2] * fa;
            if (NV == 2) a.y1[row] = (a.cs[3] * sum1 + a.cs[4] * fx) + a.cs[5] * fb;
        } else if (FUSE == 1) {           // kskipmrr.py:65-69 / :89-93 with (zeta, eta) = (c0, c1), Ar1 = sum0
            const double ay = c1 * fa + c0 * sum0;
            const double zz = c1 * fb - c0 * xr;
            const double rn = xr - ay;
            a.f_a[row] = ay;
            a.f_b[row] = zz;
            a.f_out[row] = rn;
            a.f_x[row] = fx - zz;
            acc[0] += rn * rn;
        } else {                          // kskipcg.py:53-55 / :69-71 with (alpha, beta) = (c0, c1), Ap1 = sum0
            a.f_x[row] = fx + c0 * xr;
            const double rn = fa - c0 * sum0;
            a.f_a[row] = rn;
            a.f_out[row] = rn + c1 * xr;
            acc[0] += rn * rn;
        }
    };

    // Row phase of one tile.  BND (boundary tile of the fused exchange): columns >= n_own are halo entries and are read
    // from the receive buffer with L2 loads (they were written by a peer during this kernel); everything else gathers x
    // through the read-only path.
    auto row_phase = [&](auto bnd_tag, bool staged, int nr, long long r0, const int* rp, int rofs, int q0,
                         const int* scol, const double* sval) {
        constexpr bool BND = decltype(bnd_tag)::value;
        const double* h0 = nullptr;    // halo entries of the two vectors in the receive buffer, indexed by the column itself
        const double* h1 = nullptr;
        if (BND) {
            const int bank = (int)((*a.hp->seq + 1ull) & 1ull);
            h0 = a.hrecv + PK_HALO_HDR + (size_t)(bank * 2 + 0) * a.n_halo - a.n_own;
            h1 = a.hrecv + PK_HALO_HDR + (size_t)(bank * 2 + 1) * a.n_halo - a.n_own;
        }
        if (staged) {
            if (tid < nr) {
                const int sb = rp[rofs + tid] - q0, se = rp[rofs + tid + 1] - q0;
                const long long row = r0 + tid;
                // operands of the epilogue are requested before the row loop: their latency is hidden behind it
                const double wi = (FUSE == 0 && a.w) ? __ldg(a.w + row) : 0.0;
                const double fa = (FUSE == 3) ? (a.f_a ? __ldg(a.f_a + row) : 0.0) : (FUSE ? a.f_a[row] : 0.0);
                const double fb = (FUSE == 3) ? ((NV == 2 && a.f_b) ? __ldg(a.f_b + row) : 0.0) : ((FUSE == 1) ? a.f_b[row] : 0.0);
                const double fx = (FUSE == 3) ? (NV == 2 ? __ldg(a.x1 + row) : 0.0) : (FUSE ? a.f_x[row] : 0.0);
                const double xr = FUSE ? __ldg(a.x0 + row) : 0.0;
                double sum0 = 0.0, sum1 = 0.0;
                if (!BND) {
                    constexpr int UNR = 8;
                    for (int j = sb; j < se; j += UNR) {
                        double vv[UNR], xa[UNR], xb[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            const bool ok = j + u < se;
                            const int c = ok ? scol[j + u] : 0;
                            vv[u] = ok ? sval[j + u] : 0.0;
                            xa[u] = ok ? __ldg(a.x0 + c) : 0.0;
                            if (NV == 2) xb[u] = ok ? __ldg(a.x1 + c) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < UNR; ++u) {
                            if (j + u < se) {
                                sum0 += vv[u] * xa[u];
                                if (NV == 2) sum1 += vv[u] * xb[u];
                            }
                        }
                    }
                } else {
                    // boundary tiles are a few per cent of the tiles: a plain loop keeps the register footprint of the
                    // fused-exchange variant at that of the plain kernel (the interior path above is what must be fast)
#pragma unroll 1
                    for (int j = sb; j < se; ++j) {
                        const int c = scol[j];
                        const double v = sval[j];
                        const bool own = c < n_own;
                        sum0 += v * (own ? __ldg(a.x0 + c) : __ldcg(h0 + c));
                        if (NV == 2) sum1 += v * (own ? __ldg(a.x1 + c) : __ldcg(h1 + c));
                    }
                }
                finish_row(row, sum0, sum1, wi, fa, fb, fx, xr);
            }
        } else {
            for (int r = warp; r < nr; r += NW) {
                const int sb = rp[r], se = rp[r + 1];
                double sum0 = 0.0, sum1 = 0.0;
                for (int q = sb + lane; q < se; q += 32) {
                    const int c = a.col[q];
                    const double v = a.val[q];
                    if (!BND || c < n_own) {
                        sum0 += v * __ldg(a.x0 + c);
                        if (NV == 2) sum1 += v * __ldg(a.x1 + c);
                    } else {
                        sum0 += v * __ldcg(h0 + c);
                        if (NV == 2) sum1 += v * __ldcg(h1 + c);
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    sum0 += __shfl_down_sync(0xffffffffu, sum0, off);
                    if (NV == 2) sum1 += __shfl_down_sync(0xffffffffu, sum1, off);
                }
                if (lane == 0) {
                    const long long row = r0 + r;
                    if (FUSE == 3)
                        finish_row(row, sum0, sum1, 0.0, a.f_a ? a.f_a[row] : 0.0, (NV == 2 && a.f_b) ? a.f_b[row] : 0.0,
                                   NV == 2 ? a.x1[row] : 0.0, a.x0[row]);
                    else
                        finish_row(row, sum0, sum1, (FUSE == 0 && a.w) ? a.w[row] : 0.0, FUSE ? a.f_a[row] : 0.0,
                                   (FUSE == 1) ? a.f_b[row] : 0.0, FUSE ? a.f_x[row] : 0.0, FUSE ? a.x0[row] : 0.0);
                }
            }
        }
    };

    for (long long i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % STAGES);
        if (tid == 0) {
            const long long nxt = i + STAGES - 1;
            if (nxt < my_tiles) {
                issue(nxt, nb, ne);
                if (nxt + 1 < my_tiles) endpoints(nxt + 1, nb, ne);   // consumed next iteration: latency hidden
            }
        }
        mbar_wait(&full[s], (unsigned)((i / STAGES) & 1));
        const long long tile = first_tile + i * tile_step;
        long long r0, rend;
        tile_rows(tile, r0, rend);
        const int nr = (int)((rend - r0) < BLOCK ? (rend - r0) : BLOCK);
        double* sval = st_val(s);
        int* scol = st_col(s);
        int* rp = st_rp(s);
        int q0 = meta[s].q0;
        int rofs = meta[s].ra0;
        bool staged = meta[s].bulk != 0;
        if (!staged) {
            // plain fetch of this tile (array tail or long rows)
            rofs = 0;
            for (int t = tid; t <= nr; t += BLOCK) rp[t] = a.rowptr[r0 + t];
            __syncthreads();
            const int base = rp[0], end = rp[nr];
            q0 = base;
            if (end - base <= a.cap) {
                for (int q = base + tid; q < end; q += BLOCK) {
                    scol[q - base] = a.col[q];
                    sval[q - base] = a.val[q];
                }
                staged = true;
            }
            __syncthreads();
        }
        // boundary tile of the fused exchange: the peers' entries must have landed before its halo columns are read
        const bool bnd = HALO && tile >= tiles_a;
        if (bnd && !halo_ready) wait_halo();
        if (bnd) row_phase(std::true_type{}, staged, nr, r0, rp, rofs, q0, scol, sval);
        else row_phase(std::false_type{}, staged, nr, r0, rp, rofs, q0, scol, sval);
        __syncthreads();   // everyone is done with stage s: thread 0 may refill it next iteration
    }
    if (HALO) {
        // The last block to leave advances the exchange counter (every block read it at its start) and re-arms the push
        // ticket.  It must itself have seen the peers' flags: their arrival is what licenses the NEXT exchange to reuse
        // the other bank at the peers, so a rank without boundary rows (pure sender) still consumes the flags here.
        __shared__ int leave_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            leave_last = (atomicAdd(a.hp->ticket_done, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (leave_last) {
            if (!halo_ready) wait_halo();
            if (tid == 0) {
                *a.hp->ticket = 0u;
                *a.hp->ticket_done = 0u;
                *a.hp->seq = *a.hp->seq + 1ull;
                __threadfence();
            }
        }
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

// --------------------------------------------------------------------------------------------------------------
// Pattern-compressed operator (opt-in): one 16-bit id per row selects a (offsets, values) sequence from a table held in
// shared memory.  Thread per row, left-to-right accumulation with the table's values — the same products and the same
// order as the CSR kernels, hence the same bits — but the matrix costs 2 bytes per row instead of 12 per nonzero.
// Neighbouring rows almost always share a pattern, so a warp's gathers are fully coalesced.
struct PatArgs {
    const uint16_t* id;
    const int32_t* ptr;
    const int32_t* off;
    const double* val;
    int n_pat, n_ent;
};

template <int NV, int FUSE>
__global__ void __launch_bounds__(256) k_spmv_pat(PatArgs pa, SpmvArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    extern __shared__ __align__(16) unsigned char psm[];
    double* tval = reinterpret_cast<double*>(psm);                 // [n_ent]
    int* toff = reinterpret_cast<int*>(tval + pa.n_ent);           // [n_ent]
    int* tptr = toff + pa.n_ent;                                   // [n_pat + 1]
    for (int t = threadIdx.x; t < pa.n_ent; t += 256) { tval[t] = pa.val[t]; toff[t] = pa.off[t]; }
    for (int t = threadIdx.x; t <= pa.n_pat; t += 256) tptr[t] = pa.ptr[t];
    __syncthreads();
    double acc[3] = {0.0, 0.0, 0.0};
    const double c0 = FUSE ? ra.st->coef[2 * a.cj] : 0.0;
    const double c1 = FUSE ? ra.st->coef[2 * a.cj + 1] : 0.0;
    if (FUSE && ra.dyn_cj >= 0) {          // see k_spmv_tma: device-resident k picks the side of the ping-pong pair
        const int kk = ra.st->k;
        if ((kk - a.cj + 1) & 1) {
            const double* t = a.x0;
            a.x0 = a.f_out;
            a.f_out = const_cast<double*>(t);
        }
        if (ra.dyn_last) a.reduce = (a.cj == kk) ? 1 : 0;
    }
    const long long na = a.row_hi > a.row_lo ? a.row_hi - a.row_lo : 0;
    const long long nb = a.row_hi2 > a.row_lo2 ? a.row_hi2 - a.row_lo2 : 0;
    // R rows per thread are in flight together (ids, epilogue operands and gathers of all R rows are independent):
    // this kernel moves ~20 bytes per row, so it needs thousands of rows in flight per SM to cover DRAM latency.
    constexpr int R = (FUSE == 0 && NV == 1) ? 4 : 2;
    // blocked assignment (see k_spmv_tma): block b owns rows [b*span, (b+1)*span), walked 256 at a time; the R rows a
    // thread has in flight are 256 apart, i.e. consecutive sub-tiles of the same block
    const long long total = na + nb;
    const long long span = a.blocked ? ((total + gridDim.x - 1) / gridDim.x + 255) / 256 * 256 : 0;
    const long long stride = a.blocked ? 256 : (long long)gridDim.x * 256;
    const long long t_begin = a.blocked ? (long long)blockIdx.x * span + threadIdx.x : (long long)blockIdx.x * 256 + threadIdx.x;
    const long long t_end = a.blocked ? (((long long)blockIdx.x + 1) * span < total ? ((long long)blockIdx.x + 1) * span : total) : total;
    for (long long t0 = t_begin; t0 < t_end; t0 += R * stride) {
        long long row[R];
        int sb[R], len[R];
        double wi[R], fa[R], fb[R], fx[R], xr[R], sum0[R], sum1[R];
        int id[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long t = t0 + r * stride;
            const bool ok = t < t_end;
            row[r] = ok ? (t < na ? a.row_lo + t : a.row_lo2 + (t - na)) : -1;
            id[r] = ok ? (int)pa.id[row[r]] : 0;
        }
        int maxlen = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = row[r] >= 0;
            sb[r] = tptr[id[r]];
            len[r] = ok ? tptr[id[r] + 1] - sb[r] : 0;
            maxlen = len[r] > maxlen ? len[r] : maxlen;
            wi[r] = (ok && FUSE == 0 && a.w) ? __ldg(a.w + row[r]) : 0.0;
            fa[r] = (ok && FUSE) ? a.f_a[row[r]] : 0.0;
            fb[r] = (ok && FUSE == 1) ? a.f_b[row[r]] : 0.0;
            fx[r] = (ok && FUSE) ? a.f_x[row[r]] : 0.0;
            xr[r] = (ok && FUSE) ? __ldg(a.x0 + row[r]) : 0.0;
            sum0[r] = 0.0;
            sum1[r] = 0.0;
        }
#pragma unroll 4
        for (int j = 0; j < maxlen; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (j < len[r]) {                     // left-to-right within each row: the CSR / scipy order
                    const long long c = row[r] + toff[sb[r] + j];
                    const double v = tval[sb[r] + j];
                    sum0[r] += v * __ldg(a.x0 + c);
                    if (NV == 2) sum1[r] += v * __ldg(a.x1 + c);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row[r] < 0) continue;
            const long long rw = row[r];
            if (FUSE == 0) {
                a.y0[rw] = sum0[r];
                if (NV == 2) a.y1[rw] = sum1[r];
                if (a.w) {
                    acc[0] += wi[r] * sum0[r];
                    acc[1] += sum0[r] * sum0[r];
                    acc[2] += wi[r] * wi[r];
                }
            } else if (FUSE == 1) {
                const double ay = c1 * fa[r] + c0 * sum0[r];
                const double zz = c1 * fb[r] - c0 * xr[r];
                const double rn = xr[r] - ay;
                a.f_a[rw] = ay;
                a.f_b[rw] = zz;
                a.f_out[rw] = rn;
                a.f_x[rw] = fx[r] - zz;
                acc[0] += rn * rn;
            } else {
                a.f_x[rw] = fx[r] + c0 * xr[r];
                const double rn = fa[r] - c0 * sum0[r];
                a.f_a[rw] = rn;
                a.f_out[rw] = rn + c1 * xr[r];
                acc[0] += rn * rn;
            }
        }
    }
    if (a.reduce) pk_grid_reduce<3, 256>(acc, ra);
}

// one 64-bit hash per row over its (col - row, value bits) sequence; equal rows <=> equal hashes up to collisions,
// which k_pattern_verify rules out exactly
__global__ void k_row_hash(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                           const double* __restrict__ val, long long n_rows, unsigned long long* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(rowptr[r + 1] - rowptr[r]);
        for (int q = rowptr[r]; q < rowptr[r + 1]; ++q) {
            const unsigned long long o = (unsigned long long)(long long)(col[q] - (int)r);
            const unsigned long long v = (unsigned long long)__double_as_longlong(val[q]);
            h ^= o + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xBF58476D1CE4E5B9ull;
            h ^= v + 0x94D049BB133111EBull + (h << 6) + (h >> 2);
            h *= 0x94D049BB133111EBull;
            h ^= h >> 29;
        }
        out[r] = h;
    }
}

// every row must reproduce its pattern exactly (offsets and value bits); mismatches are counted
__global__ void k_pattern_verify(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const double* __restrict__ val, long long n_rows, PatArgs pa, int* __restrict__ bad) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        const int id = pa.id[r];
        bool ok = id < pa.n_pat;
        if (ok) {
            const int sb = pa.ptr[id], se = pa.ptr[id + 1];
            ok = (se - sb) == (rowptr[r + 1] - rowptr[r]);
            for (int j = 0; ok && j < se - sb; ++j) {
                const int q = rowptr[r] + j;
                ok = (col[q] - (int)r) == pa.off[sb + j] &&
                     __double_as_longlong(val[q]) == __double_as_longlong(pa.val[sb + j]);
            }
        }
        if (!ok) atomicAdd(bad, 1);
    }
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

// A warp owns R = 4 consecutive rows at a time: every x (and x1) element it loads is used for 4 rows from registers, so
// the vector traffic through L1/L2 is a quarter of the one-row-per-warp form (with two right-hand sides the two vectors
// no longer fit L1 next to each other and the old kernel was L2-bound at 55 % of HBM peak), and 4 independent 128-bit
// loads of A are in flight per lane and iteration.  Per row the lane partition and the order of additions are unchanged
// (lane l sums columns 2l, 2l+1, 2l+64, ... then the shuffle tree), so results are bit-identical to the one-row form.
template <int NV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_gemv(GemvArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    constexpr int NW = BLOCK / 32;
    constexpr int R = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long n_groups = (a.n_rows + R - 1) / R;
    for (long long g = (long long)blockIdx.x * NW + warp; g < n_groups; g += (long long)gridDim.x * NW) {
        const long long row0 = g * R;
        const double* ar[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r < a.n_rows ? row0 + r : a.n_rows - 1;     // clamp: tail rows are discarded below
            ar[r] = a.A + row * a.lda;
        }
        double s0[R], s1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { s0[r] = 0.0; s1[r] = 0.0; }
        if (a.vec) {
            const long long n2 = a.n_cols >> 1;
            const double2* x02 = reinterpret_cast<const double2*>(a.x0);
            const double2* x12 = reinterpret_cast<const double2*>(a.x1);
#pragma unroll 2
            for (long long c = lane; c < n2; c += 32) {
                double2 av[R];
#pragma unroll
                for (int r = 0; r < R; ++r) av[r] = __ldg(reinterpret_cast<const double2*>(ar[r]) + c);
                const double2 xv = __ldg(x02 + c);
                double2 xw = make_double2(0.0, 0.0);
                if (NV == 2) xw = __ldg(x12 + c);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    s0[r] += av[r].x * xv.x;
                    s0[r] += av[r].y * xv.y;
                    if (NV == 2) {
                        s1[r] += av[r].x * xw.x;
                        s1[r] += av[r].y * xw.y;
                    }
                }
            }
            if ((a.n_cols & 1) && lane == 0) {
                const long long c = a.n_cols - 1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    s0[r] += ar[r][c] * a.x0[c];
                    if (NV == 2) s1[r] += ar[r][c] * a.x1[c];
                }
            }
        } else {
            for (long long c = lane; c < a.n_cols; c += 32) {
                const double xv = a.x0[c];
                const double xw = (NV == 2) ? a.x1[c] : 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double av = ar[r][c];
                    s0[r] += av * xv;
                    if (NV == 2) s1[r] += av * xw;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s0[r] += __shfl_down_sync(0xffffffffu, s0[r], off);
                if (NV == 2) s1[r] += __shfl_down_sync(0xffffffffu, s1[r], off);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long row = row0 + r;
                if (row >= a.n_rows) break;
                a.y0[row] = s0[r];
                if (NV == 2) a.y1[row] = s1[r];
                if (a.w) {
                    const double wi = a.w[row];
                    acc[0] += wi * s0[r];
                    acc[1] += s0[r] * s0[r];
                    acc[2] += wi * wi;
                }
            }
        }
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

// Two right-hand sides: x0 and x1 together (2 x 8 n_cols bytes) no longer fit L1 next to each other, every warp walks
// the vectors at its own pace, and the kernel above becomes L2-bound (55 % of HBM peak at 16384^2).  Here the block's
// warps walk the columns TOGETHER: the current chunk of x0 / x1 is staged in shared memory by bulk copies (2-stage ring,
// one mbarrier per stage, issued by one thread while the previous chunk is consumed) and every warp reads it from there
// for its 4 rows.  Per row, the lane partition and the order of additions are those of k_gemv (bit-identical results).
template <int NV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_gemv_staged(GemvArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    constexpr int NW = BLOCK / 32;
    constexpr int R = 4;
    constexpr int CW = 2048;                        // columns per chunk (16 KB per vector and stage)
    constexpr int STAGES = 2;
    extern __shared__ __align__(128) unsigned char gsm_raw[];
    double* xs = reinterpret_cast<double*>(gsm_raw);                 // [STAGES][NV][CW]
    __shared__ __align__(8) unsigned long long full[STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long n_groups = (a.n_rows + (long long)R * NW - 1) / ((long long)R * NW);    // 32 rows per block and pass
    const long long n_chunks = (a.n_cols + CW - 1) / CW;
    const long long my_groups = n_groups > blockIdx.x ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = my_groups * n_chunks;                    // chunk visits of this block
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long v) {                                  // thread 0: chunk visit v -> stage v % STAGES
        const int s = (int)(v % STAGES);
        const long long c0 = (v % n_chunks) * CW;
        const unsigned bytes = (unsigned)((a.n_cols - c0 < CW ? a.n_cols - c0 : CW) * sizeof(double));
        mbar_expect_tx(&full[s], bytes * NV);
        bulk_g2s(xs + ((size_t)s * NV + 0) * CW, a.x0 + c0, bytes, &full[s]);
        if (NV == 2) bulk_g2s(xs + ((size_t)s * NV + 1) * CW, a.x1 + c0, bytes, &full[s]);
    };
    if (tid == 0)
        for (long long v = 0; v < STAGES - 1 && v < total; ++v) issue(v);
    long long v = 0;
    for (long long gi = 0; gi < my_groups; ++gi) {
        const long long g = blockIdx.x + gi * (long long)gridDim.x;
        const long long row0 = (g * NW + warp) * R;
        const double2* ar2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = row0 + r < a.n_rows ? row0 + r : a.n_rows - 1;     // clamp: tail rows are discarded below
            ar2[r] = reinterpret_cast<const double2*>(a.A + row * a.lda);
        }
        double s0[R], s1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { s0[r] = 0.0; s1[r] = 0.0; }
        // Software pipeline over the whole row, independent of the x chunks: two register buffers of A (this lane's
        // columns c2 and c2 + 32, as double2); while one is consumed the other's loads — and the reload of the first —
        // are in flight.  (ncu r02: with the loads left to the compiler the 80-register budget made it load two
        // vectors, wait, compute, load two more: 62 % of HBM peak, every sample a long-scoreboard stall on the first DMUL.)
        const long long n2 = a.n_cols >> 1;
        long long c2 = lane;
        double2 a0[R], a1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a0[r] = (c2 < n2) ? __ldg(ar2[r] + c2) : make_double2(0.0, 0.0);
            a1[r] = (c2 + 32 < n2) ? __ldg(ar2[r] + c2 + 32) : make_double2(0.0, 0.0);
        }
        auto consume = [&](const double2 (&av)[R], const double2 xv, const double2 xw) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                s0[r] += av[r].x * xv.x;
                s0[r] += av[r].y * xv.y;
                if (NV == 2) {
                    s1[r] += av[r].x * xw.x;
                    s1[r] += av[r].y * xw.y;
                }
            }
        };
        for (long long ch = 0; ch < n_chunks; ++ch, ++v) {
            const int s = (int)(v % STAGES);
            if (tid == 0 && v + STAGES - 1 < total) issue(v + STAGES - 1);
            mbar_wait(&full[s], (unsigned)((v / STAGES) & 1));
            const long long cb = (ch * CW) >> 1;                                     // first double2 column of the chunk
            const long long ce = (cb + CW / 2 < n2) ? cb + CW / 2 : n2;              // one past its last
            const double2* x02 = reinterpret_cast<const double2*>(xs + ((size_t)s * NV + 0) * CW);
            const double2* x12 = reinterpret_cast<const double2*>(xs + ((size_t)s * NV + 1) * CW);
            // (a pair c2, c2 + 32 never straddles a full-chunk boundary: chunks hold a multiple of 64 double2 columns)
            while (c2 < ce) {
                consume(a0, x02[c2 - cb], NV == 2 ? x12[c2 - cb] : make_double2(0.0, 0.0));
                if (c2 + 64 < n2) {
#pragma unroll
                    for (int r = 0; r < R; ++r) a0[r] = __ldg(ar2[r] + c2 + 64);
                }
                if (c2 + 32 < ce) {
                    consume(a1, x02[c2 + 32 - cb], NV == 2 ? x12[c2 + 32 - cb] : make_double2(0.0, 0.0));
                    if (c2 + 96 < n2) {
#pragma unroll
                        for (int r = 0; r < R; ++r) a1[r] = __ldg(ar2[r] + c2 + 96);
                    }
                }
                c2 += 64;
            }
            __syncthreads();          // the stage may be refilled by the next visit's issue
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s0[r] += __shfl_down_sync(0xffffffffu, s0[r], off);
                if (NV == 2) s1[r] += __shfl_down_sync(0xffffffffu, s1[r], off);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long row = row0 + r;
                if (row >= a.n_rows) break;
                a.y0[row] = s0[r];
                if (NV == 2) a.y1[row] = s1[r];
                if (a.w) {
                    const double wi = a.w[row];
                    acc[0] += wi * s0[r];
                    acc[1] += s0[r] * s0[r];
                    acc[2] += wi * wi;
                }
            }
        }
    }
    if (a.reduce) pk_grid_reduce<3, BLOCK>(acc, ra);
}

template <int NV, int BLOCK, bool VEC>
int stream_grid(pk_ctx* ctx, size_t smem, long long n_tiles) {
    const int per_sm = pk_blocks_per_sm((const void*)k_spmv_stream<NV, BLOCK, VEC>, BLOCK, smem);
    long long g = (long long)ctx->sm_count * per_sm;
    if (g > n_tiles) g = n_tiles;
    if (g < 1) g = 1;
    return (int)g;
}

// mode 0: choose the grid and launch; 1: dry run (only report the grid); 2: launch with the grid in *grid_io
template <int NV, int BLOCK, bool VEC>
int launch_stream(pk_ctx* ctx, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int grid_cap, int mode) {
    const size_t smem = (size_t)a.cap * (sizeof(double) + sizeof(int));
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

int launch_plain_any(pk_ctx* ctx, pk_mat* m, bool two, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap,
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

// ---- TMA variant launcher ------------------------------------------------------------------------------------------
template <int NV, int BLOCK, int STAGES, bool HALO, int FUSE>
int launch_tma(pk_ctx* ctx, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int grid_cap, int mode) {
    const size_t stage_bytes = (((size_t)a.cap * 12 + (BLOCK + 8) * 4) + 127) / 128 * 128;
    const size_t smem = stage_bytes * STAGES;
    const long long n_rows = a.row_hi - a.row_lo;
    const long long n_rows2 = a.row_hi2 > a.row_lo2 ? a.row_hi2 - a.row_lo2 : 0;
    const long long n_rows3 = a.row_hi3 > a.row_lo3 ? a.row_hi3 - a.row_lo3 : 0;
    // a fused-exchange launch (HALO) runs even without rows: it still has to push and to consume the peers' flags
    if (n_rows <= 0 && n_rows2 <= 0 && n_rows3 <= 0 && !HALO) { *grid_io = 0; return PK_OK; }
    const long long n_tiles = (n_rows > 0 ? (n_rows + BLOCK - 1) / BLOCK : 0) + (n_rows2 + BLOCK - 1) / BLOCK +
                              (n_rows3 + BLOCK - 1) / BLOCK;
    int grid = *grid_io;
    if (mode != 2) {
        const int per_sm = pk_blocks_per_sm((const void*)k_spmv_tma<NV, BLOCK, STAGES, HALO, FUSE>, BLOCK, smem);
        long long g = (long long)ctx->sm_count * per_sm;
        if (g > n_tiles) g = n_tiles;
        if (g > grid_cap) g = grid_cap;
        grid = (int)(g < 1 ? 1 : g);
        *grid_io = grid;
        if (mode == 1) return PK_OK;
    } else {
        pk_blocks_per_sm((const void*)k_spmv_tma<NV, BLOCK, STAGES, HALO, FUSE>, BLOCK, smem);   // sets the smem attribute
    }
    k_spmv_tma<NV, BLOCK, STAGES, HALO, FUSE><<<grid, BLOCK, smem, ctx->stream>>>(a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("spmv(tma) launch (grid %d, smem %zu): %s", grid, smem, cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    return PK_OK;
}

template <int NV, int BLOCK>
int launch_tma_stages(pk_ctx* ctx, int stages, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap, int mode) {
    if (NV == 2 && a.fuse == 3) {
        if (a.hrecv != nullptr) return launch_tma<2, BLOCK, 2, true, 3>(ctx, a, ra, grid_io, cap, mode);
        return launch_tma<2, BLOCK, 2, false, 3>(ctx, a, ra, grid_io, cap, mode);
    }
    if (NV == 1 && a.fuse == 1) {
        if (a.hrecv != nullptr) return launch_tma<1, BLOCK, 2, true, 1>(ctx, a, ra, grid_io, cap, mode);
        return launch_tma<1, BLOCK, 2, false, 1>(ctx, a, ra, grid_io, cap, mode);
    }
    if (NV == 1 && a.fuse == 2) {
        if (a.hrecv != nullptr) return launch_tma<1, BLOCK, 2, true, 2>(ctx, a, ra, grid_io, cap, mode);
        return launch_tma<1, BLOCK, 2, false, 2>(ctx, a, ra, grid_io, cap, mode);
    }
    if (a.hrecv != nullptr) return launch_tma<NV, BLOCK, 2, true, 0>(ctx, a, ra, grid_io, cap, mode);
    switch (stages) {
        case 3: return launch_tma<NV, BLOCK, 3, false, 0>(ctx, a, ra, grid_io, cap, mode);
        case 4: return launch_tma<NV, BLOCK, 4, false, 0>(ctx, a, ra, grid_io, cap, mode);
        default: return launch_tma<NV, BLOCK, 2, false, 0>(ctx, a, ra, grid_io, cap, mode);
    }
}

int launch_tma_any(pk_ctx* ctx, pk_mat* m, bool two, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap, int mode) {
    if (m->tile_rows == 128) {
        return two ? launch_tma_stages<2, 128>(ctx, m->stages, a, ra, grid_io, cap, mode)
                   : launch_tma_stages<1, 128>(ctx, m->stages, a, ra, grid_io, cap, mode);
    }
    return two ? launch_tma_stages<2, 256>(ctx, m->stages, a, ra, grid_io, cap, mode)
               : launch_tma_stages<1, 256>(ctx, m->stages, a, ra, grid_io, cap, mode);
}

template <int NV, int FUSE>
int launch_pat(pk_ctx* ctx, pk_mat* m, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int grid_cap, int mode) {
    PatArgs pa{m->pat_id, m->pat_ptr, m->pat_off, m->pat_val, m->n_pat, m->pat_entries};
    const size_t smem = (size_t)m->pat_entries * 12 + (size_t)(m->n_pat + 1) * 4 + 16;
    const long long n_rows = (a.row_hi > a.row_lo ? a.row_hi - a.row_lo : 0) + (a.row_hi2 > a.row_lo2 ? a.row_hi2 - a.row_lo2 : 0);
    if (n_rows <= 0) { *grid_io = 0; return PK_OK; }
    int grid = *grid_io;
    if (mode != 2) {
        const int per_sm = pk_blocks_per_sm((const void*)k_spmv_pat<NV, FUSE>, 256, smem);
        long long g = (long long)ctx->sm_count * per_sm;
        const long long want = (n_rows + 511) / 512;
        if (g > want) g = want;
        if (g > grid_cap) g = grid_cap;
        grid = (int)(g < 1 ? 1 : g);
        *grid_io = grid;
        if (mode == 1) return PK_OK;
    } else {
        pk_blocks_per_sm((const void*)k_spmv_pat<NV, FUSE>, 256, smem);
    }
    k_spmv_pat<NV, FUSE><<<grid, 256, smem, ctx->stream>>>(pa, a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("spmv(pattern) launch (grid %d, smem %zu): %s", grid, smem, cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    return PK_OK;
}

int launch_pat_any(pk_ctx* ctx, pk_mat* m, bool two, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap, int mode) {
    if (two) return launch_pat<2, 0>(ctx, m, a, ra, grid_io, cap, mode);
    if (a.fuse == 1) return launch_pat<1, 1>(ctx, m, a, ra, grid_io, cap, mode);
    if (a.fuse == 2) return launch_pat<1, 2>(ctx, m, a, ra, grid_io, cap, mode);
    return launch_pat<1, 0>(ctx, m, a, ra, grid_io, cap, mode);
}

int launch_stream_any(pk_ctx* ctx, pk_mat* m, bool two, const SpmvArgs& a, PkRedArgs ra, int* grid_io, int cap,
                      int mode) {
    if (m->pat_on && a.hrecv == nullptr) return launch_pat_any(ctx, m, two, a, ra, grid_io, cap, mode);
    if (m->use_tma) return launch_tma_any(ctx, m, two, a, ra, grid_io, cap, mode);
    return launch_plain_any(ctx, m, two, a, ra, grid_io, cap, mode);
}

int launch_gemv(pk_ctx* ctx, pk_mat* m, bool two, const GemvArgs& a, PkRedArgs ra) {
    constexpr int BLOCK = 256;
    (void)m;
    static int staged_mode = -1;                  // PK_GEMV=plain: never stage x; staged: also for one right-hand side
    if (staged_mode < 0) {
        const char* e = getenv("PK_GEMV");
        staged_mode = (e && strcmp(e, "plain") == 0) ? 0 : ((e && strcmp(e, "staged") == 0) ? 2 : 1);
    }
    // x staged in shared memory: pays when the vectors do not fit L1 (two right-hand sides, long rows); needs 16-byte
    // aligned rows and an even number of columns (bulk copies move multiples of 16 bytes)
    const bool staged = a.vec && (a.n_cols & 1) == 0 && a.n_cols >= 4096 && (staged_mode == 2 || (staged_mode == 1 && two));
    if (staged) {
        const size_t smem = (size_t)2 * (two ? 2 : 1) * 2048 * sizeof(double);
        const void* kern = two ? (const void*)k_gemv_staged<2, BLOCK> : (const void*)k_gemv_staged<1, BLOCK>;
        long long want = (a.n_rows + 31) / 32;
        long long cap = (long long)ctx->sm_count * pk_blocks_per_sm(kern, BLOCK, smem);
        if (cap > ctx->red.max_blocks) cap = ctx->red.max_blocks;
        int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        if (two) k_gemv_staged<2, BLOCK><<<grid, BLOCK, smem, ctx->stream>>>(a, ra);
        else k_gemv_staged<1, BLOCK><<<grid, BLOCK, smem, ctx->stream>>>(a, ra);
    } else {
        long long want = ((a.n_rows + 3) / 4 + (BLOCK / 32) - 1) / (BLOCK / 32);      // a warp takes 4 rows at a time
        long long cap = (long long)ctx->sm_count *
                        pk_blocks_per_sm(two ? (const void*)k_gemv<2, BLOCK> : (const void*)k_gemv<1, BLOCK>, BLOCK, 0);
        if (cap > ctx->red.max_blocks) cap = ctx->red.max_blocks;
        int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        if (two) k_gemv<2, BLOCK><<<grid, BLOCK, 0, ctx->stream>>>(a, ra);
        else k_gemv<1, BLOCK><<<grid, BLOCK, 0, ctx->stream>>>(a, ra);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("gemv launch: %s", cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    return PK_OK;
}

__global__ void k_tile_max(const int32_t* __restrict__ rowptr, long long n_rows, int tile_rows, int* out) {
    const long long n_tiles = (n_rows + tile_rows - 1) / tile_rows;
    int m = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += (long long)gridDim.x * blockDim.x) {
        const long long r0 = t * tile_rows, r1 = r0 + tile_rows < n_rows ? r0 + tile_rows : n_rows;
        const int c = rowptr[r1] - rowptr[r0];
        m = c > m ? c : m;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const int o = __shfl_down_sync(0xffffffffu, m, off);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// One-time structural check of a CSR block (the scipy-based reference raises on malformed input; silently reading out
// of bounds is not an option): bit 0 rowptr[0] != 0, bit 1 rowptr not monotone, bit 2 rowptr[n] != nnz,
// bit 3 a column index outside [0, n_cols).
template <class RP>
__global__ void k_csr_validate(const RP* __restrict__ rowptr, const int32_t* __restrict__ col, long long n_rows,
                               long long n_cols, long long nnz, int* __restrict__ bad) {
    int f = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 == 0) {
        if (rowptr[0] != 0) f |= 1;
        if ((long long)rowptr[n_rows] != nnz) f |= 4;
    }
    for (long long r = t0; r < n_rows; r += stride)
        if (rowptr[r + 1] < rowptr[r]) f |= 2;
    for (long long q = t0; q < nnz; q += stride) {
        const int c = col[q];
        if (c < 0 || (long long)c >= n_cols) f |= 8;
    }
    if (f) atomicOr(bad, f);
}

}  // namespace

__global__ void k_rebase_rowptr(const long long* __restrict__ rp64, long long base, long long count, int32_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = (int32_t)(rp64[i] - base);
}

int pk_rebase_rowptr(pk_ctx* ctx, const int64_t* rp64, long long base, long long count, int32_t* out) {
    int grid = (int)((count + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    k_rebase_rowptr<<<grid, 256, 0, ctx->stream>>>((const long long*)rp64, base, count, out);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int pk_csr_validate(pk_ctx* ctx, const void* rowptr, int rowptr64, const int32_t* col, long long n_rows,
                    long long n_cols, long long nnz, int* flags) {
    int* d = nullptr;
    PK_CUDA(cudaMalloc(&d, sizeof(int)));
    PK_CUDA(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
    long long work = nnz > n_rows ? nnz : n_rows;
    int grid = (int)((work + 1023) / 1024);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    if (rowptr64)
        k_csr_validate<long long><<<grid, 256, 0, ctx->stream>>>((const long long*)rowptr, col, n_rows, n_cols, nnz, d);
    else
        k_csr_validate<int32_t><<<grid, 256, 0, ctx->stream>>>((const int32_t*)rowptr, col, n_rows, n_cols, nnz, d);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(flags, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return PK_OK;
}

// Largest number of nonzeros in any tile of `tile_rows` consecutive rows (one small pass over rowptr, blocking).
int pk_tile_max_nnz(pk_ctx* ctx, const int32_t* rowptr, long long n_rows, int tile_rows, int* result) {
    int* d = nullptr;
    PK_CUDA(cudaMalloc(&d, sizeof(int)));
    PK_CUDA(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
    const long long n_tiles = (n_rows + tile_rows - 1) / tile_rows;
    int grid = (int)((n_tiles + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    k_tile_max<<<grid, 256, 0, ctx->stream>>>(rowptr, n_rows, tile_rows, d);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(result, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return PK_OK;
}

static int spmv_impl(pk_ctx* ctx, pk_mat* m, double* x, double* y, double* x1, double* y1, PkDots dots);

// ---- row-pattern compression: C-ABI ---------------------------------------------------------------------------------
extern "C" int pk_mat_row_hashes(pk_mat* m, uint64_t* d_hash) {
    PK_REQUIRE(m && d_hash, "null argument");
    PK_REQUIRE(m->kind != MAT_DENSE && m->segs.empty(), "row patterns apply to CSR blocks with 32-bit row pointers");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    int grid = (int)((m->n_rows + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    if (grid < 1) grid = 1;
    k_row_hash<<<grid, 256, 0, ctx->stream>>>(m->rowptr, m->col, m->val, m->n_rows, (unsigned long long*)d_hash);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_mat_set_patterns(pk_mat* m, int n_pat, int n_entries, const uint16_t* d_id, const int32_t* d_ptr,
                                   const int32_t* d_off, const double* d_val) {
    PK_REQUIRE(m && d_id && d_ptr && d_off && d_val, "null argument");
    PK_REQUIRE(m->kind != MAT_DENSE && m->segs.empty(), "row patterns apply to CSR blocks with 32-bit row pointers");
    PK_REQUIRE(n_pat >= 1 && n_pat <= 65535 && n_entries >= 0 && (size_t)n_entries * 12 + (size_t)(n_pat + 1) * 4 <= 64 * 1024,
               "pattern table too large for shared memory");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    // exact verification: every row must equal its pattern bit for bit, otherwise the CSR path stays in use
    PatArgs pa{d_id, d_ptr, d_off, d_val, n_pat, n_entries};
    int* d_bad = nullptr;
    int h_bad = 0;
    PK_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    PK_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    int grid = (int)((m->n_rows + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    if (grid < 1) grid = 1;
    k_pattern_verify<<<grid, 256, 0, ctx->stream>>>(m->rowptr, m->col, m->val, m->n_rows, pa, d_bad);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bad);
    if (h_bad != 0) {
        pk_set_error("pattern table does not reproduce %d rows; CSR path kept", h_bad);
        return PK_ERR_ARG;
    }
    m->pat_id = d_id; m->pat_ptr = d_ptr; m->pat_off = d_off; m->pat_val = d_val;
    m->n_pat = n_pat; m->pat_entries = n_entries;
    m->pat_on = true;
    return PK_OK;
}

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
    ra.defer = (ctx->n_ranks > 1 && !ctx->d_p2p && !ctx->nocomm) ? 1 : 0;
    ra.p2p = ctx->nocomm ? nullptr : ctx->d_p2p;
    ra.ar_n = dots.w ? 3 + dots.extra_sums : (((dots.fuse == 1 || dots.fuse == 2) && dots.epi != EPI_KS_STEP) ? 1 : 0);
    ra.g_off = -1;
    ra.red_off = 0;
    ra.post_only = 0;
    ra.block_off = 0;
    ra.nb_total = 0;
    ra.store_only = 0;
    ra.only_rollback = ctx->ctl_only_rollback;
    ra.dyn_cj = ctx->ctl_dyn_cj;
    ra.dyn_last = ctx->ctl_dyn_last;
    ctx->spmvs += two ? 2 : 1;

    if (m->kind == MAT_DENSE) {
        if (m->distributed) {
            PK_CHECK(pk_comm_halo_start(ctx, m, x, x1));
            PK_CHECK(pk_comm_halo_wait(ctx));
        }
        GemvArgs g;
        g.A = m->dense; g.lda = m->lda; g.n_rows = m->n_rows; g.n_cols = m->n_cols;
        g.x0 = x; g.x1 = x1; g.y0 = y; g.y1 = y1; g.w = dots.w; g.reduce = dots.w ? 1 : 0;
        g.vec = (((uintptr_t)m->dense & 15) == 0 && (m->lda & 1) == 0 && ((uintptr_t)x & 15) == 0 &&
                 (!x1 || ((uintptr_t)x1 & 15) == 0)) ? 1 : 0;
        PK_CHECK(launch_gemv(ctx, m, two, g, ra));
        if (dots.w) return pk_finish_reduce(ctx, 3 + dots.extra_sums, dots.epi, -1, 0);
        return PK_OK;
    }

    SpmvArgs a;
    a.rowptr = m->rowptr; a.col = m->col; a.val = m->val;
    a.x0 = x; a.x1 = x1; a.y0 = y; a.y1 = y1; a.w = dots.w;
    a.fuse = dots.fuse; a.cj = dots.cj; a.f_a = dots.f_a; a.f_b = dots.f_b; a.f_x = dots.f_x; a.f_out = dots.f_out;
    for (int i = 0; i < 6; ++i) a.cs[i] = dots.cs[i];
    if (dots.fuse == 3) PK_REQUIRE(two && m->use_tma && !m->pat_on && m->kind != MAT_DENSE,
                                   "the Chebyshev basis needs the TMA CSR kernel (16-byte aligned CSR arrays, no pattern compression)");
    a.nnz_total = m->nnz;
    a.rowptr_len = m->n_rows + 1;
    a.base_mis = 0;
    a.row_lo2 = a.row_hi2 = a.row_lo3 = a.row_hi3 = 0;
    a.hrecv = nullptr; a.hp = nullptr; a.halo_dry = 0; a.n_own = m->n_rows; a.n_halo = m->n_halo;
    a.cap = m->tile_cap;
    {
        static int ef = -1;
        if (ef < 0) { const char* e = getenv("PK_L2_HINT"); ef = e ? atoi(e) : 1; }
        a.evict_first = ef;
        static int bl = -1;
        if (bl < 0) { const char* e = getenv("PK_TILE_ORDER"); bl = (e && strcmp(e, "blocked") == 0) ? 1 : 0; }
        a.blocked = bl;
    }
    a.reduce = (dots.w || ((dots.fuse == 1 || dots.fuse == 2) && dots.epi != EPI_KS_STEP)) ? 1 : 0;
    int grid = 0;

    const bool exchange = m->distributed && (m->n_halo > 0 || (!m->send_off.empty() && m->send_off.back() > 0));
    if (!exchange && !m->segs.empty()) {
        // block with 64-bit row pointers: one launch per row segment (< 2^31 nonzeros each, 32-bit row pointer rebased to
        // the segment), all feeding ONE reduction through disjoint block slots like the interior / boundary launches
        const int S = (int)m->segs.size();
        std::vector<SpmvArgs> as(S, a);
        std::vector<int> grids(S, 0);
        int total = 0;
        for (int i = 0; i < S; ++i) {
            const PkSeg& sg = m->segs[i];
            as[i].rowptr = sg.rp32 - sg.row_lo;           // indexed by the absolute row
            as[i].col = m->col + sg.base;
            as[i].val = m->val + sg.base;
            as[i].nnz_total = sg.nnz;
            as[i].rowptr_len = sg.row_hi + 1;
            as[i].base_mis = (int)(sg.base & 3);
            as[i].row_lo = sg.row_lo;
            as[i].row_hi = sg.row_hi;
            PK_CHECK(launch_stream_any(ctx, m, two, as[i], ra, &grids[i], ctx->red.max_blocks / S, 1));
            total += grids[i];
        }
        int off = 0;
        for (int i = 0; i < S; ++i) {
            PkRedArgs rs = ra;
            rs.block_off = off;
            rs.nb_total = total;
            rs.store_only = (i + 1 < S) ? 1 : 0;
            if (grids[i] > 0) PK_CHECK(launch_stream_any(ctx, m, two, as[i], rs, &grids[i], ctx->red.max_blocks / S, 2));
            off += grids[i];
        }
    } else if (!exchange) {
        a.row_lo = 0; a.row_hi = m->n_rows;
        PK_CHECK(launch_stream_any(ctx, m, two, a, ra, &grid, ctx->red.max_blocks, 0));
    } else if (m->halo_p2p && m->use_tma && !m->pat_on) {
        // Exchange fused into the operator kernel (ONE launch, no NCCL, no side stream): every block pushes its share of
        // my boundary entries into the peers' receive buffers over NVLink, the interior tiles run while the peers'
        // pushes are in flight, the boundary tiles (scheduled last) wait for the peers' flags in-kernel.
        a.row_lo = m->interior_lo; a.row_hi = m->interior_hi;
        a.row_lo2 = 0; a.row_hi2 = m->interior_lo;
        a.row_lo3 = m->interior_hi; a.row_hi3 = m->n_rows;
        a.hrecv = m->d_recvbuf;
        a.hp = m->d_push;
        a.halo_dry = ctx->nocomm ? 1 : 0;      // "compute only": the same kernel without push and waits (results meaningless)
        PK_CHECK(launch_stream_any(ctx, m, two, a, ra, &grid, ctx->red.max_blocks, 0));
    } else {
        // Interior rows (no halo column) run while the halo of x is in flight on the side stream; the boundary
        // rows follow the exchange.  All launches store partials into disjoint block slots of ONE reduction,
        // which the last launch finishes (fixed slot order => deterministic).
        PK_CHECK(pk_comm_halo_start(ctx, m, x, x1));
        const long long lo = m->interior_lo, hi = m->interior_hi;
        const int cap_each = ctx->red.max_blocks / 2;
        SpmvArgs ai = a, ab = a;
        ai.row_lo = lo; ai.row_hi = hi;
        if (m->use_tma || m->pat_on) {       // one launch covers the rows above and below the interior
            ab.row_lo = 0; ab.row_hi = lo; ab.row_lo2 = hi; ab.row_hi2 = m->n_rows;
        } else {                             // plain kernel: single range per launch -> treat [0,lo) then [hi,n) separately
            ab.row_lo = 0; ab.row_hi = lo;
        }
        SpmvArgs ac = a;
        ac.row_lo = hi; ac.row_hi = (m->use_tma || m->pat_on) ? hi : m->n_rows;     // third launch only for the plain kernel
        int g_int = 0, g_bnd = 0, g_c = 0;
        PK_CHECK(launch_stream_any(ctx, m, two, ai, ra, &g_int, cap_each, 1));
        PK_CHECK(launch_stream_any(ctx, m, two, ab, ra, &g_bnd, cap_each / 2, 1));
        PK_CHECK(launch_stream_any(ctx, m, two, ac, ra, &g_c, cap_each / 2, 1));
        const int total = g_int + g_bnd + g_c;
        PkRedArgs r1 = ra, r2 = ra, r3 = ra;
        r1.block_off = 0; r1.nb_total = total; r1.store_only = (g_bnd + g_c > 0) ? 1 : 0;
        r2.block_off = g_int; r2.nb_total = total; r2.store_only = (g_c > 0) ? 1 : 0;
        r3.block_off = g_int + g_bnd; r3.nb_total = total; r3.store_only = 0;
        bool waited = false;
        if (g_int > 0) PK_CHECK(launch_stream_any(ctx, m, two, ai, r1, &g_int, cap_each, 2));
        PK_CHECK(pk_comm_halo_wait(ctx));
        waited = true;
        if (g_bnd > 0) PK_CHECK(launch_stream_any(ctx, m, two, ab, r2, &g_bnd, cap_each / 2, 2));
        if (g_c > 0) PK_CHECK(launch_stream_any(ctx, m, two, ac, r3, &g_c, cap_each / 2, 2));
        if (!waited) PK_CHECK(pk_comm_halo_wait(ctx));
    }
    if (dots.w) return pk_finish_reduce(ctx, 3 + dots.extra_sums, dots.epi, -1, 0);
    if ((dots.fuse == 1 || dots.fuse == 2) && dots.epi != EPI_KS_STEP) return pk_finish_reduce(ctx, 1, dots.epi, -1, 0);
    return PK_OK;
}

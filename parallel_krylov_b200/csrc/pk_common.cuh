// Internal definitions shared by the libpkrylov translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pkrylov.h"

// ---------------------------------------------------------------------------------------------------------------
// error handling
void pk_set_error(const char* fmt, ...);
#define PK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            pk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PK_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)
#define PK_CHECK(expr)                \
    do {                              \
        int _s = (expr);              \
        if (_s != PK_OK) return _s;   \
    } while (0)
#define PK_REQUIRE(cond, msg)                                        \
    do {                                                             \
        if (!(cond)) {                                               \
            pk_set_error("%s:%d: %s", __FILE__, __LINE__, msg);      \
            return PK_ERR_ARG;                                       \
        }                                                            \
    } while (0)

#include "pk_state.h"

struct PkReduce {           // scratch for the deterministic two-stage reductions
    double* partials;       // [PK_MAX_SUMS][max_blocks]
    unsigned int* ticket;   // last-block-done counter (self-resetting)
    int max_blocks;
};

struct PkComm;  // pk_comm.cu

// Peer-mapped mailboxes for the in-kernel all-reduce (pk_device.cuh: pk_grid_reduce).  One mailbox per rank, in that
// rank's HBM, opened by every peer through CUDA IPC; layout [bank 0..1][source rank][PK_MBOX_STRIDE doubles], the
// last slot of a row holding the sequence flag.
constexpr int PK_MAX_RANKS = 16;
constexpr int PK_MBOX_PAYLOAD = 6 * (PK_KMAX + 2) + 4;      // largest all-reduce: all Gram sums of a trip
constexpr int PK_MBOX_STRIDE = PK_MBOX_PAYLOAD + 4;         // + flag (8 bytes) + padding
constexpr size_t PK_MBOX_DOUBLES = (size_t)2 * PK_MAX_RANKS * PK_MBOX_STRIDE;

struct PkP2P {
    double* mbox[PK_MAX_RANKS];       // mbox[p]: rank p's mailbox as seen from this device
    unsigned long long* seq;          // reductions completed so far (device counter; identical on every rank)
    int n_ranks, rank;
};

struct pk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;     // halo exchange stream
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_poll[2] = {nullptr, nullptr};
    int sm_count = 148;
    PkReduce red{};
    PkState* d_state = nullptr;      // device state (one solve at a time per context, like the reference's MultiGpu)
    PkState* h_state = nullptr;      // pinned mirror for polling
    int* h_flags = nullptr;          // pinned: [slot][4]
    PkComm* comm = nullptr;
    PkP2P* d_p2p = nullptr;          // non-null: dots are all-reduced inside the reducing kernels (NVLink mailboxes)
    double* my_mbox = nullptr;
    std::vector<void*> peer_mbox;    // opened IPC mappings (to close)
    int n_ranks = 1, rank = 0;
    // predicates the launch helpers copy into PkRedArgs (set by the adaptive solver around its device-conditional launches)
    int ctl_only_rollback = 0;
    int ctl_dyn_cj = -1;
    int ctl_dyn_last = 0;
    bool nocomm = false;             // measurement aid: same kernels, no halo exchange / all-reduce (numerically meaningless)
    long long launches = 0;          // kernels launched (statistics)
    long long spmvs = 0;
    // optional per-launch timing of the operator kernel (bench.py's roofline): event pairs around each application
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
};

// Halo exchange by direct NVLink stores, fused into the operator kernel (no NCCL on the data path): at its start the
// SpMV kernel of the owner pushes the entries its peers need straight into their receive buffers (peer-mapped through
// CUDA IPC) and raises a sequence flag; the peer's boundary tiles, scheduled last in its own SpMV kernel, wait on the
// flags.  Receive buffer of a rank: [32 doubles of flags: flag[bank][source]] then data[bank 0..1][vector 0..1][n_halo].
// Flags go BOTH ways between any two ranks that exchange in at least one direction (peer_mask is symmetric), so a rank
// can never run two exchanges ahead of a peer: two banks suffice even for one-directional (nonsymmetric) patterns.
constexpr int PK_HALO_HDR = 32;
struct PkHaloPush {
    double* peer_recv[PK_MAX_RANKS];     // peer q's receive buffer as mapped on this device (own buffer at [me])
    long long send_off[PK_MAX_RANKS + 1];
    long long dst_off[PK_MAX_RANKS];     // where my entries start inside q's halo (q's recv_off[me])
    long long peer_nhalo[PK_MAX_RANKS];
    int send_first[PK_MAX_RANKS];
    int send_contig[PK_MAX_RANKS];
    const int32_t* send_idx;
    unsigned long long* seq;             // exchanges completed (advanced by the last block to leave the SpMV kernel)
    unsigned int* ticket;                // blocks that have finished their share of the push
    unsigned int* ticket_done;           // blocks that have left the kernel
    int n_ranks, me;
    unsigned int peer_mask;              // bit p set: rank p and I exchange entries in at least one direction
};

enum PkMatKind : int { MAT_CSR_STREAM = 0, MAT_CSR_VECTOR = 1, MAT_DENSE = 2 };

// A block with nnz >= 2^31 (64-bit row pointers from the caller) is applied as a few row SEGMENTS of < 2^31 nonzeros each:
// per segment a 32-bit row pointer REBASED to the segment's first nonzero (owned by the operator), so the kernels keep
// 32-bit offsets and stream 4 bytes of row pointer per row instead of 8.  Segments start at multiples of the tile height.
struct PkSeg {
    long long row_lo = 0, row_hi = 0;   // rows [row_lo, row_hi)
    long long base = 0;                 // index of the segment's first nonzero in col / val
    long long nnz = 0;
    int32_t* rp32 = nullptr;            // owned: rp32[r - row_lo] = rowptr64[r] - base, r = row_lo .. row_hi
};

struct pk_mat {
    pk_ctx* ctx = nullptr;
    int kind = MAT_CSR_STREAM;
    long long n_rows = 0, n_cols = 0, nnz = 0;
    const int32_t* rowptr = nullptr;     // nullptr when the block is segmented (64-bit row pointers)
    std::vector<PkSeg> segs;             // non-empty: apply segment by segment
    const int32_t* col = nullptr;
    const double* val = nullptr;
    const double* dense = nullptr;
    long long lda = 0;
    // csr-stream tiling
    int tile_rows = 256;     // rows per tile == threads per block
    int tile_cap = 2048;     // nnz capacity of the shared-memory product buffer (per right-hand side)
    bool vec_ok = true;      // rowptr/col/val 16-byte aligned: 128-bit loads / bulk copies
    bool use_tma = true;     // TMA-pipelined kernel (needs vec_ok)
    int stages = 3;          // ring depth of the TMA pipeline
    // halo plan
    bool distributed = false;
    long long n_halo = 0;
    std::vector<long long> send_off, recv_off;   // per peer
    const int32_t* d_send_idx = nullptr;
    std::vector<char> send_contig;               // per peer: list is a contiguous run
    std::vector<int32_t> send_first;             // per peer: first index of that run
    double* d_sendbuf = nullptr;                 // owned
    // NVLink push path (replaces ncclSend/Recv when the peers' receive buffers are mapped)
    bool halo_p2p = false;
    PkHaloPush push{};
    PkHaloPush* d_push = nullptr;                // owned: device copy of `push` (the SpMV kernel reads it)
    double* d_recvbuf = nullptr;                 // owned: [PK_HALO_HDR + 4 * n_halo]
    std::vector<void*> peer_recv_maps;           // IPC mappings to close
    long long interior_lo = 0, interior_hi = 0;
    long long ld = 0;
    // Row-pattern compression (opt-in, lossless): rows that share the same (column offsets relative to the row, values)
    // sequence are stored once in a small table; the matrix becomes one 16-bit pattern id per row.  Constant-coefficient
    // stencils have a few dozen patterns (27 for the 3-D 7-point Laplacian), so A shrinks from ~84 to 2 bytes per row.
    // one-pass matrix-powers kernel (pk_matpow.cu): structure probe, cached
    bool mp_checked = false;
    int mp_rmax = 0;         // longest row
    int mp_bw = 0;           // half bandwidth max |col - row|
    // row-partitioned matrix powers (pk_mat_set_matpow_ext): ghost rows of A from the two neighbours, level-0 ghost inputs
    bool mp_ext = false;
    bool mp_dense = false;   // every row holds its full band: the two-rows-per-thread kernel applies
    // Row-partitioned dense band: a single-GPU operator over [ghost rows above | owned rows | ghost rows below] (borrowed;
    // pk_mat_set_band_ext), on which the dense-band kernels run unchanged — whatever the cut ends get wrong travels
    // bw rows per mat-vec and never reaches the owned rows within a trip.
    pk_mat* band_ext = nullptr;
    long long band_ra = 0, band_rb = 0;
    long long mp_row0 = 0, mp_n_global = 0;
    const long long* mp_halo_global = nullptr;                 // borrowed
    const int32_t* mp_g_rowptr[2] = {nullptr, nullptr};        // borrowed: ghost rows above / below (global columns)
    const int32_t* mp_g_col[2] = {nullptr, nullptr};
    const double* mp_g_val[2] = {nullptr, nullptr};
    long long mp_g_rows[2] = {0, 0};
    double* mp_gin = nullptr;                                  // owned: [chain 0..1][above, below][mp_depth]
    long long mp_depth = 0;                                    // ghost rows + bw: vector entries exchanged per side
    bool pat_on = false;
    int n_pat = 0, pat_entries = 0;
    const uint16_t* pat_id = nullptr;      // [n_rows]
    const int32_t* pat_ptr = nullptr;      // [n_pat + 1]
    const int32_t* pat_off = nullptr;      // [pat_entries]  col - row
    const double* pat_val = nullptr;       // [pat_entries]
};


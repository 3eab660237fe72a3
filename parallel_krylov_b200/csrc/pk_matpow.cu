// Matrix-powers kernel: the k-skip basis  A^l u, A^l v  (l = 1..k) of BOTH chains in ONE pass over A, for operators of
// small bandwidth (every column within +-bw of its row; the banded system of BASELINE.json configs[3] has bw = 13).
// Replaces the basis loops /root/reference/v3/cpu/kskipmrr.py:45-48 and kskipcg.py:36-39 (2k mat-vecs = 2k passes over
// A in the reference, k two-chain passes in the SpMV path of this library).
//
// A block owns a window of T = 640 consecutive rows, one row per thread, and keeps its row of A — up to RMAX (value,
// column offset) pairs — in REGISTERS for all k levels; the level vectors of the window ping-pong through shared memory.
// Level l is only valid (l-1)*bw rows inside the window on either side (a trapezoid), so the block emits
// T - 2(k-1)bw finished rows per window and neighbouring windows overlap by the ghost rows: A is re-read ~T/T_out times
// (mostly from L2, the neighbour block just had it), instead of k times from HBM.  Per row the sum runs left to right
// over the CSR entries with separately rounded products and sums, exactly like the SpMV kernels, so the basis is
// bit-identical to k sequential operator applications (tests/test_gpu_kernels.py).
#include "pk_device.cuh"
#include "pk_launch.h"

#include <stdlib.h>

#include <algorithm>

namespace {

constexpr int MP_T = 640;       // rows per window = threads per block (one block per SM: the rows of A fill the registers)
constexpr int MP_RMAX = 28;     // nonzeros per row kept in registers

struct MpArgs {
    const int32_t* rowptr;
    const int32_t* col;
    const double* val;
    long long n;                // rows (square block, not distributed)
    long long ld;               // distance between consecutive levels of a chain
    double* base0;              // chain 0: level l at base0 + l * ld, l = 0 (input) .. k
    double* base1;              // chain 1
    int bw;                     // half bandwidth
    int k;                      // levels to generate (dyn: the current k is read from PkState)
    int dyn;
    // Row-partitioned operator (EXT): the block works on the WINDOW of positions that covers its owned rows, gr ghost
    // rows of A on either side (copies of the neighbours' rows, shipped once at set-up) and bw more vector entries beyond
    // them; position p <-> global index win0 + p.  Level 0 of the window's ghost part comes from the neighbours' vectors
    // (depth gr + bw, exchanged ONCE per trip); levels are valid on the owned rows as long as (k-1) bw <= gr.
    long long n_loc;            // owned rows
    long long own_lo;           // position of the first owned row (= bw + ghost rows above)
    long long win0;             // global index of position 0 (may be negative at the first rank)
    long long n_global;
    long long row0;             // global index of the first owned row
    const long long* halo_global;   // global index of local column n_loc + h
    const int32_t* g_rowptr[2];     // ghost rows above / below: CSR with GLOBAL column indices
    const int32_t* g_col[2];
    const double* g_val[2];
    long long g_rows[2];
    const double* gin0[2];          // level 0 of chain 0 above / below the owned rows (gr + bw entries each)
    const double* gin1[2];
    long long g_in[2];              // entries available above / below
};

template <bool EXT>
__global__ void __launch_bounds__(MP_T, 1) k_matpow(MpArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const int k = a.dyn ? ra.st->k : a.k;
    if (k < 1) return;
    const int bw = a.bw;
    const int ghost = (k - 1) * bw;
    const int t_out = MP_T - 2 * ghost;                 // finished rows per window (host guarantees >= MP_T / 2 for a.k)
    const int W = MP_T + 2 * bw;                        // a level in shared memory: the window + bw entries on either side
    // the two chains are interleaved (one double2 per position): a gather is ONE 128-bit shared load for both chains
    extern __shared__ __align__(16) double2 mp_sm[];    // [buffer 0..1][W]
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * W; i += MP_T) mp_sm[i] = make_double2(0.0, 0.0);   // pads of the buffer levels >= 1 are written into
    __syncthreads();
    const long long n_tiles = (a.n + t_out - 1) / t_out;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long o0 = tile * t_out;              // finished rows [o0, o1)
        const long long o1 = (o0 + t_out < a.n) ? o0 + t_out : a.n;
        const long long s0 = o0 - ghost;                // first row of the window
        const long long row = s0 + tid;
        // ---- my row of A -> registers (column offsets relative to the row, packed 4 per register)
        double v[MP_RMAX];
        unsigned int offp[MP_RMAX / 4];
        int cnt = 0;
#pragma unroll
        for (int t = 0; t < MP_RMAX / 4; ++t) offp[t] = 0u;
#pragma unroll
        for (int t = 0; t < MP_RMAX; ++t) v[t] = 0.0;
        if (row >= 0 && row < a.n) {
            // which row of A sits at this position: an owned row, a ghost row above / below (EXT), or nothing
            const int32_t* rp = a.rowptr;
            const int32_t* cl = a.col;
            const double* vl = a.val;
            long long lr = row;                                 // row index within its CSR
            int src = 0;                                        // 0 owned (local columns), 1 ghost (global columns)
            bool have = true;
            if (EXT) {
                const long long e = row - a.own_lo;             // relative to the first owned row
                if (e >= 0 && e < a.n_loc) lr = e;
                else if (e < 0 && -e <= a.g_rows[0]) { rp = a.g_rowptr[0]; cl = a.g_col[0]; vl = a.g_val[0]; lr = a.g_rows[0] + e; src = 1; }
                else if (e >= a.n_loc && e - a.n_loc < a.g_rows[1]) { rp = a.g_rowptr[1]; cl = a.g_col[1]; vl = a.g_val[1]; lr = e - a.n_loc; src = 1; }
                else have = false;
            }
            if (have) {
                const int q0 = __ldg(rp + lr);
                cnt = __ldg(rp + lr + 1) - q0;
#pragma unroll
                for (int t = 0; t < MP_RMAX; ++t) {
                    if (t < cnt) {
                        v[t] = __ldg(vl + q0 + t);
                        int o;
                        if (!EXT) o = __ldg(cl + q0 + t) - (int)row;
                        else {
                            const long long c = __ldg(cl + q0 + t);
                            // global column -> offset from this row's own global index (win0 + row)
                            const long long gc = src ? c : (c < a.n_loc ? a.row0 + c : a.halo_global[c - a.n_loc]);
                            o = (int)(gc - (a.win0 + row));
                        }
                        offp[t >> 2] |= ((unsigned int)(o & 0xff)) << (8 * (t & 3));
                    }
                }
            }
        }
        // ---- level 0 of both chains -> shared memory (window + bw on either side; zeros outside the matrix)
        double2* P = mp_sm;
        double2* N = mp_sm + W;
        for (int i = tid; i < W; i += MP_T) {
            const long long g = s0 - bw + i;
            double u0 = 0.0, u1 = 0.0;
            if (g >= 0 && g < a.n) {
                if (!EXT) { u0 = a.base0[g]; u1 = a.base1[g]; }
                else {
                    const long long e = g - a.own_lo;
                    if (e >= 0 && e < a.n_loc) { u0 = a.base0[e]; u1 = a.base1[e]; }
                    else if (e < 0 && -e <= a.g_in[0]) { u0 = a.gin0[0][a.g_in[0] + e]; u1 = a.gin1[0][a.g_in[0] + e]; }
                    else if (e >= a.n_loc && e - a.n_loc < a.g_in[1]) { u0 = a.gin0[1][e - a.n_loc]; u1 = a.gin1[1][e - a.n_loc]; }
                }
            }
            P[i] = make_double2(u0, u1);
        }
        __syncthreads();
        for (int l = 1; l <= k; ++l) {
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int t = 0; t < MP_RMAX; ++t) {
                if (t < cnt) {                           // left to right over the CSR entries of the row
                    const int o = (int)(signed char)((offp[t >> 2] >> (8 * (t & 3))) & 0xffu);
                    const double2 pv = P[tid + bw + o];
                    y0 += v[t] * pv.x;
                    y1 += v[t] * pv.y;
                }
            }
            N[tid + bw] = make_double2(y0, y1);
            if (row >= o0 && row < o1) {
                if (!EXT) {
                    a.base0[(size_t)l * a.ld + row] = y0;
                    a.base1[(size_t)l * a.ld + row] = y1;
                } else if (row >= a.own_lo && row < a.own_lo + a.n_loc) {       // only the owned rows leave the window
                    a.base0[(size_t)l * a.ld + (row - a.own_lo)] = y0;
                    a.base1[(size_t)l * a.ld + (row - a.own_lo)] = y1;
                }
            }
            __syncthreads();
            double2* t0 = P; P = N; N = t0;
        }
        // the buffer that held level 0 carried real neighbour data in its pads; levels >= 1 of the next window must not see
        // stale pads of a different window as anything but finite numbers — they are finite, and only ever feed rows outside
        // the valid trapezoid, so no reset is needed; the barrier above already separates this window from the next load.
    }
}

// ---- dense band: two rows per thread, the next window staged by TMA --------------------------------------------------
// When EVERY row holds its full band (row r has exactly the columns max(r-bw,0) .. min(r+bw,n-1): the banded system of
// BASELINE.json configs[3]), a thread that owns the two adjacent rows 2t, 2t+1 needs the 2bw+2 consecutive entries
// x[2t-bw .. 2t+1+bw] of a level, and every one of them but the two outermost feeds BOTH rows: 2bw+2 shared loads for
// two rows instead of 2(2bw+1).  No column offsets are kept or decoded (entry e of a row is diagonal e + max(bw-r,0)),
// the row pointers follow from (n, bw) in closed form, and the window grows to 2*NT rows, which also trims the trapezoid
// overlap (1.40x -> 1.31x at k = 8, bw = 13).  A level lives in shared memory split by parity of the window position (even
// positions in E, odd in O), so that consecutive threads read consecutive 16-byte words whichever position they ask for.
// The values of a window are ONE contiguous run of the CSR value array.  Row-per-thread global loads of that run touch 32
// different sectors per warp instruction and kept the first version of this kernel L1-bound (measured: 7.9 ms against
// 9.1 ms of k_matpow), so the run of the NEXT window — and level 0 of both chains for it — is fetched by TMA bulk copies
// into shared memory WHILE the levels of the current one are computed; at the top of a window the threads move their two
// rows from shared memory to registers (27 conflict-free 128-bit loads).  DRAM time hides behind the arithmetic: the
// kernel ends up bound by the FP64 pipe (4 multiply-adds per nonzero per level, issued as DMUL + DADD for bit parity).
// Entries clipped at the matrix edge are kept as +0.0 values: they meet positions outside the matrix, which hold +0.0 at
// every level, and a running sum that starts at +0.0 is unchanged by adding (+-0.0) — results stay bit-identical to k
// chained mat-vecs (tests/test_gpu_kernels.py).
constexpr int MB_DMAX = 27;     // diagonals per row kept in registers (bw <= 13)
constexpr int MB_BWMAX = 13;
constexpr int MB_NT = 384;      // threads per block: 768-row windows, <= 168 registers per thread

struct MbLayout {
    static constexpr int T2 = 2 * MB_NT;                       // rows per window
    static constexpr int H = MB_NT + MB_BWMAX + 1;             // 16-byte words of E (and of O) per level
    static constexpr int SW = T2 + 2 * MB_BWMAX + 2;           // doubles per chain of the staged level 0 (even)
    static constexpr size_t off_levels = 16;                   // after the mbarrier
    static constexpr size_t off_stage = off_levels + sizeof(double2) * 4 * H;
    static constexpr size_t off_a = off_stage + sizeof(double) * 2 * SW;
    static constexpr size_t a_doubles = (size_t)T2 * MB_DMAX + 2 + 32;   // + slack: a predicated-off read may be speculated
    static constexpr size_t bytes = off_a + sizeof(double) * a_doubles;
};
static_assert(MbLayout::off_stage % 16 == 0 && MbLayout::off_a % 16 == 0, "TMA destinations must be 16-byte aligned");
static_assert(MbLayout::bytes <= 227 * 1024, "shared memory of one block");

// Row pointer of a dense band in closed form (rowptr[0] = 0; k_band_dense verifies the operator's own array against it).
__host__ __device__ __forceinline__ long long mb_rowptr(long long r, long long n, int bw) {
    long long q = r * (2 * bw + 1);
    const long long t = r < bw ? r : bw;                       // rows i < t lose bw - i entries at the top edge
    q -= t * bw - t * (t - 1) / 2;
    const long long m = r - (n - bw);                          // rows n - bw .. r - 1 lose 1 .. m entries at the bottom edge
    if (m > 0) q -= m * (m + 1) / 2;
    return q;
}

// The values of window rows [s0, s0 + T2) as one run of the CSR value array: source (16-byte aligned: the run starts at
// the even element at or below its first one) and byte count (a multiple of 16; an odd last element of the MATRIX is moved
// by hand, elsewhere the copy simply takes the next row's first value along).
__device__ __forceinline__ void mb_value_run(const double* val, long long n, int bw, long long s0, double* smA,
                                             const double** src, unsigned* bytes) {
    constexpr int T2 = MbLayout::T2;
    const long long f = s0 < 0 ? 0 : s0, e = s0 + T2 > n ? n : s0 + T2;
    const long long q0 = mb_rowptr(f, n, bw), q1 = mb_rowptr(e, n, bw), nnz = mb_rowptr(n, n, bw);
    const long long q_al = q0 & ~1ll;
    long long cnt = q1 - q_al;
    if (cnt & 1) {
        if (q1 < nnz) ++cnt;
        else { smA[cnt - 1] = __ldg(val + q1 - 1); --cnt; }
    }
    *src = val + q_al;
    *bytes = (unsigned)(cnt * 8);
}
__device__ __forceinline__ void mb_copy_values(double* smA, const double* src, unsigned bytes, unsigned long long* bar) {
    constexpr unsigned PIECE = 16384;
    for (unsigned o = 0; o < bytes; o += PIECE)
        bulk_g2s((char*)smA + o, (const char*)src + o, (bytes - o < PIECE) ? bytes - o : PIECE, bar);
}

// Stage window rows [s0, s0 + T2): their values, and level 0 of both chains on [s0 - bw, s0 + T2 + bw) — one thread,
// completion on `bar`.  Sources must be 16-byte aligned, so a run starts at the even element at or below its first one;
// the copy may run one element past a run's end where the array goes on (it always does for the vectors: ld > n or n even).
__device__ __forceinline__ void mb_issue_window(const MpArgs& a, int bw, long long s0, double* smA, double* stage,
                                                unsigned long long* bar) {
    constexpr int T2 = MbLayout::T2;
    unsigned bytes_a;
    const double* src_a;
    mb_value_run(a.val, a.n, bw, s0, smA, &src_a, &bytes_a);
    const long long gs = s0 - bw < 0 ? 0 : s0 - bw, ge = s0 + T2 + bw > a.n ? a.n : s0 + T2 + bw;
    const long long g_al = gs & ~1ll;
    const long long gcnt = (ge - g_al + 1) & ~1ll;
    const unsigned bytes_v = (unsigned)(gcnt * 8);
    mbar_expect_tx(bar, bytes_a + 2u * bytes_v);               // also the single arrival of this phase
    if (bytes_v) {
        bulk_g2s(stage, a.base0 + g_al, bytes_v, bar);
        bulk_g2s(stage + MbLayout::SW, a.base1 + g_al, bytes_v, bar);
    }
    mb_copy_values(smA, src_a, bytes_a, bar);
}

// My two rows (s0 + 2 tid, s0 + 2 tid + 1) of a staged window: shared memory -> registers.  Diagonal d of row r is its
// entry d - max(bw - r, 0); entries clipped by the matrix edges become +0.0.
template <int BWT>
__device__ __forceinline__ void mb_rows_to_regs(const double* smA, long long s0, long long q_al, int tid, long long n,
                                                int bw, double (&vA)[MB_DMAX], double (&vB)[MB_DMAX]) {
    constexpr int T2 = MbLayout::T2;
    const int D = 2 * bw + 1;
    const long long rA = s0 + 2 * tid, rB = rA + 1;
    const bool interior = s0 >= bw && s0 + T2 + bw <= n;           // window clear of the matrix edges: every row is full
    if (BWT && interior) {
        // the pair of rows is 2 D consecutive doubles from q = (first value of the window) + 2 tid D
        double w[2 * MB_DMAX];
        const int q = (int)(mb_rowptr(s0, n, bw) - q_al) + 2 * tid * (2 * BWT + 1);
        if ((q & 1) == 0) {
            const double2* s2 = (const double2*)(smA + q);
#pragma unroll
            for (int j = 0; j < MB_DMAX; ++j) { const double2 t2 = s2[j]; w[2 * j] = t2.x; w[2 * j + 1] = t2.y; }
        } else {
            const double2* s2 = (const double2*)(smA + q + 1);
            w[0] = smA[q];
#pragma unroll
            for (int j = 0; j < MB_DMAX - 1; ++j) { const double2 t2 = s2[j]; w[2 * j + 1] = t2.x; w[2 * j + 2] = t2.y; }
            w[2 * MB_DMAX - 1] = smA[q + 2 * MB_DMAX - 1];
        }
#pragma unroll
        for (int d = 0; d < MB_DMAX; ++d) { vA[d] = w[d]; vB[d] = w[MB_DMAX + d]; }
    } else if (interior) {
        const double* sA = smA + (int)(mb_rowptr(s0, n, bw) - q_al) + 2 * tid * D;
#pragma unroll
        for (int d = 0; d < MB_DMAX; ++d) {
            if (d < D) { vA[d] = sA[d]; vB[d] = sA[D + d]; }
            else { vA[d] = 0.0; vB[d] = 0.0; }
        }
    } else {
        int qA = 0, cA = 0, qB = 0, cB = 0;
        if (rA >= 0 && rA < n) { const long long q = mb_rowptr(rA, n, bw); qA = (int)(q - q_al); cA = (int)(mb_rowptr(rA + 1, n, bw) - q); }
        if (rB >= 0 && rB < n) { const long long q = mb_rowptr(rB, n, bw); qB = (int)(q - q_al); cB = (int)(mb_rowptr(rB + 1, n, bw) - q); }
        const int dA = rA < bw ? (int)(bw - rA) : 0, dB = rB < bw ? (int)(bw - rB) : 0;
#pragma unroll
        for (int d = 0; d < MB_DMAX; ++d) {
            const int eA = d - dA, eB = d - dB;
            vA[d] = (eA >= 0 && eA < cA) ? smA[qA + eA] : 0.0;
            vB[d] = (eB >= 0 && eB < cB) ? smA[qB + eB] : 0.0;
        }
    }
}

// BWT > 0: half bandwidth known at compile time (13: the 27-diagonal band of configs[3]) — the level loop is then
// straight-line code (28 shared loads with immediate offsets, 108 multiplies, 108 adds) and the loads are scheduled ahead
// of the arithmetic; BWT = 0 keeps bw a run-time value (uniform branches per position).
template <int BWT>
__global__ void __launch_bounds__(MB_NT, 1) k_matpow_band(MpArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    const int k = a.dyn ? ra.st->k : a.k;
    if (k < 1) return;
    using L = MbLayout;
    constexpr int NT = MB_NT, T2 = L::T2, H = L::H;
    const int bw = BWT ? BWT : a.bw;
    const int D = 2 * bw + 1;
    const int ghost = (k - 1) * bw;
    const int t_out = T2 - 2 * ghost;
    extern __shared__ __align__(16) unsigned char mb_raw[];
    unsigned long long* bar = (unsigned long long*)mb_raw;
    double2* lev = (double2*)(mb_raw + L::off_levels);          // [buffer 0..1][E | O][H], both chains interleaved per position
    double* stage = (double*)(mb_raw + L::off_stage);           // level 0 as it lies in memory: [chain][SW]
    double* smA = (double*)(mb_raw + L::off_a);                 // the values of one window, as they lie in the CSR array
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * H; i += NT) lev[i] = make_double2(0.0, 0.0);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (a.n + t_out - 1) / t_out;
    if (tid == 0 && blockIdx.x < n_tiles) mb_issue_window(a, bw, (long long)blockIdx.x * t_out - ghost, smA, stage, bar);
    unsigned phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long o0 = tile * t_out;
        const long long o1 = (o0 + t_out < a.n) ? o0 + t_out : a.n;
        const long long s0 = o0 - ghost;                // first row of the window (position 0)
        const long long rA = s0 + 2 * tid, rB = rA + 1;
        const long long f = s0 < 0 ? 0 : s0;
        const long long q_al = mb_rowptr(f, a.n, bw) & ~1ll;
        const long long g_al = (s0 - bw < 0 ? 0 : s0 - bw) & ~1ll;
        mbar_wait(bar, phase);
        phase ^= 1u;
        // ---- level 0 of both chains; position p (row s0 + p) sits at index p + bw, even indices in E, odd in O: the
        // entries a thread gathers are then at the compile-time indices 2 tid + s, s = 0 .. 2 bw + 1
        double2* P = lev;
        double2* N = lev + 2 * H;
        for (int i = tid; i < T2 + 2 * bw; i += NT) {
            const long long g = s0 - bw + i;
            double u0 = 0.0, u1 = 0.0;
            if (g >= 0 && g < a.n) { u0 = stage[g - g_al]; u1 = stage[L::SW + (g - g_al)]; }
            P[(i & 1) * H + (i >> 1)] = make_double2(u0, u1);
        }
        // ---- my two rows: shared memory -> registers
        double vA[MB_DMAX], vB[MB_DMAX];
        mb_rows_to_regs<BWT>(smA, s0, q_al, tid, a.n, bw, vA, vB);
        __syncthreads();                                // level 0 visible; everyone is done with the staged window
        if (tid == 0 && tile + gridDim.x < n_tiles)     // the next window of this block streams in behind the levels
            mb_issue_window(a, bw, (tile + gridDim.x) * t_out - ghost, smA, stage, bar);
        // my rows' own positions: index 2 tid + bw (row A) and 2 tid + bw + 1 (row B)
        const int wA = ((bw & 1) ? H : 0) + tid + (bw >> 1);
        const int wB = (((bw + 1) & 1) ? H : 0) + tid + ((bw + 1) >> 1);
        for (int l = 1; l <= k; ++l) {
            double yA0 = 0.0, yA1 = 0.0, yB0 = 0.0, yB1 = 0.0;
            const double2* Pt = P + tid;
#pragma unroll
            for (int s = 0; s <= MB_DMAX; ++s) {
                if (s > D) break;                        // uniform: positions 2 tid - bw + s, s = 0 .. 2 bw + 1
                const double2 X = Pt[(s & 1) * H + (s >> 1)];
                if (s < MB_DMAX) {
                    if (s < D) {                         // row A, diagonal s
                        yA0 += vA[s] * X.x;
                        yA1 += vA[s] * X.y;
                    }
                }
                if (s >= 1) {                            // row B, diagonal s - 1 (same position, one row further down)
                    yB0 += vB[s - 1] * X.x;
                    yB1 += vB[s - 1] * X.y;
                }
            }
            N[wA] = make_double2(yA0, yA1);
            N[wB] = make_double2(yB0, yB1);
            if (rA >= o0 && rA < o1) {
                a.base0[(size_t)l * a.ld + rA] = yA0;
                a.base1[(size_t)l * a.ld + rA] = yA1;
            }
            if (rB >= o0 && rB < o1) {
                a.base0[(size_t)l * a.ld + rB] = yB0;
                a.base1[(size_t)l * a.ld + rB] = yB1;
            }
            __syncthreads();
            double2* t0 = P; P = N; N = t0;
        }
    }
}

// ---- the k+1 steps of a k-skip MrR trip in ONE pass over a dense band -------------------------------------------------
// Once the Gram epilogue has produced (zeta_j, eta_j), j = 0..k, the steps of a trip (/root/reference/v3/cpu/kskipmrr.py:
// 63-69 and :87-93) are a fixed recurrence with no inner product in it:
//     y <- eta_j y + zeta_j (A r);   z <- eta_j z - zeta_j r;   r <- r - y;   x <- x - z;   (A r) <- A r
// — k+1 element-wise updates, each followed by ONE mat-vec of r.  The step-by-step path streams A through the SpMV
// kernel k+1 times per trip (17 of the 25 ms of a trip on the 32M band); here a block keeps the rows of A of its window
// in registers (as k_matpow_band does), r of the window ping-pongs through shared memory between the steps, and y, z, x,
// A r of a thread's two rows never leave its registers: A and the five vectors are read once and written once per trip.
// After m mat-vecs the window is valid (m bw) rows inside either end, so a window of 768 rows finishes
// 768 - 2 (k+1) bw of them.  Every row runs exactly the arithmetic of k_mrr_update / the fused SpMV epilogue
// (same products, same sums, same order), so x, r, y, z come out bit-identical to the step-by-step path.
// Windows overlap, hence r, A r, y are written to scratch vectors (the neighbours still read the old ones) and copied
// home by k_copy3; z and x are only touched on the rows a block finishes, in place.
struct StArgs {
    const double* val;
    long long n;
    int bw, k;
    const double* r; const double* ar; const double* y;         // read on whole windows
    double* z; double* x;                                       // in place, finished rows only
    double* r_out; double* ar_out; double* y_out;
    long long own_lo, own_hi;                                   // rows that may be finished (row-partitioned: the owned ones)
};

struct StLayout {
    static constexpr int T2 = MbLayout::T2, H = MbLayout::H;
    static constexpr int SW = T2 + 2;                           // doubles per staged vector run (even)
    static constexpr size_t off_coef = 16;                      // after the mbarrier: (zeta_j, eta_j), j = 0..PK_KMAX
    static constexpr size_t off_levels = off_coef + sizeof(double) * 2 * (PK_KMAX + 1) + 16;
    static constexpr size_t off_stage = off_levels + sizeof(double) * 4 * H;      // [buffer 0..1][E | O][H]
    static constexpr size_t off_a = off_stage + sizeof(double) * 5 * SW;          // r, A r, y (window), z, x (finished rows)
    static constexpr size_t bytes = off_a + sizeof(double) * MbLayout::a_doubles;
};
static_assert(StLayout::off_levels % 16 == 0 && StLayout::off_stage % 16 == 0 && StLayout::off_a % 16 == 0, "alignment");
static_assert(StLayout::bytes <= 227 * 1024, "shared memory of one block");

__device__ __forceinline__ void st_issue_window(const StArgs& a, int bw, long long s0, long long o0, long long o1,
                                                double* smA, double* stage, unsigned long long* bar) {
    constexpr int T2 = StLayout::T2, SW = StLayout::SW;
    unsigned bytes_a;
    const double* src_a;
    mb_value_run(a.val, a.n, bw, s0, smA, &src_a, &bytes_a);
    const long long f = s0 < 0 ? 0 : s0, e = s0 + T2 > a.n ? a.n : s0 + T2;
    o0 = o0 < a.own_lo ? a.own_lo : o0;
    o1 = o1 > a.own_hi ? a.own_hi : o1;
    const long long w_al = f & ~1ll, o_al = o0 & ~1ll;
    // r, A r, y live in ld-padded work vectors: a run may take one element past its end along.  x is the caller's
    // solution vector of exactly n entries: an odd last element of it (and of z, for symmetry) is moved by hand.
    const unsigned bytes_w = (unsigned)(((e - w_al + 1) & ~1ll) * 8);
    long long ocnt = o1 > o0 ? o1 - o_al : 0;
    if (ocnt & 1) {
        if (o1 < a.own_hi) ++ocnt;
        else { stage[3 * SW + ocnt - 1] = a.z[o1 - 1]; stage[4 * SW + ocnt - 1] = a.x[o1 - 1]; --ocnt; }
    }
    const unsigned bytes_o = (unsigned)(ocnt * 8);
    mbar_expect_tx(bar, bytes_a + 3u * bytes_w + 2u * bytes_o);
    bulk_g2s(stage, a.r + w_al, bytes_w, bar);
    bulk_g2s(stage + SW, a.ar + w_al, bytes_w, bar);
    bulk_g2s(stage + 2 * SW, a.y + w_al, bytes_w, bar);
    if (bytes_o) {
        bulk_g2s(stage + 3 * SW, a.z + o_al, bytes_o, bar);
        bulk_g2s(stage + 4 * SW, a.x + o_al, bytes_o, bar);
    }
    mb_copy_values(smA, src_a, bytes_a, bar);
}

template <int BWT>
__global__ void __launch_bounds__(MB_NT, 1) k_mrr_steps_band(StArgs a, PkRedArgs ra) {
    if (pk_skip(ra)) return;
    using L = StLayout;
    constexpr int NT = MB_NT, T2 = L::T2, H = L::H, SW = L::SW;
    const int k = a.k;
    const int bw = BWT ? BWT : a.bw;
    const int D = 2 * bw + 1;
    const int ghost = (k + 1) * bw;                     // k+1 mat-vecs per trip (k between the steps + the closing one)
    const int t_out = T2 - 2 * ghost;
    extern __shared__ __align__(16) unsigned char mb_raw[];
    unsigned long long* bar = (unsigned long long*)mb_raw;
    double* coef = (double*)(mb_raw + L::off_coef);
    double* lev = (double*)(mb_raw + L::off_levels);
    double* stage = (double*)(mb_raw + L::off_stage);
    double* smA = (double*)(mb_raw + L::off_a);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * H; i += NT) lev[i] = 0.0;
    for (int i = tid; i < 2 * (k + 1); i += NT) coef[i] = ra.st->coef[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (a.n + t_out - 1) / t_out;
    auto issue = [&](long long tile) {
        const long long o0 = tile * t_out, o1 = (o0 + t_out < a.n) ? o0 + t_out : a.n;
        st_issue_window(a, bw, o0 - ghost, o0, o1, smA, stage, bar);
    };
    if (tid == 0 && blockIdx.x < n_tiles) issue(blockIdx.x);
    unsigned phase = 0;
    double acc[1] = {0.0};
    // my rows' own positions in a level: index 2 tid + bw (row A) and 2 tid + bw + 1 (row B); even indices in E, odd in O
    const int wA = ((bw & 1) ? H : 0) + tid + (bw >> 1);
    const int wB = (((bw + 1) & 1) ? H : 0) + tid + ((bw + 1) >> 1);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long t0 = tile * t_out;
        const long long t1 = (t0 + t_out < a.n) ? t0 + t_out : a.n;
        const long long s0 = t0 - ghost;
        const long long o0 = t0 < a.own_lo ? a.own_lo : t0;    // finished rows of this window that are also owned
        const long long o1 = t1 > a.own_hi ? a.own_hi : t1;
        const long long rowA = s0 + 2 * tid, rowB = rowA + 1;
        const long long f = s0 < 0 ? 0 : s0;
        const long long q_al = mb_rowptr(f, a.n, bw) & ~1ll;
        const long long w_al = f & ~1ll, o_al = o0 & ~1ll;
        const bool inA = rowA >= 0 && rowA < a.n, inB = rowB >= 0 && rowB < a.n;
        const bool outA = rowA >= o0 && rowA < o1, outB = rowB >= o0 && rowB < o1;
        mbar_wait(bar, phase);
        phase ^= 1u;
        double rA = 0.0, arA = 0.0, yA = 0.0, zA = 0.0, xA = 0.0, rB = 0.0, arB = 0.0, yB = 0.0, zB = 0.0, xB = 0.0;
        if (inA) { rA = stage[rowA - w_al]; arA = stage[SW + (rowA - w_al)]; yA = stage[2 * SW + (rowA - w_al)]; }
        if (inB) { rB = stage[rowB - w_al]; arB = stage[SW + (rowB - w_al)]; yB = stage[2 * SW + (rowB - w_al)]; }
        if (outA) { zA = stage[3 * SW + (rowA - o_al)]; xA = stage[4 * SW + (rowA - o_al)]; }
        if (outB) { zB = stage[3 * SW + (rowB - o_al)]; xB = stage[4 * SW + (rowB - o_al)]; }
        double vA[MB_DMAX], vB[MB_DMAX];
        mb_rows_to_regs<BWT>(smA, s0, q_al, tid, a.n, bw, vA, vB);
        __syncthreads();                                // everyone is done with the staged window (and with the levels)
        if (tid == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x);
        int cur = 0;
        for (int j = 0; j <= k; ++j) {
            const double zeta = coef[2 * j], eta = coef[2 * j + 1];
            // the step, exactly as k_mrr_update / the FUSE == 1 epilogue of k_spmv_tma write it
            yA = eta * yA + zeta * arA;
            zA = eta * zA - zeta * rA;
            rA = rA - yA;
            xA = xA - zA;
            yB = eta * yB + zeta * arB;
            zB = eta * zB - zeta * rB;
            rB = rB - yB;
            xB = xB - zB;
            double* Nb = lev + cur * 2 * H;
            Nb[wA] = rA;
            Nb[wB] = rB;
            __syncthreads();
            // A r of my two rows from the r just published (entries 2 tid - bw + s, s = 0 .. 2 bw + 1)
            double sA = 0.0, sB = 0.0;
            const double* Pt = Nb + tid;
#pragma unroll
            for (int s = 0; s <= MB_DMAX; ++s) {
                if (s > D) break;
                const double X = Pt[(s & 1) * H + (s >> 1)];
                if (s < MB_DMAX) {
                    if (s < D) sA += vA[s] * X;
                }
                if (s >= 1) sB += vB[s - 1] * X;
            }
            arA = sA;
            arB = sB;
            cur ^= 1;
        }
        if (outA) {
            a.r_out[rowA] = rA; a.ar_out[rowA] = arA; a.y_out[rowA] = yA; a.z[rowA] = zA; a.x[rowA] = xA;
            acc[0] += rA * rA;
        }
        if (outB) {
            a.r_out[rowB] = rB; a.ar_out[rowB] = arB; a.y_out[rowB] = yB; a.z[rowB] = zB; a.x[rowB] = xB;
            acc[0] += rB * rB;
        }
    }
    pk_grid_reduce<1, MB_NT>(acc, ra);
}

// three vectors home in one launch (128-bit when everything is 16-byte aligned)
__global__ void __launch_bounds__(256) k_copy3(long long n, const double* __restrict__ s0, double* __restrict__ d0,
                                               const double* __restrict__ s1, double* __restrict__ d1,
                                               const double* __restrict__ s2, double* __restrict__ d2, PkRedArgs ra) {
    // NOT gated on the stop flag: the steps kernel in front of it may have raised the flag in its own epilogue, and its
    // results must still reach home.  (If that kernel was itself skipped the solve is over and the vectors are dead.)
    if (ra.only_rollback && ra.st->rollback == 0) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n2 = n >> 1;
    const double2 *a0 = (const double2*)s0, *a1 = (const double2*)s1, *a2 = (const double2*)s2;
    double2 *b0 = (double2*)d0, *b1 = (double2*)d1, *b2 = (double2*)d2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 u = a0[i], v = a1[i], w = a2[i];
        b0[i] = u; b1[i] = v; b2[i] = w;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { d0[n - 1] = s0[n - 1]; d1[n - 1] = s1[n - 1]; d2[n - 1] = s2[n - 1]; }
}

// Is the block a dense band of half width bw?  out[0] counts the rows that are not.
__global__ void k_band_dense(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, long long n, int bw, int* out) {
    int bad = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const long long lo = r - bw < 0 ? 0 : r - bw, hi = r + bw > n - 1 ? n - 1 : r + bw;
        const int q0 = rowptr[r], q1 = rowptr[r + 1];
        if (q1 - q0 != (int)(hi - lo + 1) || q0 != mb_rowptr(r, n, bw) || q1 != mb_rowptr(r + 1, n, bw)) bad = 1;
        else
            for (int q = q0; q < q1; ++q)
                if (col[q] != (int)(lo + (q - q0))) { bad = 1; break; }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAdd(out, 1);
}

__global__ void k_band_info(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, long long n_rows,
                            long long row0, int* out) {
    int mlen = 0, mbw = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        const int q0 = rowptr[r], q1 = rowptr[r + 1];
        mlen = (q1 - q0) > mlen ? (q1 - q0) : mlen;
        if (q1 > q0) {        // sorted or not: look at both ends and, to be safe, every entry of short rows
            for (int q = q0; q < q1 && q < q0 + 64; ++q) {
                const long long d = (long long)col[q] - (r + row0);
                const int ad = (int)(d < 0 ? -d : d) > (1 << 20) ? (1 << 20) : (int)(d < 0 ? -d : d);
                mbw = ad > mbw ? ad : mbw;
            }
            if (q1 - q0 > 64) mbw = 1 << 20;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        int o = __shfl_down_sync(0xffffffffu, mlen, off);
        mlen = o > mlen ? o : mlen;
        o = __shfl_down_sync(0xffffffffu, mbw, off);
        mbw = o > mbw ? o : mbw;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mlen);
        atomicMax(out + 1, mbw);
    }
}

}  // namespace

static int band_probe(pk_ctx* ctx, const int32_t* rowptr, const int32_t* col, long long n_rows, long long row0, int* h) {
    int* d = nullptr;
    h[0] = h[1] = 0;
    PK_CUDA(cudaMalloc(&d, 2 * sizeof(int)));
    PK_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(int), ctx->stream));
    int grid = (int)((n_rows + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    k_band_info<<<grid, 256, 0, ctx->stream>>>(rowptr, col, n_rows, row0, d);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return PK_OK;
}

extern "C" int pk_csr_band_info(pk_ctx* ctx, int64_t n_rows, int64_t row0, const int32_t* d_rowptr, const int32_t* d_col,
                                int* h_out) {
    PK_REQUIRE(ctx && h_out && (n_rows == 0 || (d_rowptr && d_col)), "null argument");
    PK_CUDA(cudaSetDevice(ctx->device));
    return band_probe(ctx, d_rowptr, d_col, n_rows, row0, h_out);
}

// One-time (cached) structure probe: longest row and half bandwidth of a square, non-distributed CSR block.
static int band_info(pk_ctx* ctx, pk_mat* m) {
    if (m->mp_checked) return PK_OK;
    m->mp_checked = true;
    m->mp_rmax = 1 << 30;
    m->mp_bw = 1 << 30;
    if (m->kind == MAT_DENSE || m->distributed || !m->segs.empty() || m->rowptr == nullptr || m->n_rows != m->n_cols) return PK_OK;
    int h[2] = {0, 0};
    PK_CHECK(band_probe(ctx, m->rowptr, m->col, m->n_rows, 0, h));
    m->mp_rmax = h[0];
    m->mp_bw = h[1];
    if (m->mp_bw >= 1 && 2 * m->mp_bw + 1 <= MB_DMAX && m->mp_rmax <= MB_DMAX && m->n_rows > 2 * m->mp_bw &&
        ((uintptr_t)m->val & 15) == 0) {           // the TMA bulk copies of the value runs need a 16-byte aligned array
        int* d = nullptr;
        int bad = 1;
        PK_CUDA(cudaMalloc(&d, sizeof(int)));
        PK_CUDA(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
        int grid = (int)((m->n_rows + 255) / 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        k_band_dense<<<grid, 256, 0, ctx->stream>>>(m->rowptr, m->col, m->n_rows, m->mp_bw, d);
        PK_CUDA(cudaGetLastError());
        PK_CUDA(cudaMemcpyAsync(&bad, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d);
        m->mp_dense = (bad == 0);
    }
    return PK_OK;
}

// Window of the kernel that will run: rows per block, before the trapezoid is taken off.
static bool band_kernel_on(const pk_mat* m) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("PK_MATPOW_BAND");
        enabled = e ? atoi(e) : 1;
    }
    return enabled && m->mp_dense && !m->distributed;
}
static int mp_window(const pk_mat* m) { return band_kernel_on(m) ? MbLayout::T2 : MP_T; }

extern "C" int pk_mat_set_matpow_ext(pk_mat* m, int half_bw, int max_row_nnz, int64_t row0, int64_t n_global,
                                     const int64_t* d_halo_global, int64_t rows_above, const int32_t* d_rowptr_above,
                                     const int32_t* d_col_above, const double* d_val_above, int64_t rows_below,
                                     const int32_t* d_rowptr_below, const int32_t* d_col_below, const double* d_val_below) {
    PK_REQUIRE(m != nullptr, "null operator");
    PK_REQUIRE(m->kind != MAT_DENSE && m->segs.empty() && m->rowptr != nullptr, "matrix powers need a CSR block with 32-bit row pointers");
    PK_REQUIRE(half_bw >= 1 && half_bw <= 127 && max_row_nnz >= 1, "bad band description");
    PK_REQUIRE(rows_above >= 0 && rows_below >= 0, "negative ghost row count");
    PK_REQUIRE(rows_above == 0 || (d_rowptr_above && d_col_above && d_val_above), "null ghost rows (above)");
    PK_REQUIRE(rows_below == 0 || (d_rowptr_below && d_col_below && d_val_below), "null ghost rows (below)");
    PK_REQUIRE(m->n_halo == 0 || d_halo_global != nullptr, "null halo map");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    const long long depth = std::max<long long>(rows_above, rows_below) + half_bw;
    PK_REQUIRE(m->n_rows >= depth, "block smaller than the ghost zone its neighbours need");
    m->mp_checked = true;
    m->mp_rmax = max_row_nnz;
    m->mp_bw = half_bw;
    m->mp_row0 = row0;
    m->mp_n_global = n_global;
    m->mp_halo_global = (const long long*)d_halo_global;
    m->mp_g_rowptr[0] = d_rowptr_above; m->mp_g_col[0] = d_col_above; m->mp_g_val[0] = d_val_above; m->mp_g_rows[0] = rows_above;
    m->mp_g_rowptr[1] = d_rowptr_below; m->mp_g_col[1] = d_col_below; m->mp_g_val[1] = d_val_below; m->mp_g_rows[1] = rows_below;
    m->mp_depth = depth;
    if (m->mp_gin) cudaFree(m->mp_gin);
    PK_CUDA(cudaMalloc(&m->mp_gin, sizeof(double) * 4 * (size_t)depth));
    PK_CUDA(cudaMemset(m->mp_gin, 0, sizeof(double) * 4 * (size_t)depth));
    m->mp_ext = true;
    return PK_OK;
}

// Can the k levels of both chains be generated in one pass over A?  (small bandwidth, short rows, and the trapezoid must
// keep at least half of the window: otherwise the two-chain SpMV passes are the better choice)
bool pk_matpow_ok(pk_ctx* ctx, pk_mat* m, int k) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("PK_MATPOW");
        enabled = e ? atoi(e) : 1;
    }
    if (!enabled || k < 2) return false;
    if (m->distributed) {
        // row-partitioned: needs the neighbours' ghost rows (pk_mat_set_matpow_ext), deep enough for this k, and NCCL for
        // the once-per-trip ghost-zone exchange
        if (!m->mp_ext || ctx->comm == nullptr) return false;
        const long long need = (long long)(k - 1) * m->mp_bw;
        if ((ctx->rank > 0 && m->mp_g_rows[0] < need) || (ctx->rank + 1 < ctx->n_ranks && m->mp_g_rows[1] < need)) return false;
    } else if (band_info(ctx, m) != PK_OK) return false;
    if (m->mp_rmax > MP_RMAX || m->mp_bw > 127 || m->mp_bw < 1) return false;
    const int win = mp_window(m);
    return win - 2 * (k - 1) * m->mp_bw >= win / 2;
}

int pk_launch_matpow(pk_ctx* ctx, pk_mat* m, int k, double* base0, double* base1, int dyn) {
    MpArgs a{};
    a.rowptr = m->rowptr; a.col = m->col; a.val = m->val;
    a.n = m->n_rows; a.ld = m->ld; a.base0 = base0; a.base1 = base1;
    a.bw = m->mp_bw; a.k = k; a.dyn = dyn;
    const bool ext = m->distributed;
    if (ext) {
        const bool has_prev = ctx->rank > 0, has_next = ctx->rank + 1 < ctx->n_ranks;
        // ONE exchange per trip: depth = ghost rows + bw entries of both level-0 vectors with each neighbour
        PK_CHECK(pk_comm_ghost_exchange(ctx, base0, base1, m->n_rows, m->mp_depth, m->mp_gin, has_prev, has_next));
        a.n_loc = m->n_rows;
        a.g_rows[0] = has_prev ? m->mp_g_rows[0] : 0;
        a.g_rows[1] = has_next ? m->mp_g_rows[1] : 0;
        a.g_in[0] = has_prev ? m->mp_depth : 0;
        a.g_in[1] = has_next ? m->mp_depth : 0;
        a.own_lo = a.g_in[0];
        a.row0 = m->mp_row0;
        a.win0 = m->mp_row0 - a.own_lo;
        a.n_global = m->mp_n_global;
        a.halo_global = m->mp_halo_global;
        for (int s = 0; s < 2; ++s) {
            a.g_rowptr[s] = m->mp_g_rowptr[s]; a.g_col[s] = m->mp_g_col[s]; a.g_val[s] = m->mp_g_val[s];
            a.gin0[s] = m->mp_gin + (size_t)(0 + s) * m->mp_depth;
            a.gin1[s] = m->mp_gin + (size_t)(2 + s) * m->mp_depth;
        }
        // the depth actually received may exceed what this k needs; rows deeper than the ghost rows only feed level 0
        a.n = a.own_lo + a.n_loc + a.g_in[1];        // size of the window
    }
    PkRedArgs ra{};
    ra.st = ctx->d_state;
    ra.only_rollback = ctx->ctl_only_rollback;
    ra.dyn_cj = -1;
    ra.dyn_last = 0;
    if (band_kernel_on(m)) {
        if ((((uintptr_t)base0 | (uintptr_t)base1) & 15) != 0) {
            pk_set_error("matrix powers on a dense band stage level 0 by TMA: the level vectors must be 16-byte aligned");
            return PK_ERR_ARG;
        }
        const size_t smem = MbLayout::bytes;
        const bool b13 = a.bw == 13;
        pk_blocks_per_sm(b13 ? (const void*)k_matpow_band<13> : (const void*)k_matpow_band<0>, MB_NT, smem);   // opt in to the size
        const int t_out = MbLayout::T2 - 2 * (k - 1) * a.bw;
        const long long n_tiles = (a.n + t_out - 1) / t_out;
        int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
        if (grid < 1) grid = 1;
        if (b13) k_matpow_band<13><<<grid, MB_NT, smem, ctx->stream>>>(a, ra);
        else k_matpow_band<0><<<grid, MB_NT, smem, ctx->stream>>>(a, ra);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            pk_set_error("matrix-powers (dense band) launch: %s", cudaGetErrorString(e));
            return PK_ERR_CUDA;
        }
        ctx->launches++;
        ctx->spmvs += 2LL * k;
        return PK_OK;
    }
    const size_t smem = sizeof(double) * 4 * (size_t)(MP_T + 2 * a.bw);
    const void* kern = ext ? (const void*)k_matpow<true> : (const void*)k_matpow<false>;
    pk_blocks_per_sm(kern, MP_T, smem);          // opts in to the dynamic shared memory size if needed
    const int t_out = MP_T - 2 * (k - 1) * a.bw;
    long long n_tiles = (a.n + t_out - 1) / t_out;
    int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    if (grid < 1) grid = 1;
    if (ext) k_matpow<true><<<grid, MP_T, smem, ctx->stream>>>(a, ra);
    else k_matpow<false><<<grid, MP_T, smem, ctx->stream>>>(a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("matrix-powers launch: %s", cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    ctx->spmvs += 2LL * k;
    return PK_OK;
}

// Can the k+1 steps of a k-skip MrR trip run as one pass (k_mrr_steps_band)?  Dense band on one GPU, and the window must
// keep at least half of its rows after k+1 mat-vecs.
bool pk_mrr_steps_ok(pk_ctx* ctx, pk_mat* m, int k) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("PK_KSTEPS");
        enabled = e ? atoi(e) : 1;
    }
    if (!enabled || k < 2 || !pk_matpow_ok(ctx, m, k) || !band_kernel_on(m)) return false;
    return MbLayout::T2 - 2 * (k + 1) * m->mp_bw >= MbLayout::T2 / 2;
}

// r, ar = A r, y, z, x: the vectors of the trip; t0..t2: three scratch vectors (free basis slots).  Runs the k+1 steps,
// the closing mat-vec and the trip-end epilogue (red[0] = r.r); leaves everything at home.
int pk_launch_mrr_steps(pk_ctx* ctx, pk_mat* m, int k, double* r, double* ar, double* y, double* z, double* x,
                        double* t0, double* t1, double* t2, int epi, long long own_lo, long long own_hi) {
    const uintptr_t al = (uintptr_t)r | (uintptr_t)ar | (uintptr_t)y | (uintptr_t)z | (uintptr_t)x | (uintptr_t)t0 | (uintptr_t)t1 | (uintptr_t)t2;
    if ((al & 15) || (own_lo & 1)) {
        pk_set_error("fused k-skip steps stage their vectors by TMA: all vectors must be 16-byte aligned");
        return PK_ERR_ARG;
    }
    StArgs a{};
    a.val = m->val; a.n = m->n_rows; a.bw = m->mp_bw; a.k = k;
    a.r = r; a.ar = ar; a.y = y; a.z = z; a.x = x;
    a.r_out = t0; a.ar_out = t1; a.y_out = t2;
    a.own_lo = own_lo;
    a.own_hi = own_hi < 0 ? a.n : own_hi;
    PkRedArgs ra{};
    ra.partials = ctx->red.partials;
    ra.ticket = ctx->red.ticket;
    ra.max_blocks = ctx->red.max_blocks;
    ra.st = ctx->d_state;
    ra.epi = epi;
    ra.defer = (ctx->n_ranks > 1 && !ctx->d_p2p && !ctx->nocomm) ? 1 : 0;     // NCCL all-reduce + scalar kernel follow
    ra.p2p = ctx->nocomm ? nullptr : ctx->d_p2p;                              // or: all-reduced inside this kernel
    ra.ar_n = 1;
    ra.g_off = -1;
    ra.only_rollback = ctx->ctl_only_rollback;
    ra.dyn_cj = -1;
    const size_t smem = StLayout::bytes;
    const bool b13 = a.bw == 13;
    pk_blocks_per_sm(b13 ? (const void*)k_mrr_steps_band<13> : (const void*)k_mrr_steps_band<0>, MB_NT, smem);
    const int t_out = StLayout::T2 - 2 * (k + 1) * a.bw;
    const long long n_tiles = (a.n + t_out - 1) / t_out;
    int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    if (grid < 1) grid = 1;
    if (grid > ctx->red.max_blocks) grid = ctx->red.max_blocks;
    if (b13) k_mrr_steps_band<13><<<grid, MB_NT, smem, ctx->stream>>>(a, ra);
    else k_mrr_steps_band<0><<<grid, MB_NT, smem, ctx->stream>>>(a, ra);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("fused k-skip steps launch: %s", cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    ctx->spmvs += k + 1;
    PK_CHECK(pk_finish_reduce(ctx, 1, epi, -1, 0));
    PkRedArgs rc{};
    rc.st = ctx->d_state;
    rc.only_rollback = ctx->ctl_only_rollback;
    rc.dyn_cj = -1;
    const long long n_own = a.own_hi - a.own_lo;
    long long want = (n_own / 2 + 255) / 256;
    int cgrid = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
    if (cgrid < 1) cgrid = 1;
    k_copy3<<<cgrid, 256, 0, ctx->stream>>>(n_own, t0 + a.own_lo, r + a.own_lo, t1 + a.own_lo, ar + a.own_lo,
                                           t2 + a.own_lo, y + a.own_lo, rc);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        pk_set_error("copy-back launch: %s", cudaGetErrorString(e));
        return PK_ERR_CUDA;
    }
    ctx->launches++;
    return PK_OK;
}

// ---- row-partitioned dense band -----------------------------------------------------------------------------------------
// The dense-band kernels run UNCHANGED on a single-GPU operator that covers [ghost rows above | owned rows | ghost rows
// below] of this rank (the rows cut at its two ends make it a dense band of its own).  The vectors of the solve carry the
// same ghost zones in pads around their owned part; ONE exchange per trip (depth (k+1) bw of r, A r, y) refreshes them.
// Whatever the cut ends and the stale zones beyond the exchanged depth get wrong moves bw rows inward per mat-vec: after
// the k+1 mat-vecs of a trip it has not reached the owned rows, which therefore equal the single-GPU results bit for bit.
extern "C" int pk_mat_set_band_ext(pk_mat* m, pk_mat* ext, int64_t rows_above, int64_t rows_below) {
    PK_REQUIRE(m != nullptr, "null operator");
    if (ext == nullptr) {
        m->band_ext = nullptr;
        m->band_ra = m->band_rb = 0;
        return PK_OK;
    }
    PK_REQUIRE(m->distributed && m->ctx == ext->ctx, "the extended operator belongs to a row-partitioned operator of the same context");
    PK_REQUIRE(!ext->distributed && ext->kind != MAT_DENSE && ext->segs.empty() && ext->n_rows == ext->n_cols, "extended operator: square single-GPU CSR");
    PK_REQUIRE(rows_above >= 0 && rows_below >= 0 && (rows_above & 1) == 0 && (rows_below & 1) == 0, "ghost row counts must be even");
    PK_REQUIRE(ext->n_rows == m->n_rows + rows_above + rows_below, "extended operator: owned + ghost rows");
    pk_ctx* ctx = m->ctx;
    PK_CUDA(cudaSetDevice(ctx->device));
    PK_CHECK(band_info(ctx, ext));
    if (!ext->mp_dense) {
        pk_set_error("extended operator is not a dense band of <= %d diagonals", MB_DMAX);
        return PK_ERR_UNSUPPORTED;
    }
    // one leading dimension for both: the solve's vectors hold owned rows + both ghost zones
    long long ld = m->ld > ext->n_rows + m->n_halo ? m->ld : ext->n_rows + m->n_halo;
    ld = (ld + 31) / 32 * 32;
    m->ld = ld;
    ext->ld = ld;
    m->band_ext = ext;
    m->band_ra = rows_above;
    m->band_rb = rows_below;
    return PK_OK;
}

long long pk_band_ext_depth(pk_ctx* ctx, pk_mat* m, int k) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("PK_BAND_DIST");
        enabled = e ? atoi(e) : 1;
    }
    if (!enabled || !m->distributed || m->band_ext == nullptr || ctx->comm == nullptr) return 0;
    pk_mat* ext = m->band_ext;
    if (!pk_mrr_steps_ok(ctx, ext, k)) return 0;
    long long depth = (long long)(k + 1) * ext->mp_bw;
    depth += depth & 1;                                   // pads keep the 16-byte alignment of the owned part
    const bool has_prev = ctx->rank > 0, has_next = ctx->rank + 1 < ctx->n_ranks;
    if ((has_prev && m->band_ra < depth) || (has_next && m->band_rb < depth) || m->n_rows < depth) return 0;
    if ((!has_prev && m->band_ra != 0) || (!has_next && m->band_rb != 0)) return 0;
    return depth;
}

extern "C" int pk_mat_matpow_info(pk_ctx* ctx, pk_mat* mat, int k, int* kind, int* window_rows) {
    PK_REQUIRE(ctx && mat && kind && window_rows, "null argument");
    PK_CUDA(cudaSetDevice(ctx->device));
    const bool ok = pk_matpow_ok(ctx, mat, k);
    *kind = !ok ? 0 : (band_kernel_on(mat) ? 2 : 1);
    *window_rows = ok ? mp_window(mat) : 0;
    return PK_OK;
}

// C-ABI building block (tests / benchmarks): levels 1..k of both chains from level 0 at d_base0 / d_base1.
extern "C" int pk_matpow(pk_ctx* ctx, pk_mat* mat, int k, double* d_base0, double* d_base1) {
    PK_REQUIRE(ctx && mat && d_base0 && d_base1, "null argument");
    PK_REQUIRE(k >= 1 && k <= PK_KMAX, "k out of range");
    PK_CUDA(cudaSetDevice(ctx->device));
    if (!pk_matpow_ok(ctx, mat, k < 2 ? 2 : k)) {
        pk_set_error("operator not eligible for the one-pass matrix-powers kernel (needs a square single-GPU CSR block, "
                     "rows of <= %d nonzeros, half bandwidth bw with W - 2(k-1)bw >= W/2 for the window W = 640 rows, or 768 "
                     "for a dense band; PK_MATPOW=0 disables it)", MP_RMAX);
        return PK_ERR_UNSUPPORTED;
    }
    PK_CUDA(cudaMemsetAsync(&ctx->d_state->done, 0, sizeof(int), ctx->stream));
    return pk_launch_matpow(ctx, mat, k, d_base0, d_base1, 0);
}

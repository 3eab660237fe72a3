"""Row-partitioned multi-GPU layer: one process per GPU, ``torch.distributed`` for the plumbing, NCCL (inside
libpkrylov) on the data path.

Replaces ``MultiGpu`` of /root/reference/v3/gpu/mpi/common.py:46-171.  The reference keeps every vector replicated on
every rank and, per mat-vec, broadcasts the full x to each GPU (``memcpyPeer``, :144), gathers the pieces (:156) and
``comm.Allgather``s the result (:163).  Here vectors are SHARDED like the rows of A, a mat-vec exchanges only the
entries of x the local block references (the *halo*), and dot products are local partials + one small all-reduce
(precedent: /root/reference/v1/processes/adaptivekskipmrr.py:104-116).

The halo plan is computed with torch ops on whatever device the index arrays live on, so the same code is exercised
on CPU tensors over ``gloo`` (tests/test_dist_gloo.py) and on CUDA tensors over ``nccl``.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .._core import Context, Operator, _ptr
from .._lib import PK_IPC_HANDLE_BYTES, PkError, check


def _all_gather_int(value: int, group, device) -> list:
    world = dist.get_world_size(group)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [int(o.item()) for o in out]


def _comm_device(group) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def row_offsets_from_local(n_local: int, group=None) -> list:
    """Global row offsets [0, n_0, n_0+n_1, ...] of a contiguous block-row partition (the reference's partition,
    /root/reference/v3/gpu/mpi/common.py:104-131, without its N % P == 0 requirement)."""
    counts = _all_gather_int(n_local, group, _comm_device(group))
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    return offs


def build_halo_plan(rowptr: torch.Tensor, col_global: torch.Tensor, row_offsets: Sequence[int], rank: int,
                    group=None) -> dict:
    """From a local CSR block with GLOBAL column indices derive
      * ``col_local``   — columns renumbered to [owned 0..n_rows) | halo n_rows..n_rows+n_halo),
      * ``recv_off``    — per peer, which slice of the halo tail it fills (halo sorted by global index ⇒ by owner),
      * ``send_idx`` / ``send_off`` — per peer, which owned entries this rank must send (local indices),
      * ``interior``    — the longest run of rows that reference no halo column (overlapped with the exchange).
    One collective round (sizes) + point-to-point index lists; runs once per operator."""
    dev = col_global.device
    world = len(row_offsets) - 1
    row0, row1 = int(row_offsets[rank]), int(row_offsets[rank + 1])
    n_rows = row1 - row0
    colg = col_global.to(torch.int64)
    ext_mask = (colg < row0) | (colg >= row1)
    ext_cols = torch.unique(colg[ext_mask])                       # sorted ascending
    n_halo = int(ext_cols.numel())
    offs_t = torch.tensor(list(row_offsets), dtype=torch.int64, device=dev)
    owner = torch.searchsorted(offs_t, ext_cols, right=True) - 1   # ascending because ext_cols is sorted
    recv_counts = torch.bincount(owner, minlength=world)[:world] if n_halo else torch.zeros(world, dtype=torch.int64, device=dev)
    recv_off = [0]
    for c in recv_counts.tolist():
        recv_off.append(recv_off[-1] + int(c))
    # renumber columns
    col_local = colg - row0
    if n_halo:
        pos = torch.searchsorted(ext_cols, colg[ext_mask])
        col_local[ext_mask] = n_rows + pos
    col_local = col_local.to(torch.int32)
    # boundary rows -> interior run
    if n_halo:
        ext_pos = torch.nonzero(ext_mask, as_tuple=False).flatten()
        brow = torch.unique(torch.searchsorted(rowptr.to(torch.int64)[1:].contiguous(), ext_pos, right=True))
        edges = torch.cat([torch.tensor([-1], device=dev, dtype=torch.int64), brow.to(torch.int64),
                           torch.tensor([n_rows], device=dev, dtype=torch.int64)])
        gaps = edges[1:] - edges[:-1] - 1
        g = int(torch.argmax(gaps).item())
        interior = (int(edges[g].item()) + 1, int(edges[g + 1].item()))
    else:
        interior = (0, n_rows)

    # Tell every owner which of its entries we need.  ONE all-gather of the (padded) halo lists instead of pairwise
    # isend/irecv: point-to-point transfers open a new NCCL channel per peer pair on first use (seconds at 8 ranks — most
    # of the operator set-up time in r01), an all-gather rides on the communicator the process group already has.
    # Every rank then keeps, per peer, the wanted entries that fall into its own row range (lists are sorted, so each
    # peer's wishes are one contiguous slice found by two binary searches).
    cdev = _comm_device(group) if world > 1 else dev
    counts_mat = [torch.zeros(world, dtype=torch.int64, device=cdev) for _ in range(world)]
    if world > 1:
        dist.all_gather(counts_mat, recv_counts.to(cdev), group=group)
    else:
        counts_mat[0] = recv_counts.to(cdev)
    send_counts = [int(counts_mat[p][rank].item()) for p in range(world)]     # what peer p wants from me
    send_off = [0]
    for c in send_counts:
        send_off.append(send_off[-1] + c)
    send_idx = torch.zeros(send_off[-1], dtype=torch.int64, device=cdev)
    if world > 1:
        halo_sizes = [int(counts_mat[p].sum().item()) for p in range(world)]
        width = max(max(halo_sizes), 1)
        mine = torch.full((width,), -1, dtype=torch.int64, device=cdev)
        mine[:n_halo] = ext_cols.to(cdev)
        wanted = [torch.empty(width, dtype=torch.int64, device=cdev) for _ in range(world)]
        dist.all_gather(wanted, mine, group=group)
        for p in range(world):
            if p == rank or send_counts[p] == 0:
                continue
            lst = wanted[p][:halo_sizes[p]]
            lo_i = int(torch.searchsorted(lst, torch.tensor([row0], dtype=torch.int64, device=cdev)).item())
            hi_i = int(torch.searchsorted(lst, torch.tensor([row1], dtype=torch.int64, device=cdev)).item())
            if hi_i - lo_i != send_counts[p]:
                raise PkError("halo plan: peer's request list and counts disagree")
            send_idx[send_off[p]:send_off[p + 1]] = lst[lo_i:hi_i]
    send_idx = (send_idx - row0).to(torch.int32)
    if send_idx.numel() and (int(send_idx.min()) < 0 or int(send_idx.max()) >= n_rows):
        raise PkError("halo plan: a peer asked for rows this rank does not own")
    return {"col_local": col_local, "n_halo": n_halo, "recv_off": recv_off, "send_off": send_off,
            "send_idx": send_idx, "interior": interior, "halo_global": ext_cols, "n_rows": n_rows}


def band_ext_csr(pieces, first: int, n_ext: int):
    """CSR of the extended operator of a row-partitioned dense band (csrc/pk_matpow.cu, pk_mat_set_band_ext): `pieces` are
    consecutive row blocks (rowptr rebased to 0, GLOBAL column indices, values) — ghost rows above, owned rows, ghost rows
    below; columns are renumbered from `first` (the global index of the first row) and entries outside [0, n_ext) — the
    ones the outermost ghost rows lose — are dropped.  Pure tensor code (any device): -> (rowptr int64, col int64, val)."""
    dev = pieces[0][2].device
    cnt = torch.cat([(p[0][1:] - p[0][:-1]).to(torch.int64) for p in pieces])
    colx = torch.cat([p[1].to(torch.int64) for p in pieces]) - int(first)
    valx = torch.cat([p[2] for p in pieces])
    if cnt.numel() != n_ext:
        raise PkError("extended operator: the row blocks do not add up")
    inside = (colx >= 0) & (colx < n_ext)
    if not bool(inside.all()):
        row_of = torch.repeat_interleave(torch.arange(n_ext, device=dev), cnt)
        cnt = torch.zeros(n_ext, dtype=torch.int64, device=dev).index_add_(0, row_of[inside], torch.ones_like(row_of[inside]))
        colx, valx = colx[inside], valx[inside]
    rp = torch.zeros(n_ext + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt, 0, out=rp[1:])
    return rp, colx, valx


class DistOperator(Operator):
    """A contiguous row block of A on this rank's GPU plus its halo plan."""

    @classmethod
    def from_local_csr(cls, rowptr: torch.Tensor, col_global: torch.Tensor, val: torch.Tensor, n_global: int,
                       group=None, ctx: Optional[Context] = None, row_offsets: Optional[Sequence[int]] = None
                       ) -> "DistOperator":
        ctx = ctx or Context.get()
        ctx.init_comm(group)
        dev = ctx.torch_device
        rank, world = ctx.rank, ctx.n_ranks
        prof = os.environ.get("PK_SETUP_PROF") and rank == 0       # developer switch: where does operator set-up time go
        import time as _time
        marks = [("start", _time.perf_counter())]

        def mark(tag):
            if prof:
                torch.cuda.synchronize()
                marks.append((tag, _time.perf_counter()))
        h2d = sum(t.numel() * t.element_size() for t in (rowptr, col_global, val) if not t.is_cuda)
        rowptr = rowptr.to(dev, torch.int32).contiguous()
        col_global = col_global.to(dev)
        val = val.to(dev, torch.float64).contiguous()
        n_rows = rowptr.numel() - 1
        if row_offsets is None:
            row_offsets = row_offsets_from_local(n_rows, group) if world > 1 else [0, n_rows]
        if row_offsets[-1] != n_global:
            raise PkError(f"row blocks cover {row_offsets[-1]} rows, A has {n_global} columns")
        mark("h2d")
        plan = build_halo_plan(rowptr, col_global, row_offsets, rank, group)
        col_local = plan["col_local"].contiguous()
        mark("halo plan")
        op = cls(ctx)
        op.tensors = {"rowptr": rowptr, "col": col_local, "val": val}
        op.n_rows = n_rows
        op.n_global = int(n_global)
        op.row0 = int(row_offsets[rank])
        op.row_offsets = list(row_offsets)
        op.nnz = int(val.numel())
        op.kind = "csr"
        op.n_halo = plan["n_halo"]
        op.h2d_bytes = h2d
        op.plan = plan
        with torch.cuda.device(ctx.device):
            check(ctx.lib.pk_mat_csr(ctx.handle, C.byref(op.handle), n_rows, n_rows + op.n_halo, op.nnz,
                                     _ptr(rowptr), _ptr(col_local), _ptr(val)), "pk_mat_csr")
            mark("pk_mat_csr")
            if world > 1:
                send_idx_d = plan["send_idx"].to(dev).contiguous()
                send_idx_h = plan["send_idx"].cpu().contiguous()
                op.tensors["send_idx"] = send_idx_d
                so = np.asarray(plan["send_off"], dtype=np.int64)
                ro = np.asarray(plan["recv_off"], dtype=np.int64)
                check(ctx.lib.pk_mat_set_halo(op.handle, world, so.ctypes.data_as(C.c_void_p),
                                              ro.ctypes.data_as(C.c_void_p), _ptr(send_idx_d),
                                              C.c_void_p(send_idx_h.data_ptr()), plan["interior"][0],
                                              plan["interior"][1]), "pk_mat_set_halo")
                mark("set_halo")
                op._setup_matpow(rowptr, col_global, val, plan, group, world, rank, list(row_offsets))
                mark("matpow probe")
                # default: halo exchange fused into the SpMV kernel over NVLink peer memory; PK_HALO=nccl keeps
                # ncclSend/ncclRecv on a side stream (also the fallback when the peers' buffers cannot be mapped)
                op.halo_path = "nccl"
                if ctx.fused_allreduce and os.environ.get("PK_HALO", "p2p") != "nccl":
                    op._open_halo_push(group, world, rank, plan)
                mark("halo push buffers (IPC)")
        if prof:
            print("[pk setup] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms" for a, b in zip(marks, marks[1:])),
                  file=sys.stderr, flush=True)
        return op

    def _setup_matpow(self, rowptr, col_global, val, plan, group, world, rank, row_offsets):
        """Row-partitioned one-pass matrix powers (csrc/pk_matpow.cu): if the GLOBAL operator has a small bandwidth (every
        column within +-bw <= 127 of its row, rows of <= 28 nonzeros) and each rank only exchanges with its two neighbours,
        ship copies of the neighbours' 16*bw boundary rows of A to this rank once, so that a k-skip trip needs ONE exchange
        of depth 17*bw instead of one per basis level (redundant ghost-zone scheme).  Silently skipped otherwise."""
        ctx = self.ctx
        dev = ctx.torch_device
        self.matpow_ghost_rows = 0
        if os.environ.get("PK_MATPOW", "1") in ("0", ""):
            return
        n_rows, row0, n_global = self.n_rows, int(row_offsets[rank]), int(row_offsets[-1])
        colg32 = col_global.to(torch.int32).contiguous()
        info = (C.c_int * 2)()
        check(ctx.lib.pk_csr_band_info(ctx.handle, n_rows, row0, _ptr(rowptr), _ptr(colg32), info), "pk_csr_band_info")
        ro = plan["recv_off"]
        far = any(ro[p + 1] > ro[p] and abs(p - rank) != 1 for p in range(world))
        cdev = _comm_device(group)
        t = torch.tensor([info[0], info[1], -n_rows, 1 if far else 0], dtype=torch.int64, device=cdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        rmax, bw, n_min, any_far = int(t[0]), int(t[1]), -int(t[2]), int(t[3])
        if any_far or rmax > 28 or bw < 1 or bw > 127:
            return
        gr = min(16 * bw, (n_min - bw) // bw * bw)           # ghost rows per side; every rank must own gr + bw rows
        if gr < bw:
            return
        rp64 = rowptr.to(torch.int64)

        def rows_slice(lo, hi):
            a, b = int(rp64[lo]), int(rp64[hi])
            return ((rp64[lo:hi + 1] - a).to(torch.int32).cpu().numpy(), colg32[a:b].cpu().numpy(), val[a:b].cpu().numpy())

        mine = {"top": rows_slice(0, gr), "bot": rows_slice(n_rows - gr, n_rows)}
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        empty = (np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64))
        above = everyone[rank - 1]["bot"] if rank > 0 else empty
        below = everyone[rank + 1]["top"] if rank + 1 < world else empty
        keep = {}
        for tag, (rp_g, col_g, val_g) in (("above", above), ("below", below)):
            keep[tag] = (torch.from_numpy(np.ascontiguousarray(rp_g)).to(dev), torch.from_numpy(np.ascontiguousarray(col_g)).to(dev),
                         torch.from_numpy(np.ascontiguousarray(val_g)).to(dev))
        halo_global = plan["halo_global"].to(dev, torch.int64).contiguous()
        self.tensors["matpow_ghosts"] = (keep, halo_global)
        ra = gr if rank > 0 else 0
        rb = gr if rank + 1 < world else 0
        check(ctx.lib.pk_mat_set_matpow_ext(self.handle, bw, rmax, row0, n_global, _ptr(halo_global),
                                            ra, _ptr(keep["above"][0]), _ptr(keep["above"][1]), _ptr(keep["above"][2]),
                                            rb, _ptr(keep["below"][0]), _ptr(keep["below"][1]), _ptr(keep["below"][2])),
              "pk_mat_set_matpow_ext")
        self.matpow_ghost_rows = gr
        self._setup_band_ext(rowptr, colg32, val, keep, gr, rmax, bw, group, world, rank, row0, n_global)

    def _setup_band_ext(self, rowptr, colg32, val, keep, gr, rmax, bw, group, world, rank, row0, n_global):
        """Row-partitioned DENSE band (csrc/pk_matpow.cu, pk_mat_set_band_ext): a single-GPU operator over
        [ghost rows above | owned rows | ghost rows below] of this rank, columns renumbered from its first row and cut at
        its two ends.  The dense-band matrix-powers kernel and the fused k-skip MrR trip then run on it unchanged, with one
        ghost-zone exchange per trip; the owned rows come out exactly as on one GPU.  All ranks take the same decision."""
        ctx = self.ctx
        dev = ctx.torch_device
        self.band_ext_op = None
        cdev = _comm_device(group)
        # a full band of half width bw has n (2 bw + 1) - bw (bw + 1) entries: anything else keeps the general kernels and
        # never pays for the extended copy (the C side re-checks the structure row by row)
        tot = torch.tensor([int(val.numel())], dtype=torch.int64, device=cdev)
        dist.all_reduce(tot, group=group)
        full = int(tot.item()) == n_global * (2 * bw + 1) - bw * (bw + 1)
        want = (os.environ.get("PK_BAND_DIST", "1") not in ("0", "") and full and 2 * bw + 1 <= 27 and rmax <= 27
                and gr % 2 == 0)
        ext_op = None
        ok = 0
        if want:
            try:
                ra = gr if rank > 0 else 0
                rb = gr if rank + 1 < world else 0
                n_ext = self.n_rows + ra + rb
                first = row0 - ra                                   # global index of the extended operator's first row
                pieces = []
                if ra:
                    pieces.append(keep["above"])
                pieces.append((rowptr, colg32, val))
                if rb:
                    pieces.append(keep["below"])
                rp, colx, valx = band_ext_csr(pieces, first, n_ext)
                if int(rp[-1]) < 2 ** 31:
                    ext_op = Operator.from_csr_tensors(rp.to(torch.int32), colx.to(torch.int32).contiguous(),
                                                       valx.contiguous(), n_ext, ctx)
                    rc = ctx.lib.pk_mat_set_band_ext(self.handle, ext_op.handle, ra, rb)
                    ok = 1 if rc == 0 else 0
            except (PkError, RuntimeError):
                ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=cdev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        if int(t.item()) == 1:
            self.band_ext_op = ext_op                               # keeps the extended CSR arrays alive
        else:
            check(ctx.lib.pk_mat_set_band_ext(self.handle, None, 0, 0), "pk_mat_set_band_ext")

    def _open_halo_push(self, group, world, rank, plan):
        """Map the peers' halo receive buffers (CUDA IPC) so that the halo is exchanged by direct NVLink stores from
        inside libpkrylov instead of ncclSend/ncclRecv."""
        ctx = self.ctx
        dev = ctx.torch_device
        hb = C.create_string_buffer(PK_IPC_HANDLE_BYTES)
        check(ctx.lib.pk_mat_halo_p2p_handle(self.handle, hb), "pk_mat_halo_p2p_handle")
        cdev = _comm_device(group)
        mine = torch.tensor(list(hb.raw), dtype=torch.uint8, device=cdev)
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        ro = torch.tensor(plan["recv_off"], dtype=torch.int64, device=cdev)
        allro = [torch.zeros_like(ro) for _ in range(world)]
        dist.all_gather(allro, ro, group=group)
        dst_off = np.asarray([int(allro[q][rank].item()) for q in range(world)], dtype=np.int64)
        nhalo = np.asarray([int(allro[q][world].item()) for q in range(world)], dtype=np.int64)
        raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
        rc = ctx.lib.pk_mat_halo_p2p_open(self.handle, raw, dst_off.ctypes.data_as(C.c_void_p),
                                          nhalo.ctypes.data_as(C.c_void_p))
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=cdev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) != 1:                      # all ranks must take the same path
            check(ctx.lib.pk_mat_halo_p2p_disable(self.handle), "pk_mat_halo_p2p_disable")
            if rank == 0:
                print("parallel_krylov_b200: peer mapping of the halo receive buffers failed on some rank; "
                      "using the NCCL halo exchange", file=sys.stderr)
        else:
            self.halo_path = "p2p"

    @classmethod
    def from_local_dense(cls, local_a: torch.Tensor, group=None, ctx: Optional[Context] = None) -> "DistOperator":
        """Dense row block (local_N x N), the reference's ndarray branch (v3/gpu/mpi/common.py:124-125).  Every column
        is referenced, so the halo is the whole rest of x: columns are permuted to [owned | others]."""
        ctx = ctx or Context.get()
        ctx.init_comm(group)
        dev = ctx.torch_device
        rank, world = ctx.rank, ctx.n_ranks
        h2d = 0 if local_a.is_cuda else local_a.numel() * 8
        a = local_a.to(dev, torch.float64)
        n_rows, n_global = int(a.shape[0]), int(a.shape[1])
        row_offsets = row_offsets_from_local(n_rows, group) if world > 1 else [0, n_rows]
        row0, row1 = row_offsets[rank], row_offsets[rank + 1]
        perm = torch.cat([torch.arange(row0, row1, device=dev), torch.arange(0, row0, device=dev),
                          torch.arange(row1, n_global, device=dev)])
        a = a[:, perm].contiguous()
        op = cls(ctx)
        op.tensors = {"dense": a}
        op.n_rows, op.n_global, op.row0, op.row_offsets = n_rows, n_global, row0, list(row_offsets)
        op.nnz = a.numel()
        op.kind = "dense"
        op.n_halo = n_global - n_rows
        op.h2d_bytes = h2d
        with torch.cuda.device(ctx.device):
            check(ctx.lib.pk_mat_dense(ctx.handle, C.byref(op.handle), n_rows, n_global, _ptr(a), int(a.stride(0))),
                  "pk_mat_dense")
            if world > 1:
                # peer p sends its whole block; we send ours to everyone
                send_off, recv_off = [0], [0]
                for p in range(world):
                    cnt = row_offsets[p + 1] - row_offsets[p]
                    send_off.append(send_off[-1] + (n_rows if p != rank else 0))
                    recv_off.append(recv_off[-1] + (cnt if p != rank else 0))
                idx = torch.arange(n_rows, dtype=torch.int32).repeat(world - 1)
                idx_d = idx.to(dev)
                op.tensors["send_idx"] = idx_d
                so = np.asarray(send_off, dtype=np.int64)
                ro = np.asarray(recv_off, dtype=np.int64)
                check(ctx.lib.pk_mat_set_halo(op.handle, world, so.ctypes.data_as(C.c_void_p),
                                              ro.ctypes.data_as(C.c_void_p), _ptr(idx_d),
                                              C.c_void_p(idx.data_ptr()), 0, 0), "pk_mat_set_halo")
        return op

    @classmethod
    def from_any_local(cls, local_A, group=None, ctx: Optional[Context] = None) -> "DistOperator":
        if isinstance(local_A, DistOperator):
            return local_A
        if isinstance(local_A, np.ndarray):
            return cls.from_local_dense(torch.from_numpy(np.ascontiguousarray(local_A, dtype=np.float64)), group, ctx)
        if isinstance(local_A, torch.Tensor) and local_A.layout == torch.strided:
            return cls.from_local_dense(local_A, group, ctx)
        if isinstance(local_A, torch.Tensor) and local_A.layout == torch.sparse_csr:
            return cls.from_local_csr(local_A.crow_indices(), local_A.col_indices(), local_A.values(),
                                      local_A.shape[1], group, ctx)
        if isinstance(local_A, (tuple, list)) and len(local_A) == 4:
            rp, ci, va, n = local_A
            as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            return cls.from_local_csr(as_t(rp), as_t(ci), as_t(va), int(n), group, ctx)
        if hasattr(local_A, "tocsr"):
            m = local_A.tocsr()
            if not m.has_sorted_indices:
                m = m.sorted_indices()
            return cls.from_local_csr(torch.from_numpy(np.ascontiguousarray(m.indptr)),
                                      torch.from_numpy(np.ascontiguousarray(m.indices)),
                                      torch.from_numpy(np.ascontiguousarray(m.data, dtype=np.float64)),
                                      m.shape[1], group, ctx)
        raise PkError(f"unsupported local_A type {type(local_A)!r}")


def solve_dist(method: str, comm, local_A, b, x=None, tol=1e-05, maxiter=None, k=0, *, gather_x: bool = True, **kw):
    """Shared body of the ``(comm, local_A, b, ...)`` entry points (/root/reference/v3/gpu/mpi/cg.py:10).

    ``comm``: a torch.distributed process group (None = WORLD) standing in for the mpi4py communicator.
    ``b`` / ``x``: full length N on every rank (reference convention) or just this rank's rows.
    Returns on EVERY rank (the reference returns on rank 0 and calls ``exit(0)`` elsewhere, cg.py:59-68): x is the
    full-length solution (``gather_x=True``) or this rank's slice."""
    from .._core import solve
    if not dist.is_initialized():
        raise PkError("torch.distributed is not initialised (launch with torchrun, backend nccl)")
    import time as _time
    _t0 = _time.perf_counter()
    op = DistOperator.from_any_local(local_A, comm)
    ctx = op.ctx
    _t1 = _time.perf_counter()
    n, N = op.n_rows, op.n_global
    lo = op.row0

    def local_part(v):
        if v is None:
            return None
        size = v.size if isinstance(v, np.ndarray) else v.numel()
        if size == N and N != n:
            return v[lo:lo + n]
        if size == n:
            return v
        raise PkError(f"vector has {size} entries; expected N={N} or the local {n}")

    b_loc = local_part(b)
    x_loc = local_part(x) if isinstance(x, (np.ndarray, torch.Tensor)) else None
    if maxiter is None:
        maxiter = N
    x_out, info = solve(method, op, b_loc, x=x_loc, tol=tol, maxiter=maxiter, k=k, ctx=ctx, **kw)
    if os.environ.get("PK_SETUP_PROF") and ctx.rank == 0:
        torch.cuda.synchronize()
        print(f"[pk solve_dist] operator {1e3 * (_t1 - _t0):.1f} ms, solve() {1e3 * (_time.perf_counter() - _t1):.1f} ms "
              f"(loop {1e3 * info['time']:.1f} ms, pk_solve wall {1e3 * info['wall_time']:.1f} ms)", file=sys.stderr, flush=True)
    if gather_x and ctx.n_ranks > 1:
        sizes = [op.row_offsets[p + 1] - op.row_offsets[p] for p in range(ctx.n_ranks)]
        if len(set(sizes)) == 1:
            full = torch.empty(N, dtype=torch.float64, device=ctx.torch_device)
            xs = x_out.contiguous()
            torch.cuda.current_stream(ctx.device).synchronize()
            check(ctx.lib.pk_allgather(ctx.handle, _ptr(xs), _ptr(full), n), "pk_allgather")
            ctx.sync()
        else:
            mx = max(sizes)                      # uneven blocks: pad to the largest, gather, trim
            pad = torch.zeros(mx, dtype=torch.float64, device=ctx.torch_device)
            pad[:n] = x_out
            parts = [torch.empty(mx, dtype=torch.float64, device=ctx.torch_device) for _ in sizes]
            dist.all_gather(parts, pad, group=comm)
            full = torch.cat([p[:s] for p, s in zip(parts, sizes)])
        x_out = full
    return x_out, info
